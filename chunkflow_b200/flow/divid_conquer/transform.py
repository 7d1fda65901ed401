"""Test-time augmentation for host patch plug-ins (reference
chunkflow/flow/divid_conquer/transform.py:114-156): 8 transform sequences = the product of
(Lazy | Transpose) x (Lazy | FlipLR) x (Lazy | FlipUD); the patch backend runs on every transformed
copy, the outputs are transformed back and averaged (inferencer.py:422-431).

Two modes:

``mode='reference'`` (default, what ``augment=True`` / ``--augment`` selects) restates the reference
LITERALLY, so that results are identical to the reference's: its FlipLR / FlipUD call ``np.fliplr`` /
``np.flipud`` on ``arr[..., z, :, :]`` (transform.py:33-36,48-51).  For the 5-D (B, C, z, y, x) buffers the
Inferencer passes, that slice is 4-D (B, C, y, x): ``fliplr`` reverses axis 1, the CHANNEL axis, and
``flipud`` reverses axis 0, the BATCH axis -- no spatial flip happens.  ``backward`` applies the inverse
steps in FORWARD order (transform.py:147-156), which is harmless because these steps commute.

``mode='spatial'`` (explicit opt-in, ``augment='spatial'``) is the evidently intended augmentation:
spatial flips along x and y, undone in reverse order.  It does NOT reproduce the reference's numbers for
a real network; for flip-equivariant backends such as ``identity`` both modes agree.

The device network path implements the same two modes inside the kernels (CFB_AUGMENT_REFERENCE /
CFB_AUGMENT_SPATIAL in include/chunkflow_b200.h).
"""
from itertools import product
from typing import List

import numpy as np


def _transpose(a):
    return np.swapaxes(a, -1, -2)


def _flip_x(a):
    return a[..., ::-1]


def _flip_y(a):
    return a[..., ::-1, :]


def _keep(a):
    return a


def _flip_axis_of_slice(a, slice_axis):
    """Reverse axis `slice_axis` of ``a[..., z, :, :]`` -- expressed on the full array (the z axis sits at -3)."""
    nd_slice = a.ndim - 1
    # axes of the slice, in order, are all axes of `a` except -3
    axes = [ax for ax in range(a.ndim) if ax != a.ndim - 3]
    assert len(axes) == nd_slice
    target = axes[slice_axis]
    index = [slice(None)] * a.ndim
    index[target] = slice(None, None, -1)
    return a[tuple(index)]


def _ref_flipud(a):
    a = np.asarray(a)
    return _flip_axis_of_slice(a, 0)


def _ref_fliplr_any(a):
    a = np.asarray(a)
    return _flip_axis_of_slice(a, 1)


class TransformSequences:
    def __init__(self, mode: str = 'reference'):
        if mode in (True, 'reference'):
            steps = ((_keep, _transpose), (_keep, _ref_fliplr_any), (_keep, _ref_flipud))
            self.reverse_on_backward = False   # the reference applies the inverse steps in forward order
        elif mode == 'spatial':
            steps = ((_keep, _transpose), (_keep, _flip_x), (_keep, _flip_y))
            self.reverse_on_backward = True
        else:
            raise ValueError(f"unknown augmentation mode {mode!r}")
        self.mode = 'reference' if mode is True else mode
        self.transform_sequences = list(product(*steps))
        assert len(self.transform_sequences) == 8

    def forward(self, arr: np.ndarray) -> List[np.ndarray]:
        outs = []
        for seq in self.transform_sequences:
            a = arr
            for step in seq:
                a = step(a)
            outs.append(np.ascontiguousarray(a))
        return outs

    def backward(self, transformed_arrays: List[np.ndarray]) -> List[np.ndarray]:
        assert len(transformed_arrays) == len(self.transform_sequences)
        outs = []
        for a, seq in zip(transformed_arrays, self.transform_sequences):
            for step in (reversed(seq) if self.reverse_on_backward else seq):  # every step is an involution
                a = step(a)
            outs.append(np.ascontiguousarray(a))
        return outs
