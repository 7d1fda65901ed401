"""Test-time augmentation for host patch plugins: the 8 combinations of
{transpose y<->x} x {flip x} x {flip y}, averaged (reference
chunkflow/flow/divid_conquer/transform.py:114-156).

Deliberate deviation (documented in DESIGN.md): the reference's FlipLR / FlipUD call
``np.fliplr`` / ``np.flipud`` on ``arr[..., z, :, :]`` which, for the 5-D (B, C, z, y, x)
buffers the Inferencer passes, reverse the CHANNEL and BATCH axes -- no spatial flip happens
and output channels get permuted (transform.py:33-36,48-51).  Here the flips are the
intended spatial ones, and ``backward`` undoes each sequence in REVERSE order (the
reference applies the inverse steps in forward order, transform.py:147-156, which is only
correct because its flips commute with the transpose).  For flip-equivariant backends such
as ``identity`` both give the same result.
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


class TransformSequences:
    def __init__(self):
        self.transform_sequences = list(product((_keep, _transpose), (_keep, _flip_x), (_keep, _flip_y)))
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
            for step in reversed(seq):  # every step is an involution
                a = step(a)
            outs.append(np.ascontiguousarray(a))
        return outs
