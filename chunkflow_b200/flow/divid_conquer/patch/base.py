"""PatchInferencer plugin base (reference: chunkflow/flow/divid_conquer/patch/base.py:6-74)."""
import numpy as np

from .patch_mask import PatchMask


class PatchInferencerBase(object):
    """Geometry + bump mask shared by every patch backend.

    A backend maps ``(B, Cin, z, y, x) float32 in [0,1]`` to
    ``(B, Cout, oz, oy, ox) float32`` ALREADY cropped and multiplied by
    ``output_patch_mask`` (reference patch/pytorch.py:112-113).
    """

    def __init__(self, input_patch_size: tuple, output_patch_size: tuple, output_patch_overlap: tuple,
                 num_output_channels: int, dtype: str = "float32"):
        if output_patch_size is None:
            output_patch_size = input_patch_size
        assert len(output_patch_overlap) == 3
        assert len(input_patch_size) == 3
        assert len(output_patch_size) == 3
        self.input_patch_size = tuple(input_patch_size)
        self.output_patch_size = tuple(output_patch_size)
        self.output_patch_overlap = tuple(output_patch_overlap)
        self.num_output_channels = num_output_channels
        self.crop_margin = tuple((i - o) // 2 for i, o in zip(input_patch_size, output_patch_size))
        self.input_patch_overlap = tuple(o + 2 * c for o, c in zip(output_patch_overlap, self.crop_margin))
        self.input_patch_stride = tuple(p - o for p, o in zip(input_patch_size, self.input_patch_overlap))
        self.output_patch_stride = tuple(p - o for p, o in zip(output_patch_size, output_patch_overlap))
        self.output_patch_mask = PatchMask(self.output_patch_size, self.output_patch_overlap, dtype=dtype)
        self.output_patch_mask_numpy = self.output_patch_mask

    @property
    def compute_device(self) -> str:
        raise NotImplementedError

    def __call__(self, input_patch: np.ndarray) -> np.ndarray:
        raise NotImplementedError("this function should be overloaded by the inherited class!")

    def _reshape_patch_to_5d(self, input_patch):
        assert isinstance(input_patch, np.ndarray)
        if input_patch.ndim == 3:
            input_patch = input_patch.reshape((1, 1) + input_patch.shape)
        elif input_patch.ndim == 4:
            input_patch = input_patch.reshape((1,) + input_patch.shape)
        return input_patch

    def _crop_output_patch(self, output_patch: np.ndarray):
        cz, cy, cx = self.crop_margin
        return output_patch[:, :self.num_output_channels,
                            cz:output_patch.shape[-3] - cz,
                            cy:output_patch.shape[-2] - cy,
                            cx:output_patch.shape[-1] - cx]
