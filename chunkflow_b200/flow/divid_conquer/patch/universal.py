"""User plugin backend (reference: chunkflow/flow/divid_conquer/patch/universal.py:8-69).

The model file defines ``class PatchInferencer`` with ``__init__(self, model_weight_file,
output_patch_mask)`` and ``__call__(self, input_patch) -> np.ndarray`` (already masked).
"""
import platform

import numpy as np

from chunkflow_b200.lib import load_source

from .base import PatchInferencerBase


class Universal(PatchInferencerBase):
    def __init__(self, convnet_model: str, convnet_weight_path: str, input_patch_size: tuple,
                 output_patch_size: tuple, output_patch_overlap: tuple, num_output_channels: int = 1,
                 dtype: str = "float32", bump: str = "wu"):
        assert bump == "wu"
        super().__init__(input_patch_size, output_patch_size, output_patch_overlap, num_output_channels, dtype=dtype)
        net_source = load_source(convnet_model)
        assert hasattr(net_source, "PatchInferencer")
        self.patch_inferencer = net_source.PatchInferencer(convnet_weight_path, self.output_patch_mask)

    @property
    def compute_device(self):
        if hasattr(self.patch_inferencer, "compute_device"):
            return self.patch_inferencer.compute_device
        return platform.processor()

    def __call__(self, input_patch):
        input_patch = self._reshape_patch_to_5d(input_patch)
        output_patch = self.patch_inferencer(input_patch)
        assert isinstance(output_patch, np.ndarray)
        return output_patch
