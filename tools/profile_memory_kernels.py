"""One pass over the HBM-bound kernels of the hot path and of the operators either side, for `ncu` (never a bench number):

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_memory_kernels.csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
        python tools/profile_memory_kernels.py
    python tools/summarize_memory_kernels.py gpurun_out/r02_memory_kernels.csv > profiles/r02_memory_kernels.md

identity_blend / weight_volume / normalize / any_nonzero (identity backend, 96x1024x1024 chunk, 12 patches in flight),
extract_patches / blend_patches (fp32 `simt` network path and the host plug-in path), halo_add (multi-GPU halo planes),
normalize-contrast / maskout / crop-margin / quantize on a 512x1024x1024 device chunk.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkflow_b200 import Chunk, Inferencer, _native  # noqa: E402
from chunkflow_b200.chunk.device import DeviceChunk  # noqa: E402

rng = np.random.default_rng(0)
geo = dict(input_patch_size=(32, 256, 256), output_patch_overlap=(8, 64, 64), num_output_channels=3)
big = rng.integers(1, 255, size=(96, 1024, 1024), dtype=np.uint8)
small = rng.integers(1, 255, size=(56, 448, 448), dtype=np.uint8)
ident = Inferencer(None, None, framework="identity", batch_size=12, **geo)
simt = Inferencer(None, None, framework="b200", batch_size=4, precision="simt", **geo)


class HostIdentity:
    compute_device = "host"

    def __init__(self, mask):
        self.mask = mask

    def __call__(self, patch):
        return np.repeat(patch * self.mask, 3, axis=1)


plug = Inferencer(HostIdentity(_native.make_patch_mask((32, 256, 256), (8, 64, 64))), None, framework="prebuilt", batch_size=4, **geo)
img = DeviceChunk(torch.from_numpy(rng.integers(0, 256, size=(512, 1024, 1024), dtype=np.uint8)).cuda())
aff = DeviceChunk(torch.rand((3, 256, 1024, 1024), device="cuda"), layer_type="affinity_map")
mask = DeviceChunk(torch.from_numpy(rng.integers(0, 2, size=(128, 512, 512), dtype=np.uint8)).cuda(), voxel_size=(2, 2, 2))
a = torch.rand(100 << 20, device="cuda")
b = torch.rand(100 << 20, device="cuda")


def work():
    ident(Chunk(big))
    simt(Chunk(small))
    plug(Chunk(small))
    _native.halo_add_device(a.data_ptr(), b.data_ptr(), a.numel(), torch.cuda.current_stream().cuda_stream)
    x = DeviceChunk(img.tensor.clone(), voxel_size=(1, 1, 1))
    x.normalize_contrast()
    m8 = DeviceChunk(img.tensor[:256].clone(), voxel_size=(1, 1, 1))
    mask.maskout(m8)
    mf = DeviceChunk(aff.tensor.clone(), voxel_size=(1, 1, 1), layer_type="affinity_map")
    mask.maskout(mf)
    mf.crop_margin((8, 64, 64))
    aff.quantize(mode="xy")
    torch.cuda.synchronize()


work()   # warm-up (allocations, cached tables)
torch.cuda.cudart().cudaProfilerStart()
work()
torch.cuda.cudart().cudaProfilerStop()
print("done")
