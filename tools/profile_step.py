"""One small inference for profiling under ncu (never a bench number).  The first call autotunes; the
second one runs between cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees only it."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkflow_b200 import Chunk, Inferencer

shape = tuple(int(v) for v in os.environ.get("CFB_PROFILE_CHUNK", "64,512,512").split(","))
batch = int(os.environ.get("CFB_BENCH_BATCH", 12))
rng = np.random.default_rng(0)
img = rng.integers(0, 256, size=shape, dtype=np.uint8)
inf = Inferencer(None, None, (32, 256, 256), output_patch_overlap=(8, 64, 64), num_output_channels=3, framework="b200",
                 batch_size=batch, precision=os.environ.get("CHUNKFLOW_B200_PRECISION"))
out = inf(Chunk(img))          # warm-up + tile autotune
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = inf(Chunk(img))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", out.shape, inf.timing)
