"""BASELINE config #5 on real GPUs: one chunk split across the ranks along z with an NCCL halo
exchange of partial sums.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_split_chunk.py [--shape Z Y X] [--check]

Each rank reports its z-range; with --check rank 0 also runs the whole chunk alone and compares.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chunkflow_b200 import Chunk, Inferencer  # noqa: E402
from chunkflow_b200 import distributed as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=(128, 512, 512))
    ap.add_argument("--patch", type=int, nargs=3, default=(32, 256, 256))
    ap.add_argument("--overlap", type=int, nargs=3, default=(8, 64, 64))
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(20260922)
    img = rng.integers(0, 256, size=tuple(args.shape), dtype=np.uint8)   # same chunk on every rank (shared host chunk)
    inf = Inferencer(None, None, tuple(args.patch), output_patch_overlap=tuple(args.overlap), num_output_channels=3,
                     framework="b200", batch_size=args.batch_size, mask_output_chunk=True, device=local)
    D.infer_chunk_split(inf, Chunk(img))   # warm-up (autotune, allocations)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    part = D.infer_chunk_split(inf, Chunk(img))
    torch.cuda.synchronize(); dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    info = {"rank": rank, "z_range": None if part is None else [int(part.voxel_offset[0]), int(part.voxel_offset[0] + part.shape[1])]}
    err = None
    if args.check:
        whole = inf(Chunk(img)).array if rank == 0 else None
        gathered = [None] * world
        dist.all_gather_object(gathered, None if part is None else (int(part.voxel_offset[0]), part.array))
        if rank == 0:
            full = np.concatenate([g[1] for g in sorted((g for g in gathered if g is not None), key=lambda t: t[0])], axis=1)
            err = float(np.abs(full - whole).max())
    infos = [None] * world
    dist.all_gather_object(infos, info)
    if rank == 0:
        print(json.dumps({"config": "one chunk split along z with NCCL halo exchange", "shape": list(args.shape), "n_gpus": world,
                          "seconds": float(dt.item()), "mvoxels_per_s": float(np.prod(args.shape) / dt.item() / 1e6),
                          "ranks": infos, "max_abs_vs_single_gpu": err}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
