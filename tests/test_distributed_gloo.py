"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chunk sharding and the z-slab
split with halo exchange of partial sums.  The per-rank slab arithmetic is
injected from the oracle here; on GPUs it is ``cfb_infer_slab_device`` (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chunkflow_b200 import distributed as D
from oracle import inferencer_oracle as O

PATCH, OVERLAP = (8, 32, 32), (2, 8, 8)


def test_plan_z_slabs_covers_every_row_once():
    for chunk_z, world in [(512, 8), (40, 2), (20, 3), (8, 4), (1024, 8), (66, 2)]:
        pz, ov = 32 if chunk_z > 100 else 8, 8 if chunk_z > 100 else 2
        starts = D.axis_patch_starts(chunk_z, pz, ov)
        slabs = D.plan_z_slabs(chunk_z, pz, ov, world)
        rows = [r for s in slabs for r in range(s.row_begin, s.row_end)]
        assert rows == list(range(len(starts)))
        active = [s for s in slabs if not s.empty]
        assert active[0].own_z0 == 0 and active[-1].own_z1 == chunk_z
        for a, b in zip(active, active[1:]):
            assert a.own_z1 == b.own_z0 and b.z0 < a.z1       # contiguous ownership, overlapping extents
        sizes = [s.row_end - s.row_begin for s in slabs]
        assert max(sizes) - min(sizes) <= 1


def test_chunks_for_rank_round_robin():
    got = sorted(k for r in range(3) for k in D.chunks_for_rank(8, r, 3))
    assert got == list(range(8)) and D.chunks_for_rank(8, 1, 3) == [1, 4, 7]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, img, expected, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chunkflow_b200.chunk import Chunk

        class HostGeometry:   # the attributes infer_chunk_split reads from an Inferencer
            input_patch_size, input_patch_overlap = PATCH, OVERLAP
            output_crop_margin = output_patch_crop_margin = (0, 0, 0)

        def compute_partial(sub):
            s, w = O.infer_chunk(sub, input_patch_size=PATCH, output_patch_overlap=OVERLAP, num_output_channels=2,
                                 framework="identity", raw_sums=True)
            return torch.from_numpy(s), torch.from_numpy(w)

        part = D.infer_chunk_split(HostGeometry(), Chunk(img, voxel_offset=(5, 0, 0)), compute_partial=compute_partial)
        z0 = part.voxel_offset[0] - 5
        ref = expected[:, z0:z0 + part.shape[1]]
        q.put((rank, z0, part.shape[1], float(np.abs(part.array - ref).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_split_chunk_with_halo_exchange_matches_whole_chunk(world):
    rng = np.random.default_rng(42)
    img = rng.integers(1, 255, size=(27, 40, 44), dtype=np.uint8)   # 4 z-rows, the last one clamped
    expected, _ = O.infer_chunk(img, input_patch_size=PATCH, output_patch_overlap=OVERLAP, num_output_channels=2,
                                framework="identity")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, img, expected, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = 0
    for rank, z0, nz, err in results:
        assert z0 == covered and err <= 2e-6, (rank, z0, nz, err)
        covered += nz
    assert covered == img.shape[0]
