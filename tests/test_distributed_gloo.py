"""world_size>1 gloo tests (CPU) of the multi-GPU host logic: chunk sharding and the z-slab
split with the halo exchange of partial sums.  The per-rank slab arithmetic is
injected from the oracle here; on GPUs it is ``cfb_infer_slab_device`` (tests/test_gpu_parity.py,
tests/test_gpu_multi.py)."""
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
    for chunk_z, world in [(512, 8), (40, 2), (20, 3), (8, 4), (1024, 8), (66, 2), (100, 4)]:
        pz, ov = (32, 8) if chunk_z > 100 else ((32, 16) if chunk_z == 100 else (8, 2))
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


def test_halo_transfers_reach_non_neighbours():
    """plan_z_slabs(100, 32, 16, 4) gives rank1 = [32, 80) and rank3 = [68, 100): planes 68..79 are owned
    by rank 1 and touched by rank 2 AND rank 3 -- every contributor must appear in the exchange plan."""
    slabs = D.plan_z_slabs(100, 32, 16, 4)
    assert [(s.z0, s.z1) for s in slabs] == [(0, 48), (32, 80), (64, 96), (68, 100)]
    assert [(s.own_z0, s.own_z1) for s in slabs] == [(0, 48), (48, 80), (80, 96), (96, 100)]
    plan = D.halo_transfers(slabs)
    assert (3, 1, 68, 80) in plan and (2, 1, 64, 80) in plan and (1, 0, 32, 48) in plan and (3, 2, 80, 96) in plan
    # every plane of every slab reaches its owner exactly once
    for s in slabs:
        for z in range(s.z0, s.z1):
            owner = next(o.rank for o in slabs if o.own_z0 <= z < o.own_z1)
            n = sum(1 for c, o, lo, hi in plan if c == s.rank and o == owner and lo <= z < hi)
            assert n == (0 if owner == s.rank else 1), (s.rank, z)


def test_chunks_for_rank_round_robin():
    got = sorted(k for r in range(3) for k in D.chunks_for_rank(8, r, 3))
    assert got == list(range(8)) and D.chunks_for_rank(8, 1, 3) == [1, 4, 7]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Ramp(torch.nn.Module):
    """A patch backend whose output depends on the position INSIDE the patch and on the channel (so that a wrong or
    missing halo contribution changes the blended result, unlike the identity backend whose weighted average is the
    input whatever subset of patches is summed)."""
    def forward(self, x):
        z = torch.linspace(0.1, 0.9, x.shape[-3]).view(1, 1, -1, 1, 1)
        y = torch.linspace(0.2, 0.8, x.shape[-2]).view(1, 1, 1, -1, 1)
        out = torch.cat([x * z, x * y * 0.5 + 0.25 * z], dim=1)
        return out


def _worker(rank, world, port, img, patch, overlap, expected, q, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chunkflow_b200.chunk import Chunk

        class HostGeometry:   # the attributes infer_chunk_split reads from an Inferencer
            input_patch_size, input_patch_overlap = patch, overlap
            output_patch_size, output_patch_overlap = patch, overlap
            output_crop_margin = output_patch_crop_margin = (0, 0, 0)
            mask_myelin_threshold = None

        def compute_partial(sub):
            if not np.any(sub):   # the oracle short-cuts an all-zero chunk itself; a slab may be all zero in a non-zero chunk
                return torch.zeros((2,) + sub.shape)
            kw = dict(framework="identity") if mode == "identity" else dict(framework="pytorch", model=_Ramp())
            s, _ = O.infer_chunk(sub, input_patch_size=patch, output_patch_overlap=overlap, num_output_channels=2,
                                 raw_sums=True, **kw)
            return torch.from_numpy(s)

        chunk = Chunk(img[None] if mode == "identity" else img, voxel_offset=(5, 0, 0))   # 4-D (1, z, y, x) input is accepted
        part = D.infer_chunk_split(HostGeometry(), chunk, compute_partial=compute_partial)
        if part is None:
            q.put((rank, -1, 0, 0.0))
            return
        z0 = part.voxel_offset[0] - 5
        ref = expected[:, z0:z0 + part.shape[1]]
        q.put((rank, z0, part.shape[1], float(np.abs(part.array - ref).max())))
    finally:
        dist.destroy_process_group()


def _run(world, img, patch, overlap, expected, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, img, patch, overlap, expected, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = 0
    for rank, z0, nz, err in results:
        if z0 < 0:
            continue
        assert z0 == covered and err <= 2e-6, (rank, z0, nz, err)
        covered += nz
    assert covered == img.shape[0]


@pytest.mark.parametrize("world", [2, 3])
def test_split_chunk_with_halo_exchange_matches_whole_chunk(world):
    rng = np.random.default_rng(42)
    img = rng.integers(1, 255, size=(27, 40, 44), dtype=np.uint8)   # 4 z-rows, the last one clamped
    expected, _ = O.infer_chunk(img, input_patch_size=PATCH, output_patch_overlap=OVERLAP, num_output_channels=2,
                                framework="identity")
    _run(world, img, PATCH, OVERLAP, expected, "identity")


def test_split_chunk_non_neighbour_overlap_position_dependent_backend():
    """z = 100, patch 32, overlap 16, world 4: the clamped last row (rank 3) overlaps rank 1's planes.  With a
    position-dependent backend a dropped contribution shows up as a wrong value (it did not with `identity`)."""
    patch, overlap = (32, 16, 16), (16, 4, 4)
    rng = np.random.default_rng(43)
    img = rng.integers(1, 255, size=(100, 28, 28), dtype=np.uint8)
    expected, _ = O.infer_chunk(img, input_patch_size=patch, output_patch_overlap=overlap, num_output_channels=2,
                                framework="pytorch", model=_Ramp())
    _run(4, img, patch, overlap, expected, "ramp")


def test_split_chunk_all_zero_input_gives_zeros():
    img = np.zeros((27, 40, 44), np.uint8)
    _run(2, img, PATCH, OVERLAP, np.zeros((2, 27, 40, 44), np.float32), "identity")
