"""Multi-GPU operation of the inference hot path.

The reference has no in-process parallelism: chunks are independent units handed to one
process per GPU by a task queue (flow/flow.py:584-620, distributed/kubernetes/deploy.yml:37).

* Many chunks (BASELINE config #4): :func:`chunks_for_rank` -- chunk k runs on rank k % world,
  no communication.
* One oversized chunk (config #5): :func:`infer_chunk_split` -- the patch GRID is cut into
  contiguous slabs of z-rows, every rank runs its rows on its own sub-chunk, neighbouring slabs
  overlap by the patch overlap in z, and the un-normalised partial sums (C channels) plus the
  partial weight sums of the overlapping planes are exchanged with the two neighbours
  (``torch.distributed`` send/recv: NCCL over NVLink on GPUs, gloo in the CPU tests) and added
  before each rank normalises its own planes.  This is the only collective on the path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def chunks_for_rank(num_chunks: int, rank: int, world: int) -> List[int]:
    """Static round-robin assignment of independent chunks to ranks."""
    return list(range(rank, num_chunks, world))


def axis_patch_starts(size: int, patch: int, overlap: int) -> List[int]:
    """Patch starts along one axis: ``range(0, size - overlap, stride)`` with the last start clamped
    back into the chunk (reference inferencer.py:268-271)."""
    stride = patch - overlap
    starts = []
    for i in range(0, size - overlap, stride):
        if i + patch > size:
            i = size - patch
        starts.append(i)
    return starts


@dataclass(frozen=True)
class Slab:
    rank: int
    row_begin: int      # first z-row of the patch grid owned by this rank
    row_end: int        # one past the last z-row
    z0: int             # first input plane this rank reads (== first output plane it touches)
    z1: int             # one past the last plane
    own_z0: int         # planes [own_z0, own_z1) of the final result are reported by this rank
    own_z1: int

    @property
    def empty(self) -> bool:
        return self.row_end <= self.row_begin


def plan_z_slabs(chunk_z: int, patch_z: int, overlap_z: int, world: int) -> List[Slab]:
    """Contiguous, balanced split of the z patch rows over `world` ranks."""
    starts = axis_patch_starts(chunk_z, patch_z, overlap_z)
    n = len(starts)
    slabs, begin = [], 0
    for r in range(world):
        rows = n // world + (1 if r < n % world else 0)
        end = begin + rows
        if rows == 0:
            slabs.append(Slab(r, begin, begin, 0, 0, 0, 0))
            continue
        z0, z1 = starts[begin], starts[end - 1] + patch_z
        slabs.append(Slab(r, begin, end, z0, z1, z0, z1))
        begin = end
    # ownership: a plane shared by two slabs is reported by the LOWER rank
    active = [s for s in slabs if not s.empty]
    fixed = {}
    for i, s in enumerate(active):
        own_z0 = s.z0 if i == 0 else max(s.z0, active[i - 1].z1)
        fixed[s.rank] = Slab(s.rank, s.row_begin, s.row_end, s.z0, s.z1, min(own_z0, s.z1), s.z1)
    return [fixed.get(s.rank, s) for s in slabs]


def exchange_halo(partial, weight, slabs: Sequence[Slab], rank: int, group=None):
    """Add the overlapping planes of the neighbouring slabs into `partial` (C, z, y, x) and
    `weight` (z, y, x) -- torch tensors covering planes [slab.z0, slab.z1) -- in place.

    Lower-rank contributions are added first so that every rank ends up with the same sum.
    """
    import torch
    import torch.distributed as dist

    me = slabs[rank]
    if me.empty:
        return
    active = [s for s in slabs if not s.empty]
    idx = [s.rank for s in active].index(rank)
    ops, recv_bufs = [], []
    for nb_idx in (idx - 1, idx + 1):
        if nb_idx < 0 or nb_idx >= len(active):
            continue
        other = active[nb_idx]
        lo, hi = max(me.z0, other.z0), min(me.z1, other.z1)
        if hi <= lo:
            continue
        sl = slice(lo - me.z0, hi - me.z0)
        send_p = partial[:, sl].contiguous()
        send_w = weight[sl].contiguous()
        recv_p, recv_w = torch.empty_like(send_p), torch.empty_like(send_w)
        ops += [dist.P2POp(dist.isend, send_p, other.rank, group), dist.P2POp(dist.isend, send_w, other.rank, group),
                dist.P2POp(dist.irecv, recv_p, other.rank, group), dist.P2POp(dist.irecv, recv_w, other.rank, group)]
        recv_bufs.append((sl, recv_p, recv_w, other.rank < rank))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for sl, recv_p, recv_w, other_is_lower in recv_bufs:
        if other_is_lower:   # (lower + mine) on both sides
            partial[:, sl] = recv_p + partial[:, sl]
            weight[sl] = recv_w + weight[sl]
        else:
            partial[:, sl] += recv_p
            weight[sl] += recv_w


def infer_chunk_split(inferencer, input_chunk, group=None, compute_partial: Optional[Callable] = None):
    """Config #5: one chunk split across the ranks of `group` along z.

    Returns this rank's part of the result as a ``Chunk`` covering output planes
    [own_z0, own_z1) (``None`` for a rank without rows).  Requires ``mask_output_chunk=True``
    geometry (crop margin 0) and equal input / output patch size.

    ``compute_partial(sub_array) -> (partial, weight)`` is injectable for the CPU (gloo) tests;
    by default the slab runs on this rank's GPU through ``cfb_infer_slab_device``.
    """
    import torch
    import torch.distributed as dist
    from chunkflow_b200.chunk import Chunk

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    arr = input_chunk.array
    pz, ovz = inferencer.input_patch_size[0], inferencer.input_patch_overlap[0]
    assert tuple(inferencer.output_crop_margin) == (0, 0, 0) and tuple(inferencer.output_patch_crop_margin) == (0, 0, 0)
    slabs = plan_z_slabs(arr.shape[-3], pz, ovz, world)
    me = slabs[rank]
    if me.empty:
        exchange_halo(None, None, slabs, rank, group)
        return None
    sub = np.ascontiguousarray(arr[me.z0:me.z1])
    if compute_partial is None:
        eng = inferencer.engine
        dev = torch.device("cuda", inferencer.device)
        d_in = torch.from_numpy(sub).to(dev)
        shape = eng.output_shape(sub.shape)
        partial = torch.empty(shape, dtype=torch.float32, device=dev)
        weight = torch.empty(shape[1:], dtype=torch.float32, device=dev)
        n_rows = me.row_end - me.row_begin
        stream = torch.cuda.current_stream(dev).cuda_stream
        eng.infer_slab_device(d_in.data_ptr(), sub.dtype, sub.shape, 0, n_rows, partial.data_ptr(), weight.data_ptr(), stream)
    else:
        partial, weight = compute_partial(sub)
    exchange_halo(partial, weight, slabs, rank, group)
    if compute_partial is None:
        eng.normalize_device(partial.data_ptr(), weight.data_ptr(), partial.shape, stream)
        torch.cuda.current_stream(dev).synchronize()
    else:
        partial /= weight
    own = partial[:, me.own_z0 - me.z0:me.own_z1 - me.z0]
    out = own.cpu().numpy() if hasattr(own, "cpu") else np.asarray(own)
    off = tuple(input_chunk.voxel_offset)
    return Chunk(np.ascontiguousarray(out), voxel_offset=(off[0] + me.own_z0, off[1], off[2]), voxel_size=input_chunk.voxel_size)
