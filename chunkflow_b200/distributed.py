"""Multi-GPU operation of the inference hot path.

The reference has no in-process parallelism: chunks are independent units handed to one
process per GPU by a task queue (flow/flow.py:584-620, distributed/kubernetes/deploy.yml:37).

* Many chunks (BASELINE config #4): :func:`chunks_for_rank` -- chunk k runs on rank k % world,
  no communication.
* One oversized chunk (config #5): :func:`infer_chunk_split` -- the patch GRID is cut into
  contiguous slabs of z-rows, every rank runs its rows on its own sub-chunk
  (``cfb_infer_slab_device``: un-normalised partial sums).  Every output plane has ONE owner (the lowest
  rank whose slab covers it); every other rank whose slab touches the plane sends its partial sums of that
  plane to the owner (``torch.distributed`` P2P: NCCL over NVLink on GPUs, gloo in the CPU tests), the owner
  adds them in rank order (``cfb_halo_add_device``), computes the weight volume of its planes locally
  (pure geometry, ``cfb_weight_volume_device`` -- no weight halo travels) and normalises.  This is the only
  exchange on the path; any number of slabs may overlap a plane (short trailing slabs whose last row is
  clamped back into the chunk overlap non-neighbours).
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
    # ownership: a plane covered by several slabs is reported by the LOWEST rank covering it
    fixed, covered_to = {}, 0
    for s in slabs:
        if s.empty:
            continue
        own_z0 = min(max(s.z0, covered_to), s.z1)
        fixed[s.rank] = Slab(s.rank, s.row_begin, s.row_end, s.z0, s.z1, own_z0, s.z1)
        covered_to = max(covered_to, s.z1)
    return [fixed.get(s.rank, s) for s in slabs]


def halo_transfers(slabs: Sequence[Slab]) -> List[Tuple[int, int, int, int]]:
    """(sender rank, owner rank, z_lo, z_hi) for every pair of slabs where the sender's slab touches planes the
    owner reports: the complete exchange plan, identical on every rank, ordered by (owner, sender)."""
    out = []
    for o in slabs:
        if o.empty or o.own_z1 <= o.own_z0:
            continue
        for c in slabs:
            if c.empty or c.rank == o.rank:
                continue
            lo, hi = max(o.own_z0, c.z0), min(o.own_z1, c.z1)
            if hi > lo:
                out.append((c.rank, o.rank, lo, hi))
    return out


def exchange_halo(partial, slabs: Sequence[Slab], rank: int, group=None, add: Optional[Callable] = None) -> int:
    """Complete the planes this rank OWNS: receive the partial sums of every other slab that touches them and add
    them (lower ranks first) into `partial` (C, z, y, x), a torch tensor covering planes [slab.z0, slab.z1); send this
    rank's planes that somebody else owns.  `add(dst, src)` performs ``dst += src`` (default: torch; on GPUs the
    caller passes the library's kernel).  Returns the bytes this rank sent."""
    import torch
    import torch.distributed as dist

    me = slabs[rank]
    ops, recvs, keep, sent = [], [], [], 0
    for sender, owner, lo, hi in halo_transfers(slabs):
        if me.empty:
            break
        if sender == rank:
            for c in range(partial.shape[0]):   # one contiguous run of planes per channel
                t = partial[c, lo - me.z0:hi - me.z0]
                ops.append(dist.P2POp(dist.isend, t, owner, group))
                sent += t.numel() * t.element_size()
        elif owner == rank:
            buf = torch.empty((partial.shape[0], hi - lo) + tuple(partial.shape[2:]), dtype=partial.dtype, device=partial.device)
            for c in range(partial.shape[0]):
                ops.append(dist.P2POp(dist.irecv, buf[c], sender, group))
            recvs.append((sender, lo, hi, buf))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for sender, lo, hi, buf in sorted(recvs, key=lambda r: r[0]):
        for c in range(partial.shape[0]):
            dst = partial[c, lo - me.z0:hi - me.z0]
            if add is None:
                dst += buf[c]
            else:
                add(dst, buf[c])
        keep.append(buf)
    return sent


def infer_chunk_split(inferencer, input_chunk, group=None, compute_partial: Optional[Callable] = None,
                      to_host: bool = True, timing: Optional[dict] = None):
    """Config #5: one chunk split across the ranks of `group` along z.

    Returns this rank's part of the result as a ``Chunk`` covering output planes [own_z0, own_z1)
    (``None`` for a rank without rows); with ``to_host=False`` the array stays a torch tensor on the
    GPU.  Requires ``mask_output_chunk=True`` geometry (crop margin 0) and equal input / output patch
    size.  Same results and errors as the single-GPU call: all-zero shortcut, the ``< 1.0001``
    assertion, myelin masking.

    ``compute_partial(sub_array) -> partial`` (un-normalised sums of this rank's rows on its sub-chunk)
    is injectable for the CPU (gloo) tests; by default the slab runs on this rank's GPU through
    ``cfb_infer_slab_device``.
    """
    import torch
    import torch.distributed as dist
    from chunkflow_b200 import _native
    from chunkflow_b200.chunk import Chunk

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    arr = input_chunk.array
    if arr.ndim == 4:
        assert arr.shape[0] == 1, 'one input channel'
        arr = arr[0]
    assert arr.ndim == 3
    pz, ovz = inferencer.input_patch_size[0], inferencer.input_patch_overlap[0]
    assert tuple(inferencer.output_crop_margin) == (0, 0, 0) and tuple(inferencer.output_patch_crop_margin) == (0, 0, 0)
    slabs = plan_z_slabs(arr.shape[0], pz, ovz, world)
    me = slabs[rank]
    on_gpu = compute_partial is None
    dev = torch.device("cuda", inferencer.device) if on_gpu else torch.device("cpu")
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if (on_gpu and timing is not None) else None
    partial = None
    if not me.empty:
        sub = np.ascontiguousarray(arr[me.z0:me.z1])
        if on_gpu:
            eng = inferencer.engine
            stream = torch.cuda.current_stream(dev).cuda_stream
            d_in = torch.from_numpy(sub).to(dev)
            shape = eng.output_shape(sub.shape)
            partial = torch.empty(shape, dtype=torch.float32, device=dev)
            if ev:
                ev[0].record()
            eng.infer_slab_device(d_in.data_ptr(), sub.dtype, sub.shape, 0, me.row_end - me.row_begin, partial.data_ptr(), 0, stream)
            flag[0] = int(eng.slab_nonzero(stream))
            if ev:
                ev[1].record()
        else:
            partial = compute_partial(sub)
            flag[0] = int(bool(np.any(sub)))
    elif ev:
        ev[0].record(); ev[1].record()
    # the reference returns zeros for an all-zero chunk (inferencer.py:387-393): every rank must agree
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    all_zero = int(flag.item()) == 0
    add = None
    if on_gpu:
        def add(dst, src):  # noqa: E306  (contiguous runs of planes)
            _native.halo_add_device(dst.data_ptr(), src.data_ptr(), dst.numel(), torch.cuda.current_stream(dev).cuda_stream)
    sent = exchange_halo(partial, slabs, rank, group, add)
    if ev:
        ev[2].record()
    if me.empty or me.own_z1 <= me.own_z0:
        return None
    own = partial[:, me.own_z0 - me.z0:me.own_z1 - me.z0]
    C = own.shape[0]
    drop_myelin = bool(inferencer.mask_myelin_threshold)
    if on_gpu:
        # normalise the owned planes (a contiguous run per channel only if the slab IS the owned range: copy otherwise)
        own = own.contiguous()
        weight = torch.empty(own.shape[1:], dtype=torch.float32, device=dev)
        eng.weight_volume_device(arr.shape, me.own_z0, me.own_z1, weight.data_ptr(), invert=True, stream=stream)
        try:
            eng.normalize_device(own.data_ptr(), weight.data_ptr(), own.shape, stream, weight_is_inverse=True, all_zero_input=all_zero)
        except _native.NativeError as err:
            if err.code == _native.ERR_OUTPUT_RANGE:
                raise AssertionError('output buffer should not be greater than 1') from err
            raise
        if ev:
            ev[3].record()
            torch.cuda.current_stream(dev).synchronize()
            timing.update(compute_ms=ev[0].elapsed_time(ev[1]), exchange_ms=ev[1].elapsed_time(ev[2]),
                          normalize_ms=ev[2].elapsed_time(ev[3]), halo_bytes_sent=sent)
        if drop_myelin:
            own = own[:-1]
        out = own.cpu().numpy() if to_host else own
    else:
        # CPU (test) path: the weight volume from the same geometry, in numpy
        from chunkflow_b200.flow.divid_conquer.patch.patch_mask import make_patch_mask
        own = own.clone() if hasattr(own, "clone") else np.array(own)
        w = _weight_planes_numpy(inferencer, arr.shape, me.own_z0, me.own_z1, make_patch_mask)
        own = own / torch.from_numpy(w) if hasattr(own, "numpy") else own / w
        if all_zero:
            own = own * 0
        assert bool((own < 1.0001).all()), 'output buffer should not be greater than 1'
        if drop_myelin:
            assert C == 4
            own = own[:-1] * (own[-1] < inferencer.mask_myelin_threshold)
        out = own.numpy() if hasattr(own, "numpy") else np.asarray(own)
    off = tuple(input_chunk.voxel_offset)
    offset = (off[0] + me.own_z0, off[1], off[2])
    if isinstance(out, np.ndarray):
        return Chunk(np.ascontiguousarray(out), voxel_offset=offset, voxel_size=input_chunk.voxel_size)
    from chunkflow_b200.chunk.device import DeviceChunk
    return DeviceChunk(out, voxel_offset=offset, voxel_size=input_chunk.voxel_size)


def _weight_planes_numpy(inferencer, chunk_zyx, z0: int, z1: int, make_patch_mask) -> np.ndarray:
    """Planes [z0, z1) of the weight volume of the FULL chunk (reference inferencer.py:294-333), host arithmetic for
    the gloo tests only."""
    ps, ov = tuple(inferencer.output_patch_size), tuple(inferencer.output_patch_overlap)
    mask = np.asarray(make_patch_mask(ps, ov), dtype=np.float32)
    w = np.zeros((z1 - z0,) + tuple(chunk_zyx[1:]), np.float32)
    for sz in axis_patch_starts(chunk_zyx[0], ps[0], ov[0]):
        lo, hi = max(sz, z0), min(sz + ps[0], z1)
        if hi <= lo:
            continue
        for sy in axis_patch_starts(chunk_zyx[1], ps[1], ov[1]):
            for sx in axis_patch_starts(chunk_zyx[2], ps[2], ov[2]):
                w[lo - z0:hi - z0, sy:sy + ps[1], sx:sx + ps[2]] += mask[lo - sz:hi - sz]
    return w
