"""BASELINE config #5 on real GPUs: ONE chunk split over several ranks (z-slabs of patch rows), partial sums of the
overlapping planes sent to their owner over NCCL and added by the library's kernel.  Needs >= 2 GPUs (skipped otherwise;
the host logic of the exchange is covered on CPU by tests/test_distributed_gloo.py)."""
import os
import socket

import numpy as np
import pytest

from conftest import MODEL_FILE

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, img, kw, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from chunkflow_b200 import Chunk, Inferencer
        from chunkflow_b200 import distributed as D
        inf = Inferencer(kw.pop("model"), None, kw.pop("input_patch_size"), device=rank, **kw)
        tm = {}
        part = D.infer_chunk_split(inf, Chunk(img, voxel_offset=(7, 0, 0)), timing=tm)
        if part is None:
            q.put((rank, -1, None, tm))
        else:
            q.put((rank, part.voxel_offset[0] - 7, part.array, tm))
    finally:
        dist.destroy_process_group()


def _run_split(world, img, **kw):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, img, dict(kw), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    parts = [(z0, arr) for _, z0, arr, _ in results if z0 >= 0]
    parts.sort(key=lambda t: t[0])
    covered = 0
    for z0, arr in parts:
        assert z0 == covered
        covered += arr.shape[1]
    assert covered == img.shape[0]
    return np.concatenate([a for _, a in parts], axis=1), [r[3] for r in results]


@pytest.mark.skipif(_gpu_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_split_chunk_over_nccl_matches_single_gpu_and_oracle(world, unet_model):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from chunkflow_b200 import Chunk, Inferencer
    from oracle import inferencer_oracle as O
    rng = np.random.default_rng(57)
    # 6 z-rows (the last one clamped): with 4 ranks the trailing one-row slabs overlap a non-neighbour
    img = rng.integers(0, 256, size=(36, 72, 80), dtype=np.uint8)
    geo = dict(input_patch_size=(8, 32, 32), output_patch_overlap=(2, 8, 8), num_output_channels=3)
    got, timings = _run_split(world, img, model=MODEL_FILE, framework="b200", batch_size=6, **geo)
    single = Inferencer(MODEL_FILE, None, (8, 32, 32), output_patch_overlap=(2, 8, 8), num_output_channels=3, framework="b200",
                        batch_size=6)(Chunk(img)).array
    oracle, _ = O.infer_chunk(img, framework="pytorch", model=unet_model, **geo)
    print("split vs single GPU", np.abs(got - single).max(), "| vs oracle", np.abs(got - oracle).max(), "| timing", timings)
    assert np.abs(got - single).max() <= 2e-6
    assert np.abs(got - oracle).max() <= 2e-4
    assert any(t.get("halo_bytes_sent", 0) > 0 for t in timings)
