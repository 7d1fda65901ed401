"""Where does the time of the pageable end-to-end call go?  (development probe, GPU box)"""
import os
import sys
import time
import threading

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def first_touch(nbytes, threads):
    a = np.empty(nbytes, np.uint8)
    step = nbytes // threads
    def work(i):
        a[i * step:(i + 1) * step] = 1
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts: t.start()
    for t in ts: t.join()
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    del a
    return dt, time.perf_counter() - t1


if __name__ == "__main__":
    for f in ("enabled", "defrag", "shmem_enabled"):
        try:
            print("THP", f, open(f"/sys/kernel/mm/transparent_hugepage/{f}").read().strip())
        except Exception as e:
            print("THP", f, e)
    n = 12 * (1 << 30)
    for th in (1, 8, 16, 32):
        dt, free = first_touch(n if th > 1 else n // 8, th)
        print(f"first touch {(n if th > 1 else n // 8) / 2**30:.1f} GiB with {th} threads: {dt:.3f} s, free {free:.3f} s")
    import torch
    from chunkflow_b200 import Chunk, Inferencer
    shape = tuple(int(v) for v in os.environ.get("CFB_PROBE_CHUNK", "1024,1024,1024").split(","))
    img = np.random.default_rng(0).integers(0, 256, size=shape, dtype=np.uint8)
    inf = Inferencer(None, None, (32, 256, 256), output_patch_overlap=(8, 64, 64), num_output_channels=3, framework="b200",
                     batch_size=12)
    out = inf(Chunk(img)); del out
    for tag in ("staged", "staged"):
        t0 = time.perf_counter()
        out = inf(Chunk(img))
        t1 = time.perf_counter()
        print(tag, "pageable call", round(t1 - t0, 3), "s timing", inf.timing)
        t2 = time.perf_counter()
        del out
        print("   free of the result", round(time.perf_counter() - t2, 3), "s")
    pin = torch.empty(inf.engine.output_shape(shape), dtype=torch.float32, pin_memory=True)
    out = inf(Chunk(img), output_buffer=pin.numpy())
    t0 = time.perf_counter()
    out = inf(Chunk(img), output_buffer=pin.numpy())
    print("pinned output, pageable input", round(time.perf_counter() - t0, 3), "s timing", inf.timing)
