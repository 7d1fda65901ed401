"""Generate golden vectors by running the REAL reference (imported from /root/reference with
stubbed I/O modules, see oracle/reference_harness.py).  Run in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box; these small fixtures can.
"""
import hashlib
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.reference_harness import import_reference  # noqa: E402

MODEL_FILE = os.path.join(ROOT, "chunkflow_b200", "convnet", "unet3l.py")


def state_dict_digest(state) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(np.ascontiguousarray(state[k].detach().cpu().numpy().astype(np.float32)).tobytes())
    return h.hexdigest()


def main():
    import torch
    torch.cuda.is_available = lambda: False  # the reference moves to CUDA whenever it can (pytorch.py:41-46)
    Inferencer, Chunk, PatchMask = import_reference()
    quiet = io.StringIO()

    # 1. identity backend, non-aligned chunk, chunk-level mask (mirrors test_non_aligned_input_chunk)
    rng = np.random.default_rng(20260922)
    img = rng.integers(1, 255, size=(18, 56, 60), dtype=np.uint8)
    with redirect_stdout(quiet):
        with Inferencer(None, None, (8, 32, 32), output_patch_overlap=(2, 8, 8), num_output_channels=2,
                        batch_size=5, framework="identity", mask_output_chunk=True) as inf:
            out = inf(Chunk(img, voxel_offset=(3, 5, 7)))
            slices = [[[s.start, s.stop] for s in pair[0]] + [[s.start, s.stop] for s in pair[1]]
                      for pair in inf.patch_slices_list]
    np.savez_compressed(os.path.join(HERE, "identity_nonaligned.npz"), input=img, output=out.array,
                        voxel_offset=np.array(out.voxel_offset), patch_slices=np.array(slices))

    # 2. identity backend, aligned, no chunk mask (mirrors test_aligned_patch_num)
    img2 = rng.integers(1, 255, size=(6 * 2 + 2, 24 * 2 + 8, 24 * 2 + 8), dtype=np.uint8)
    with redirect_stdout(quiet):
        with Inferencer(None, None, (8, 32, 32), output_patch_overlap=(2, 8, 8), num_output_channels=2,
                        patch_num=(2, 2, 2), framework="identity", batch_size=3, mask_output_chunk=False) as inf:
            out2 = inf(Chunk(img2))
    np.savez_compressed(os.path.join(HERE, "identity_aligned.npz"), input=img2, output=out2.array,
                        voxel_offset=np.array(out2.voxel_offset))

    # 3. the 3-level U-Net through the reference's own `-f pytorch` CPU path
    img3 = rng.integers(0, 256, size=(12, 40, 48), dtype=np.uint8)
    with redirect_stdout(quiet):
        with Inferencer(MODEL_FILE, None, (8, 32, 32), output_patch_overlap=(2, 8, 8), num_output_channels=3,
                        batch_size=1, framework="pytorch", mask_output_chunk=True) as inf:
            out3 = inf(Chunk(img3))
            digest = state_dict_digest(inf.patch_inferencer.model.state_dict())
    np.savez_compressed(os.path.join(HERE, "unet3l_small.npz"), input=img3, output=out3.array)

    # 4. README config #1 geometry: sin chunk 64x256x256 is too big to store; keep its patch grid + sparse samples
    grids = {}
    for size, patch, ov in [((64, 256, 256), (20, 256, 256), (4, 64, 64)),
                            ((66, 455, 457), (32, 256, 256), (4, 64, 64)),
                            ((128, 512, 512), (20, 256, 256), (4, 64, 64)),
                            ((40, 300, 257), (32, 256, 256), (8, 64, 64))]:
        with redirect_stdout(quiet):
            inf = Inferencer(None, None, patch, output_patch_overlap=ov, num_output_channels=1,
                             framework="identity", mask_output_chunk=True)
            inf.input_size = size
            inf._construct_patch_slices_list((0, 0, 0))
        grids["x".join(map(str, size)) + "_" + "x".join(map(str, patch)) + "_" + "x".join(map(str, ov))] = [
            [s.start for s in pair[0]] for pair in inf.patch_slices_list]
    # patch-mask known answers (fp32) straight from the reference class
    masks = {}
    for ps, ov in [((20, 256, 256), (4, 64, 64)), ((32, 256, 256), (8, 64, 64)), ((10, 128, 128), (2, 32, 32)),
                   ((8, 32, 32), (2, 8, 8))]:
        m = np.asarray(PatchMask(ps, ov))
        masks["x".join(map(str, ps)) + "_" + "x".join(map(str, ov))] = dict(
            sha256=hashlib.sha256(m.tobytes()).hexdigest(), min=float(m.min()), max=float(m.max()),
            sum=float(m.sum(dtype=np.float64)), corner=float(m[0, 0, 0]), probe=float(m[1, ps[1] // 3, ps[2] // 5]))
    with open(os.path.join(HERE, "geometry.json"), "w") as f:
        json.dump(dict(patch_grids=grids, patch_masks=masks, unet3l_state_sha256=digest,
                       numpy=np.__version__, torch=torch.__version__), f, indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
