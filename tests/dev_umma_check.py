"""(test infrastructure -- it uses the oracle, so it lives under tests/)  Development check of the tcgen05 path on a GPU box: every case runs in its own subprocess
with a timeout so that a trap / protocol bug cannot take the whole session down."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %(root)r)
from chunkflow_b200 import _native
prec, cin, cout, size = %(prec)d, %(cin)d, %(cout)d, %(size)r
rng = np.random.default_rng(cin * 1000 + cout)
x = rng.standard_normal((cin,) + size).astype(np.float32)
w = (rng.standard_normal((cout, cin, 3, 3, 3)) / np.sqrt(27 * cin)).astype(np.float32)
b = rng.standard_normal(cout).astype(np.float32)
eng = _native.Engine(input_patch_size=(8, 32, 32), output_patch_size=(8, 32, 32), output_patch_overlap=(2, 8, 8),
                     output_crop_margin=(0, 0, 0), framework=_native.FRAMEWORK_IDENTITY, precision=prec)
got = eng.debug_conv3(x, w, b, relu=True)
xt, wt = torch.from_numpy(x), torch.from_numpy(w)
if prec == 2:
    xt, wt = xt.half().float(), wt.half().float()
ref = torch.relu(torch.nn.functional.conv3d(xt[None], wt, torch.from_numpy(b), padding=1))[0].numpy()
err = np.abs(got - ref)
print(json.dumps(dict(prec=prec, cin=cin, cout=cout, size=size, max_abs=float(err.max()), mean_abs=float(err.mean()),
                      ref_absmax=float(np.abs(ref).max()), bad_frac=float((err > 1e-3).mean()))))
'''

NET = r'''
import sys, json, numpy as np
sys.path.insert(0, %(root)r)
from chunkflow_b200 import Chunk, Inferencer
from chunkflow_b200.lib import load_source
from oracle import inferencer_oracle as O
mf = %(root)r + "/chunkflow_b200/convnet/unet3l.py"
rng = np.random.default_rng(3)
img = rng.integers(0, 256, size=%(chunk)r, dtype=np.uint8)
kw = dict(input_patch_size=%(patch)r, output_patch_overlap=%(ov)r, num_output_channels=3)
inf = Inferencer(mf, None, framework="b200", batch_size=%(batch)d, precision=%(prec)d, **kw)
out = inf(Chunk(img))
ref, _ = O.infer_chunk(img, framework="pytorch", model=load_source(mf).load_model(None), **kw)
err = np.abs(out.array - ref)
print(json.dumps(dict(net=True, prec=%(prec)d, chunk=%(chunk)r, max_abs=float(err.max()), mean_abs=float(err.mean()), timing=inf.timing)))
'''


def run(code, timeout=120):
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout)
        out = (r.stdout.strip().splitlines() or ["<no stdout>"])[-1]
        if r.returncode != 0:
            out += " | rc=%d | " % r.returncode + " ".join(r.stderr.strip().splitlines()[-3:])[:600]
        return out
    except subprocess.TimeoutExpired:
        return "TIMEOUT"


if __name__ == "__main__":
    cases = [(16, 16, (3, 8, 40)), (16, 16, (4, 16, 70)), (32, 32, (3, 12, 20)), (16, 32, (2, 6, 128)),
             (16, 16, (5, 20, 256)), (64, 64, (6, 8, 8)), (64, 32, (4, 16, 16)), (32, 16, (3, 8, 130)), (32, 64, (3, 9, 64))]
    precs = [int(a) for a in sys.argv[1:]] or [2, 1]
    for prec in precs:
        for cin, cout, size in cases:
            print(run(CASE % dict(root=ROOT, prec=prec, cin=cin, cout=cout, size=size)), flush=True)
    for prec in precs:
        print(run(NET % dict(root=ROOT, prec=prec, chunk=(12, 40, 48), patch=(8, 32, 32), ov=(2, 8, 8), batch=4)), flush=True)
        print(run(NET % dict(root=ROOT, prec=prec, chunk=(36, 256, 256), patch=(20, 256, 256), ov=(4, 64, 64), batch=2), timeout=300), flush=True)
