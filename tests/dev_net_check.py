"""Whole-network parity of a precision mode at real patch sizes (each case in a subprocess)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dev_umma_check import NET, ROOT, run

if __name__ == "__main__":
    precs = [int(a) for a in sys.argv[1:]] or [1]
    for prec in precs:
        print(run(NET % dict(root=ROOT, prec=prec, chunk=(36, 256, 256), patch=(20, 256, 256), ov=(4, 64, 64), batch=2), timeout=400), flush=True)
        print(run(NET % dict(root=ROOT, prec=prec, chunk=(40, 300, 260), patch=(32, 256, 256), ov=(8, 64, 64), batch=3), timeout=400), flush=True)
