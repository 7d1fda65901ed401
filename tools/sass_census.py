"""Per-kernel census of the Blackwell-specific SASS opcodes of the shipped library (tracked evidence that the hot path is
tcgen05 / TMEM / TMA code):   python tools/sass_census.py > profiles/r02_sass_census.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chunkflow_b200", "_native", "libchunkflow_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCSHIFT", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "REDG", "ATOMG", "FFMA", "HMMA"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    per[cur][o] += 1
    print("# SASS opcode census of chunkflow_b200/_native/libchunkflow_b200.so (sm_100a), `cuobjdump -sass`, round 2\n")
    print("UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4 (e4m3), UTCSHIFT = tcgen05.shift, LDTM / STTM = tcgen05.ld / st,")
    print("UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, REDG = red.global.\n")
    print("| kernel | " + " | ".join(OPS) + " |")
    print("|---|" + "---|" * len(OPS))
    tot = Counter()
    for fn, c in per.items():
        if not any(c[o] for o in OPS if o not in ("FFMA", "SYNCS", "REDG", "ATOMG")) and "cfb" not in fn:
            continue
        name = demangle(fn)
        name = re.sub(r"cfb::\(anonymous namespace\)::|cfb::", "", name)
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        if not any(c.values()):
            continue
        print("| %s | " % name + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
        tot.update(c)
    print("| **total** | " + " | ".join(str(tot[o]) for o in OPS) + " |")


if __name__ == "__main__":
    main()
