"""ncu long-format CSV (gpu__time_duration, dram bytes, dram throughput) -> one markdown row per kernel: launches, time,
DRAM bytes per launch, achieved DRAM GB/s (bytes / time) against the measured HBM peak of MEASURED_PEAKS.json."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
        "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "%": 1.0}


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"cfb::\(anonymous namespace\)::|cfb::<unnamed>::|<?unnamed>::|cfb::", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    peak = 6584.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
    head = rows[0]
    ki, idi, mi, vi, ui = (head.index(n) for n in ("Kernel Name", "ID", "Metric Name", "Metric Value", "Metric Unit"))
    launches = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[idi], {"kernel": short(r[ki])})
        d[r[mi]] = float(r[vi].replace(",", "") or 0) * UNIT.get(r[ui], 1.0)
    agg = OrderedDict()
    for d in launches.values():
        if "gpu__time_duration.sum" not in d:
            continue
        a = agg.setdefault(d["kernel"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"]
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    print("| kernel | launches | ms per launch | DRAM MB per launch (rd+wr) | achieved DRAM GB/s | of measured HBM peak (%.0f GB/s) | ncu dram throughput %% |" % peak)
    print("|---|---|---|---|---|---|---|")
    for k, (n, t, b, pct) in agg.items():
        gbs = b / t / 1e9 if t > 0 else 0.0
        print("| %s | %d | %.4f | %.1f | %.0f | %.1f %% | %.1f |" % (k, n, t / n * 1e3, b / n / 1e6, gbs, 100 * gbs / peak, pct / n))


if __name__ == "__main__":
    main()
