"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, share).

    python tools/summarize_launches.py gpurun_out/launches.csv [bench_under_ncu.json] > profiles/<name>.md
"""
import csv
import json
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"cfb::\(anonymous namespace\)::|cfb::<unnamed>::|<?unnamed>::|cfb::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("(bool)", "").replace("(int)", "")


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
    head = rows[0]
    k_name, k_metric, k_val, k_unit = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= k_val or r[k_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[k_val].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[k_unit].replace("second", "s").replace("n", "n"), None)
        if scale is None:
            scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[r[k_unit]]
        n = short(r[k_name])
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + v * scale)
    total = sum(t for _, t in agg.values())
    print("| kernel | launches | ncu total ms | ncu share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.1f} | {100 * t / total:.1f} % |")
    print(f"| **total** | {sum(c for c, _ in agg.values())} | {total:.1f} | 100 % |")
    conv = sum(t for n, (_, t) in agg.items() if n.startswith("conv3_"))
    line = f"\n3x3x3 tcgen05 convolution kernels: ncu share {100 * conv / total:.1f} %"
    if len(sys.argv) > 2:
        d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
        k = d["kernel_ms_per_chunk"]
        live = sum(v for n, v in k.items() if n[:3] in ("enc", "dec") and n != "enc0.0")
        line += f"  vs  live CUDA-event share {100 * live / sum(k.values()):.1f} % (live per-class ms of the same run: {json.dumps(k)})."
    print(line)


if __name__ == "__main__":
    main()
