"""Summarise an `ncu --set full` report of tools/profile_step.py: one row per kernel launch (layer order) + traffic JSON.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/summarize_ncu.py /tmp/raw.csv profiles/<summary>.md profiles/<traffic>.json <patches per launch> "<source note>"
"""
import csv
import json
import re
import sys

LAYERS = ["enc0.0", "enc0.2", "pool0", "enc1.0", "enc1.2", "pool1", "enc2.0", "enc2.2", "up1", "dec1.0", "dec1.2", "up0", "dec0.0",
          "dec0.2+head+blend", "normalize"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"cfb::\(anonymous namespace\)::|cfb::<unnamed>::|<?unnamed>::|cfb::", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    raw, md, tj, patches, note = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    rows = list(csv.reader(open(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(head)}

    def val(r, name, scale=True):
        i = col[name]
        v = float(r[i].replace(",", "") or 0)
        return v * UNIT.get(units[i], 1.0) if scale else v

    out = ["| layer | kernel | grid x block | ms | tensor pipe active % | TC smem wavefronts % | LSU smem wavefronts % | DRAM MB (rd+wr) | DRAM % | SM % | regs |",
           "|---|---|---|---|---|---|---|---|---|---|---|"]
    layers = {}
    names = LAYERS if len(data) >= len(LAYERS) else [n for n in LAYERS if not n.startswith("pool")]  # round 2: pooling is fused into enc0.2 / enc1.2
    for layer, r in zip(names, data):
        k = short(r[col["Kernel Name"]])
        ms = val(r, "gpu__time_duration.sum")
        dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False)
        out.append("| %s | %s | %s x %s | %.5g | %.4g | %.4g | %.4g | %.0f | %.4g | %.4g | %d |" % (
            layer, k, r[col["launch__grid_size"]], r[col["launch__block_size"]], ms, tp,
            val(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", False),
            val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", False), dram / 1e6,
            val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
            val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed", False), int(val(r, "launch__registers_per_thread", False))))
        layers[layer] = {"kernel": k, "dram_bytes_per_patch": dram / patches, "tensor_pipe_active_pct": tp, "gpu_time_ms": ms}
    open(md, "w").write("# " + note + "\n\n" + "\n".join(out) + "\n")
    json.dump({"source": note, "layers": layers}, open(tj, "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
