"""Summarise an `ncu --page raw --csv` export into a compact markdown table for profiles/.

    python scripts/summarize_ncu.py gpurun_out/r1_train_raw.csv > profiles/r1_train_ncu_summary.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time us", 1e-3),
    ("dram__bytes_read.sum", "dram rd MB", 1e-6),
    ("dram__bytes_write.sum", "dram wr MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
]


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, units, data = r, rows[i + 1], rows[i + 2:]
            break
    if hdr is None:
        print("no ncu table in", path)
        return
    idx = {}
    for key, _, _ in COLS:
        for j, h in enumerate(hdr):
            if h == key or h.endswith(key):
                idx[key] = j
                break
    kn = hdr.index("Kernel Name")
    names = [c[1] for c in COLS if c[0] in idx]
    print("| kernel | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for r in data:
        if len(r) <= kn:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("seald::", "")
        vals = []
        for key, _, scale in COLS:
            if key not in idx:
                continue
            v = to_float(r[idx[key]])
            u = units[idx[key]]
            if v is None:
                vals.append("-")
                continue
            if key.startswith("gpu__time_duration"):
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
            elif key.startswith("dram__bytes"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            vals.append(("%.1f" % v) if abs(v) < 1e6 else ("%.3g" % v))
        print("| %s | %s |" % (name[:70], " | ".join(vals)))


if __name__ == "__main__":
    main(sys.argv[1])
