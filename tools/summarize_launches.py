"""Summarise an ncu CSV launch list (gpu__time_duration, tensor-pipe activity, DRAM bytes) with per-launch labels.
usage: summarize_launches.py file.csv "title" label1,label2,..."""
import csv
import sys


def load(f):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ii, gi = (h.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    ui = h.index("Metric Unit")
    d = {}
    for r in rows[1:]:
        e = d.setdefault(int(r[ii]), {"k": r[ki].split("(")[0].replace("void ", "").replace("damc::", ""), "grid": r[gi]})
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("dram__bytes"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(r[ui], 1e-6)
        if r[mi].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        e[r[mi]] = v
    return [d[k] for k in sorted(d)]


def main():
    f, title = sys.argv[1], sys.argv[2]
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    L = load(f)
    tot = sum(e["gpu__time_duration.sum"] for e in L)
    print(title)
    for i, e in enumerate(L):
        t = e["gpu__time_duration.sum"]
        rd, wr = e.get("dram__bytes_read.sum", 0.0), e.get("dram__bytes_write.sum", 0.0)
        tp = e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        nm = names[i] if names and i < len(names) else ""
        print(f"{nm:24s} {e['k'][:32]:32s} grid {e['grid']:>12s} {t:8.1f} us {100 * t / tot:5.1f}%  dram rd {rd:7.1f} MB wr {wr:7.1f} MB"
              f" ({(rd + wr) / t if t else 0:5.2f} TB/s)  tensor-pipe {tp:5.1f}%")
    print(f"total {tot:.1f} us")


if __name__ == "__main__":
    main()
