"""Per-launch table (in launch order) from an ncu --csv metrics log: time, DRAM bytes and throughput, tensor-pipe %.  For profiles/."""
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"b200::(\(anonymous namespace\)::)?", "", name)
    return re.sub(r"\(.*", "", name)[:52]


def main(path, peak_gbs=6552.0):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    launches = {}
    order = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        key = r[col["ID"]]
        if key not in launches:
            launches[key] = {"name": short(r[col["Kernel Name"]])}
            order.append(key)
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)      # -> us
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        launches[key][m] = v
    tot_t = tot_b = 0.0
    print(f"{'id':>4s} {'kernel':52s} {'time us':>9s} {'DRAM MB':>9s} {'TB/s':>6s} {'% of HBM peak':>13s} {'tensor pipe %':>13s}")
    for k in order:
        d = launches[k]
        t = d.get("gpu__time_duration.sum", 0.0)
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        tot_t += t
        tot_b += b
        bw = b / (t * 1e-6) / 1e12 if t else 0.0
        print(f"{k:>4s} {d['name']:52s} {t:9.1f} {b / 1e6:9.1f} {bw:6.2f} {100 * bw * 1e3 / peak_gbs:13.1f} {tp:13.1f}")
    print(f"total {tot_t / 1e3:.2f} ms, DRAM {tot_b / 1e9:.2f} GB")


if __name__ == "__main__":
    main(sys.argv[1])
