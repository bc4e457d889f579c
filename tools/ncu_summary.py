"""Compact per-kernel summary of an .ncu-rep (ncu --set full) or of a launch-list csv, for profiles/."""
import collections, csv, re, subprocess, sys

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "gld_req"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "gld_sect"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "gst_req"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "gst_sect"),
        ("smsp__inst_executed.sum", "inst")]


def short(name):
    m = re.search(r"conv_gemm_kernel<(\d+), (\d+), (\d+)>", name)
    return f"conv_gemm<{m.group(1)},{m.group(2)},epi{m.group(3)}>" if m else re.sub(r"\(.*", "", name)[:48]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        out = [short(r[idx["Kernel Name"]])]
        for k, lbl in KEYS:
            if k in idx:
                v = r[idx[k]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                out.append(f"{lbl}={v}{units[idx[k]] if units[idx[k]] not in ('', 'inst', 'sector') else ''}")
        print("  ".join(out))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu].startswith("n") else (v * 1e3 if r[mu].startswith("m") else v)
        a = agg[short(r[kn])]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'total_ms':>9s} {'avg_us':>9s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {n:5d} {t / 1e3:9.3f} {t / n:9.2f} {t / tot:6.3f}")


if __name__ == "__main__":
    (full if sys.argv[1].endswith(".ncu-rep") else launches)(sys.argv[1])
