"""Warp-stall breakdown of the first kernel in an .ncu-rep (source page): totals per reason and the top SASS lines."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    data.append(r)
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(f(r, "# Samples") for r in data)
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("samples", int(tot), {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    st = {s.replace("stall_", ""): int(f(r, s)) for s in stalls if f(r, s) > 0.15 * f(r, "# Samples")}
    print(f"{int(f(r, '# Samples')):5d} exec={int(f(r, 'Instructions Executed')):8d} {r[idx['Source']][:70]:70s} {st}")
