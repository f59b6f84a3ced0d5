"""Top stall sites of a kernel's hottest region from an .ncu-rep (read here, no GPU):
   python bench_tools/ncu_stalls.py gpurun_out/x.ncu-rep [n]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; I = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > I["# Samples"]]
base = int(data[0][I["Address"]], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def f(r, k):
    try: return float(r[I[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
agg = collections.Counter()
for r in data:
    for s in stalls: agg[s] += f(r, s)
print("samples", int(tot), {k.replace("stall_", ""): int(v) for k, v in agg.most_common(9)})
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    print(f"{int(r[I['Address']], 16) - base:#7x} {r[I['Source']][:64]:64s} {int(f(r, '# Samples')):4d}",
          {s.replace("stall_", ""): int(f(r, s)) for s in stalls if f(r, s) >= 3})
