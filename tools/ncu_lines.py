"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
agg = defaultdict(lambda: [0, 0, 0, ""])  # samples, insts, local?, text
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    is_src = r[2] == "-"   # source line summary row has no address
    if not is_src:
        continue
    s = int(r[hdr["# Samples"]] or 0)
    i = int(r[hdr["Instructions Executed"]] or 0)
    key = (cur_file, line)
    agg[key][0] += s
    agg[key][1] += i
    agg[key][3] = r[1].strip()[:110]
    tot_s += s
    tot_i += i
print(f"total samples {tot_s}  total warp-instructions {tot_i}")
print("--- by stall samples")
for (f, l), (s, i, _, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*s/tot_s:5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins  {f}:{l}  {t}")

if len(sys.argv) > 3:
    # bucket report: "name:lo-hi,name:lo-hi" on potus_kernel.cu
    print("--- buckets (potus_kernel.cu line ranges)")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(v[0] for (f, l), v in agg.items() if f == "potus_kernel.cu" and lo <= l <= hi)
        i = sum(v[1] for (f, l), v in agg.items() if f == "potus_kernel.cu" and lo <= l <= hi)
        print(f"{name:28s} {100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins")
    s = sum(v[0] for (f, l), v in agg.items() if f != "potus_kernel.cu")
    i = sum(v[1] for (f, l), v in agg.items() if f != "potus_kernel.cu")
    print(f"{'other files (ptx wrappers)':28s} {100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins")
