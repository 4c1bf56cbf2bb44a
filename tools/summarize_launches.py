"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/profile_step.py: per-kernel totals
of the last Euler step and of the DAC decode.   python tools/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    try:
        rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", "")), row.get("Grid Size", "")))
    except Exception:
        pass
names = [x[0] for x in rows]
i0 = [i for i, n in enumerate(names) if "vectok_silu" in n][-1]
i1 = [i for i, n in enumerate(names) if "advance_step" in n][-1]
step = rows[i0:i1 + 1]
tot = sum(x[1] for x in step)
print(f"one Euler step: {len(step)} launches, sum of kernel durations {tot / 1e3:.1f} us (cold-cache, serialised)")
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, g in step:
    k = re.sub(r"\(.*", "", n).replace("void ", "").replace("foley::", "")
    agg[k][0] += 1
    agg[k][1] += v
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1000:9.1f} us {100 * v / tot:5.1f}%  n={c:4d} avg={v / c / 1000:7.2f} us  {k[:70]}")
print("GEMM launches by (tile, grid):")
g2 = collections.defaultdict(lambda: [0, 0.0])
for n, v, g in step:
    if "gemm_tcgen05" in n:
        key = (re.sub(r".*gemm_tcgen05_kernel<([^>]*)>.*", r"\1", n), g)
        g2[key][0] += 1
        g2[key][1] += v
for k, (c, v) in sorted(g2.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1000:9.1f} us n={c:4d} avg={v / c / 1000:7.2f} us  <{k[0]}> grid {k[1]}")
dac = rows[i1 + 1:]
j = [i for i, (n, v, g) in enumerate(dac) if "dac_transpose_in" in n]
if j:
    d = dac[j[-1]:]
    print(f"DAC decode: {len(d)} launches, sum {sum(x[1] for x in d) / 1e3:.1f} us")
    for n, v, g in sorted(d, key=lambda x: -x[1])[:8]:
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("foley::", "")
        print(f"   {v / 1000:8.1f} us grid {g} {k[:60]}")
