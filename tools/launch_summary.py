"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total per kernel, and the kernels of the last
timed step (from the last features_i16_kernel<0> launch on).  Usage: launch_summary.py launches.csv"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = OrderedDict(), {}
for r in rows:
    k = r[ik]; v = float(r[iv].replace(",", "")) / 1e6
    tot[k] = tot.get(k, 0.0) + v; cnt[k] = cnt.get(k, 0) + 1
s = sum(tot.values())
print(f"{len(rows)} launches, {s:.1f} ms of kernel time (cold-cache, serialised by the profiler)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:14]:
    print(f"{v:10.3f} ms {cnt[k]:5d}x {v / s * 100:5.1f}%  {k[:70]}")
cands = [i for i, r in enumerate(rows) if "features_i16_kernel<0>" in r[ik]]
if not cands:
    sys.exit(0)
last = max(cands)
step = rows[last:]
ss = sum(float(r[iv].replace(",", "")) for r in step) / 1e6
print("last timed step:")
for r in step:
    v = float(r[iv].replace(",", "")) / 1e6
    print(f"   {v:9.3f} ms {v / ss * 100:5.1f}%  {r[ik][:60]}")
