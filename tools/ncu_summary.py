"""Summarise an .ncu-rep (raw page + source page) without a GPU: key utilisation metrics, stall mix, top SASS hotspots."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__pipe_tensor_subpipe_imma",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_bytes.sum ",
        "launch__registers_per_thread ", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "sm__warps_active.avg.pct",
        "smem", "shared"]
for h, u, v in zip(hdr, units, vals):
    hh = h + " "
    if any(w in hh for w in want) and v not in ("", "0"):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}; data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data); ti = sum(f(r, "Instructions Executed") for r in data)
print(f"\nwarp instructions {ti:.3e}; samples {tot:.0f}")
agg = {k: sum(f(r, k) for r in data) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
print("stalls:", ", ".join(f"{k[6:]} {v / tot * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
c = Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
    c[op.split(".")[0]] += f(r, "Instructions Executed")
print("mix:", ", ".join(f"{op} {v / ti * 100:.1f}%" for op, v in c.most_common(16)))
print("top sample sites:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    st = sorted(((k, f(r, k)) for k in agg), key=lambda kv: -kv[1])[0]
    print(f"  {f(r, '# Samples') / tot * 100:5.1f}% inst {f(r, 'Instructions Executed') / ti * 100:5.2f}% thr {f(r, 'Avg. Threads Executed'):4.1f}  {r[ix['Source']][:64]:64s} {st[0][6:]}")
