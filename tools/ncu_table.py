"""Per-kernel table out of an .ncu-rep (or its `--page raw --csv` dump) with many kernels (ncu --set full): duration, DRAM bytes and achieved GB/s, pipe
utilisation, issue activity, registers, occupancy.  Usage: ncu_table.py rep [hbm_peak_GBps]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6526.5
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(raw)) if r]
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def col(name):
    for h in hdr:
        if h == name or h.endswith("." + name) or h.endswith(name):
            return ix[h]
    return None
def val(r, name, scale_units=True):
    c = col(name)
    if c is None or r[c] == "":
        return float("nan")
    try:
        v = float(r[c].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[c]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}
    return v * mult.get(u, 1.0) if scale_units else v
print(f"{'kernel':58s} {'ms':>9s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'GB/s':>8s} {'of HBM':>7s} {'tensor%':>8s} {'alu%':>6s} {'fma%':>6s} {'fp64%':>6s} {'issue%':>7s} {'regs':>5s} {'warps%':>7s}")
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("tmg::", "")[:58]
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / t / 1e9 if t > 0 else float("nan")
    print(f"{name:58s} {t * 1e3:9.3f} {rd / 1e6:11.1f} {wr / 1e6:11.1f} {gbs:8.1f} {gbs / peak:7.3f} "
          f"{val(r, 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', False):8.1f} "
          f"{val(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', False):6.1f} "
          f"{val(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', False):6.1f} "
          f"{val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', False):6.1f} "
          f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active', False):7.1f} "
          f"{val(r, 'launch__registers_per_thread', False):5.0f} {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active', False):7.1f}")
