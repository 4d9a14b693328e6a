"""One-row-per-kernel CSV of an .ncu-rep (read here, no GPU) for profiles/: duration, tensor / XU pipe utilisation,
DRAM bytes, achieved DRAM bandwidth, L2 hit rate, registers, top stalls.
    python tools/ncu_table.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_<what>.csv"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")
def to_bytes(r, k):
    v = val(r, k)
    u = units[col[k]] if k in col else ""
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_us(r, k):
    v = val(r, k)
    u = units[col[k]] if k in col else ""
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
w = csv.writer(sys.stdout)
w.writerow(["kernel", "duration_us", "tensor_pipe_active_pct", "xu_pipe_pct", "issue_active_pct", "dram_read_MB", "dram_write_MB",
            "dram_GBps", "l2_hit_pct", "regs", "top_stalls"])
for r in rows[2:]:
    if len(r) < len(h):
        continue
    us = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    st = {}
    for k, i in col.items():
        if "smsp__pcsamp_warps_issue_stalled" in k and not k.endswith("_not_issued"):
            try:
                st[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i].replace(",", ""))
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    top = " ".join(f"{k}:{v / tot * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("fitv2::", "")
    w.writerow([name, f"{us:.2f}", f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}",
                f"{val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f}",
                f"{val(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f}",
                f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}", f"{(rd + wr) / us / 1e3:.0f}", f"{val(r, 'lts__t_sector_hit_rate.pct'):.1f}",
                f"{val(r, 'launch__registers_per_thread'):.0f}", top])
