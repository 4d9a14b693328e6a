"""Summarise an .ncu-rep (read here, no GPU): duration, pipe utilisation, DRAM traffic, stall reasons.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for ri, row in enumerate(rows[2:]):
    if len(sys.argv) > 2 and ri != int(sys.argv[2]):
        continue
    d = dict(zip(h, row))
    print("==", ri, d.get("Kernel Name", "")[:100])
    for k in KEYS:
        if k in d:
            print(f"  {k:90s} {d[k]}")
    st = {}
    for k, v in d.items():
        if "smsp__pcsamp_warps_issue_stalled" in k and not k.endswith("_not_issued"):
            try:
                st[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v.replace(",", ""))
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    print("  stalls:", ", ".join(f"{k} {v / tot * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
