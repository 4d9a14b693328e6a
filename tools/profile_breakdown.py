"""Per-kernel-class device time of one NFE (CUDA events on the launching stream, via fitv2_profile_*).
    python tools/profile_breakdown.py [workload] [batch] [operand]   -> gpurun_out/breakdown_<workload>.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS, flops_per_forward_row
from fitv2_b200 import FiT, make_grid

wl = sys.argv[1] if len(sys.argv) > 1 else "xl256"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
operand = sys.argv[3] if len(sys.argv) > 3 else "bf16"
kw, (hp, wp), rope_kw = WORKLOADS[wl]
torch.manual_seed(0)
m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora",
        operand_dtype=operand, **kw, **rope_kw).randomize_zero_init_(1).cuda().eval()
R, N = 2 * n, hp * wp
g = torch.Generator().manual_seed(0)
x = torch.randn(R, N, 16, generator=g).cuda()
t = torch.full((R,), 0.3).cuda()
y = torch.randint(0, 1001, (R,), generator=g).cuda()
grid, mask = make_grid(R, hp, wp).cuda(), torch.ones(R, N).cuda()
for _ in range(3):
    m(x, t, y, grid, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    m(x, t, y, grid, mask)
e1.record()
torch.cuda.synchronize()
nfe_ms = e0.elapsed_time(e1) / reps
m.profile("all")
for _ in range(reps):
    m(x, t, y, grid, mask)
prof = m.profile_read()
m.profile(None)
D, L = kw["hidden_size"], kw["depth"]
Hm = (int(D * 4.0) * 2) // 3
M = R * N
flops = dict(qkv_gemm=2.0 * M * D * 3 * D, proj_gemm=2.0 * M * D * D, gateup_gemm=2.0 * M * D * 2 * Hm, fc2_gemm=2.0 * M * Hm * D,
             attention=4.0 * R * N * N * D)
out = dict(workload=wl, rows=R, tokens=N, operand=operand, nfe_ms=nfe_ms, model_tflops=flops_per_forward_row(kw, N) * R / nfe_ms / 1e9, classes={})
tot = 0.0
for name, (ms, cnt) in prof.items():
    per_nfe = ms / reps
    tot += per_nfe
    us = ms / max(cnt, 1) * 1e3
    row = dict(ms_per_nfe=round(per_nfe, 4), launches_per_nfe=cnt // reps, us_per_launch=round(us, 2))
    if name in flops:
        row["tflops"] = round(flops[name] / (us * 1e-6) / 1e12, 1)
    if name == "ln_modulate" and us > 0:
        row["gbs"] = round(M * D * 6 / (us * 1e-6) / 1e9, 1)      # read fp32 + write 16-bit
    out["classes"][name] = row
out["sum_of_classes_ms"] = tot
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"breakdown_{wl}_{operand}.json"), "w") as f:
    json.dump(out, f, indent=1)
