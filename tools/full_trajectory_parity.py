"""Final latents after the FULL 250-step CFG Euler trajectory: fitv2_b200 on the GPU against the CPU fp32 oracle
(north_star: "final latents after the full trajectory within a stated tolerance").  One-off measurement, ~2-4 minutes of
CPU time for the oracle at batch 1; the result goes to gpurun_out/full_trajectory_parity.json (copied to profiles/).

    python tools/full_trajectory_parity.py [steps=250] [depth=36]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import fitv2_oracle as O                      # checker
from fitv2_b200 import FiT, EulerCFGSampler, make_grid

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 250
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 36
torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 1)
kw = dict(hidden_size=1152, depth=depth, num_heads=16, adaln_lora_dim=288)
torch.manual_seed(0)
m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora", **kw).randomize_zero_init_(1)
sd = {k: v.detach().clone().float() for k, v in m.state_dict().items()}
m = m.cuda().eval()
cfg = O.FiTConfig(**kw)
n, hp, wp = 1, 16, 16
g = torch.Generator().manual_seed(0)
z0 = torch.randn(n, hp * wp, 16, generator=g)
y = torch.randint(0, 1000, (n,), generator=g)
grid, mask = make_grid(n, hp, wp), torch.ones(n, hp * wp)
res = {}
for operand in ("bf16", "fp16"):
    mm = m if operand == "bf16" else None
    if mm is None:
        torch.manual_seed(0)
        mm = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora",
                 operand_dtype="fp16", **kw).randomize_zero_init_(1).cuda().eval()
    smp = EulerCFGSampler(mm, y.cuda(), grid.cuda(), mask.cuda(), steps, 1.5)
    res[operand] = smp.sample(z0.cuda()).cpu()
t0 = time.time()
zr = O.euler_cfg_sample(cfg, sd, z0, y, grid, mask, None, steps, 1.5)
cpu_s = time.time() - t0
out = dict(steps=steps, depth=depth, samples=n, tokens=hp * wp, cfg_scale=1.5, oracle_cpu_seconds=round(cpu_s, 1),
           ref_abs_max=float(zr.abs().max()), ref_rms=float(zr.pow(2).mean().sqrt()))
for operand, z in res.items():
    d = z - zr
    out[operand] = dict(max_abs_over_max_abs=float(d.abs().max() / zr.abs().max()), rms_rel=float(d.pow(2).mean().sqrt() / zr.pow(2).mean().sqrt()),
                        finite=bool(torch.isfinite(z).all()))
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "full_trajectory_parity.json"), "w") as f:
    json.dump(out, f, indent=1)
