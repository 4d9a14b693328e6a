"""Per-kernel-class device time (CUDA events in the running forward, fitv2_profile_*) of the generic variants at the headline
shape (64 rows x 256 tokens, XL width, depth 8): the GELU Mlp (fc1 + tanh-GELU in the plain GEMM epilogue, fc2 with K = 4608) and
the SwiGLU modulation MLPs (adaln_type 'swiglu', fp32 FMA linears) next to the FiTv2 default.
    python tools/variant_timing.py   -> gpurun_out/variant_timing.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fitv2_b200 import FiT, make_grid

R, hp, wp, depth = 64, 16, 16, 8
N = hp * wp
g = torch.Generator().manual_seed(0)
x = torch.randn(R, N, 16, generator=g).cuda()
t = torch.full((R,), 0.3).cuda()
y = torch.randint(0, 1001, (R,), generator=g).cuda()
grid, mask = make_grid(R, hp, wp).cuda(), torch.ones(R, N).cuda()
base = dict(hidden_size=1152, depth=depth, num_heads=16, learn_sigma=False, use_sit=True, q_norm="layernorm", k_norm="layernorm")
cases = {
    "fitv2_default": dict(use_swiglu=True, adaln_type="lora", adaln_lora_dim=288),
    "gelu_mlp": dict(use_swiglu=False, adaln_type="lora", adaln_lora_dim=288),
    "gelu_mlp_plain_epilogue": dict(use_swiglu=False, adaln_type="lora", adaln_lora_dim=288),
    "adaln_swiglu": dict(use_swiglu=True, adaln_type="swiglu"),
    "rope_v": dict(use_swiglu=True, adaln_type="lora", adaln_lora_dim=288, add_rel_pe_to_v=True),
}
out = {}
for name, kw in cases.items():
    torch.manual_seed(0)
    m = FiT(**base, **kw).randomize_zero_init_(1).cuda().eval()
    if name.endswith("plain_epilogue"):
        m.set_option("gelu_epi", 1)
    for _ in range(3):
        m(x, t, y, grid, mask)
    m.profile("all")
    reps = 5
    for _ in range(reps):
        m(x, t, y, grid, mask)
    prof = m.profile_read()
    m.profile(None)
    out[name] = {k: dict(us_per_launch=round(ms / max(c, 1) * 1e3, 2), launches_per_nfe=c // reps, ms_per_nfe=round(ms / reps, 4)) for k, (ms, c) in prof.items()}
    del m
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "variant_timing.json"), "w") as f:
    json.dump(out, f, indent=1)
