"""Small deterministic workload for ncu: XL/2-width, depth-2 FiTv2, 64 rows x 256 tokens, 2 forwards."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS
from fitv2_b200 import FiT, make_grid

wl = sys.argv[1] if len(sys.argv) > 1 else "xl256"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kw, (hp, wp), rope_kw = WORKLOADS[wl]
kw = dict(kw, depth=depth)
torch.manual_seed(0)
m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora",
        **kw, **rope_kw).randomize_zero_init_(1).cuda().eval()
R, N = 64, hp * wp
g = torch.Generator().manual_seed(0)
x = torch.randn(R, N, 16, generator=g).cuda()
t = torch.full((R,), 0.3).cuda()
y = torch.randint(0, 1001, (R,), generator=g).cuda()
grid, mask = make_grid(R, hp, wp).cuda(), torch.ones(R, N).cuda()
for _ in range(2):
    out = m(x, t, y, grid, mask)
torch.cuda.synchronize()
# the sampler-side elementwise kernels (fused CFG + Euler update; SDE Euler-Maruyama step), at the headline state size
from fitv2_b200 import EulerCFGSampler, Sampler, create_transport
n = R // 2
smp = EulerCFGSampler(m, y[:n], grid[:n], mask[:n], 250, 1.5)
z = torch.randn(n, N, 16, device="cuda")
for i in range(2):
    smp._step(z, smp.t_table[i], smp.dsig[i:i + 1])
fn = Sampler(create_transport()).sample_sde(diffusion_form="sigma", num_steps=3)
xs = fn(torch.randn(R, N, 16, device="cuda"), lambda xx, tt, **k: xx * 0.5)
torch.cuda.synchronize()
# the kernels of the model variants: FiTv1 configuration (no q/k norm -> generic QKV epilogue + online-max attention, adaLN 'normal',
# learn_sigma, (B, C, N) layout transposes) at depth 1, same row count
torch.manual_seed(0)
m1 = FiT(hidden_size=1152, depth=1, num_heads=16, learn_sigma=True, use_swiglu=True, use_swiglu_large=True).randomize_zero_init_(1).cuda().eval()
o1 = m1(x.transpose(1, 2).contiguous(), t, y, grid, mask)
# adaptive dopri5 pieces (fitv2_lincomb / fitv2_scaled_rms) on the headline state size
ys = Sampler(create_transport()).sample_ode(sampling_method="dopri5", num_steps=2)(torch.randn(n, N, 16, device="cuda") * 0.3, lambda xx, tt, **k: -xx)
torch.cuda.synchronize()
print("ok", float(out.abs().max()), float(z.abs().max()), float(xs[-1].abs().max()), float(o1.abs().max()), float(ys[-1].abs().max()))
