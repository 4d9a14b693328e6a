"""Generate tests/golden/swiglu_large_xl_d1.pt from the REAL reference (build container only; TEST INFRASTRUCTURE).

`use_swiglu_large=True` (modules.py:248-249) widens the SwiGLU hidden size from (int(D*mlp_ratio)*2)//3 to
int(D*mlp_ratio).  The script builds the reference FiT (through the timm shim of oracle/make_golden.py) at depth 1,
checks the oracle is bit-equal on CPU fp32 and stores inputs + the reference velocity for the GPU parity test.

Usage:  python oracle/make_swiglu_large_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as G                      # noqa: E402
from oracle import fitv2_oracle as O         # noqa: E402


def main():
    FiT = G.install_reference()
    cfg = O.FiTConfig(depth=1, use_swiglu_large=True)
    kw = G.ref_kwargs(cfg)
    kw["use_swiglu_large"] = True
    torch.manual_seed(0)
    m = FiT(**kw).eval()
    sd = O.redraw_zero_params({k: v.clone() for k, v in m.state_dict().items()}, 1)
    m.load_state_dict(sd)
    assert sd["blocks.0.mlp.fc1_g.weight"].shape == (4608, 1152) and cfg.mlp_hidden == 4608
    g = torch.Generator().manual_seed(11)
    R, hp, wp = 3, 10, 20
    x = torch.randn(R, hp * wp, 16, generator=g)
    t = torch.rand(R, generator=g)
    y = torch.randint(0, 1001, (R,), generator=g)
    grid, mask = O.make_grid(R, hp, wp), torch.ones(R, hp * wp)
    with torch.no_grad():
        ref = m(x, t, y, grid, mask.clone(), None)
    ora = O.forward(cfg, sd, x, t, y, grid, mask)
    assert torch.equal(ref, ora), float((ref - ora).abs().max())
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "swiglu_large_xl_d1.pt")
    torch.save(dict(x=x, t=t, y=y, hp=hp, wp=wp, init_seed=0, redraw_seed=1, v_ref=ref.clone(),
                    fc1_g_abs_sum=float(sd["blocks.0.mlp.fc1_g.weight"].double().abs().sum())), out)
    print("oracle bit-equal to the reference with use_swiglu_large=True; wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
