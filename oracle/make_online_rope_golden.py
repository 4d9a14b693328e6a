"""Pin the oracle's online-RoPE path (fit_model.py:212-214, rope.py:234-274) against the REAL reference
(build container only; needs /root/reference).  Writes tests/golden/xl_depth2_online.pt.

    python oracle/make_online_rope_golden.py
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as MG                      # timm shim + reference import helpers
from oracle import fitv2_oracle as O


def main():
    FiT = MG.install_reference()
    torch.set_grad_enabled(False)
    notes = []
    from fit.model.rope import VisionRotaryEmbedding
    # ---- frequency tables for every rule the online mode supports ----
    g = torch.Generator().manual_seed(5)
    grid = torch.randint(0, 24, (4, 2, 50), generator=g)
    size = torch.tensor([[10, 20], [16, 16], [8, 24], [20, 10]])[:, None, :]
    for cf in ("linear", "ntk-aware", "ntk-by-parts"):
        for decouple in (False, True):
            for hd in (72, 96):
                ref = VisionRotaryEmbedding(head_dim=hd, custom_freqs=cf, online_rope=True, decouple=decouple, ori_max_pe_len=16)
                rc, rs = ref.online_get_2d_rope_from_grid(grid, size)
                cfg = O.FiTConfig(hidden_size=hd * 16, num_heads=16, custom_freqs=cf, decouple=decouple, ori_max_pe_len=16, online_rope=True)
                oc, os_ = O.rope_cos_sin_online(cfg, grid, size)
                assert torch.equal(rc, oc) and torch.equal(rs, os_), (cf, decouple, hd)
    notes.append("online RoPE tables: 3 rules x decouple x head_dim {72, 96} bit-equal")
    # ---- depth-2 XL/2 forward on a mixed-aspect padded batch with per-sample dynamic NTK scale ----
    cfg = O.FiTConfig(hidden_size=1152, depth=2, num_heads=16, adaln_lora_dim=288, custom_freqs="ntk-aware", decouple=True,
                      ori_max_pe_len=16, max_pe_len_h=16, max_pe_len_w=16, online_rope=True)
    kw = MG.ref_kwargs(cfg)
    kw["online_rope"] = True
    torch.manual_seed(0)
    m = FiT(**kw).eval()
    sd = O.redraw_zero_params({k: v.clone() for k, v in m.state_dict().items()}, 1)
    m.load_state_dict(sd)
    layouts = [(10, 20), (16, 16), (8, 24), (20, 10)]
    x, grid, mask = MG.mixed_padded_batch(cfg, layouts, 256, seed=9)
    size = torch.tensor(layouts)[:, None, :]
    t = torch.tensor([0.1, 0.4, 0.7, 0.95])
    y = torch.tensor([1, 500, 1000, 7])
    ref = m(x, t, y, grid, mask, size)
    ours = O.forward(cfg, sd, x, t, y, grid, mask, size)
    assert torch.equal(ref, ours)
    notes.append(f"xl_depth2_online: forward with online_rope (ntk-aware, decouple, per-sample size) bit-equal (|out|max {float(ref.abs().max()):.4f})")
    dst = os.path.join(ROOT, "tests", "golden", "xl_depth2_online.pt")
    torch.save(dict(x=x, t=t, y=y, grid=grid, mask=mask, size=size, out=ref, layouts=layouts), dst)
    with open(os.path.join(ROOT, "tests", "golden", "README.md"), "a") as f:
        f.write("\n## xl_depth2_online.pt (oracle/make_online_rope_golden.py)\n\n")
        for n in notes:
            f.write(f"* {n}\n")
    print("\n".join(notes))


if __name__ == "__main__":
    main()
