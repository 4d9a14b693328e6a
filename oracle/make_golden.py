"""Generate tests/golden/*.pt from the REAL reference (run in the build container only).

TEST INFRASTRUCTURE.  Needs /root/reference (read-only); the GPU box does not
have it, so the outputs are committed.  What it does:

  1. installs a 30-line ``timm`` shim (timm is not installed) that
     restates ``timm.layers.mlp.SwiGLU`` / ``Mlp`` creation order + forward, and
     drops the stray ``save_attention`` kwarg that makes ``FiT.__init__`` raise
     at HEAD;
  2. imports ``fit.model.fit_model.FiT`` from /root/reference and checks that
     ``oracle.fitv2_oracle`` is BIT-EQUAL to it on CPU fp32: the init under a
     seed, RoPE tables for every frequency rule, forward, forward_with_cfg,
     unpatchify and the Euler/CFG loop body;
  3. writes small fixtures (inputs, seeds, reference outputs, weight checksums)
     that ``tests/test_oracle_golden.py`` replays without the reference.

Usage:  python oracle/make_golden.py [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = os.environ.get("FITV2_REFERENCE", "/root/reference")


def install_reference():
    """timm shim + F1 fix; returns the reference FiT class."""
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        layers = types.ModuleType("timm.layers")
        mlp = types.ModuleType("timm.layers.mlp")
        data = types.ModuleType("timm.data")

        class SwiGLU(nn.Module):  # timm.layers.mlp.SwiGLU semantics (norm=Identity, drop=0)
            def __init__(self, in_features, hidden_features=None, out_features=None, bias=True, **kw):
                super().__init__()
                out_features = out_features or in_features
                hidden_features = hidden_features or in_features
                self.fc1_g = nn.Linear(in_features, hidden_features, bias=bias)
                self.fc1_x = nn.Linear(in_features, hidden_features, bias=bias)
                self.act = nn.SiLU()
                self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)

            def forward(self, x):
                return self.fc2(self.act(self.fc1_g(x)) * self.fc1_x(x))

        class Mlp(nn.Module):
            def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, bias=True, **kw):
                super().__init__()
                self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
                self.act = act_layer()
                self.fc2 = nn.Linear(hidden_features, out_features or in_features, bias=bias)

            def forward(self, x):
                return self.fc2(self.act(self.fc1(x)))

        mlp.SwiGLU, mlp.Mlp = SwiGLU, Mlp
        data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
        data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
        timm.layers, timm.data, layers.mlp = layers, data, mlp
        sys.modules.update({"timm": timm, "timm.layers": layers, "timm.layers.mlp": mlp, "timm.data": data})
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import fit.model.modules as rm
    if not getattr(rm.Attention, "_f1_fixed", False):
        orig = rm.Attention.__init__

        def patched(self, *a, save_attention=False, **kw):
            orig(self, *a, **kw)
        rm.Attention.__init__ = patched
        rm.Attention._f1_fixed = True
    from fit.model.fit_model import FiT
    return FiT


def ref_kwargs(cfg):
    kw = dict(context_size=cfg.context_size, patch_size=cfg.patch_size, in_channels=cfg.in_channels,
              hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio,
              class_dropout_prob=0.1, num_classes=cfg.num_classes, learn_sigma=False, use_sit=True,
              use_swiglu=True, use_swiglu_large=False, use_checkpoint=False, q_norm="layernorm",
              k_norm="layernorm", qk_norm_weight=False, rel_pos_embed="rope", abs_pos_embed=None,
              adaln_type="lora", adaln_lora_dim=cfg.adaln_lora_dim, custom_freqs=cfg.custom_freqs,
              online_rope=False)
    if cfg.custom_freqs != "normal":
        kw.update(max_pe_len_h=cfg.max_pe_len_h, max_pe_len_w=cfg.max_pe_len_w, decouple=cfg.decouple,
                  ori_max_pe_len=cfg.ori_max_pe_len)
    return kw


def build_reference(FiT, cfg, init_seed=0, redraw_seed=1):
    from oracle import fitv2_oracle as O
    torch.manual_seed(init_seed)
    m = FiT(**ref_kwargs(cfg)).eval()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    sd = O.redraw_zero_params(sd0, redraw_seed)
    m.load_state_dict(sd)
    return m, sd0, sd


def mixed_padded_batch(cfg, layouts, target_len, seed):
    """fit/data/in1k_latent_dataset.py:54-69 padded layout: feature 0, grid 0, mask 0 past seq_len."""
    from oracle import fitv2_oracle as O
    g = torch.Generator().manual_seed(seed)
    B = len(layouts)
    x = torch.zeros(B, target_len, cfg.token_channels)
    grid = torch.zeros(B, 2, target_len, dtype=torch.long)
    mask = torch.zeros(B, target_len)
    for b, (h, w) in enumerate(layouts):
        n = h * w
        x[b, :n] = torch.randn(n, cfg.token_channels, generator=g)
        grid[b, :, :n] = O.make_grid(1, h, w)[0]
        mask[b, :n] = 1
    return x, grid, mask


def checksum(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--skip-xl", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    from oracle import fitv2_oracle as O
    FiT = install_reference()
    torch.set_grad_enabled(False)
    report = []

    # ---- 1. RoPE frequency rules -------------------------------------------------
    from fit.model.rope import VisionRotaryEmbedding, rotate_half as ref_rot
    rope_cases = []
    for (dh, cf, h, w, dec, ori) in [
        (72, "normal", None, None, False, None), (96, "normal", None, None, False, None),
        (72, "ntk-aware", 10, 20, True, 16), (72, "ntk-aware", 32, 32, True, 16), (72, "ntk-aware", 10, 20, False, 16),
        (72, "linear", 10, 20, True, 16), (72, "ntk-aware-pro1", 20, 20, False, 16),
        (72, "ntk-aware-pro2", 20, 20, False, 16), (72, "ntk-by-parts", 24, 40, True, 16),
        (72, "yarn", 24, 40, True, 16), (96, "yarn", 32, 32, False, 16), (32, "ntk-aware", 6, 12, True, 4),
    ]:
        kw = dict(head_dim=dh, custom_freqs=cf)
        if cf != "normal":
            kw.update(max_pe_len_h=h, max_pe_len_w=w, decouple=dec, ori_max_pe_len=ori)
        r = VisionRotaryEmbedding(**kw)
        cfg = O.FiTConfig(hidden_size=dh * 2, num_heads=2, custom_freqs=cf, max_pe_len_h=h, max_pe_len_w=w,
                          decouple=dec, ori_max_pe_len=ori)
        fh, fw, mag = O.rope_setup(cfg)
        assert torch.equal(fh, r.freqs_h) and torch.equal(fw, r.freqs_w), (dh, cf)
        gh, gw = (h or 16), (w or 16)
        grid = O.make_grid(2, gh, gw)
        rc, rs = r.get_cached_2d_rope_from_grid(grid)
        oc, os_ = O.rope_cos_sin(cfg, grid)
        assert torch.equal(rc, oc) and torch.equal(rs, os_), (dh, cf)
        rope_cases.append(dict(head_dim=dh, custom_freqs=cf, max_pe_len_h=h, max_pe_len_w=w, decouple=dec,
                               ori_max_pe_len=ori, freqs_h=r.freqs_h.clone(), freqs_w=r.freqs_w.clone(), mag=mag,
                               grid_hw=(gh, gw), cos_tok=rc[0, [0, 1, gw + 1, gh * gw - 1]].clone(),
                               sin_tok=rs[0, [0, 1, gw + 1, gh * gw - 1]].clone()))
    xr = torch.arange(8.0)
    assert torch.equal(ref_rot(xr), O.rotate_half(xr))
    report.append(f"rope: {len(rope_cases)} rule cases bit-equal")
    torch.save(dict(cases=rope_cases, rotate_half_0to7=ref_rot(xr)), os.path.join(args.out, "rope_kat.pt"))

    # ---- 2. tiny models: full state_dict stored ---------------------------------
    tiny = []
    for name, cfgkw, layouts, tl in [
        ("tiny_normal", dict(hidden_size=96, depth=2, num_heads=4, adaln_lora_dim=24), [(4, 4), (4, 4), (4, 4)], 16),
        ("tiny_ntk_padded", dict(hidden_size=128, depth=3, num_heads=4, adaln_lora_dim=32, custom_freqs="ntk-aware",
                                 max_pe_len_h=6, max_pe_len_w=12, decouple=True, ori_max_pe_len=4),
         [(6, 12), (12, 6), (8, 8), (4, 10)], 80),
    ]:
        cfg = O.FiTConfig(**cfgkw)
        m, sd0, sd = build_reference(FiT, cfg)
        osd0 = O.reference_init_state_dict(cfg, 0)
        assert list(osd0.keys()) == list(sd0.keys())
        for k in sd0:
            assert torch.equal(sd0[k], osd0[k]), k
        x, grid, mask = mixed_padded_batch(cfg, layouts, tl, seed=7)
        B = x.shape[0]
        t = torch.tensor([0.1, 0.5, 0.9, 1.0][:B])
        y = torch.tensor([7, 1000, 999, 0][:B])
        ref = m(x, t, y=y, grid=grid, mask=mask, size=None)
        got = O.forward(cfg, sd, x, t, y, grid, mask)
        assert torch.equal(ref, got), (name, (ref - got).abs().max())
        ref2 = m(x, t, y=y, grid=grid, mask=mask * 2, size=None)      # raw-mask multiply (fit_model.py:230)
        assert torch.equal(ref2, O.forward(cfg, sd, x, t, y, grid, mask * 2))
        refb = m(x, t, y=y, grid=grid, mask=mask.bool(), size=None)
        assert torch.equal(refb, O.forward(cfg, sd, x, t, y, grid, mask.bool()))
        tiny.append(dict(name=name, cfg=cfgkw, state_dict=sd, x=x, t=t, y=y, grid=grid, mask=mask, out=ref,
                         out_mask2=ref2))
        report.append(f"{name}: init + forward bit-equal (|out|max {float(ref.abs().max()):.4f})")
    torch.save(tiny, os.path.join(args.out, "tiny_models.pt"))

    # ---- 3. XL-width shallow model: forward_with_cfg, unpatchify, Euler loop -----
    cfg = O.FiTConfig(depth=2, **{k: v for k, v in O.XL2.items() if k != "depth"}, custom_freqs="ntk-aware",
                      max_pe_len_h=10, max_pe_len_w=20, decouple=True, ori_max_pe_len=16)
    m, sd0, sd = build_reference(FiT, cfg)
    osd = O.synthetic_state_dict(cfg)
    for k in sd:
        assert torch.equal(sd[k], osd[k]), k
    x, grid, mask = mixed_padded_batch(cfg, [(10, 20), (16, 16), (8, 24), (20, 10)], 256, seed=11)
    t = torch.tensor([0.1, 0.5, 0.9, 0.3])
    y = torch.tensor([7, 1000, 999, 1000])
    ref = m(x, t, y=y, grid=grid, mask=mask, size=None)
    got = O.forward(cfg, sd, x, t, y, grid, mask)
    assert torch.equal(ref, got), (ref - got).abs().max()
    refc = m.forward_with_cfg(x, t, y, grid, mask, None, 1.5)
    assert torch.equal(refc, O.forward_with_cfg(cfg, sd, x, t, y, grid, mask, None, 1.5))
    refp = m.forward_with_cfg(x, t, y, grid, mask, None, 4.0, scale_pow=2.0)
    assert torch.equal(refp, O.forward_with_cfg(cfg, sd, x, t, y, grid, mask, None, 4.0, scale_pow=2.0))
    lat = torch.randn(2, 200, 16, generator=torch.Generator().manual_seed(3))
    assert torch.equal(m.unpatchify(lat, (20, 40)), O.unpatchify(cfg, lat, (20, 40)))
    torch.save(dict(cfg=dict(depth=2, custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20, decouple=True,
                             ori_max_pe_len=16), x=x, t=t, y=y, grid=grid, mask=mask, out=ref, out_cfg=refc,
                    out_cfg_pow=refp, weight_checksum=checksum(sd), unpatchify_in=lat[:, :8].clone(),
                    unpatchify_out=m.unpatchify(lat, (20, 40))[:, :, :2].clone()),
               os.path.join(args.out, "xl_depth2_padded.pt"))
    report.append(f"xl_depth2_padded: forward/forward_with_cfg/unpatchify bit-equal (|out|max {float(ref.abs().max()):.4f})")

    # ---- 4. config 1: XL/2 depth 36, one CFG Euler step, batch 2 (4 rows) -------
    if not args.skip_xl:
        cfg = O.FiTConfig(**O.XL2)
        m, sd0, sd = build_reference(FiT, cfg)
        osd = O.synthetic_state_dict(cfg)
        for k in sd:
            assert torch.equal(sd[k], osd[k]), k
        n, N = 2, 256
        torch.manual_seed(0)                       # global_seed*world+rank = 0 (sample_fitv2_ddp.py:54-55)
        z = torch.randn(n, N, 16)
        yl = torch.randint(0, 1000, (n,))
        grid = O.make_grid(n, 16, 16)
        mask = torch.ones(n, N)
        y2 = torch.cat([yl, torch.full((n,), 1000)], 0)
        grid2, mask2 = torch.cat([grid, grid], 0), torch.cat([mask, mask], 0)
        sig = torch.linspace(0, 1, 251)
        outs = []
        zz = z
        for idx in (0, 1):
            z_in = torch.cat([zz, zz], 0)
            ts = sig[idx].expand(2 * n)
            v2 = m(z_in, ts, y=y2, grid=grid2, mask=mask2, size=None)
            o2 = O.forward(cfg, sd, z_in, ts, y2, grid2, mask2)
            assert torch.equal(v2, o2), (v2 - o2).abs().max()
            cond, uncond = v2.chunk(2, dim=0)
            v = uncond + 1.5 * (cond - uncond)
            zz = zz + (sig[idx + 1] - sig[idx]) * v
            assert torch.equal(zz, O.cfg_euler_update(z_in[:n], v2, 1.5, sig[idx], sig[idx + 1]))
            outs.append((v2.clone(), zz.clone()))
        # a mid-trajectory NFE (t = 0.5) so the KAT is not only t≈0
        ts = torch.full((2 * n,), 0.5)
        vmid = m(torch.cat([z, z], 0), ts, y=y2, grid=grid2, mask=mask2, size=None)
        torch.save(dict(z=z, y=yl, v_step0=outs[0][0], z_step0=outs[0][1], v_step1=outs[1][0], z_step1=outs[1][1],
                        v_t05=vmid, weight_checksum=checksum(sd), n_params=sum(v.numel() for v in sd.values())),
                   os.path.join(args.out, "xl_config1.pt"))
        report.append(f"xl_config1: depth-36 init + 2 Euler steps bit-equal, params {sum(v.numel() for v in sd.values())}")

    with open(os.path.join(args.out, "README.md"), "w") as f:
        f.write("# Golden fixtures\n\nGenerated by `python oracle/make_golden.py` in the build container from the REAL "
                "reference classes at /root/reference (timm shim + save_attention fix).  The reference ships no tests "
                "or vectors of its own (SURVEY.md F5), so these outputs of the reference code are what pins the oracle.\n\n"
                + "\n".join(f"* {r}" for r in report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
