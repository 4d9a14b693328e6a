"""The oracle against the fixtures generated from the REAL reference code (oracle/make_golden.py).

The reference ships no tests or vectors; tests/golden/*.pt are outputs of the
reference's own classes, captured in the build container, and they pin the oracle here.  CPU only.
Tolerance: 2e-6 relative (identical arithmetic; BLAS blocking may differ between hosts).
"""
import os

import pytest
import torch

from oracle import fitv2_oracle as O

TOL = 2e-6


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_rope_known_answers(golden_dir):
    """SURVEY.md A.4 values captured from fit/model/rope.py."""
    fh, fw, mag = O.rope_setup(O.FiTConfig(hidden_size=1152, num_heads=16))
    assert fh[:3].tolist() == pytest.approx([1.0, 0.5994842052, 0.3593813777], rel=1e-6)
    assert float(fh[17]) == pytest.approx(1.6681010311e-4, rel=1e-6) and mag == 1.0
    cfg = O.FiTConfig(hidden_size=1152, num_heads=16, custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20,
                      decouple=True, ori_max_pe_len=16)
    fh2, fw2, _ = O.rope_setup(cfg)
    assert torch.equal(fh2, fh)
    assert fw2[:3].tolist() == pytest.approx([1.0, 0.5916668177, 0.3500695825], rel=1e-6)
    assert float(fw2[17]) == pytest.approx(1.3344809122e-4, rel=1e-6)
    cfg = O.FiTConfig(hidden_size=1152, num_heads=16, custom_freqs="ntk-aware", max_pe_len_h=32, max_pe_len_w=32,
                      decouple=True, ori_max_pe_len=16)
    assert O.rope_setup(cfg)[0][:3].tolist() == pytest.approx([1.0, 0.5755328536, 0.3312380910], rel=1e-6)
    f3 = O.rope_setup(O.FiTConfig(hidden_size=2304, num_heads=24))[0]
    assert f3[:3].tolist() == pytest.approx([1.0, 0.6812920570, 0.4641588926], rel=1e-6)
    cfg = O.FiTConfig(hidden_size=1152, num_heads=16, custom_freqs="ntk-aware-pro2", max_pe_len_h=20, max_pe_len_w=20,
                      decouple=False, ori_max_pe_len=16)
    assert O.rope_setup(cfg)[2] == pytest.approx(1.09794450, rel=1e-6)
    cfg.custom_freqs = "ntk-aware-pro1"
    assert O.rope_setup(cfg)[2] == pytest.approx(1.15346527, rel=1e-6)
    cfg.custom_freqs = "yarn"
    assert O.rope_setup(cfg)[2] == pytest.approx(1.02231431, rel=1e-6)
    assert O.rotate_half(torch.arange(8.0)).tolist() == [-1, 0, -3, 2, -5, 4, -7, 6]


def test_rope_rule_fixtures(golden_dir):
    kat = torch.load(os.path.join(golden_dir, "rope_kat.pt"))
    assert torch.equal(kat["rotate_half_0to7"], O.rotate_half(torch.arange(8.0)))
    for c in kat["cases"]:
        cfg = O.FiTConfig(hidden_size=c["head_dim"] * 2, num_heads=2, custom_freqs=c["custom_freqs"],
                          max_pe_len_h=c["max_pe_len_h"], max_pe_len_w=c["max_pe_len_w"], decouple=c["decouple"],
                          ori_max_pe_len=c["ori_max_pe_len"])
        fh, fw, mag = O.rope_setup(cfg)
        assert torch.allclose(fh, c["freqs_h"], rtol=1e-6, atol=0) and torch.allclose(fw, c["freqs_w"], rtol=1e-6, atol=0)
        assert mag == pytest.approx(c["mag"], rel=1e-6)
        gh, gw = c["grid_hw"]
        cos, sin = O.rope_cos_sin(cfg, O.make_grid(1, gh, gw))
        idx = [0, 1, gw + 1, gh * gw - 1]
        assert torch.allclose(cos[0, idx], c["cos_tok"], atol=2e-6) and torch.allclose(sin[0, idx], c["sin_tok"], atol=2e-6)


def test_grid_layout():
    g = O.make_grid(1, 10, 20)[0]
    assert g[0, :22].tolist() == list(range(20)) + [0, 1]          # w index runs fastest
    assert g[1, :22].tolist() == [0] * 20 + [1, 1]                 # h index
    cfg = O.FiTConfig(hidden_size=1152, num_heads=16, custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20,
                      decouple=True, ori_max_pe_len=16)
    cos, _ = O.rope_cos_sin(cfg, O.make_grid(1, 10, 20))
    assert cos[0, 21, 0:4].tolist() == pytest.approx([.5403023362, .5403023362, .8256267309, .8256267309], abs=1e-6)
    assert cos[0, 21, 36:40].tolist() == pytest.approx([.5403023362, .5403023362, .8300121427, .8300121427], abs=1e-6)


def test_tiny_models_match_reference(golden_dir):
    for fx in torch.load(os.path.join(golden_dir, "tiny_models.pt")):
        cfg = O.FiTConfig(**fx["cfg"])
        out = O.forward(cfg, fx["state_dict"], fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"])
        assert rel(out, fx["out"]) < TOL, fx["name"]
        out2 = O.forward(cfg, fx["state_dict"], fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"] * 2)
        assert rel(out2, fx["out_mask2"]) < TOL
        assert torch.equal(out2, out * 2)                                   # raw-mask multiply (fit_model.py:230)
        assert bool((out[fx["mask"] == 0] == 0).all())                      # pad rows exactly zero
        # the init restatement reproduces the reference init + re-draw bit-for-bit
        sd = O.synthetic_state_dict(cfg)
        assert all(torch.equal(sd[k], fx["state_dict"][k]) for k in sd)


def _checksum_ok(sd, want):
    for k, (s, a) in want.items():
        assert float(sd[k].double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-9), k
        assert float(sd[k].double().abs().sum()) == pytest.approx(a, rel=1e-9), k


def test_xl_depth2_padded_matches_reference(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "xl_depth2_padded.pt"))
    cfg = O.FiTConfig(hidden_size=1152, num_heads=16, adaln_lora_dim=288, **fx["cfg"])
    sd = O.synthetic_state_dict(cfg)
    _checksum_ok(sd, fx["weight_checksum"])
    a = (fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"])
    out = O.forward(cfg, sd, *a)
    assert rel(out, fx["out"]) < TOL
    assert rel(O.forward_with_cfg(cfg, sd, *a, None, 1.5), fx["out_cfg"]) < TOL
    assert rel(O.forward_with_cfg(cfg, sd, *a, None, 4.0, scale_pow=2.0), fx["out_cfg_pow"]) < TOL
    # forward_with_cfg: channels 0..11 guided + duplicated, 12..15 pass through per half (fit_model.py:253-275)
    oc = O.forward_with_cfg(cfg, sd, *a, None, 1.5)
    assert torch.equal(oc[:2, :, :12], oc[2:, :, :12])
    x2 = torch.cat([fx["x"][:2], fx["x"][:2]])
    base = O.forward(cfg, sd, x2, *a[1:])
    assert torch.equal(oc[:, :, 12:], base[:, :, 12:])
    # padding invariance: sample 0 (10x20 = 200 tokens) alone, unpadded
    solo = O.forward(cfg, sd, fx["x"][:1, :200], fx["t"][:1], fx["y"][:1], fx["grid"][:1, :, :200], fx["mask"][:1, :200])
    assert rel(solo, out[:1, :200]) < 1e-5
    lat = torch.randn(2, 200, 16, generator=torch.Generator().manual_seed(3))
    up = O.unpatchify(cfg, lat, (20, 40))
    assert up.shape == (2, 4, 20, 40) and torch.equal(up[:, :, :2], fx["unpatchify_out"])


def test_xl_config1_matches_reference(golden_dir):
    """BASELINE.json configs[0]: XL/2 depth 36, one CFG Euler step, 256 tokens, batch 2 (4 rows)."""
    fx = torch.load(os.path.join(golden_dir, "xl_config1.pt"))
    cfg = O.FiTConfig(**O.XL2)
    sd = O.synthetic_state_dict(cfg)
    assert sum(v.numel() for v in sd.values()) == fx["n_params"] == 671045776
    _checksum_ok(sd, fx["weight_checksum"])
    n, N = 2, 256
    torch.manual_seed(0)
    z = torch.randn(n, N, 16)
    y = torch.randint(0, 1000, (n,))
    assert torch.equal(z, fx["z"]) and torch.equal(y, fx["y"])
    grid, mask = O.make_grid(n, 16, 16), torch.ones(n, N)
    z1 = O.euler_cfg_sample(cfg, sd, z, y, grid, mask, None, steps=250, cfg_scale=1.5, first_steps=1)
    assert rel(z1, fx["z_step0"]) < TOL
    sig = torch.linspace(0, 1, 251)
    assert torch.equal(O.cfg_euler_update(z, fx["v_step0"], 1.5, sig[0], sig[1]), fx["z_step0"])
    assert torch.equal(O.cfg_euler_update(fx["z_step0"], fx["v_step1"], 1.5, sig[1], sig[2]), fx["z_step1"])


def test_flops_formula():
    """304.62 / 236.17 / 1348.35 / 1329.63 GFLOP per forward row."""
    xl, b3 = O.FiTConfig(**O.XL2), O.FiTConfig(**O.B3_2)
    assert O.flops_per_forward_row(xl, 256) / 1e9 == pytest.approx(304.62, abs=0.01)
    assert O.flops_per_forward_row(xl, 200) / 1e9 == pytest.approx(236.17, abs=0.01)
    assert O.flops_per_forward_row(xl, 1024) / 1e9 == pytest.approx(1348.35, abs=0.01)
    assert O.flops_per_forward_row(b3, 256) / 1e9 == pytest.approx(1329.63, abs=0.01)


def test_online_rope_forward_golden(golden_dir):
    """oracle online_rope path == the real reference's output (oracle/make_online_rope_golden.py)."""
    fx = torch.load(os.path.join(golden_dir, "xl_depth2_online.pt"))
    cfg = O.FiTConfig(hidden_size=1152, depth=2, num_heads=16, adaln_lora_dim=288, custom_freqs="ntk-aware", decouple=True,
                      ori_max_pe_len=16, max_pe_len_h=16, max_pe_len_w=16, online_rope=True)
    sd = O.synthetic_state_dict(cfg)
    out = O.forward(cfg, sd, fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"], fx["size"])
    assert rel(out, fx["out"]) < TOL
    # the product's host-side frequency rule == the oracle's
    from fitv2_b200.rope import online_rope_frequencies
    for cf in ("linear", "ntk-aware", "ntk-by-parts"):
        for dec in (False, True):
            fh, fw = online_rope_frequencies(72, cf, 10000.0, dec, 16, fx["size"])
            s = fx["size"].reshape(-1, 2)
            sh, sw = (s[:, 0], s[:, 1]) if dec else (torch.max(s[:, 0], s[:, 1]),) * 2
            assert torch.equal(fh, O.rope_1d_freqs_online(cf, 10000.0, 36, sh, 16).float())
            assert torch.equal(fw, O.rope_1d_freqs_online(cf, 10000.0, 36, sw, 16).float())


def test_swiglu_large_golden(golden_dir):
    """use_swiglu_large=True (modules.py:248-249: SwiGLU hidden int(D*mlp_ratio) = 4608): the oracle replays the REAL
    reference's depth-1 forward (oracle/make_swiglu_large_golden.py); fitv2_b200.FiT draws the same weights."""
    from fitv2_b200 import FiT, make_grid
    fx = torch.load(os.path.join(golden_dir, "swiglu_large_xl_d1.pt"))
    cfg = O.FiTConfig(depth=1, use_swiglu_large=True)
    assert cfg.mlp_hidden == 4608 and O.FiTConfig(depth=1).mlp_hidden == 3072
    torch.manual_seed(fx["init_seed"])
    m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, use_swiglu_large=True, q_norm="layernorm", k_norm="layernorm",
            adaln_type="lora", depth=1, hidden_size=1152, num_heads=16, adaln_lora_dim=288).randomize_zero_init_(fx["redraw_seed"])
    sd = m.state_dict()
    assert m.mlp_hidden == 4608 and sd["blocks.0.mlp.fc2.weight"].shape == (1152, 4608)
    assert float(sd["blocks.0.mlp.fc1_g.weight"].double().abs().sum()) == pytest.approx(fx["fc1_g_abs_sum"], rel=1e-12)
    R = fx["x"].shape[0]
    out = O.forward(cfg, sd, fx["x"], fx["t"], fx["y"], make_grid(R, fx["hp"], fx["wp"]), torch.ones(R, fx["hp"] * fx["wp"]))
    assert rel(out, fx["v_ref"]) < TOL


def test_ctor_variant_goldens(golden_dir):
    """The remaining constructor switches (oracle/make_ctor_variant_goldens.py, outputs of the REAL reference classes): the class
    defaults (GELU Mlp, learn_sigma, (B,C,N), adaLN 'normal'), bias-free qkv / ffn without rotation, rotation of v under 'XPOS',
    SwiGLU modulation MLPs.  The oracle regenerates the weights from the seeds and replays forward / forward_with_cfg."""
    import hashlib
    cases = torch.load(os.path.join(golden_dir, "ctor_variants_xl.pt"))
    assert [c["name"] for c in cases] == ["defaults", "nobias_norope", "rope_v", "adaln_swiglu"]
    for c in cases:
        cfg = O.FiTConfig(**c["oracle_kwargs"])
        sd = O.synthetic_state_dict(cfg)
        assert list(sd.keys()) == c["keys"], c["name"]
        a = (c["x"], c["t"], c["y"], c["grid"], c["mask"])
        assert rel(O.forward(cfg, sd, *a), c["out"]) < TOL, c["name"]
        assert rel(O.forward_with_cfg(cfg, sd, *a, None, 1.5), c["out_cfg"]) < TOL, c["name"]
    d = {c["name"]: c for c in cases}
    assert "blocks.0.mlp.fc1.weight" in d["defaults"]["keys"] and "blocks.0.mlp.fc1_g.weight" not in d["defaults"]["keys"]
    assert "blocks.0.attn.qkv.bias" not in d["nobias_norope"]["keys"] and "blocks.0.mlp.fc2.bias" not in d["nobias_norope"]["keys"]
    assert "blocks.0.attn.proj.bias" in d["nobias_norope"]["keys"]          # proj keeps its bias (modules.py:151)
    assert "blocks.0.adaLN_modulation.fc1_g.weight" in d["adaln_swiglu"]["keys"] and "final_layer.adaLN_modulation.fc2.bias" in d["adaln_swiglu"]["keys"]
