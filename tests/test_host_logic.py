"""Host-side logic of the drop-in (CPU only): constructor contract, state_dict keys, init parity with the
reference scheme, RoPE frequency rules, weight packing, C-ABI library exports and error behaviour."""
import ctypes
import os
import re

import pytest
import torch

from oracle import fitv2_oracle as O
from fitv2_b200 import FiT, FitV2Error, _lib
from fitv2_b200.rope import rope_frequencies

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora")
XL1 = dict(hidden_size=1152, depth=1, num_heads=16, adaln_lora_dim=288)


def test_rope_rules_match_oracle():
    cases = [(72, "normal", None, None, False, None), (96, "normal", None, None, False, None),
             (72, "ntk-aware", 10, 20, True, 16), (72, "ntk-aware", 32, 32, True, 16), (72, "ntk-aware", 10, 20, False, 16),
             (72, "linear", 10, 20, True, 16), (72, "ntk-aware-pro1", 20, 20, False, 16), (72, "ntk-aware-pro2", 20, 20, False, 16),
             (72, "ntk-by-parts", 24, 40, True, 16), (72, "yarn", 24, 40, True, 16), (96, "yarn", 32, 32, False, 16)]
    for dh, cf, h, w, dec, ori in cases:
        fh, fw, mag = rope_frequencies(dh, cf, 10000.0, h, w, dec, ori)
        cfg = O.FiTConfig(hidden_size=dh * 2, num_heads=2, custom_freqs=cf, max_pe_len_h=h, max_pe_len_w=w, decouple=dec, ori_max_pe_len=ori)
        ofh, ofw, omag = O.rope_setup(cfg)
        assert torch.equal(fh, ofh) and torch.equal(fw, ofw) and mag == omag, (dh, cf)
    with pytest.raises(ValueError):
        rope_frequencies(72, "bogus")
    with pytest.raises(ValueError):
        rope_frequencies(72, "yarn")            # missing max_pe_len / ori_max_pe_len


def test_constructor_contract_and_rejections():
    m = FiT(**KW, **XL1, context_size=256, patch_size=2, in_channels=4, mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000,
            use_swiglu_large=False, use_checkpoint=False, qk_norm_weight=False, rel_pos_embed="rope", abs_pos_embed=None,
            custom_freqs="normal", online_rope=False)     # every key of configs/fitv2/config_fitv2_xl.yaml:26-47 + injected keys
    assert m.in_channels == 4 and m.dtype == torch.float32 and m.mlp_hidden == 3072 and m.head_dim == 72
    assert m.get_attention_maps() is None and m.disable_attention_visualization() is None     # fit_model.py:301-330 with save_attention off
    with pytest.raises(NotImplementedError):
        m.enable_attention_visualization()
    for bad in (dict(adaln_type="bogus"), dict(online_rope=True), dict(q_norm="batchnorm"), dict(norm_type="none"),
                dict(num_heads=18), dict(operand_dtype="fp8"), dict(adaln_bias=False), dict(save_attention=True),
                dict(use_swiglu=False, mlp_ratio=3.9)):
        with pytest.raises(NotImplementedError):
            FiT(**{**KW, **XL1, **bad})
    with pytest.raises(AssertionError):                    # fit_model.py:68: assert not (learn_sigma and use_sit)
        FiT(**{**KW, **XL1, "learn_sigma": True})
    with pytest.raises(FitV2Error):                        # no CPU fallback
        m(torch.zeros(1, 16, 16), torch.zeros(1), torch.zeros(1, dtype=torch.long), torch.zeros(1, 2, 16, dtype=torch.long), torch.ones(1, 16))


# The literal `params:` blocks of the reference's model configs (configs/fitv2/config_fitv2_{xl,3B,hr_xl}.yaml and
# configs/fit/config_fit_xl.yaml, `network_config.params`), as text: instantiate_from_config (fit/utils/utils.py:76-93) calls
# `FiT(**params)` after sample_fitv2_ddp.py:75-99 has injected the resolution-dependent RoPE keys.
_YAML_PARAMS = {
    "config_fitv2_xl": """
      context_size: 256
      patch_size: 2
      in_channels: 4
      hidden_size: 1152
      depth: 36
      num_heads: 16
      mlp_ratio: 4.0
      class_dropout_prob: 0.1
      num_classes: 1000
      learn_sigma: false
      use_sit: true
      use_swiglu: true
      use_swiglu_large: false
      use_checkpoint: false
      q_norm: layernorm
      k_norm: layernorm
      qk_norm_weight: false
      rel_pos_embed: rope
      abs_pos_embed: null
      adaln_type: lora
      adaln_lora_dim: 288
    """,
    "config_fitv2_3B": """
      context_size: 256
      patch_size: 2
      in_channels: 4
      hidden_size: 2304
      depth: 40
      num_heads: 24
      mlp_ratio: 4.0
      class_dropout_prob: 0.1
      num_classes: 1000
      learn_sigma: false
      use_sit: true
      use_swiglu: true
      use_swiglu_large: false
      q_norm: layernorm
      k_norm: layernorm
      qk_norm_weight: false
      rel_pos_embed: rope
      abs_pos_embed: null
      adaln_type: lora
      adaln_lora_dim: 576
    """,
    "config_fitv2_hr_xl": """
      context_size: 1024
      patch_size: 2
      in_channels: 4
      hidden_size: 1152
      depth: 36
      num_heads: 16
      mlp_ratio: 4.0
      class_dropout_prob: 0.1
      num_classes: 1000
      learn_sigma: false
      use_sit: true
      use_swiglu: true
      use_swiglu_large: false
      use_checkpoint: true
      q_norm: layernorm
      k_norm: layernorm
      qk_norm_weight: false
      rel_pos_embed: rope
      custom_freqs: ntk-aware
      decouple: true
      ori_max_pe_len: 16
      online_rope: true
      abs_pos_embed: null
      adaln_type: lora
      adaln_lora_dim: 288
    """,
    "config_fit_xl": """
      context_size: 256
      patch_size: 2
      in_channels: 4
      hidden_size: 1152
      depth: 28
      num_heads: 16
      mlp_ratio: 4.0
      class_dropout_prob: 0.1
      num_classes: 1000
      learn_sigma: true
      use_swiglu: true
      use_swiglu_large: true
      rel_pos_embed: rope
    """,
}


def _parse_params(text):
    """Flat `key: scalar` YAML (omegaconf is not installed): ints, floats, booleans, null, strings."""
    out = {}
    for line in text.strip().splitlines():
        k, v = [p.strip() for p in line.split(":", 1)]
        if v in ("true", "false"): out[k] = v == "true"
        elif v in ("null", "~"): out[k] = None
        elif re.fullmatch(r"-?\d+", v): out[k] = int(v)
        elif re.fullmatch(r"-?\d+\.\d*", v): out[k] = float(v)
        else: out[k] = v
    return out


def test_yaml_params_match_the_reference_configs():
    """The embedded text is the reference's own (checked where /root/reference exists, i.e. in the build container)."""
    ref = "/root/reference/configs"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    for name, text in _YAML_PARAMS.items():
        sub = "fit" if name == "config_fit_xl" else "fitv2"
        with open(os.path.join(ref, sub, name + ".yaml")) as f:
            src = f.read()
        block = src.split("network_config:")[1].split("params:")[1]
        got = {}
        for line in block.splitlines()[1:]:
            if not line.strip() or line.strip().startswith("#"):
                continue
            if len(line) - len(line.lstrip()) < 6:            # dedent: end of the params block
                break
            got.update(_parse_params(line.split("#")[0]))
        assert got == _parse_params(text), name


@pytest.mark.parametrize("name", sorted(_YAML_PARAMS))
def test_constructor_from_the_reference_yaml(name):
    """`FiT(**params)` exactly as instantiate_from_config builds it, with the keys sample_fitv2_ddp.py:75-99 injects for a
    160x320 ntk-aware decoupled run (the hr_xl config carries use_checkpoint: true, which inference accepts as a no-op)."""
    params = _parse_params(_YAML_PARAMS[name])
    params["depth"] = 1                                                       # CPU test: one block is enough for the contract
    injected = dict(custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20, decouple=True, ori_max_pe_len=16, online_rope=False)
    if name == "config_fit_xl":                                               # the FiTv1 scripts do not touch the RoPE keys
        injected = {}
    m = FiT(**{**params, **injected})
    assert not m.online_rope
    full = _parse_params(_YAML_PARAMS[name])
    assert m.hidden_size == full["hidden_size"] and m.num_heads == full["num_heads"] and m.in_channels == 4
    assert m.learn_sigma == full["learn_sigma"] and m.use_sit == full.get("use_sit", False)
    assert m.out_channels == (8 if full["learn_sigma"] else 4)
    cfg = O.FiTConfig(hidden_size=full["hidden_size"], depth=1, num_heads=full["num_heads"], adaln_lora_dim=full.get("adaln_lora_dim") or 0,
                      learn_sigma=full["learn_sigma"], use_sit=full.get("use_sit", False), adaln_type=full.get("adaln_type", "normal"),
                      q_norm=full.get("q_norm"), k_norm=full.get("k_norm"), use_swiglu_large=full["use_swiglu_large"])
    assert list(m.state_dict().keys()) == list(O.reference_init_state_dict(cfg, 0).keys())


def test_variant_state_dict_keys_and_init_parity():
    """FiTv1 layout (adaLN 'normal', learn_sigma, no q/k norm) and the weighted norm kinds: parameter names, shapes and the
    init under a seed equal the oracle's (which make_variant_goldens.py pins bit-equal to the real reference classes)."""
    for kw, okw in ((dict(learn_sigma=True, use_sit=False, use_swiglu=True, use_swiglu_large=True),
                     dict(learn_sigma=True, use_sit=False, adaln_type="normal", q_norm=None, k_norm=None, use_swiglu_large=True, adaln_lora_dim=0)),
                    (dict(**KW, adaln_lora_dim=288, norm_type="rmsnorm", qk_norm_weight=True),
                     dict(adaln_lora_dim=288, norm_type="rmsnorm", qk_norm_weight=True))):
        torch.manual_seed(0)
        m = FiT(hidden_size=1152, depth=1, num_heads=16, **kw)
        ref = O.reference_init_state_dict(O.FiTConfig(hidden_size=1152, depth=1, num_heads=16, **okw), 0)
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        assert all(torch.equal(sd[k], ref[k]) for k in ref)
        P = m.pack_weights(torch.device("cpu"))
        if not m.use_sit:
            assert "GLOBAL_ADALN_W" not in P and P["NORMAL_ADALN_W"].shape == (1, 6 * 1152, 1152) and P["FINAL_LINEAR_W"].shape == (32, 1152)
        else:
            assert P["NORM1_W"].shape == (1, 1152) and P["Q_NORM_W"].shape == (1, 72) and P["K_NORM_W"].shape == (1, 72)


def test_ctor_variants_state_dict_init_and_packing(golden_dir):
    """The remaining constructor switches of fit_model.py:25-65 (goldens of oracle/make_ctor_variant_goldens.py): parameter names
    and the init under a seed equal the reference's, and the packed kernel-side weights reproduce the reference arithmetic."""
    F = torch.nn.functional
    cases = {c["name"]: c for c in torch.load(os.path.join(golden_dir, "ctor_variants_xl.pt"))}
    models = {}
    for name, c in cases.items():
        torch.manual_seed(0)
        m = FiT(**c["kwargs"])
        ref = O.reference_init_state_dict(O.FiTConfig(**c["oracle_kwargs"]), 0)
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys()) == c["keys"], name
        assert all(torch.equal(sd[k], ref[k]) for k in ref), name
        m.randomize_zero_init_(1)
        models[name] = (m, m.pack_weights(torch.device("cpu")))
    D = 1152
    # class defaults: timm Mlp, fc1 alone in the GATEUP slots (no interleave), hidden 4 D
    m, P = models["defaults"]
    assert not m.use_swiglu and m.mlp_hidden == 4 * D and not m.use_sit and m.learn_sigma and m.adaln_type == "normal"
    assert P["GATEUP_W"].shape == (2, 4 * D, D) and P["GATEUP_B"].shape == (2, 4 * D) and P["FC2_W"].shape == (2, D, 4 * D)
    assert torch.equal(P["GATEUP_W"][1].float(), m.blocks[1].mlp.fc1.weight.bfloat16().float()) and torch.equal(P["GATEUP_B"][1], m.blocks[1].mlp.fc1.bias)
    assert "GLOBAL_ADALN_W" not in P and "NORMAL_ADALN_W" in P and "FINAL_ADALN_W" in P
    # bias-free qkv / ffn: zero biases bound in their place; proj keeps its own; no rotation = zero frequencies
    m, P = models["nobias_norope"]
    assert m.blocks[0].attn.qkv.bias is None and m.blocks[0].mlp.fc1_g.bias is None and m.blocks[0].mlp.fc2.bias is None
    for k in ("QKV_B", "GATEUP_B", "FC2_B", "ROPE_FREQS_H", "ROPE_FREQS_W"):
        assert not bool(P[k].any()), k
    assert P["QKV_B"].shape == (1, 3 * D) and P["GATEUP_B"].shape == (1, 2 * 3072) and P["FC2_B"].shape == (1, D)
    assert bool(P["PROJ_B"].any()) and not m.rotates and m.rel_pos_embed is None
    # rotation of v: 'XPOS' is lower-cased and rotates like 'rope' (modules.py:153,170)
    m, P = models["rope_v"]
    assert m.rotates and m.add_rel_pe_to_v and bool(P["ROPE_FREQS_H"].any())
    # SwiGLU modulation MLPs: the launch sequence of the C side (g, silu(g); (x-linear) * silu(g); fc2) from the packed slots
    m, P = models["adaln_swiglu"]
    assert set(P) >= {"SG_G_W", "SG_X_B", "SG_FC2_W", "FSG_G_W", "FSG_FC2_B"} and not ({"FINAL_ADALN_W", "GLOBAL_ADALN_W", "NORMAL_ADALN_W"} & set(P))
    assert P["SG_G_W"].shape == (1, 864, D) and P["SG_FC2_W"].shape == (1, 6 * D, 864) and P["FSG_X_W"].shape == (576, D) and P["FSG_FC2_W"].shape == (2 * D, 576)
    cfg, sd = O.FiTConfig(**cases["adaln_swiglu"]["oracle_kwargs"]), {k: v.detach() for k, v in m.state_dict().items()}
    c = torch.randn(3, D)
    g = c @ P["SG_G_W"][0].t() + P["SG_G_B"][0]
    mod = ((c @ P["SG_X_W"][0].t() + P["SG_X_B"][0]) * F.silu(g)) @ P["SG_FC2_W"][0].t() + P["SG_FC2_B"][0]
    assert torch.allclose(mod, O.block_modulation(cfg, sd, c, 0, 0.0), atol=1e-6)
    assert set(P) == set(_lib.WEIGHT_SLOTS) - {"GLOBAL_ADALN_W", "GLOBAL_ADALN_B", "LORA_A_W", "LORA_A_B", "LORA_B_W", "LORA_B_B", "FINAL_ADALN_W",
                                               "FINAL_ADALN_B", "NORMAL_ADALN_W", "NORMAL_ADALN_B", "NORM1_W", "NORM2_W", "NORM_FINAL_W", "Q_NORM_W", "K_NORM_W"}
    # a reference-default model: FiT() itself constructs (depth 28, GELU Mlp) -- only the meta device is touched here
    with torch.device("meta"):
        big = FiT()
    assert big.depth == 28 and not big.use_swiglu and big.out_channels == 8


def test_state_dict_keys_and_init_parity():
    torch.manual_seed(0)
    m = FiT(**KW, **XL1)
    cfg = O.FiTConfig(**XL1)
    ref = O.reference_init_state_dict(cfg, 0)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]) for k in ref)
    # SURVEY.md A.3 names
    for k in ("x_embedder.proj.weight", "t_embedder.mlp.0.weight", "t_embedder.mlp.2.bias", "y_embedder.embedding_table.weight",
              "global_adaLN_modulation.1.weight", "blocks.0.attn.qkv.weight", "blocks.0.attn.proj.bias", "blocks.0.mlp.fc1_g.weight",
              "blocks.0.mlp.fc1_x.bias", "blocks.0.mlp.fc2.weight", "blocks.0.adaLN_modulation.1.weight", "blocks.0.adaLN_modulation.2.bias",
              "final_layer.linear.weight", "final_layer.adaLN_modulation.1.bias"):
        assert k in sd
    m.randomize_zero_init_(1)
    syn = O.synthetic_state_dict(cfg)
    assert all(torch.equal(m.state_dict()[k], syn[k]) for k in syn)
    # checkpoint round trip with the reference's strict=False loader semantics
    m2 = FiT(**KW, **XL1)
    missing, unexpected = m2.load_state_dict(syn, strict=False)
    assert not missing and not unexpected and m2._packed is None
    assert torch.equal(m2.unpatchify(torch.arange(2 * 4 * 16.).reshape(2, 4, 16), (4, 4)),
                       O.unpatchify(cfg, torch.arange(2 * 4 * 16.).reshape(2, 4, 16), (4, 4)))


def test_weight_packing_layout():
    torch.manual_seed(0)
    m = FiT(**KW, **XL1).randomize_zero_init_(1)
    P = m.pack_weights(torch.device("cpu"))
    assert set(P) == set(_lib.WEIGHT_SLOTS[:27])           # the FiTv2 family binds the first 27 slots; the rest belong to the variants
    D, Hm = 1152, 3072
    assert P["QKV_W"].shape == (1, 3 * D, D) and P["QKV_W"].dtype == torch.bfloat16
    assert P["GATEUP_W"].shape == (1, 2 * Hm, D) and P["GATEUP_B"].shape == (1, 2 * Hm)
    g, u = m.blocks[0].mlp.fc1_g, m.blocks[0].mlp.fc1_x
    W, B = P["GATEUP_W"][0].float(), P["GATEUP_B"][0]
    for t in (0, 5, 23):                                   # tile t: [128 gate rows | 128 up rows]
        assert torch.equal(W[256 * t:256 * t + 128], g.weight[128 * t:128 * t + 128].bfloat16().float())
        assert torch.equal(W[256 * t + 128:256 * t + 256], u.weight[128 * t:128 * t + 128].bfloat16().float())
        assert torch.equal(B[256 * t:256 * t + 128], g.bias[128 * t:128 * t + 128])
        assert torch.equal(B[256 * t + 128:256 * t + 256], u.bias[128 * t:128 * t + 128])
    # emulate the fused epilogue from the packed layout and compare with timm SwiGLU semantics
    x = torch.randn(5, D)
    acc = x @ W.t() + B
    hid = torch.cat([torch.nn.functional.silu(acc[:, 256 * t:256 * t + 128]) * acc[:, 256 * t + 128:256 * t + 256] for t in range(Hm // 128)], 1)
    ref = torch.nn.functional.silu(x @ g.weight.bfloat16().float().t() + g.bias) * (x @ u.weight.bfloat16().float().t() + u.bias)
    assert torch.allclose(hid, ref, atol=1e-5)
    assert P["LORA_A_W"].shape == (1, 288, D) and P["LORA_B_W"].shape == (1, 6 * D, 288) and P["ROPE_FREQS_H"].shape == (18,)
    m16 = FiT(**KW, **XL1, operand_dtype="fp16")
    assert m16.pack_weights(torch.device("cpu"))["FC2_W"].dtype == torch.float16
    # the four adaLN matrices are tf32 operands of csrc/cond_tc.cuh: rounded to nearest, low 13 mantissa bits zero
    for slot, ref in (("GLOBAL_ADALN_W", m.global_adaLN_modulation[1].weight), ("FINAL_ADALN_W", m.final_layer.adaLN_modulation[1].weight),
                      ("LORA_A_W", m.blocks[0].adaLN_modulation[1].weight[None]), ("LORA_B_W", m.blocks[0].adaLN_modulation[2].weight[None])):
        w = P[slot]
        assert w.dtype == torch.float32 and w.shape == ref.shape
        assert int((w.view(torch.int32) & 0x1FFF).abs().max()) == 0
        assert float(((w - ref).abs() / ref.abs().clamp_min(1e-30)).max()) <= 2.0 ** -11
    assert torch.equal(P["T_MLP2_W"], m.t_embedder.mlp[2].weight)            # everything else stays fp32-exact


def test_library_exports_every_declared_symbol(built_lib):
    with open(os.path.join(ROOT, "include", "fitv2_b200.h")) as f:
        hdr = f.read()
    declared = sorted(set(re.findall(r"\b(fitv2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 14
    lib = ctypes.CDLL(built_lib)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/fitv2_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    # enum order in the header == python slot table
    enum = re.findall(r"FITV2_W_([A-Z0-9_]+)\s*(?:=\s*0)?,", hdr.split("enum fitv2_weight")[1].split("FITV2_W_COUNT")[0])
    assert enum == _lib.WEIGHT_SLOTS


def test_c_abi_argument_errors(built_lib):
    lib = _lib.load()
    assert b"sm_100a" in lib.fitv2_version()
    h = ctypes.c_void_p()
    bad = _lib.FitV2Config(1000, 1, 16, 72, 3072, 288, 16, 1001, 0, 1.0, 1.0)       # hidden != heads*head_dim
    assert lib.fitv2_create(ctypes.byref(bad), ctypes.byref(h)) == -1 and b"hidden_size" in lib.fitv2_last_error()
    bad = _lib.FitV2Config(1024, 1, 16, 64, 3072, 288, 16, 1001, 0, 1.0, 1.0)       # head_dim 64 not built
    assert lib.fitv2_create(ctypes.byref(bad), ctypes.byref(h)) == -1 and b"head_dim" in lib.fitv2_last_error()
    assert lib.fitv2_cfg_euler(None, None, 1.5, 0.004, None, 1, 1, 16, None) == -1
    assert lib.fitv2_cfg_combine(None, None, 1.5, 1, 1, 16, 12, None) == -1
    if not torch.cuda.is_available():
        ok = _lib.FitV2Config(1152, 1, 16, 72, 3072, 288, 16, 1001, 0, 1.0, 1.0)
        assert lib.fitv2_create(ctypes.byref(ok), ctypes.byref(h)) == -3              # no device: CUDA error, not a fallback


def test_init_from_ckpt_rules(tmp_path):
    """fit/utils/eval_utils.py:12-71: safetensors / torch files, `_orig_mod.` prefix handling, ignore_keys, strict=False."""
    import torch
    from safetensors.torch import save_file
    from fitv2_b200 import FiT, init_from_ckpt
    kw = dict(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora",
              hidden_size=144, depth=1, num_heads=2, adaln_lora_dim=8)
    torch.manual_seed(0)
    src = FiT(**kw).randomize_zero_init_(1)
    sd = {k: v.detach().clone().contiguous() for k, v in src.state_dict().items()}
    # 1. plain safetensors (model_ema.safetensors of the released checkpoints)
    f1 = str(tmp_path / "model_ema.safetensors")
    save_file(sd, f1)
    torch.manual_seed(5)
    m = FiT(**kw)
    missing, unexpected = init_from_ckpt(m, f1, ignore_keys=None, verbose=True)
    assert missing == [] and unexpected == []
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    # 2. checkpoint saved from a torch.compile'd model: `_orig_mod.` prefix is stripped
    f2 = str(tmp_path / "compiled.safetensors")
    save_file({f"_orig_mod.{k}": v for k, v in sd.items()}, f2)
    m2 = FiT(**kw)
    assert init_from_ckpt(m2, f2) == ([], [])
    assert all(torch.equal(v, sd[k]) for k, v in m2.state_dict().items())
    # 3. torch.save file, ignore_keys as regular expressions, strict=False semantics
    f3 = str(tmp_path / "ckpt.pt")
    torch.save(dict(sd, extra_buffer=torch.zeros(1)), f3)
    m3 = FiT(**kw)
    before = m3.final_layer.linear.weight.detach().clone()
    missing, unexpected = init_from_ckpt(m3, f3, ignore_keys=[r"final_layer\.linear"])
    assert sorted(missing) == ["final_layer.linear.bias", "final_layer.linear.weight"] and unexpected == ["extra_buffer"]
    assert torch.equal(m3.final_layer.linear.weight, before)
    assert torch.equal(m3.x_embedder.proj.weight, sd["x_embedder.proj.weight"])


def test_pretrain_ckpt_ignore_keys_and_finetune(tmp_path):
    """fit_model.py:114-115,159-170,291-299: `pretrain_ckpt` loads a checkpoint on top of the init inside the constructor, skipping
    every key that CONTAINS one of `ignore_keys`; `finetune` (anything but 'full') freezes all parameters except those whose name
    contains one of the same strings."""
    from safetensors.torch import save_file
    torch.manual_seed(0)
    a = FiT(**KW, **XL1).randomize_zero_init_(1)
    path = str(tmp_path / "model_ema.safetensors")
    save_file({k: v.detach().contiguous() for k, v in a.state_dict().items()}, path)
    torch.manual_seed(5)
    b = FiT(**KW, **XL1, pretrain_ckpt=path, ignore_keys=["final_layer.linear", "y_embedder"], finetune="partial")
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if "final_layer.linear" in k:
            assert not bool(sb[k].any()), k                                  # left at its (zero) init
        elif "y_embedder" in k:
            assert not torch.equal(sa[k], sb[k]), k                          # left at its own N(0, 0.02) draw
        else:
            assert torch.equal(sa[k], sb[k]), k
    grads = {n for n, p in b.named_parameters() if p.requires_grad}
    assert grads == {n for n, _ in b.named_parameters() if "final_layer.linear" in n or "y_embedder" in n} and len(grads) == 3
    c = FiT(**KW, **XL1, pretrain_ckpt=path, finetune="full")
    assert all(p.requires_grad for p in c.parameters()) and all(torch.equal(sa[k], v) for k, v in c.state_dict().items())


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line on the same metric / unit, with
    `impl`, `cpu_baseline` and a zero-copy `e2e`, no GPU needed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-budget", "10"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/sec" and line["higher_is_better"] is True
    assert line["metric"].startswith("images/sec FiTv2-XL/2 256^2 250-step ODE CFG 1.5")
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["vs_baseline"] is None and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == dict(value=line["value"], unit="images/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_missing_extension_fails_loudly():
    """No CPU / PyTorch fallback: with the shared object absent the loader raises (it never routes around the CUDA path)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from fitv2_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.FitV2Error as e:\n    print('LOUD', 'no CPU or PyTorch fallback' in str(e))\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root,
                         env={**os.environ, "FITV2_B200_LIB": "/nonexistent/libfitv2_b200.so"})
    assert out.returncode == 0 and out.stdout.strip().endswith("LOUD True"), out.stdout + out.stderr[-1000:]


def test_deepcopy_and_option_forwarding_are_per_instance():
    """A copied module owns its own (lazily created) C handle and workspace; the environment -> option table only names options
    the library declares in its header."""
    import copy
    torch.manual_seed(0)
    m = FiT(**KW, **XL1)
    m._handle, m._packed = object(), {"x": 1}                  # pretend the original has run
    c = copy.deepcopy(m)
    assert c._handle is None and c._packed is None and c._workspace is None
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    m._handle = None                                           # (the fake handle must not reach fitv2_destroy)
    with open(os.path.join(ROOT, "include", "fitv2_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r'"([a-z0-9_]+)" \(', hdr.split("fitv2_set_option")[0].split("Per-handle tuning switches")[1]))
    assert {name for name, _ in _lib._ENV_OPTIONS.values()} <= declared, ({name for name, _ in _lib._ENV_OPTIONS.values()} - declared)
