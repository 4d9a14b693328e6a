"""CPU fp32 oracle for the FiTv2 denoising hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch restatement of the reference algorithm
(DogyunPark/FiTv2).  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  The product path (``fitv2_b200``) never does.

Parity pinning: the reference ships NO tests / golden vectors.
This restatement is pinned instead by ``oracle/make_golden.py``, which imports
the real reference classes from ``/root/reference`` (with a timm shim and the
``save_attention`` fix), checks bit-equality with this file on
CPU and writes the fixtures under ``tests/golden/``.

Reference lines followed (paths relative to the reference repo):
  * forward ................ fit/model/fit_model.py:189-233
  * forward_with_cfg ....... fit/model/fit_model.py:235-275
  * unpatchify ............. fit/model/fit_model.py:171-187
  * initialize_weights ..... fit/model/fit_model.py:117-157
  * PatchEmbedder .......... fit/model/modules.py:34-37
  * TimestepEmbedder ....... fit/model/modules.py:52-76
  * LabelEmbedder .......... fit/model/modules.py:101-106
  * Attention .............. fit/model/modules.py:159-207
  * FiTBlock ............... fit/model/modules.py:270-274
  * FinalLayer ............. fit/model/modules.py:292-296
  * modulate ............... fit/model/utils.py:6-7
  * LayerNorm (no affine) .. fit/model/norms.py:41-42
  * RoPE frequency rules ... fit/model/rope.py:24-53,134-170,173-231
  * RoPE cached lookup ..... fit/model/rope.py:308-333 ; rotate_half :107-111
  * SwiGLU ................. timm.layers.mlp.SwiGLU (third party, un-vendored,
                             unpinned in requirements.txt:12): fc2(silu(fc1_g(x)) * fc1_x(x));
                             call site fit/model/modules.py:247-251
  * Euler / CFG loop ....... sample_fitv2_ddp.py:257-314
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# configuration
# --------------------------------------------------------------------------- #
@dataclass
class FiTConfig:
    """Constructor contract of fit.model.fit_model.FiT (fit_model.py:25-65),
    restricted to the FiTv2 family the hot path covers."""
    context_size: int = 256
    patch_size: int = 2
    in_channels: int = 4
    hidden_size: int = 1152
    depth: int = 36
    num_heads: int = 16
    mlp_ratio: float = 4.0
    num_classes: int = 1000
    class_dropout_prob: float = 0.1
    adaln_lora_dim: int = 288
    rope_theta: float = 10000.0
    custom_freqs: str = "normal"
    max_pe_len_h: Optional[int] = None
    max_pe_len_w: Optional[int] = None
    decouple: bool = False
    ori_max_pe_len: Optional[int] = None
    time_shifting: int = 1
    max_cached_len: int = 256  # rope.py:126
    online_rope: bool = False  # fit_model.py:57,212-214: per-sample frequencies from `size`
    use_swiglu_large: bool = False  # modules.py:248-249: SwiGLU hidden = int(hidden*mlp_ratio) instead of 2/3 of it
    # FiTv1-style variants (configs/fit/config_fit_xl.yaml:20-36) and the other norm kinds of norms.py:35-50
    learn_sigma: bool = False       # fit_model.py:78: out_channels = 2 * in_channels
    use_sit: bool = True            # fit_model.py:204,231: False = (B, C, N) tensors in and out
    adaln_type: str = "lora"        # 'lora' (modules.py:259-264) or 'normal' (modules.py:254-258, no global adaLN)
    norm_type: str = "layernorm"    # block norms: 'layernorm' (no affine) | 'w_layernorm' | 'rmsnorm'
    q_norm: Optional[str] = "layernorm"
    k_norm: Optional[str] = "layernorm"
    qk_norm_weight: bool = False    # modules.py:141-144: promotes 'layernorm' to 'w_layernorm'
    # the remaining constructor switches of fit_model.py:25-65 (no shipped config sets them; the class DEFAULTS do)
    use_swiglu: bool = True         # modules.py:246-253: False = timm Mlp, fc2(GELU_tanh(fc1(x))), hidden int(D*mlp_ratio)
    qkv_bias: bool = True           # modules.py:140
    ffn_bias: bool = True           # modules.py:248-253
    rel_pos_embed: Optional[str] = "rope"   # modules.py:153,170: q / k are rotated only for 'rope' / 'xpos' (lower-cased)
    add_rel_pe_to_v: bool = False   # modules.py:171-172: v is rotated too

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def mlp_hidden(self) -> int:
        # modules.py:246-251 : (int(hidden*mlp_ratio) * 2) // 3, or the full int(hidden*mlp_ratio) with swiglu_large
        full = int(self.hidden_size * self.mlp_ratio)
        if not self.use_swiglu:
            return full                                                  # modules.py:253 (timm Mlp)
        return full if self.use_swiglu_large else (full * 2) // 3

    @property
    def rotates(self) -> bool:
        """modules.py:153,170: `rel_pos_embed.lower() in ['rope', 'xpos']`."""
        return self.rel_pos_embed is not None and self.rel_pos_embed.lower() in ("rope", "xpos")

    @property
    def adaln_hidden(self) -> Tuple[int, int]:
        """adaln_type 'swiglu': SwiGLU hidden of the block / final modulation (modules.py:265-268,285)."""
        return (self.hidden_size // 4) * 3, self.hidden_size // 2

    @property
    def token_channels(self) -> int:
        return self.in_channels * self.patch_size ** 2

    @property
    def out_token_channels(self) -> int:
        return self.token_channels * (2 if self.learn_sigma else 1)

    def norm_kind(self, which: str) -> Optional[str]:
        """Resolved create_norm name of 'block' / 'q' / 'k' (modules.py:141-147)."""
        if which == "block":
            return self.norm_type.lower()
        n = self.q_norm if which == "q" else self.k_norm
        if n is None or n == "" or n.lower() == "none":
            return None
        n = n.lower()
        return "w_layernorm" if (n == "layernorm" and self.qk_norm_weight) else n


XL2 = dict(hidden_size=1152, depth=36, num_heads=16, adaln_lora_dim=288)   # configs/fitv2/config_fitv2_xl.yaml:26-47
B3_2 = dict(hidden_size=2304, depth=40, num_heads=24, adaln_lora_dim=576)  # configs/fitv2/config_fitv2_3B.yaml:30-47


# --------------------------------------------------------------------------- #
# RoPE frequency rules (rope.py)
# --------------------------------------------------------------------------- #
def _find_correction_factor(num_rotations, dim, base=10000, max_position_embeddings=2048):
    # rope.py:24-25
    return (dim * math.log(max_position_embeddings / (num_rotations * 2 * math.pi))) / (2 * math.log(base))


def _find_correction_range(low_rot, high_rot, dim, base=10000, max_position_embeddings=2048):
    # rope.py:27-30
    low = math.floor(_find_correction_factor(low_rot, dim, base, max_position_embeddings))
    high = math.ceil(_find_correction_factor(high_rot, dim, base, max_position_embeddings))
    return max(low, 0), min(high, dim - 1)


def _linear_ramp_mask(lo, hi, dim):
    # rope.py:32-38
    if lo == hi:
        hi += 0.001
    linear_func = (torch.arange(dim, dtype=torch.float32) - lo) / (hi - lo)
    return torch.clamp(linear_func, 0, 1)


def rope_1d_freqs(custom_freqs: str, theta: float, dim: int, max_pe_len, ori_max_pe_len: int) -> torch.Tensor:
    """rope.py:173-231 (get_1d_rope_freqs).  ``dim`` = head_dim // 2."""
    assert isinstance(ori_max_pe_len, int)
    max_pe_len = torch.tensor(max_pe_len)
    scale = torch.clamp_min(max_pe_len / ori_max_pe_len, 1.0)
    ar = torch.arange(0, dim, 2).float() / dim
    if custom_freqs == "linear":
        return 1.0 / (scale * theta ** ar)
    if custom_freqs in ("ntk-aware", "ntk-aware-pro1", "ntk-aware-pro2"):
        newbase = theta * scale ** (dim / (dim - 2))          # rope.py:40-42
        return (1.0 / torch.pow(newbase.view(-1, 1), ar.to(scale))).squeeze()
    if custom_freqs == "ntk-by-parts":
        beta_0, beta_1, gamma_0, gamma_1 = 1.25, 0.75, 16, 2
        freqs_base = 1.0 / (theta ** ar)
        freqs_linear = 1.0 / (scale * theta ** ar)
        newbase = theta * scale ** (dim / (dim - 2))
        freqs_ntk = (1.0 / torch.pow(newbase.view(-1, 1), ar.to(scale))).squeeze()
        low, high = _find_correction_range(beta_0, beta_1, dim, theta, ori_max_pe_len)
        m = (1 - _linear_ramp_mask(low, high, dim // 2).to(scale)) * 1
        freqs = freqs_linear * (1 - m) + freqs_ntk * m
        low, high = _find_correction_range(gamma_0, gamma_1, dim, theta, ori_max_pe_len)
        m = (1 - _linear_ramp_mask(low, high, dim // 2).to(scale)) * 1
        return freqs * (1 - m) + freqs_base * m
    if custom_freqs == "yarn":
        beta_fast, beta_slow = 32, 1
        freqs_extrapolation = 1.0 / (theta ** ar)
        freqs_interpolation = 1.0 / (scale * theta ** ar)
        low, high = _find_correction_range(beta_fast, beta_slow, dim, theta, ori_max_pe_len)
        m = (1 - _linear_ramp_mask(low, high, dim // 2).to(scale).float()) * 1
        return freqs_interpolation * (1 - m) + freqs_extrapolation * m
    raise ValueError(f"Unknown custom_freqs {custom_freqs!r}")


def rope_setup(cfg: FiTConfig) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """rope.py:134-170.  Returns (freqs_h, freqs_w, magnitude) with freqs of
    shape (head_dim//4,) in fp32 and ``magnitude`` the cos/sin scale the cached
    lookup applies (rope.py:320-331)."""
    dim = cfg.head_dim // 2
    assert dim % 2 == 0
    cf = cfg.custom_freqs.lower()
    theta = cfg.rope_theta
    mag = 1.0
    if cf == "normal":
        fh = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
        fw = fh.clone()
    else:
        if cfg.decouple:
            fh = rope_1d_freqs(cf, theta, dim, cfg.max_pe_len_h, cfg.ori_max_pe_len)
            fw = rope_1d_freqs(cf, theta, dim, cfg.max_pe_len_w, cfg.ori_max_pe_len)
        else:
            m = max(cfg.max_pe_len_h, cfg.max_pe_len_w)
            fh = rope_1d_freqs(cf, theta, dim, m, cfg.ori_max_pe_len)
            fw = rope_1d_freqs(cf, theta, dim, m, cfg.ori_max_pe_len)
        lmax = max(cfg.max_pe_len_h, cfg.max_pe_len_w)
        scale = torch.clamp_min(torch.tensor(lmax) / cfg.ori_max_pe_len, 1.0)
        mscale = torch.where(scale <= 1.0, torch.tensor(1.0), 0.1 * torch.log(scale) + 1.0)  # rope.py:44-48

        def proportion(l_test, l_train):  # rope.py:50-53 (note the 2x quirk)
            l_test = l_test * 2
            return torch.where(torch.tensor(l_test / l_train) <= 1.0, torch.tensor(1.0),
                               torch.sqrt(torch.log(torch.tensor(l_test)) / torch.log(torch.tensor(l_train))))
        if cf == "yarn":
            mag = float(mscale)
        elif cf == "ntk-aware-pro1":
            mag = float(proportion(lmax, cfg.ori_max_pe_len))
        elif cf == "ntk-aware-pro2":
            mag = float(proportion(cfg.max_pe_len_h * cfg.max_pe_len_w, cfg.ori_max_pe_len ** 2))
    return fh.float(), fw.float(), mag


def rope_cos_sin(cfg: FiTConfig, grid: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """rope.py:308-333 (get_cached_2d_rope_from_grid).  grid (B,2,N) int64,
    grid[:,0] = w index, grid[:,1] = h index.  Returns cos, sin (B,N,head_dim)."""
    fh, fw, mag = rope_setup(cfg)
    pos = torch.arange(cfg.max_cached_len)
    # rope.py:165-170: pos * freqs in fp32 (int64 * fp32 -> fp32), each freq repeated twice
    th = (pos[:, None] * fh[None, :]).repeat_interleave(2, dim=-1)
    tw = (pos[:, None] * fw[None, :]).repeat_interleave(2, dim=-1)
    freqs = torch.cat([th[grid[:, 1]], tw[grid[:, 0]]], dim=-1)
    cos, sin = freqs.cos(), freqs.sin()
    if mag != 1.0:
        # the reference multiplies by a 0-dim fp32 tensor (rope.py:320-328)
        cos, sin = cos * torch.tensor(mag), sin * torch.tensor(mag)
    return cos, sin


def rope_1d_freqs_online(custom_freqs: str, theta: float, dim: int, sizes: torch.Tensor, ori_max_pe_len: int) -> torch.Tensor:
    """rope.py:173-231 (get_1d_rope_freqs) with a TENSOR ``max_pe_len`` of per-sample lengths (B,) -> (B, dim//2).
    Only the rules that work in the reference's online mode: 'normal' has no branch there (ValueError) and the
    magnitude rules (yarn, ntk-aware-pro1/2) read attributes the online constructor never sets (rope.py:143-160)."""
    assert isinstance(ori_max_pe_len, int)
    scale = torch.clamp_min(sizes / ori_max_pe_len, 1.0)                      # (B,) fp32
    ar = torch.arange(0, dim, 2).float() / dim
    if custom_freqs == "linear":
        return 1.0 / (scale[:, None] * (theta ** ar)[None, :])               # einsum('..., f -> ... f') outer product
    newbase = theta * scale ** (dim / (dim - 2))                              # rope.py:40-42
    freqs_ntk = (1.0 / torch.pow(newbase.view(-1, 1), ar.to(scale).float())).squeeze()
    if custom_freqs == "ntk-aware":
        return freqs_ntk.reshape(sizes.shape[0], -1)
    if custom_freqs == "ntk-by-parts":
        beta_0, beta_1, gamma_0, gamma_1 = 1.25, 0.75, 16, 2
        freqs_base = 1.0 / (theta ** ar)
        freqs_linear = 1.0 / (scale[:, None] * (theta ** ar.to(scale).float())[None, :])
        low, high = _find_correction_range(beta_0, beta_1, dim, theta, ori_max_pe_len)
        m = (1 - _linear_ramp_mask(low, high, dim // 2).to(scale)) * 1
        freqs = freqs_linear * (1 - m) + freqs_ntk * m
        low, high = _find_correction_range(gamma_0, gamma_1, dim, theta, ori_max_pe_len)
        m = (1 - _linear_ramp_mask(low, high, dim // 2).to(scale)) * 1
        return (freqs * (1 - m) + freqs_base * m).reshape(sizes.shape[0], -1)
    raise ValueError(f"Unknown modality {custom_freqs}. online_rope supports linear, ntk-aware, ntk-by-parts")


def rope_cos_sin_online(cfg: FiTConfig, grid: torch.Tensor, size: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """rope.py:234-274 (online_get_2d_rope_from_grid).  size (B,1,2) int64, h first, w last."""
    dim = cfg.head_dim // 2
    cf = cfg.custom_freqs.lower()
    size = size.squeeze()
    if size.dim() == 1:
        size = size[None]
    if cfg.decouple:
        fh = rope_1d_freqs_online(cf, cfg.rope_theta, dim, size[:, 0], cfg.ori_max_pe_len)
        fw = rope_1d_freqs_online(cf, cfg.rope_theta, dim, size[:, 1], cfg.ori_max_pe_len)
    else:
        smax = torch.max(size[:, 0], size[:, 1])
        fh = rope_1d_freqs_online(cf, cfg.rope_theta, dim, smax, cfg.ori_max_pe_len)
        fw = fh
    tw = (grid[:, 0][..., None] * fw[:, None, :]).repeat_interleave(2, dim=-1)
    th = (grid[:, 1][..., None] * fh[:, None, :]).repeat_interleave(2, dim=-1)
    freqs = torch.cat([th, tw], dim=-1)
    return freqs.cos(), freqs.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """rope.py:107-111: pairs (x0,x1) -> (-x1,x0)."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


# --------------------------------------------------------------------------- #
# weights: reference init order + synthetic non-zero re-draw
# --------------------------------------------------------------------------- #
class _SwiGLU(nn.Module):
    """Creation order of timm.layers.mlp.SwiGLU: fc1_g, fc1_x, fc2."""
    def __init__(self, d, h, bias=True, out=None):
        super().__init__()
        self.fc1_g = nn.Linear(d, h, bias=bias)
        self.fc1_x = nn.Linear(d, h, bias=bias)
        self.fc2 = nn.Linear(h, out or d, bias=bias)


class _Mlp(nn.Module):
    """Creation order of timm.layers.mlp.Mlp: fc1, fc2 (act / drop / norm hold no parameters)."""
    def __init__(self, d, h, bias=True):
        super().__init__()
        self.fc1 = nn.Linear(d, h, bias=bias)
        self.fc2 = nn.Linear(h, d, bias=bias)


class _Norm(nn.Module):
    """create_norm (norms.py:35-50): a ``weight`` of ones for 'w_layernorm' / 'rmsnorm', nothing otherwise."""
    def __init__(self, kind: Optional[str], dim: int):
        super().__init__()
        if kind in ("w_layernorm", "rmsnorm", "w_rmsnorm"):
            self.weight = nn.Parameter(torch.ones(dim))


class _Attn(nn.Module):
    def __init__(self, cfg: "FiTConfig"):
        super().__init__()
        d = cfg.hidden_size
        self.qkv = nn.Linear(d, 3 * d, bias=cfg.qkv_bias)      # modules.py:140
        self.q_norm = _Norm(cfg.norm_kind("q"), cfg.head_dim)   # modules.py:146-147
        self.k_norm = _Norm(cfg.norm_kind("k"), cfg.head_dim)
        self.proj = nn.Linear(d, d)         # modules.py:151


class _Block(nn.Module):
    def __init__(self, cfg: FiTConfig):
        super().__init__()
        d = cfg.hidden_size
        self.norm1 = _Norm(cfg.norm_kind("block"), d)           # modules.py:239-240
        self.norm2 = _Norm(cfg.norm_kind("block"), d)
        self.attn = _Attn(cfg)                                  # modules.py:242-247
        if cfg.use_swiglu:
            self.mlp = _SwiGLU(d, cfg.mlp_hidden, cfg.ffn_bias)    # modules.py:247-251
        else:
            self.mlp = _Mlp(d, cfg.mlp_hidden, cfg.ffn_bias)       # modules.py:253
        if cfg.adaln_type == "swiglu":                          # modules.py:265-268
            self.adaLN_modulation = _SwiGLU(d, cfg.adaln_hidden[0], True, 6 * d)
        elif cfg.adaln_type == "lora":
            self.adaLN_modulation = nn.Sequential(              # modules.py:259-264
                nn.SiLU(), nn.Linear(d, cfg.adaln_lora_dim), nn.Linear(cfg.adaln_lora_dim, 6 * d))
        else:
            self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(d, 6 * d))   # modules.py:254-258


class _Final(nn.Module):
    def __init__(self, cfg: FiTConfig):
        super().__init__()
        d = cfg.hidden_size
        self.norm_final = _Norm(cfg.norm_kind("block"), d)      # modules.py:282
        self.linear = nn.Linear(d, cfg.out_token_channels)      # modules.py:283
        if cfg.adaln_type == "swiglu":                          # modules.py:284-285
            self.adaLN_modulation = _SwiGLU(d, cfg.adaln_hidden[1], True, 2 * d)
        else:
            self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(d, 2 * d))  # modules.py:287-290


class _Skeleton(nn.Module):
    """Parameter skeleton that creates sub-modules in the reference's order
    (fit_model.py:84-112) so that RNG consumption under a fixed seed, the
    ``apply`` traversal order and ``named_parameters()`` order all coincide."""
    def __init__(self, cfg: FiTConfig):
        super().__init__()
        d = cfg.hidden_size

        class _X(nn.Module):
            def __init__(s):
                super().__init__()
                s.proj = nn.Linear(cfg.token_channels, d)
        class _T(nn.Module):
            def __init__(s):
                super().__init__()
                s.mlp = nn.Sequential(nn.Linear(256, d), nn.SiLU(), nn.Linear(d, d))
        class _Y(nn.Module):
            def __init__(s):
                super().__init__()
                s.embedding_table = nn.Embedding(cfg.num_classes + (cfg.class_dropout_prob > 0), d)
        self.x_embedder = _X()
        self.t_embedder = _T()
        self.y_embedder = _Y()
        if cfg.adaln_type == "lora":                                        # fit_model.py:97-103
            self.global_adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(d, 6 * d))
        else:
            self.global_adaLN_modulation = None
        self.blocks = nn.ModuleList([_Block(cfg) for _ in range(cfg.depth)])
        self.final_layer = _Final(cfg)


def reference_init_state_dict(cfg: FiTConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """fit_model.py:117-157 under ``torch.manual_seed(seed)``: xavier-uniform
    Linears with zero bias, N(0,0.02) label table and t-MLP weights, zeroed
    adaLN output layers / global adaLN / final linear."""
    torch.manual_seed(seed)
    m = _Skeleton(cfg)

    def _basic_init(mod):
        if isinstance(mod, nn.Linear):
            nn.init.xavier_uniform_(mod.weight)
            if mod.bias is not None:
                nn.init.constant_(mod.bias, 0)
    m.apply(_basic_init)
    w = m.x_embedder.proj.weight.data
    nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
    nn.init.constant_(m.x_embedder.proj.bias, 0)
    nn.init.normal_(m.y_embedder.embedding_table.weight, std=0.02)
    nn.init.normal_(m.t_embedder.mlp[0].weight, std=0.02)
    nn.init.normal_(m.t_embedder.mlp[2].weight, std=0.02)
    last = (lambda mod: mod.fc2) if cfg.adaln_type == "swiglu" else (lambda mod: mod[-1])   # fit_model.py:139-153
    for blk in m.blocks:
        nn.init.constant_(last(blk.adaLN_modulation).weight, 0)
        nn.init.constant_(last(blk.adaLN_modulation).bias, 0)
    if m.global_adaLN_modulation is not None:
        nn.init.constant_(m.global_adaLN_modulation[-1].weight, 0)
        nn.init.constant_(m.global_adaLN_modulation[-1].bias, 0)
    nn.init.constant_(last(m.final_layer.adaLN_modulation).weight, 0)
    nn.init.constant_(last(m.final_layer.adaLN_modulation).bias, 0)
    nn.init.constant_(m.final_layer.linear.weight, 0)
    nn.init.constant_(m.final_layer.linear.bias, 0)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def redraw_zero_params(sd: Dict[str, torch.Tensor], seed: int = 1, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """the reference init makes the model output identically
    zero; every all-zero tensor is re-drawn N(0, std^2) from a dedicated CPU
    generator, in state_dict (== named_parameters) order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if v.is_floating_point() and not bool(v.any()):
            out[k] = torch.randn(v.shape, generator=g, dtype=torch.float32) * std
        else:
            out[k] = v
    return out


def synthetic_state_dict(cfg: FiTConfig, init_seed: int = 0, redraw_seed: int = 1, std: float = 0.02):
    return redraw_zero_params(reference_init_state_dict(cfg, init_seed), redraw_seed, std)


# --------------------------------------------------------------------------- #
# forward math
# --------------------------------------------------------------------------- #
def _q(x: torch.Tensor, quant: Optional[str]) -> torch.Tensor:
    """Operand rounding used ONLY to forecast the 16-bit GEMM-operand error of
    the CUDA path on CPU (quant=None is the oracle proper)."""
    if quant is None:
        return x
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[quant]
    return x.to(dt).to(torch.float32)


def _linear(x, sd, name, quant=None):
    return F.linear(_q(x, quant), _q(sd[name + ".weight"], quant), sd.get(name + ".bias"))   # bias-free linears: qkv_bias / ffn_bias False


def _swiglu_mlp(x, sd, p, quant=None):
    """timm SwiGLU: fc2(silu(fc1_g(x)) * fc1_x(x))."""
    return _linear(F.silu(_linear(x, sd, p + ".fc1_g", quant)) * _linear(x, sd, p + ".fc1_x", quant), sd, p + ".fc2", quant)


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: int = 10000) -> torch.Tensor:
    """modules.py:52-71 (cos first, then sin; freqs built in fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = t[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(dtype=t.dtype)


def conditioning(cfg: FiTConfig, sd, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """fit_model.py:202-209 -> c (B, D)."""
    ts = cfg.time_shifting
    t = torch.clamp(ts * t / (1 + (ts - 1) * t), max=1.0).float()
    te = timestep_embedding(t)
    te = _linear(F.silu(_linear(te, sd, "t_embedder.mlp.0")), sd, "t_embedder.mlp.2")
    return te + sd["y_embedder.embedding_table.weight"][y]


def block_modulation(cfg: FiTConfig, sd, c: torch.Tensor, i: int, global_adaln) -> torch.Tensor:
    """modules.py:254-264,271 -> (B, 6D)."""
    s = F.silu(c)
    p = f"blocks.{i}.adaLN_modulation"
    if cfg.adaln_type == "swiglu":                                       # modules.py:265-268: SwiGLU on c itself (no SiLU in front)
        return _swiglu_mlp(c, sd, p) + global_adaln
    if cfg.adaln_type == "normal":
        return _linear(s, sd, p + ".1") + global_adaln
    return _linear(_linear(s, sd, p + ".1"), sd, p + ".2") + global_adaln


def layer_norm(x: torch.Tensor) -> torch.Tensor:
    """norms.py:41-42: nn.LayerNorm(dim, eps=1e-6, elementwise_affine=False)."""
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def apply_norm(kind: Optional[str], x: torch.Tensor, sd, name: str) -> torch.Tensor:
    """create_norm(kind, dim)(x) (norms.py:35-50, 53-77); ``name`` is the module path that owns ``weight``."""
    if kind is None:
        return x
    if kind == "layernorm":
        return layer_norm(x)
    if kind == "w_layernorm":                                            # nn.LayerNorm(dim, eps, bias=False)
        return F.layer_norm(x, (x.shape[-1],), weight=sd[name + ".weight"], eps=1e-6)
    if kind in ("rmsnorm", "w_rmsnorm"):                                 # norms.py:72-77
        xf = x.float()
        out = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).type_as(x)
        return out * sd[name + ".weight"]
    raise NotImplementedError(kind)


def modulate(x, shift, scale):
    """fit/model/utils.py:6-7."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def attention(cfg: FiTConfig, sd, i: int, x, mask, cos, sin, quant=None, taps=None):
    """modules.py:159-207."""
    B, N, C = x.shape
    H, dh = cfg.num_heads, cfg.head_dim
    qkv = _linear(x, sd, f"blocks.{i}.attn.qkv", quant).reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    q = apply_norm(cfg.norm_kind("q"), q, sd, f"blocks.{i}.attn.q_norm")  # :168
    k = apply_norm(cfg.norm_kind("k"), k, sd, f"blocks.{i}.attn.k_norm")
    if cfg.rotates:                                                      # :170
        if cfg.add_rel_pe_to_v:
            v = v * cos + rotate_half(v) * sin                           # :171-172
        q = q * cos + rotate_half(q) * sin                               # :173
        k = k * cos + rotate_half(k) * sin                               # :174
    if taps is not None:
        taps["q"], taps["k"], taps["v"] = q.clone(), k.clone(), v.clone()
    am = mask[:, None, None, :]
    am = (am == am.transpose(-2, -1))                                    # :176-177 segment-id equality
    keep = torch.not_equal(mask, torch.zeros_like(mask)).to(mask)        # :178
    o = F.scaled_dot_product_attention(_q(q, quant), _q(k, quant), _q(v, quant), attn_mask=am)   # :181-184
    o = o.transpose(1, 2).reshape(B, N, C)
    o = o * keep[..., None]                                              # :204
    if taps is not None:
        taps["attn_out"] = o.clone()
    return _linear(o, sd, f"blocks.{i}.attn.proj", quant)                # :205


def swiglu(cfg: FiTConfig, sd, i: int, x, quant=None):
    """timm SwiGLU: fc2(silu(fc1_g(x)) * fc1_x(x)); with use_swiglu=False timm Mlp: fc2(GELU_tanh(fc1(x))) (modules.py:253)."""
    p = f"blocks.{i}.mlp"
    if not cfg.use_swiglu:
        return _linear(F.gelu(_linear(x, sd, p + ".fc1", quant), approximate="tanh"), sd, p + ".fc2", quant)
    g = _linear(x, sd, p + ".fc1_g", quant)
    u = _linear(x, sd, p + ".fc1_x", quant)
    return _linear(F.silu(g) * u, sd, p + ".fc2", quant)


def forward(cfg: FiTConfig, sd, x, t, y, grid, mask, size=None, quant: Optional[str] = None, taps=None):
    """fit_model.py:189-233 with use_sit=True, adaln_type='lora' (online_rope per cfg).

    x (B,N,p*p*C) float, t (B,) float, y (B,) int64, grid (B,2,N) int64,
    mask (B,N) float/bool segment ids.  Returns (B,N,p*p*C)."""
    x = x.float()
    maskf = mask.to(x.dtype) if mask.dtype != torch.bool else mask
    c = conditioning(cfg, sd, t.to(x.dtype), y)
    if not cfg.use_sit:
        x = x.transpose(1, 2)                                            # :204 'B C N -> B N C'
    bn = cfg.norm_kind("block")
    h = _linear(x, sd, "x_embedder.proj")                                # :206
    if cfg.online_rope:
        cos, sin = rope_cos_sin_online(cfg, grid, size)                  # :212-214
    else:
        cos, sin = rope_cos_sin(cfg, grid)                               # :216
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    g_adaln = _linear(F.silu(c), sd, "global_adaLN_modulation.1") if cfg.adaln_type == "lora" else 0.0   # :218-221
    if taps is not None:
        taps["c"], taps["x0"] = c.clone(), h.clone()
        if cfg.adaln_type == "lora":
            taps["global_adaln"] = g_adaln.clone()
    for i in range(cfg.depth):
        sh1, sc1, g1, sh2, sc2, g2 = block_modulation(cfg, sd, c, i, g_adaln).chunk(6, dim=1)
        a = attention(cfg, sd, i, modulate(apply_norm(bn, h, sd, f"blocks.{i}.norm1"), sh1, sc1), maskf, cos, sin, quant,
                      taps if (taps is not None and i == 0) else None)
        h = h + g1.unsqueeze(1) * a                                      # modules.py:272
        m = swiglu(cfg, sd, i, modulate(apply_norm(bn, h, sd, f"blocks.{i}.norm2"), sh2, sc2), quant)
        h = h + g2.unsqueeze(1) * m                                      # modules.py:273
        if taps is not None and i == 0:
            taps["x1"] = h.clone()
    if cfg.adaln_type == "swiglu":
        shift, scale = _swiglu_mlp(c, sd, "final_layer.adaLN_modulation").chunk(2, dim=1)
    else:
        shift, scale = _linear(F.silu(c), sd, "final_layer.adaLN_modulation.1").chunk(2, dim=1)
    out = _linear(modulate(apply_norm(bn, h, sd, "final_layer.norm_final"), shift, scale), sd, "final_layer.linear")  # modules.py:292-296
    out = out * maskf[..., None]                                         # fit_model.py:230
    return out if cfg.use_sit else out.transpose(1, 2)                   # :231-232 'B N C -> B C N'



def forward_with_cfg(cfg: FiTConfig, sd, x, t, y, grid, mask, size, cfg_scale, scale_pow=0.0, quant=None):
    """fit_model.py:235-275 (use_sit=True): CFG on the first 3*p*p channels only."""
    half = x[: len(x) // 2]
    combined = torch.cat([half, half], dim=0)
    out = forward(cfg, sd, combined, t, y, grid, mask, size, quant)
    c_cfg = 3 * cfg.patch_size * cfg.patch_size
    if cfg.use_sit:
        eps, rest = out[:, :, :c_cfg], out[:, :, c_cfg:]
    else:
        eps, rest = out[:, :c_cfg], out[:, c_cfg:]                       # :256-257
    cond, uncond = torch.split(eps, len(eps) // 2, dim=0)
    if scale_pow == 0.0:
        real = cfg_scale
    else:
        step = (1 - torch.cos(((1 - torch.clamp_max(t, 1.0)) ** scale_pow) * torch.pi)) * 1 / 2
        real = ((cfg_scale - 1) * step + 1)[: len(x) // 2].view(-1, 1, 1)
    half_eps = uncond + real * (cond - uncond)
    eps = torch.cat([half_eps, half_eps], dim=0)
    return torch.cat([eps, rest], dim=2 if cfg.use_sit else 1)


def unpatchify(cfg: FiTConfig, x: torch.Tensor, hw) -> torch.Tensor:
    """fit_model.py:171-187 (use_sit=True): (B,(h w),(c p1 p2)) -> (B,c,h*p1,w*p2)."""
    h, w = hw
    p = cfg.patch_size
    B = x.shape[0]
    x = x.reshape(B, h // p, w // p, -1, p, p)          # b h w c p1 p2
    return x.permute(0, 3, 1, 4, 2, 5).reshape(B, -1, h, w)


# --------------------------------------------------------------------------- #
# sampler loop (sample_fitv2_ddp.py:257-314)
# --------------------------------------------------------------------------- #
def make_grid(n: int, n_patch_h: int, n_patch_w: int) -> torch.Tensor:
    """sample_fitv2_ddp.py:263-268: grid[:,0] = w index, grid[:,1] = h index."""
    gh = torch.arange(n_patch_h, dtype=torch.long)
    gw = torch.arange(n_patch_w, dtype=torch.long)
    g = torch.meshgrid(gw, gh, indexing="xy")
    return torch.cat([g[0].reshape(1, -1), g[1].reshape(1, -1)], dim=0).repeat(n, 1, 1)


def cfg_euler_update(z, v2, cfg_scale: float, sigma_cur, sigma_next):
    """sample_fitv2_ddp.py:310-314: chunk, uncond + s*(cond-uncond), z + dsigma*v
    evaluated in exactly that order (no FMA contraction)."""
    cond, uncond = v2.chunk(2, dim=0)
    v = uncond + cfg_scale * (cond - uncond)
    return z + (sigma_next - sigma_cur) * v


def euler_cfg_sample(cfg: FiTConfig, sd, z, y, grid, mask, size, steps: int, cfg_scale: float,
                     quant=None, first_steps: Optional[int] = None):
    """sample_fitv2_ddp.py:273-314 with CFG enabled.  z (n,N,16); y (n,)."""
    n = z.shape[0]
    y2 = torch.cat([y, torch.full((n,), cfg.num_classes, dtype=y.dtype)], 0)
    grid2, mask2 = torch.cat([grid, grid], 0), torch.cat([mask, mask], 0)
    size2 = None if size is None else torch.cat([size, size], 0)
    sigmas = torch.linspace(0, 1, steps + 1)
    todo = steps if first_steps is None else first_steps
    for idx in range(todo):
        z_in = torch.cat([z, z], 0)
        ts = sigmas[idx].expand(z_in.shape[0])
        v2 = forward(cfg, sd, z_in, ts, y2, grid2, mask2, size2, quant)
        z = cfg_euler_update(z, v2, cfg_scale, sigmas[idx], sigmas[idx + 1])
    return z


def flops_per_forward_row(cfg: FiTConfig, n_tokens: int) -> float:
    """SURVEY.md §8(d) formula: 2*MAC, attention included, per sample row."""
    D, L, Hm, N, lora = cfg.hidden_size, cfg.depth, cfg.mlp_hidden, n_tokens, cfg.adaln_lora_dim
    mac = N * (L * (4 * D * D + 3 * D * Hm + 2 * N * D) + 32 * D) \
        + L * (D * lora + 6 * D * lora) + 256 * D + D * D + 6 * D * D + 2 * D * D
    return 2.0 * mac


def perturb_norm_weights(sd: Dict[str, torch.Tensor], seed: int = 2, std: float = 0.2) -> Dict[str, torch.Tensor]:
    """Synthetic weights for the weighted norms (they are initialised to ones, which would hide a missing multiply):
    every ``*norm*.weight`` vector becomes 1 + std * N(0,1) from a dedicated CPU generator, in state_dict order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight") and v.dim() == 1 and "norm" in k.rsplit(".", 2)[-2]:
            out[k] = 1.0 + std * torch.randn(v.shape, generator=g, dtype=torch.float32)
        else:
            out[k] = v
    return out
