"""Host-side 2-D RoPE frequency rules of the FiTv2 drop-in (construct-time, fp32, CPU).

Mirrors ``fit/model/rope.py::VisionRotaryEmbedding.__init__`` / ``get_1d_rope_freqs`` (rope.py:119-231):
the per-axis inverse frequencies and the cos/sin magnitude are computed once here in fp32 with the
same operation order as the reference; the per-token ``pos * freq`` angles and cos/sin tables are
built on the GPU by ``rope_table_kernel`` (csrc/pointwise.cuh).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

_RULES = ("normal", "linear", "ntk-aware", "ntk-aware-pro1", "ntk-aware-pro2", "ntk-by-parts", "yarn")


def _correction_range(low_rot, high_rot, dim, base, max_pos):
    # rope.py:24-30
    def factor(n_rot):
        return (dim * math.log(max_pos / (n_rot * 2 * math.pi))) / (2 * math.log(base))
    return max(math.floor(factor(low_rot)), 0), min(math.ceil(factor(high_rot)), dim - 1)


def _ramp(lo, hi, dim):
    # rope.py:32-38
    if lo == hi:
        hi += 0.001
    return torch.clamp((torch.arange(dim, dtype=torch.float32) - lo) / (hi - lo), 0, 1)


def axis_freqs(rule: str, theta: float, dim: int, max_pe_len: int, ori_max_pe_len: int) -> torch.Tensor:
    """rope.py:173-231 — inverse frequencies (dim // 2,) for one axis."""
    if not isinstance(ori_max_pe_len, int):
        raise TypeError("ori_max_pe_len must be an int (rope.py:175)")
    scale = torch.clamp_min(torch.tensor(max_pe_len) / ori_max_pe_len, 1.0)
    expo = torch.arange(0, dim, 2).float() / dim
    base_pow = theta ** expo

    def ntk():
        newbase = theta * scale ** (dim / (dim - 2))                 # rope.py:40-42
        return (1.0 / torch.pow(newbase.view(-1, 1), expo.to(scale))).squeeze()

    if rule == "linear":
        return 1.0 / (scale * base_pow)
    if rule in ("ntk-aware", "ntk-aware-pro1", "ntk-aware-pro2"):
        return ntk()
    if rule == "ntk-by-parts":
        f_base, f_lin, f_ntk = 1.0 / base_pow, 1.0 / (scale * base_pow), ntk()
        lo, hi = _correction_range(1.25, 0.75, dim, theta, ori_max_pe_len)
        m = (1 - _ramp(lo, hi, dim // 2).to(scale)) * 1
        f = f_lin * (1 - m) + f_ntk * m
        lo, hi = _correction_range(16, 2, dim, theta, ori_max_pe_len)
        m = (1 - _ramp(lo, hi, dim // 2).to(scale)) * 1
        return f * (1 - m) + f_base * m
    if rule == "yarn":
        f_ext, f_int = 1.0 / base_pow, 1.0 / (scale * base_pow)
        lo, hi = _correction_range(32, 1, dim, theta, ori_max_pe_len)
        m = (1 - _ramp(lo, hi, dim // 2).to(scale).float()) * 1
        return f_int * (1 - m) + f_ext * m
    raise ValueError(f"Unknown custom_freqs {rule!r}; supported: {_RULES}")


def rope_frequencies(head_dim: int, custom_freqs: str = "normal", theta: float = 10000.0,
                     max_pe_len_h: Optional[int] = None, max_pe_len_w: Optional[int] = None,
                     decouple: bool = False, ori_max_pe_len: Optional[int] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """rope.py:134-160 — returns (freqs_h, freqs_w, magnitude); magnitude is the factor the cached
    lookup multiplies cos/sin with (rope.py:320-331): yarn -> mscale, ntk-aware-pro1/2 -> proportion1/2."""
    dim = head_dim // 2
    if dim % 2:
        raise ValueError("head_dim // 2 must be even (rope.py:137)")
    rule = custom_freqs.lower()
    if rule not in _RULES:
        raise ValueError(f"Unknown custom_freqs {custom_freqs!r}; supported: {_RULES}")
    if rule == "normal":
        f = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
        return f.float(), f.clone().float(), 1.0
    if max_pe_len_h is None or max_pe_len_w is None or ori_max_pe_len is None:
        raise ValueError("custom_freqs != 'normal' needs max_pe_len_h, max_pe_len_w and ori_max_pe_len")
    if decouple:
        fh = axis_freqs(rule, theta, dim, max_pe_len_h, ori_max_pe_len)
        fw = axis_freqs(rule, theta, dim, max_pe_len_w, ori_max_pe_len)
    else:
        longest = max(max_pe_len_h, max_pe_len_w)
        fh = axis_freqs(rule, theta, dim, longest, ori_max_pe_len)
        fw = axis_freqs(rule, theta, dim, longest, ori_max_pe_len)
    longest = max(max_pe_len_h, max_pe_len_w)
    scale = torch.clamp_min(torch.tensor(longest) / ori_max_pe_len, 1.0)
    mag = 1.0
    if rule == "yarn":                                                   # rope.py:44-48
        mag = float(torch.where(scale <= 1.0, torch.tensor(1.0), 0.1 * torch.log(scale) + 1.0))
    elif rule in ("ntk-aware-pro1", "ntk-aware-pro2"):                   # rope.py:50-53,158-160
        if rule == "ntk-aware-pro1":
            l_test, l_train = longest * 2, ori_max_pe_len
        else:
            l_test, l_train = max_pe_len_h * max_pe_len_w * 2, ori_max_pe_len ** 2
        mag = float(torch.where(torch.tensor(l_test / l_train) <= 1.0, torch.tensor(1.0),
                                torch.sqrt(torch.log(torch.tensor(l_test)) / torch.log(torch.tensor(l_train)))))
    return fh.float(), fw.float(), mag


_ONLINE_RULES = ("linear", "ntk-aware", "ntk-by-parts")


def online_axis_freqs(rule: str, theta: float, dim: int, sizes: torch.Tensor, ori_max_pe_len: int) -> torch.Tensor:
    """rope.py:173-231 with a tensor ``max_pe_len``: per-sample axis lengths (B,) -> inverse frequencies (B, dim // 2),
    fp32 on the CPU in the reference's operation order (the rules its online mode supports, rope.py:234-274)."""
    if not isinstance(ori_max_pe_len, int):
        raise TypeError("ori_max_pe_len must be an int (rope.py:175)")
    if rule not in _ONLINE_RULES:
        raise ValueError(f"Unknown modality {rule}. online_rope supports {_ONLINE_RULES}")
    scale = torch.clamp_min(sizes / ori_max_pe_len, 1.0)
    expo = torch.arange(0, dim, 2).float() / dim
    base_pow = theta ** expo
    f_lin = 1.0 / (scale[:, None] * base_pow[None, :])
    if rule == "linear":
        return f_lin
    newbase = theta * scale ** (dim / (dim - 2))
    f_ntk = (1.0 / torch.pow(newbase.view(-1, 1), expo.to(scale).float())).reshape(sizes.shape[0], -1)
    if rule == "ntk-aware":
        return f_ntk
    f_base = 1.0 / base_pow
    lo, hi = _correction_range(1.25, 0.75, dim, theta, ori_max_pe_len)
    m = (1 - _ramp(lo, hi, dim // 2).to(scale)) * 1
    f = f_lin * (1 - m) + f_ntk * m
    lo, hi = _correction_range(16, 2, dim, theta, ori_max_pe_len)
    m = (1 - _ramp(lo, hi, dim // 2).to(scale)) * 1
    return f * (1 - m) + f_base * m


def online_rope_frequencies(head_dim: int, custom_freqs: str, theta: float, decouple: bool, ori_max_pe_len: int,
                            size: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """rope.py:243-253: per-sample (freqs_h, freqs_w), each (B, head_dim // 4) fp32 CPU, from size (B,1,2) = (h, w)."""
    dim = head_dim // 2
    size = size.detach().to("cpu", torch.int64).reshape(-1, 2)
    rule = custom_freqs.lower()
    if decouple:
        return (online_axis_freqs(rule, theta, dim, size[:, 0], ori_max_pe_len).float().contiguous(),
                online_axis_freqs(rule, theta, dim, size[:, 1], ori_max_pe_len).float().contiguous())
    f = online_axis_freqs(rule, theta, dim, torch.max(size[:, 0], size[:, 1]), ori_max_pe_len).float().contiguous()
    return f, f.clone()
