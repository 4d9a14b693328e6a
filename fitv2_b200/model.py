"""``fitv2_b200.FiT`` — drop-in for ``fit.model.fit_model.FiT`` on the sampling hot path.

Same constructor keywords (fit_model.py:25-65), same ``state_dict`` keys and shapes (so
``init_from_ckpt`` / ``load_state_dict`` of a reference checkpoint work, eval_utils.py:12-71), same
``forward`` / ``forward_with_cfg`` / ``unpatchify`` signatures (fit_model.py:171-275).  The math runs
in hand-written sm_100a CUDA behind the C ABI of ``include/fitv2_b200.h``; PyTorch only owns the
device memory and the stream.  There is no CPU path and no PyTorch fallback: unsupported
configurations and non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .rope import rope_frequencies, online_rope_frequencies, _ONLINE_RULES


class _Holder(nn.Module):
    """Parameter container (never called)."""


class _NormWeight(nn.Module):
    """``norm.weight`` of a weighted LayerNorm / RMSNorm (fit/model/norms.py:35-50): ones, never called."""

    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


# ------------------------------------------------------------------------------------------------
# torch.library registration: the forward is a registered operator (with a fake / meta kernel), so torch.jit.trace,
# fvcore's FlopCountAnalysis (sample_fitv2_ddp.py:197-213) and torch.export see a node with a data dependence from x to the
# output instead of an opaque ctypes call writing into torch.empty.  The module instance is passed by registry id.
# ------------------------------------------------------------------------------------------------
_MODELS: "weakref.WeakValueDictionary[int, FiT]" = weakref.WeakValueDictionary()


@torch.library.custom_op("fitv2_b200::forward", mutates_args=(), device_types="cuda")
def _forward_op(x: torch.Tensor, t: torch.Tensor, y: torch.Tensor, grid: torch.Tensor, mask: torch.Tensor, model_id: int,
                rows: int) -> torch.Tensor:
    model = _MODELS.get(model_id)
    if model is None:
        raise _lib.FitV2Error(f"fitv2_b200::forward: model {model_id} is gone")
    return model._run(x, t, y, grid, mask, rows)


@_forward_op.register_fake
def _(x, t, y, grid, mask, model_id, rows):
    model = _MODELS.get(model_id)
    co = model.out_token_channels if model is not None else x.shape[-1]
    if model is not None and not model.use_sit:
        return x.new_empty((rows, co, x.shape[2]), dtype=torch.float32)
    return x.new_empty((rows, x.shape[1], co), dtype=torch.float32)


def _seq(*mods):
    return nn.Sequential(*mods)


def _swiglu_holder(d: int, hidden: int, out: int) -> nn.Module:
    """Parameters of a timm SwiGLU in its creation order (fc1_g, fc1_x, fc2)."""
    m = _Holder()
    m.fc1_g, m.fc1_x, m.fc2 = nn.Linear(d, hidden), nn.Linear(d, hidden), nn.Linear(hidden, out)
    return m


class FiT(nn.Module):
    """FiT / FiTv2 transformer with the constructor contract of fit_model.py:25-65: the FiTv2 family (use_sit, SwiGLU, adaLN-LoRA,
    layernorm q/k norm, 2-D RoPE) on the fast kernels, every other switch the reference class can be built with (FiTv1 layout,
    GELU Mlp, adaLN 'normal' / 'swiglu', weighted / RMS / absent norms, bias-free qkv / ffn, no rotation, rotation of v) on their
    generic variants."""

    def __init__(self, context_size: int = 256, patch_size: int = 2, in_channels: int = 4, hidden_size: int = 1152,
                 depth: int = 28, num_heads: int = 16, mlp_ratio: float = 4.0, class_dropout_prob: float = 0.1,
                 num_classes: int = 1000, learn_sigma: bool = True, use_sit: bool = False, use_checkpoint: bool = False,
                 use_swiglu: bool = False, use_swiglu_large: bool = False, rel_pos_embed: Optional[str] = "rope",
                 norm_type: str = "layernorm", q_norm: Optional[str] = None, k_norm: Optional[str] = None,
                 qk_norm_weight: bool = False, qkv_bias: bool = True, ffn_bias: bool = True, adaln_bias: bool = True,
                 adaln_type: str = "normal", adaln_lora_dim: int = None, rope_theta: float = 10000.0,
                 custom_freqs: str = "normal", max_pe_len_h: Optional[int] = None, max_pe_len_w: Optional[int] = None,
                 decouple: bool = False, ori_max_pe_len: Optional[int] = None, online_rope: bool = False,
                 add_rel_pe_to_v: bool = False, pretrain_ckpt: str = None, ignore_keys: list = None,
                 finetune: str = None, time_shifting: int = 1, save_attention: bool = False,
                 operand_dtype: str = "bf16", **kwargs):
        super().__init__()
        # ---- reject what the kernels do not implement (no silent fallback) ----
        unsupported = []
        assert not (learn_sigma and use_sit)                                # fit_model.py:68
        if adaln_type not in ("lora", "normal", "swiglu"): unsupported.append("adaln_type must be 'lora', 'normal' or 'swiglu'")
        if adaln_type == "lora" and not adaln_lora_dim: unsupported.append("adaln_type='lora' needs adaln_lora_dim")
        try:
            self.block_norm = _lib.norm_code(norm_type)
            self.q_norm_code = _lib.norm_code(q_norm, qk_norm_weight)
            self.k_norm_code = _lib.norm_code(k_norm, qk_norm_weight)
        except NotImplementedError as e:
            unsupported.append(str(e))
            self.block_norm = self.q_norm_code = self.k_norm_code = _lib.NORM_LAYERNORM
        if self.block_norm == _lib.NORM_NONE: unsupported.append("norm_type must be 'layernorm', 'w_layernorm' or 'rmsnorm'")
        # adaln_bias=False cannot be constructed in the reference either: initialize_weights calls nn.init.constant_ on the
        # missing bias (fit_model.py:141-153), so there is no behaviour to reproduce
        if not adaln_bias: unsupported.append("adaln_bias must be True (the reference constructor itself fails without it)")
        if online_rope and (custom_freqs.lower() not in _ONLINE_RULES or not isinstance(ori_max_pe_len, int)):
            unsupported.append("online_rope=True needs custom_freqs in ('linear', 'ntk-aware', 'ntk-by-parts') and ori_max_pe_len "
                               "(the reference's online mode has no 'normal' branch and never sets the yarn / ntk-aware-pro magnitudes)")
        # use_checkpoint (activation checkpointing, fit_model.py:222-226) changes nothing in a no-grad forward: accepted, unused
        # save_attention: the reference's own constructor raises on it (Attention.__init__ has no such keyword), and
        # the attention kernels never materialise the attention matrix
        if save_attention: unsupported.append("save_attention=True")
        if patch_size ** 2 * in_channels != 16: unsupported.append("patch_size**2 * in_channels must be 16")
        if hidden_size % num_heads or hidden_size // num_heads not in (72, 96): unsupported.append("head_dim must be 72 or 96")
        if operand_dtype not in ("bf16", "fp16"): unsupported.append("operand_dtype must be 'bf16' or 'fp16'")
        if unsupported:
            raise NotImplementedError("fitv2_b200.FiT does not implement: " + "; ".join(unsupported))

        self.context_size, self.hidden_size, self.depth = context_size, hidden_size, depth
        self.learn_sigma, self.use_sit, self.use_checkpoint = learn_sigma, use_sit, use_checkpoint
        self.mlp_ratio, self.class_dropout_prob, self.num_classes = mlp_ratio, class_dropout_prob, num_classes
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels   # fit_model.py:78
        self.patch_size, self.num_heads = patch_size, num_heads
        self.adaln_type, self.adaln_lora_dim = adaln_type, adaln_lora_dim
        self.online_rope, self.time_shifting, self.save_attention = online_rope, time_shifting, False
        self.head_dim = hidden_size // num_heads
        self.use_swiglu, self.use_swiglu_large = bool(use_swiglu), bool(use_swiglu_large)
        if not use_swiglu:                                                  # modules.py:253: timm Mlp, fc2(GELU_tanh(fc1(x)))
            self.mlp_hidden = int(hidden_size * mlp_ratio)
            if self.mlp_hidden % 256: raise NotImplementedError(f"fitv2_b200.FiT: Mlp hidden {self.mlp_hidden} must be a multiple of 256")
        else:
            self.mlp_hidden = int(hidden_size * mlp_ratio) if use_swiglu_large else (int(hidden_size * mlp_ratio) * 2) // 3   # modules.py:246-251
        if self.mlp_hidden % 128: raise NotImplementedError(f"fitv2_b200.FiT: SwiGLU hidden {self.mlp_hidden} must be a multiple of 128")
        # modules.py:153,170-174: q / k (and v with add_rel_pe_to_v) are rotated only for rel_pos_embed 'rope' / 'xpos' (lower-cased);
        # anything else leaves them as they are -- here: zero frequencies (cos = 1, sin = 0 exactly) and magnitude 1
        self.rel_pos_embed = None if rel_pos_embed is None else str(rel_pos_embed).lower()
        self.rotates = self.rel_pos_embed in ("rope", "xpos")
        self.add_rel_pe_to_v = bool(add_rel_pe_to_v)
        self.operand_dtype = operand_dtype
        self.rope_args = dict(head_dim=self.head_dim, custom_freqs=custom_freqs, theta=rope_theta,
                              max_pe_len_h=max_pe_len_h, max_pe_len_w=max_pe_len_w, decouple=decouple,
                              ori_max_pe_len=ori_max_pe_len)
        if online_rope:                                                     # frequencies come from `size` at call time
            self.rope_args.update(max_pe_len_h=max_pe_len_h or ori_max_pe_len, max_pe_len_w=max_pe_len_w or ori_max_pe_len)
        rope_frequencies(**self.rope_args)                                  # validate early
        self._online_key, self._online_freqs = None, None

        # ---- parameters: same names / shapes / creation order as the reference (fit_model.py:84-112, modules.py:239-264) ----
        D, C = hidden_size, in_channels * patch_size ** 2
        wn = lambda code, dim: _NormWeight(dim) if code in (_lib.NORM_WLAYERNORM, _lib.NORM_RMSNORM) else _Holder()
        self.x_embedder = _Holder(); self.x_embedder.proj = nn.Linear(C, D)
        self.t_embedder = _Holder(); self.t_embedder.mlp = _seq(nn.Linear(256, D), nn.SiLU(), nn.Linear(D, D))
        self.y_embedder = _Holder()
        self.y_embedder.embedding_table = nn.Embedding(num_classes + (class_dropout_prob > 0), D)
        self.global_adaLN_modulation = _seq(nn.SiLU(), nn.Linear(D, 6 * D)) if adaln_type == "lora" else None
        blocks = []
        for _ in range(depth):
            blk = _Holder()
            blk.norm1, blk.norm2 = wn(self.block_norm, D), wn(self.block_norm, D)
            blk.attn = _Holder(); blk.attn.qkv = nn.Linear(D, 3 * D, bias=qkv_bias)
            blk.attn.q_norm, blk.attn.k_norm = wn(self.q_norm_code, self.head_dim), wn(self.k_norm_code, self.head_dim)
            blk.attn.proj = nn.Linear(D, D)
            blk.mlp = _Holder()
            if use_swiglu:
                blk.mlp.fc1_g = nn.Linear(D, self.mlp_hidden, bias=ffn_bias); blk.mlp.fc1_x = nn.Linear(D, self.mlp_hidden, bias=ffn_bias)
            else:
                blk.mlp.fc1 = nn.Linear(D, self.mlp_hidden, bias=ffn_bias)
            blk.mlp.fc2 = nn.Linear(self.mlp_hidden, D, bias=ffn_bias)
            if adaln_type == "lora":
                blk.adaLN_modulation = _seq(nn.SiLU(), nn.Linear(D, adaln_lora_dim), nn.Linear(adaln_lora_dim, 6 * D))
            elif adaln_type == "swiglu":                                    # modules.py:265-268: timm SwiGLU(D -> (D//4)*3 -> 6D)
                blk.adaLN_modulation = _swiglu_holder(D, (D // 4) * 3, 6 * D)
            else:
                blk.adaLN_modulation = _seq(nn.SiLU(), nn.Linear(D, 6 * D))
            blocks.append(blk)
        self.blocks = nn.ModuleList(blocks)
        self.final_layer = _Holder()
        self.final_layer.norm_final = wn(self.block_norm, D)
        self.final_layer.linear = nn.Linear(D, patch_size * patch_size * self.out_channels)
        if adaln_type == "swiglu":                                          # modules.py:284-285
            self.final_layer.adaLN_modulation = _swiglu_holder(D, D // 2, 2 * D)
        else:
            self.final_layer.adaLN_modulation = _seq(nn.SiLU(), nn.Linear(D, 2 * D))
        # C handle, packed kernel-side weights and workspace are created lazily by the first CUDA forward
        self._handle, self._handle_device, self._packed, self._workspace, self._ws_shape = None, None, None, None, None
        self.initialize_weights(pretrain_ckpt=pretrain_ckpt, ignore=ignore_keys)
        if finetune is not None:                                            # fit_model.py:114-115
            self.finetune(type=finetune, unfreeze=ignore_keys)
        _MODELS[id(self)] = self

    @property
    def out_token_channels(self) -> int:
        return self.patch_size * self.patch_size * self.out_channels

    # ------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------
    def initialize_weights(self, pretrain_ckpt=None, ignore=None):
        """Same scheme as fit_model.py:117-157 (xavier Linears, N(0,0.02) tables, zeroed adaLN/final); with ``pretrain_ckpt`` the
        checkpoint is loaded on top (fit_model.py:159-170: every key that CONTAINS an ``ignore`` string is left at its init)."""
        def _basic_init(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.apply(_basic_init)
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        last = (lambda m: m.fc2) if self.adaln_type == "swiglu" else (lambda m: m[-1])   # fit_model.py:139-153
        for blk in self.blocks:
            nn.init.constant_(last(blk.adaLN_modulation).weight, 0)
            nn.init.constant_(last(blk.adaLN_modulation).bias, 0)
        if self.global_adaLN_modulation is not None:                        # fit_model.py:146-148 (adaln_type 'lora' only)
            nn.init.constant_(self.global_adaLN_modulation[-1].weight, 0)
            nn.init.constant_(self.global_adaLN_modulation[-1].bias, 0)
        for lin in (last(self.final_layer.adaLN_modulation), self.final_layer.linear):
            nn.init.constant_(lin.weight, 0)
            nn.init.constant_(lin.bias, 0)
        if pretrain_ckpt is not None:
            from .checkpoint import init_from_ckpt
            keys = list(self.state_dict().keys())
            ignore_keys = sorted({key for ign in (ignore or []) for key in keys if ign in key})
            init_from_ckpt(self, pretrain_ckpt, ignore_keys, verbose=True)

    def finetune(self, type, unfreeze):
        """fit_model.py:291-299: 'full' leaves everything trainable; otherwise only parameters whose name contains one of the
        ``unfreeze`` strings keep requires_grad (bookkeeping for the training scripts; the forward here never builds a graph)."""
        if type == "full":
            return
        for _, p in self.named_parameters():
            p.requires_grad = False
        for unf in (unfreeze or []):
            for name, p in self.named_parameters():
                if unf in name:
                    p.requires_grad = True

    @torch.no_grad()
    def randomize_zero_init_(self, seed: int = 1, std: float = 0.02):
        """Benchmark / parity helper: the reference init makes the output identically zero (all adaLN
        outputs and the final linear are zero-initialised), so every all-zero tensor is re-drawn
        N(0, std^2) from a dedicated CPU generator in ``state_dict()`` order."""
        g = torch.Generator().manual_seed(seed)
        for _, p in self.state_dict().items():
            if p.is_floating_point() and not bool(p.any()):
                p.copy_((torch.randn(p.shape, generator=g, dtype=torch.float32) * std).to(p.dtype))
        self._packed = None
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._packed = None
        return out

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._packed = None
        self._workspace = None
        self._ws_shape = None
        self._online_key = None
        return out

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().fitv2_destroy(self._handle)
        self._handle, self._handle_device, self._packed, self._workspace, self._ws_shape = None, None, None, None, None

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    # ------------------------------------------------------------------------------------------
    # C-ABI plumbing
    # ------------------------------------------------------------------------------------------
    def _require_cuda(self):
        dev = self.device
        if dev.type != "cuda":
            raise _lib.FitV2Error("fitv2_b200.FiT runs on CUDA (sm_100a) only; move the module with .to('cuda'). "
                                  "There is no CPU fallback.")
        return dev

    @torch.no_grad()
    def _ensure_packed(self):
        if self._packed is not None:
            return
        dev = self._require_cuda()
        lib = _lib.load()
        fh, fw, mag = rope_frequencies(**self.rope_args)
        if self._handle is not None and self._handle_device != dev:         # .to('cuda:1'): a handle belongs to one device
            self._drop_handle()
        with torch.cuda.device(dev):
            if self._handle is None:
                cfg = _lib.FitV2Config(self.hidden_size, self.depth, self.num_heads, self.head_dim, self.mlp_hidden,
                                       self.adaln_lora_dim or 0, self.in_channels * self.patch_size ** 2,
                                       self.y_embedder.embedding_table.weight.shape[0],
                                       _lib.OPERAND_FP16 if self.operand_dtype == "fp16" else _lib.OPERAND_BF16,
                                       float(self.time_shifting), float(mag) if self.rotates else 1.0,
                                       out_channels=self.out_token_channels,
                                       adaln_type={"lora": _lib.ADALN_LORA, "normal": _lib.ADALN_NORMAL, "swiglu": _lib.ADALN_SWIGLU}[self.adaln_type],
                                       block_norm=self.block_norm, q_norm=self.q_norm_code, k_norm=self.k_norm_code,
                                       channels_first=0 if self.use_sit else 1,
                                       mlp_type=_lib.MLP_SWIGLU if self.use_swiglu else _lib.MLP_GELU,
                                       rope_v=1 if (self.add_rel_pe_to_v and self.rotates) else 0)
                h = C.c_void_p()
                _lib.check(lib.fitv2_create(C.byref(cfg), C.byref(h)), "fitv2_create")
                self._handle, self._handle_device = h, dev
                _lib.apply_env_options(h)                                   # FITV2_* tuning switches -> fitv2_set_option
            op = torch.float16 if self.operand_dtype == "fp16" else torch.bfloat16
            P = self.pack_weights(dev)
            for name, tns in P.items():
                want = op if name in _lib.OP16_SLOTS else torch.float32
                assert tns.dtype == want and tns.is_contiguous(), name
                _lib.check(lib.fitv2_bind_weight(self._handle, _lib.SLOT[name], C.c_void_p(tns.data_ptr()), tns.numel()),
                           f"fitv2_bind_weight({name})")
            self._packed = P

    def set_option(self, name: str, value: int):
        """Per-handle tuning switch (include/fitv2_b200.h: fitv2_set_option)."""
        self._ensure_packed()
        _lib.check(_lib.load().fitv2_set_option(self._handle, name.encode(), int(value)), f"fitv2_set_option({name})")
        if name == "ws_guard":
            self._ws_shape = None                                           # the workspace is re-sized at the next call

    def workspace_layout(self):
        """[(offset, bytes)] of the buffers the last forward used inside ``self._workspace`` (fitv2_debug_layout)."""
        off, size = (C.c_int64 * 40)(), (C.c_int64 * 40)()
        n = _lib.load().fitv2_debug_layout(self._handle, off, size, 40)
        if n < 0:
            _lib.check(n, "fitv2_debug_layout")
        return [(int(off[i]), int(size[i])) for i in range(n)]

    def check_device_errors(self):
        """Raise if a kernel of an earlier call saw an out-of-range class label (the reference raises an IndexError /
        device assert there).  Non-blocking; meaningful after a synchronisation."""
        if self._handle is not None:
            _lib.check(_lib.load().fitv2_poll_error(self._handle), "fitv2_poll_error")

    @torch.no_grad()
    def pack_weights(self, dev) -> dict:
        """Kernel-side weight layouts (include/fitv2_b200.h, enum fitv2_weight): fp32 conditioning weights,
        per-block weights stacked over depth, 16-bit GEMM operands, fc1_g/fc1_x interleaved in 128-row groups
        so that one 256-wide GEMM tile holds matching gate and up columns."""
        fh, fw, _ = rope_frequencies(**self.rope_args)
        if not self.rotates:
            fh, fw = torch.zeros_like(fh), torch.zeros_like(fw)
        op = torch.float16 if self.operand_dtype == "fp16" else torch.bfloat16
        f32 = lambda p: p.detach().to(device=dev, dtype=torch.float32).contiguous()
        blocks = self.blocks
        stack32 = lambda get: torch.stack([get(b).detach().to(device=dev, dtype=torch.float32) for b in blocks]).contiguous()

        def bias32(get, n):
            """Stacked bias of a per-block Linear; zeros for a bias-free one (qkv_bias / ffn_bias False)."""
            if get(blocks[0]).bias is None:
                return torch.zeros(len(blocks), n, dtype=torch.float32, device=dev)
            return stack32(lambda b: get(b).bias)

        def tf32(w):
            # adaLN weights are tf32 operands of the tensor-pipe conditioning linears (csrc/cond_tc.cuh): round them to
            # nearest once here (the tensor pipe itself would truncate the 13 low mantissa bits)
            if os.environ.get("FITV2_COND") == "simt":
                return w
            return ((w.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        stack16 = lambda get: torch.stack([get(b).detach().to(device=dev, dtype=torch.float32).to(op) for b in blocks]).contiguous()
        Hm = self.mlp_hidden

        def gateup(b, attr):
            if getattr(b.mlp.fc1_g, attr) is None:                          # ffn_bias False
                return torch.zeros(2 * Hm, dtype=torch.float32, device=dev)
            g = getattr(b.mlp.fc1_g, attr).detach().to(device=dev, dtype=torch.float32)
            u = getattr(b.mlp.fc1_x, attr).detach().to(device=dev, dtype=torch.float32)
            g = g.reshape(Hm // 128, 128, *g.shape[1:])
            u = u.reshape(Hm // 128, 128, *u.shape[1:])
            return torch.cat([g, u], dim=1).reshape(2 * Hm, *g.shape[2:])

        P = {
            "X_EMBED_W": f32(self.x_embedder.proj.weight), "X_EMBED_B": f32(self.x_embedder.proj.bias),
            "T_MLP0_W": f32(self.t_embedder.mlp[0].weight), "T_MLP0_B": f32(self.t_embedder.mlp[0].bias),
            "T_MLP2_W": f32(self.t_embedder.mlp[2].weight), "T_MLP2_B": f32(self.t_embedder.mlp[2].bias),
            "Y_TABLE": f32(self.y_embedder.embedding_table.weight),
            "FINAL_LINEAR_W": f32(self.final_layer.linear.weight), "FINAL_LINEAR_B": f32(self.final_layer.linear.bias),
            "QKV_W": stack16(lambda b: b.attn.qkv.weight), "QKV_B": bias32(lambda b: b.attn.qkv, 3 * self.hidden_size),
            "PROJ_W": stack16(lambda b: b.attn.proj.weight), "PROJ_B": stack32(lambda b: b.attn.proj.bias),
            "FC2_W": stack16(lambda b: b.mlp.fc2.weight), "FC2_B": bias32(lambda b: b.mlp.fc2, self.hidden_size),
            "ROPE_FREQS_H": fh.to(dev).contiguous(), "ROPE_FREQS_W": fw.to(dev).contiguous(),
        }
        if self.use_swiglu:
            P["GATEUP_W"] = torch.stack([gateup(b, "weight").to(op) for b in blocks]).contiguous()
            P["GATEUP_B"] = torch.stack([gateup(b, "bias") for b in blocks]).contiguous()
        else:                                                               # GELU Mlp: the slots hold fc1 alone, no interleave
            P["GATEUP_W"] = stack16(lambda b: b.mlp.fc1.weight)
            P["GATEUP_B"] = bias32(lambda b: b.mlp.fc1, Hm)
        if self.adaln_type == "lora":
            P.update({
                "GLOBAL_ADALN_W": tf32(f32(self.global_adaLN_modulation[1].weight)),
                "GLOBAL_ADALN_B": f32(self.global_adaLN_modulation[1].bias),
                "LORA_A_W": tf32(stack32(lambda b: b.adaLN_modulation[1].weight)), "LORA_A_B": stack32(lambda b: b.adaLN_modulation[1].bias),
                "LORA_B_W": tf32(stack32(lambda b: b.adaLN_modulation[2].weight)), "LORA_B_B": stack32(lambda b: b.adaLN_modulation[2].bias),
            })
        elif self.adaln_type == "normal":                                   # modules.py:254-258: one Linear(D -> 6D) per block
            P.update({"NORMAL_ADALN_W": tf32(stack32(lambda b: b.adaLN_modulation[1].weight)),
                      "NORMAL_ADALN_B": stack32(lambda b: b.adaLN_modulation[1].bias)})
        if self.adaln_type == "swiglu":                                     # modules.py:265-268,284-285 (fp32 FMA kernels: no tf32 rounding)
            fa = self.final_layer.adaLN_modulation
            for tag, name in (("G", "fc1_g"), ("X", "fc1_x"), ("FC2", "fc2")):
                P[f"SG_{tag}_W"] = stack32(lambda b: getattr(b.adaLN_modulation, name).weight)
                P[f"SG_{tag}_B"] = stack32(lambda b: getattr(b.adaLN_modulation, name).bias)
                P[f"FSG_{tag}_W"], P[f"FSG_{tag}_B"] = f32(getattr(fa, name).weight), f32(getattr(fa, name).bias)
        else:
            P.update({"FINAL_ADALN_W": tf32(f32(self.final_layer.adaLN_modulation[1].weight)),
                      "FINAL_ADALN_B": f32(self.final_layer.adaLN_modulation[1].bias)})
        if self.block_norm != _lib.NORM_LAYERNORM:
            P.update({"NORM1_W": stack32(lambda b: b.norm1.weight), "NORM2_W": stack32(lambda b: b.norm2.weight),
                      "NORM_FINAL_W": f32(self.final_layer.norm_final.weight)})
        if self.q_norm_code in (_lib.NORM_WLAYERNORM, _lib.NORM_RMSNORM):
            P["Q_NORM_W"] = stack32(lambda b: b.attn.q_norm.weight)
        if self.k_norm_code in (_lib.NORM_WLAYERNORM, _lib.NORM_RMSNORM):
            P["K_NORM_W"] = stack32(lambda b: b.attn.k_norm.weight)
        return P

    def _ensure_workspace(self, rows: int, tokens: int):
        if self._ws_shape == (rows, tokens):
            return
        lib = _lib.load()
        need = lib.fitv2_workspace_bytes(self._handle, rows, tokens)
        if need <= 0:
            _lib.check(int(need), "fitv2_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need:
            # FITV2_POISON_WORKSPACE=empty leaves the scratch memory unwritten (compute-sanitizer initcheck runs)
            alloc = torch.empty if os.environ.get("FITV2_POISON_WORKSPACE") == "empty" else torch.zeros
            self._workspace = alloc(int(need), dtype=torch.uint8, device=self.device)
            _lib.check(lib.fitv2_set_workspace(self._handle, C.c_void_p(self._workspace.data_ptr()), self._workspace.numel()),
                       "fitv2_set_workspace")
        if os.environ.get("FITV2_POISON_WORKSPACE") == "1":             # tests: every byte NaN, so a read of scratch memory that the
            self._workspace.fill_(0xFF)                                 # current call did not write shows up in the result
        self._ws_shape = (rows, tokens)

    def _run(self, x: torch.Tensor, t, y, grid, mask, rows: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x fp32 contiguous, (x_rows, N, C) (use_sit) or (x_rows, C, N), with x_rows == rows or rows // 2 (implicit CFG
        duplication).  Returns (rows, N, C_out) / (rows, C_out, N)."""
        self._ensure_packed()
        dev = self.device
        for name, tns in (("x", x), ("t", t), ("y", y), ("grid", grid), ("mask", mask)):
            if tns.device != dev:
                raise _lib.FitV2Error(f"{name} is on {tns.device}, the model on {dev}")
        if self.use_sit:
            x_rows, tokens, ch = x.shape
        else:
            x_rows, ch, tokens = x.shape
        assert x.dtype == torch.float32 and x.is_contiguous()
        t = t.to(torch.float32).contiguous()
        y = y.to(torch.int64).contiguous()
        grid = grid.to(torch.int64).contiguous()
        mask = mask.to(torch.float32).contiguous()
        if ch != self.in_channels * self.patch_size ** 2 or t.shape != (rows,) or y.shape != (rows,) or grid.shape != (rows, 2, tokens) \
                or mask.shape != (rows, tokens):
            raise ValueError(f"shape mismatch: rows={rows} tokens={tokens} x{tuple(x.shape)} t{tuple(t.shape)} y{tuple(y.shape)} "
                             f"grid{tuple(grid.shape)} mask{tuple(mask.shape)}")
        co = self.out_token_channels
        if out is None:
            out = torch.empty((rows, tokens, co) if self.use_sit else (rows, co, tokens), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self._ensure_workspace(rows, tokens)
            st = torch.cuda.current_stream(dev).cuda_stream
            lib = _lib.load()
            _lib.check(lib.fitv2_forward(self._handle, C.c_void_p(x.data_ptr()), x_rows, C.c_void_p(t.data_ptr()),
                                         C.c_void_p(y.data_ptr()), C.c_void_p(grid.data_ptr()), C.c_void_p(mask.data_ptr()),
                                         C.c_void_p(out.data_ptr()), rows, tokens, C.c_void_p(st)), "fitv2_forward")
        return out

    # ------------------------------------------------------------------------------------------
    # reference-facing API
    # ------------------------------------------------------------------------------------------
    def _set_online_rope(self, size, rows: int):
        """online_rope (fit_model.py:212-214): per-sample frequencies from ``size`` (B,1,2).  The cache is keyed on the VALUES
        of ``size`` (a fresh tensor per batch usually reuses the freed allocation of the previous one, so a storage key would
        hand a new batch the old frequencies); a sampling loop should bind once through ``EulerCFGSampler(size=...)``."""
        if not self.online_rope or not self.rotates:
            return
        if size is None:
            raise ValueError("online_rope=True needs `size` (B, 1, 2) = (h, w) patches per sample (fit_model.py:212-213)")
        if size.reshape(-1, 2).shape[0] != rows:
            raise ValueError(f"size has {size.reshape(-1, 2).shape[0]} rows, the batch {rows}")
        self._ensure_packed()
        key = tuple(size.reshape(-1).tolist())
        if key != self._online_key:
            a = self.rope_args
            fh, fw = online_rope_frequencies(a["head_dim"], a["custom_freqs"], a["theta"], a["decouple"], a["ori_max_pe_len"], size)
            self._online_freqs = (fh.to(self.device), fw.to(self.device))  # kept alive until replaced: the forward reads them
            self._online_key = key
        fh, fw = self._online_freqs
        _lib.check(_lib.load().fitv2_set_online_rope(self._handle, C.c_void_p(fh.data_ptr()), C.c_void_p(fw.data_ptr()), rows),
                   "fitv2_set_online_rope")

    @torch.no_grad()
    def forward(self, x, t, y, grid, mask, size=None):
        """fit_model.py:189-233.  x (B,N,p*p*C) (use_sit) or (B,p*p*C,N), t (B,), y (B,), grid (B,2,N), mask (B,N); ``size``
        (B,1,2) is read only with online_rope=True.  Returns the model output (B,N,p*p*C_out) / (B,p*p*C_out,N) in x.dtype.
        Dispatched as the registered operator ``fitv2_b200::forward`` (traceable, see the top of this file)."""
        xf = x.to(torch.float32).contiguous()
        if xf.is_cuda:
            self._set_online_rope(size, x.shape[0])
            if _MODELS.get(id(self)) is not self:                           # e.g. a module restored by pickle (no __init__ ran)
                _MODELS[id(self)] = self
            out = torch.ops.fitv2_b200.forward(xf, t, y, grid, mask, id(self), x.shape[0])
        else:
            out = self._run(xf, t, y, grid, mask, rows=x.shape[0])          # raises: no CPU path
        return out if x.dtype == torch.float32 else out.to(x.dtype)

    @torch.no_grad()
    def forward_with_cfg(self, x, t, y, grid, mask, size, cfg_scale, scale_pow=0.0):
        """fit_model.py:235-275: forward on cat([x[:B], x[:B]]), CFG on the first 3*p*p channels only."""
        rows = x.shape[0]
        half = rows // 2
        xf = x[:half].to(torch.float32).contiguous()
        self._set_online_rope(size, rows)
        out = self._run(xf, t, y, grid, mask, rows=rows)               # implicit cat([half, half])
        c_cfg = 3 * self.patch_size * self.patch_size
        scale_ptr, scale = None, float(cfg_scale)
        if scale_pow != 0.0:
            tt = t.to(torch.float32)
            step = (1 - torch.cos(((1 - torch.clamp_max(tt, 1.0)) ** scale_pow) * torch.pi)) * 1 / 2
            per = ((cfg_scale - 1) * step + 1)[:half].contiguous()
            scale_ptr = C.c_void_p(per.data_ptr())
        lib = _lib.load()
        # (2B, N, C): guided channels are the first c_cfg of every token; (2B, C, N) (use_sit = False): the first c_cfg * N
        # elements of every sample -- the same kernel with one "token" of C * N channels
        tokens, chans, cc = (out.shape[1], out.shape[2], c_cfg) if self.use_sit else (1, out.shape[1] * out.shape[2], c_cfg * out.shape[2])
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.fitv2_cfg_combine(C.c_void_p(out.data_ptr()), scale_ptr, scale, half, tokens, chans,
                                             cc, C.c_void_p(st)), "fitv2_cfg_combine")
        return out if x.dtype == torch.float32 else out.to(x.dtype)

    def unpatchify(self, x, hw, scaling_factor: float = 1.0):
        """fit_model.py:171-187 (use_sit): (B,(h w),(c p1 p2)) -> (B,c,h*p1,w*p2), optionally fused with the latent
        scaling ``samples / vae.config.scaling_factor`` of sample_fitv2_ddp.py:320-321 (one kernel, bit-exact).
        CUDA fp32 tensors go through the C ABI (``fitv2_unpatchify_scale``); anything else (the reference also calls
        this on CPU tensors in its data tools) is the same index permutation as a tensor view."""
        h, w = hw
        p = self.patch_size
        B = x.shape[0]
        if not self.use_sit:                                                # fit_model.py:184-186: 'b (c p1 p2) (h w)' input
            x = x.transpose(1, 2)
        if x.is_cuda and x.dtype == torch.float32:
            x = x.contiguous()
            ch = x.shape[2] // (p * p)
            out = torch.empty(B, ch, h, w, dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                st = torch.cuda.current_stream(x.device).cuda_stream
                _lib.check(_lib.load().fitv2_unpatchify_scale(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), float(scaling_factor),
                                                              B, h // p, w // p, ch, p, C.c_void_p(st)), "fitv2_unpatchify_scale")
            return out
        x = x.reshape(B, h // p, w // p, -1, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, -1, h, w)
        return x if scaling_factor == 1.0 else x / scaling_factor

    # ------------------------------------------------------------------------------------------
    # attention-map API of the reference class (fit_model.py:301-330)
    # ------------------------------------------------------------------------------------------
    def get_attention_maps(self):
        """fit_model.py:301-316: ``None`` unless save_attention is enabled -- which it never is here (the attention kernels keep
        S and P in tensor / shared memory and never materialise the (B, H, N, N) matrix)."""
        return None

    def enable_attention_visualization(self):
        raise NotImplementedError("fitv2_b200.FiT does not materialise attention maps (save_attention); use the reference model "
                                  "for visualisation")

    def disable_attention_visualization(self):
        """fit_model.py:325-330: nothing to switch off."""
        self.save_attention = False

    # ------------------------------------------------------------------------------------------
    # introspection used by tests / bench
    # ------------------------------------------------------------------------------------------
    def kernel_launches(self) -> int:
        return int(_lib.load().fitv2_kernel_launches(self._handle)) if self._handle is not None else 0

    def profile(self, classes=None):
        """Enable (list of names from _lib.PROFILE_CLASSES, or 'all') / disable (None) CUDA-event timing per
        kernel class inside fitv2_forward."""
        self._ensure_packed()
        if classes is None:
            mask = 0
        elif classes == "all":
            mask = (1 << len(_lib.PROFILE_CLASSES)) - 1
        else:
            mask = sum(1 << _lib.PROFILE_CLASSES.index(c) for c in classes)
        _lib.check(_lib.load().fitv2_profile_set(self._handle, mask), "fitv2_profile_set")

    def profile_read(self) -> dict:
        """{class: (total_ms, launches)} accumulated since the last read (synchronises on the events)."""
        n = len(_lib.PROFILE_CLASSES)
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(_lib.load().fitv2_profile_read(self._handle, ms, cnt), "fitv2_profile_read")
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(_lib.PROFILE_CLASSES)}

    def debug_tap(self, name: str) -> torch.Tensor:
        """Copy an internal buffer of the LAST forward (see _lib.TAPS)."""
        rows, tokens = self._ws_shape
        D, H, dh, Hm, L = self.hidden_size, self.num_heads, self.head_dim, self.mlp_hidden, self.depth
        op = torch.float16 if self.operand_dtype == "fp16" else torch.bfloat16
        tv = (tokens + 7) // 8 * 8
        shapes = dict(c=((rows, D), torch.float32), gmod=((rows, 6 * D), torch.float32), mod=((L, rows, 6 * D), torch.float32),
                      fmod=((rows, 2 * D), torch.float32), x_res=((rows, tokens, D), torch.float32),
                      q=((rows, H, tokens, dh), op), k=((rows, H, tokens, dh), op), vt=((rows, H, dh, tv), op),
                      attn_out=((rows, tokens, D), op), h=((rows, tokens, D), op), hidden=((rows, tokens, Hm), op),
                      rope_cos=((dh // 2, rows, tokens), torch.float32), rope_sin=((dh // 2, rows, tokens), torch.float32),
                      seg_uniform=((rows,), torch.int32))
        shape, dt = shapes[name]
        dst = torch.empty(shape, dtype=dt, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.load().fitv2_debug_tap(self._handle, _lib.TAPS[name], C.c_void_p(dst.data_ptr()),
                                               dst.numel() * dst.element_size(), C.c_void_p(st)), f"tap {name}")
        return dst

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().fitv2_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def __deepcopy__(self, memo):
        # the C handle / workspace are per instance: a copy gets its own lazily
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_handle", "_handle_device", "_packed", "_workspace", "_ws_shape", "_online_key", "_online_freqs"):
                setattr(new, k, None)
            else:
                setattr(new, k, copy.deepcopy(v, memo))
        _MODELS[id(new)] = new
        return new
