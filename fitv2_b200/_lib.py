"""ctypes binding of the C-ABI shared library (include/fitv2_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is NO fallback:
if the shared object is missing or a call fails, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FITV2_B200_LIB", os.path.join(_HERE, "libfitv2_b200.so"))   # override: A/B kernel experiments

OPERAND_BF16, OPERAND_FP16 = 0, 1

# enum fitv2_weight (include/fitv2_b200.h) — keep in the same order
WEIGHT_SLOTS = [
    "X_EMBED_W", "X_EMBED_B", "T_MLP0_W", "T_MLP0_B", "T_MLP2_W", "T_MLP2_B", "Y_TABLE",
    "GLOBAL_ADALN_W", "GLOBAL_ADALN_B", "LORA_A_W", "LORA_A_B", "LORA_B_W", "LORA_B_B",
    "FINAL_ADALN_W", "FINAL_ADALN_B", "FINAL_LINEAR_W", "FINAL_LINEAR_B",
    "QKV_W", "QKV_B", "PROJ_W", "PROJ_B", "GATEUP_W", "GATEUP_B", "FC2_W", "FC2_B",
    "ROPE_FREQS_H", "ROPE_FREQS_W",
    "NORMAL_ADALN_W", "NORMAL_ADALN_B", "NORM1_W", "NORM2_W", "NORM_FINAL_W", "Q_NORM_W", "K_NORM_W",
    "SG_G_W", "SG_G_B", "SG_X_W", "SG_X_B", "SG_FC2_W", "SG_FC2_B",
    "FSG_G_W", "FSG_G_B", "FSG_X_W", "FSG_X_B", "FSG_FC2_W", "FSG_FC2_B",
]
SLOT = {n: i for i, n in enumerate(WEIGHT_SLOTS)}
OP16_SLOTS = {"QKV_W", "PROJ_W", "GATEUP_W", "FC2_W"}

NORM_NONE, NORM_LAYERNORM, NORM_WLAYERNORM, NORM_RMSNORM = 0, 1, 2, 3       # FITV2_NORM_*
ADALN_LORA, ADALN_NORMAL, ADALN_SWIGLU = 0, 1, 2                            # FITV2_ADALN_*
MLP_SWIGLU, MLP_GELU = 0, 1                                                 # FITV2_MLP_*


def norm_code(name, weight: bool = False) -> int:
    """fit/model/norms.py:35-50 (create_norm) names -> FITV2_NORM_*; ``weight`` is the ``qk_norm_weight`` promotion of
    modules.py:141-144 ('layernorm' -> 'w_layernorm')."""
    if name is None or name == "" or str(name).lower() == "none":
        return NORM_NONE
    name = str(name).lower()
    if name == "layernorm":
        return NORM_WLAYERNORM if weight else NORM_LAYERNORM
    if name == "w_layernorm":
        return NORM_WLAYERNORM
    if name in ("rmsnorm", "w_rmsnorm"):
        return NORM_RMSNORM
    raise NotImplementedError(f"Unknown norm_type: '{name}'")


# fitv2_set_option names <- FITV2_* environment variables (read when a handle is created, so tests / A/B tools can still use
# the environment; the library itself never calls getenv)
_ENV_OPTIONS = {
    "FITV2_PDL": ("pdl", int), "FITV2_ATTN": ("attn", lambda v: {"tm": 1, "ws": 2, "general": 3, "v1": 3, "tm1": 4}.get(v, 0)),
    "FITV2_ATTN_EARLY": ("attn_early", int), "FITV2_LN_THREADS": ("ln_threads", int),
    "FITV2_LN_WIDE_SINGLE": ("ln_wide_single", int), "FITV2_BN_RESID": ("bn_resid", int), "FITV2_QKV": ("qkv_heads", int),
    "FITV2_RESID_T": ("resid_t", int), "FITV2_BN_RESID_T": ("bn_resid_t", int),
    "FITV2_COND": ("cond", lambda v: 1 if v == "simt" else 0), "FITV2_L2_PERSIST_MB": ("l2_persist_mb", int),
    "FITV2_FINAL_TC": ("final_tc", int), "FITV2_GELU_EPI": ("gelu_epi", int), "FITV2_VERBOSE": ("verbose", int),
}


def apply_env_options(handle):
    lib = load()
    for env, (name, conv) in _ENV_OPTIONS.items():
        v = os.environ.get(env)
        if v is not None and v != "":
            check(lib.fitv2_set_option(handle, name.encode(), int(conv(v))), f"fitv2_set_option({name})")


PROFILE_CLASSES = ["conditioning", "ln_modulate", "qkv_gemm", "attention", "proj_gemm", "gateup_gemm", "fc2_gemm", "embed_final"]

TAPS = dict(c=0, gmod=1, mod=2, fmod=3, x_res=4, q=5, k=6, vt=7, attn_out=8, h=9, hidden=10,
            rope_cos=11, rope_sin=12, seg_uniform=13)


class FitV2Config(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int32), ("depth", C.c_int32), ("num_heads", C.c_int32), ("head_dim", C.c_int32),
        ("mlp_hidden", C.c_int32), ("lora_dim", C.c_int32), ("token_channels", C.c_int32),
        ("num_embeddings", C.c_int32), ("operand_dtype", C.c_int32), ("time_shifting", C.c_float),
        ("rope_magnitude", C.c_float),
        ("out_channels", C.c_int32), ("adaln_type", C.c_int32), ("block_norm", C.c_int32), ("q_norm", C.c_int32),
        ("k_norm", C.c_int32), ("channels_first", C.c_int32), ("mlp_type", C.c_int32), ("rope_v", C.c_int32),
    ]

    def __init__(self, *a, **kw):
        # FiTv2 defaults for the variant fields (so that positional 11-field constructions keep meaning FiTv2)
        full = dict(out_channels=0, adaln_type=ADALN_LORA, block_norm=NORM_LAYERNORM, q_norm=NORM_LAYERNORM,
                    k_norm=NORM_LAYERNORM, channels_first=0, mlp_type=MLP_SWIGLU, rope_v=0)
        full.update(kw)
        super().__init__(*a, **full)


class FitV2Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libfitv2_b200.so; raises FitV2Error (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FitV2Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fitv2_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.fitv2_last_error.restype = C.c_char_p
    lib.fitv2_version.restype = C.c_char_p
    lib.fitv2_create.argtypes = [C.POINTER(FitV2Config), C.POINTER(vp)]
    lib.fitv2_destroy.argtypes = [vp]
    lib.fitv2_destroy.restype = None
    lib.fitv2_bind_weight.argtypes = [vp, i32, vp, i64]
    lib.fitv2_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.fitv2_poll_error.argtypes = [vp]
    lib.fitv2_rk_stage.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i64, vp]
    lib.fitv2_lincomb.argtypes = [vp, vp, C.POINTER(vp), vp, i32, i64, vp]
    lib.fitv2_scaled_rms.argtypes = [vp, vp, vp, vp, f32, f32, i64, vp]
    lib.fitv2_set_online_rope.argtypes = [vp, vp, vp, i32]
    lib.fitv2_workspace_bytes.argtypes = [vp, i32, i32]
    lib.fitv2_workspace_bytes.restype = i64
    lib.fitv2_set_workspace.argtypes = [vp, vp, i64]
    lib.fitv2_forward.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.fitv2_cfg_combine.argtypes = [vp, vp, f32, i32, i32, i32, i32, vp]
    lib.fitv2_cfg_euler.argtypes = [vp, vp, f32, f32, vp, i32, i32, i32, vp]
    lib.fitv2_sde_step.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.fitv2_sde_drift.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.fitv2_scaled_add.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.fitv2_heun_combine.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.fitv2_tweedie.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.fitv2_unpatchify_scale.argtypes = [vp, vp, f32, i32, i32, i32, i32, i32, vp]
    lib.fitv2_pack_uint8.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.fitv2_debug_gemm.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.fitv2_debug_attention.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.fitv2_debug_tap.argtypes = [vp, i32, vp, i64, vp]
    lib.fitv2_debug_layout.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), i32]
    lib.fitv2_kernel_launches.argtypes = [vp]
    lib.fitv2_kernel_launches.restype = i64
    lib.fitv2_profile_set.argtypes = [vp, C.c_uint32]
    lib.fitv2_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().fitv2_last_error().decode("utf-8", "replace")
        raise FitV2Error(f"{what or 'fitv2 call'} failed ({rc}): {msg}")


EXPORTED_SYMBOLS = [
    "fitv2_last_error", "fitv2_version", "fitv2_create", "fitv2_destroy", "fitv2_bind_weight",
    "fitv2_set_online_rope", "fitv2_workspace_bytes", "fitv2_set_workspace", "fitv2_forward", "fitv2_cfg_combine", "fitv2_cfg_euler",
    "fitv2_set_option", "fitv2_poll_error", "fitv2_rk_stage", "fitv2_lincomb", "fitv2_scaled_rms",
    "fitv2_sde_step", "fitv2_sde_drift", "fitv2_scaled_add", "fitv2_heun_combine", "fitv2_tweedie", "fitv2_unpatchify_scale", "fitv2_pack_uint8",
    "fitv2_debug_gemm", "fitv2_debug_attention", "fitv2_debug_tap", "fitv2_debug_layout", "fitv2_kernel_launches",
    "fitv2_profile_set", "fitv2_profile_read",
]
