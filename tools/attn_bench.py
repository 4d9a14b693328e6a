"""Micro-benchmark of the attention kernel alone through the C ABI (fitv2_debug_attention).
    python tools/attn_bench.py [R,T,dh[,reps]] ...     (FITV2_B200_LIB selects the library, FITV2_ATTN the kernel: tm / ws / general)
Default shapes: 64,256,72 (headline), 64,1024,72 (512x512), 64,256,96 (3B/2)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fitv2_b200 import _lib

specs = [a for a in sys.argv[1:] if "," in a] or ["64,256,72", "64,1024,72", "64,256,96"]
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr())
for spec in specs:
    parts = [int(v) for v in spec.split(",")]
    R, T, dh = parts[:3]
    reps = parts[3] if len(parts) > 3 else 30
    H = 16 if dh == 72 else 24
    cfg = _lib.FitV2Config(H * dh, 1, H, dh, 3072 if dh == 72 else 6144, 288 if dh == 72 else 576, 16, 1001, 0, 1.0, 1.0)
    h = C.c_void_p()
    _lib.check(lib.fitv2_create(C.byref(cfg), C.byref(h)))
    _lib.apply_env_options(h)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    _lib.check(lib.fitv2_set_workspace(h, p(ws), ws.numel()))
    g = torch.Generator(device="cuda").manual_seed(0)
    ln = lambda x: torch.nn.functional.layer_norm(x, (dh,))
    q, k = [ln(torch.randn(R, H, T, dh, generator=g, device="cuda")).bfloat16() for _ in range(2)]
    tv = (T + 7) // 8 * 8
    vt = torch.zeros(R, H, dh, tv, dtype=torch.bfloat16, device="cuda")
    vt[..., :T] = torch.randn(R, H, dh, T, generator=g, device="cuda").bfloat16()
    mask = torch.ones(R, T, device="cuda")
    if os.environ.get("ATTN_MASKED"):                      # mixed-aspect padded batch: lengths 200 / 256 / 192 / 200, pad id 0
        for r in range(R):
            mask[r, (200, 256, 192, 200)[r % 4] * T // 256:] = 0
    out = torch.empty(R, T, H * dh, dtype=torch.bfloat16, device="cuda")
    call = lambda: _lib.check(lib.fitv2_debug_attention(h, p(q), p(k), p(vt), p(mask), p(out), R, T, None, None, None))
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3      # includes the tiny seg_uniform kernel in front of every launch
    fl = 4.0 * R * H * T * T * dh
    # accuracy against the fp32 formula on the first sample (also catches a broken experiment build)
    qf, kf, vf = q[:1].float(), k[:1].float(), vt[:1, :, :, :T].float().transpose(-1, -2)
    am = mask[:1, None, None, :]
    ref = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, attn_mask=(am == am.transpose(-2, -1)))
    ref = ref.transpose(1, 2).reshape(1, T, H * dh) * (mask[:1] != 0)[..., None]
    err = float((out[:1].float() - ref).abs().max() / ref.abs().max())
    print(f"attention R={R} T={T} dh={dh} lib={os.path.basename(_lib.LIB_PATH)} attn={os.environ.get('FITV2_ATTN', 'auto')}: "
          f"{us:.1f} us/launch  {fl / us / 1e6:.0f} TFLOP/s  max-rel-err {err:.2e}", flush=True)
    lib.fitv2_destroy(h)
