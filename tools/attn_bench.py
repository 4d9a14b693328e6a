"""Micro-benchmark of the attention kernel alone through the C ABI (fitv2_debug_attention).
    python tools/attn_bench.py [R] [T] [dh] [reps]    (FITV2_B200_LIB / FITV2_ATTN select the variant)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fitv2_b200 import _lib

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dh = int(sys.argv[3]) if len(sys.argv) > 3 else 72
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
H = 16 if dh == 72 else 24
lib = _lib.load()
cfg = _lib.FitV2Config(H * dh, 1, H, dh, 3072 if dh == 72 else 6144, 288 if dh == 72 else 576, 16, 1001, 0, 1.0, 1.0)
h = C.c_void_p()
_lib.check(lib.fitv2_create(C.byref(cfg), C.byref(h)))
_lib.apply_env_options(h)
ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
_lib.check(lib.fitv2_set_workspace(h, C.c_void_p(ws.data_ptr()), ws.numel()))
g = torch.Generator().manual_seed(0)
ln = lambda x: torch.nn.functional.layer_norm(x, (dh,))
q, k = [ln(torch.randn(R, H, T, dh, generator=g)).bfloat16().cuda() for _ in range(2)]
tv = (T + 7) // 8 * 8
vt = torch.zeros(R, H, dh, tv, dtype=torch.bfloat16)
vt[..., :T] = torch.randn(R, H, dh, T, generator=g).bfloat16()
vt = vt.cuda()
mask = torch.ones(R, T)
if os.environ.get("ATTN_MASKED"):                      # mixed-aspect padded batch: lengths 200 / 256 / 192 / 200, pad id 0
    for r in range(R):
        mask[r, (200, 256, 192, 200)[r % 4]:] = 0
mask = mask.cuda()
out = torch.empty(R, T, H * dh, dtype=torch.bfloat16, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
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
print(f"attention R={R} T={T} dh={dh} lib={os.path.basename(_lib.LIB_PATH)} attn={os.environ.get('FITV2_ATTN', 'ws')}: "
      f"{us:.1f} us/launch  {fl / us / 1e6:.0f} TFLOP/s")
