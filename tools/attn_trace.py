"""Timeline of CTA 0 of the attention kernel (needs a lib built with -DFITV2_ATTN_TRACE; timing experiments only).
    FITV2_B200_LIB=$PWD/ab/lib_TRACE.so python tools/attn_trace.py [T]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fitv2_b200 import _lib

R, T, dh, H = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 72, 16
lib = _lib.load()
cfg = _lib.FitV2Config(H * dh, 1, H, dh, 3072, 288, 16, 1001, 0, 1.0, 1.0)
h = C.c_void_p()
_lib.check(lib.fitv2_create(C.byref(cfg), C.byref(h)))
_lib.apply_env_options(h)
ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
_lib.check(lib.fitv2_set_workspace(h, C.c_void_p(ws.data_ptr()), ws.numel()))
g = torch.Generator().manual_seed(0)
ln = lambda x: torch.nn.functional.layer_norm(x, (dh,))
q, k = [ln(torch.randn(R, H, T, dh, generator=g)).bfloat16().cuda() for _ in range(2)]
vt = torch.randn(R, H, dh, T, generator=g).bfloat16().cuda()
mask = torch.ones(R, T, device="cuda")
out = torch.empty(R, T, H * dh, dtype=torch.bfloat16, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
call = lambda: _lib.check(lib.fitv2_debug_attention(h, p(q), p(k), p(vt), p(mask), p(out), R, T, None, None, None))
for _ in range(3):
    call()
torch.cuda.synchronize()
lib.fitv2_debug_attn_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
_lib.check(lib.fitv2_debug_attn_trace(None, None, 1))
call()
tr = np.zeros((20, 512, 2), dtype=np.uint64)
n = np.zeros(20, dtype=np.uint32)
_lib.check(lib.fitv2_debug_attn_trace(tr.ctypes.data, n.ctypes.data, 0))
ev = []
for role in range(20):
    for i in range(min(int(n[role]), 512)):
        ev.append((int(tr[role, i, 1]), role, int(tr[role, i, 0])))
ev.sort()
t0 = ev[0][0]
names = {0: "TMA", 1: "MMAa", 2: "MMAb"}
roles = set(sys.argv[2].split(",")) if len(sys.argv) > 2 and sys.argv[2] else None
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 260
shown = 0
for t, role, tag in ev:
    nm = names.get(role, f"w{role}")
    if roles and nm not in roles:
        continue
    print(f"{t - t0:8d}  {nm:4s} {tag}")
    shown += 1
    if shown >= limit:
        break
print("total span", ev[-1][0] - t0, "events", len(ev))
