"""Bring-up probe for the GPU box: runs every kernel family in isolation (one subprocess per stage so a
trap in one stage cannot poison the others) and writes logs / small tensors under gpurun_out/.

    python tools/gpu_probe.py all            # every stage
    python tools/gpu_probe.py gemm_small     # one stage
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

STAGES = ["env", "cfg_euler", "gemm_small", "gemm_shapes", "attention", "forward_depth1", "forward_padded",
          "forward_fp16", "forward_3b", "timing"]


def _p(t):
    return C.c_void_p(t.data_ptr())


def _model(depth=1, operand="bf16", width="xl", **extra):
    import torch
    from fitv2_b200 import FiT
    kw = dict(hidden_size=1152, num_heads=16, adaln_lora_dim=288) if width == "xl" else \
        dict(hidden_size=2304, num_heads=24, adaln_lora_dim=576)
    torch.manual_seed(0)
    m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm",
            adaln_type="lora", depth=depth, operand_dtype=operand, **kw, **extra).randomize_zero_init_(1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.cuda().eval(), sd, dict(depth=depth, **kw, **extra)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def stage_env():
    import torch
    print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total",
                          "--format=csv"], capture_output=True, text=True).stdout)
    from fitv2_b200 import _lib
    print(_lib.load().fitv2_version().decode())
    print("cpu_count", os.cpu_count())


def stage_cfg_euler():
    import torch
    from fitv2_b200 import _lib
    from oracle import fitv2_oracle as O
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (B, N, Cc) in [(2, 256, 16), (32, 256, 16), (3, 200, 16), (1, 7, 16)]:
        z = torch.randn(B, N, Cc, generator=g)
        v2 = torch.randn(2 * B, N, Cc, generator=g)
        sig = torch.linspace(0, 1, 251)
        for idx in (0, 100, 249):
            ref = O.cfg_euler_update(z, v2, 1.5, sig[idx], sig[idx + 1])
            zc, vc = z.cuda().clone(), v2.cuda()
            ds = float(sig[idx + 1] - sig[idx])
            _lib.check(lib.fitv2_cfg_euler(_p(zc), _p(vc), 1.5, ds, None, B, N, Cc, None))
            torch.cuda.synchronize()
            print(f"cfg_euler B={B} N={N} idx={idx} bit_exact={torch.equal(zc.cpu(), ref)} maxdiff={float((zc.cpu()-ref).abs().max()):.3e}")
    out = torch.randn(4, 50, 16, generator=g)
    ref = out.clone()
    c, u = ref[:2, :, :12], ref[2:, :, :12]
    gd = u + 1.5 * (c - u)
    ref[:2, :, :12], ref[2:, :, :12] = gd, gd
    oc = out.cuda()
    _lib.check(lib.fitv2_cfg_combine(_p(oc), None, 1.5, 2, 50, 16, 12, None))
    print("cfg_combine bit_exact", torch.equal(oc.cpu(), ref))


def _gemm_case(lib, h, M, N, K, bn, op, save=None, seed=0):
    import torch
    from fitv2_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, generator=g, device="cuda")).to(op)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(op)
    bias = torch.randn(N, generator=g, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.fitv2_debug_gemm(h, 3, _p(a), _p(w), _p(bias), _p(out), M, N, K, bn, None), "debug_gemm")
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    err = rel(out, ref)
    nan = int(torch.isnan(out).sum())
    print(f"gemm M={M} N={N} K={K} bn={bn} {op}: max-rel-err {err:.3e} nans {nan}", flush=True)
    if save and (err > 1e-3 or nan):
        torch.save(dict(out=out.cpu(), ref=ref.cpu()), os.path.join(OUT, save))
        d = (out - ref).abs()
        print("   err by 32-row block:", [f"{float(x):.2e}" for x in d.view(M // 32, 32, N).amax(dim=(1, 2))[:8]])
        print("   err by 16-col block:", [f"{float(x):.2e}" for x in d.view(M, N // 16, 16).amax(dim=(0, 2))[:16]])
    return err


def _handle(operand="bf16"):
    import torch
    from fitv2_b200 import _lib
    lib = _lib.load()
    cfg = _lib.FitV2Config(1152, 1, 16, 72, 3072, 288, 16, 1001, 1 if operand == "fp16" else 0, 1.0, 1.0)
    h = C.c_void_p()
    _lib.check(lib.fitv2_create(C.byref(cfg), C.byref(h)))
    _lib.apply_env_options(h)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    _lib.check(lib.fitv2_set_workspace(h, _p(ws), ws.numel()))
    return lib, h, ws


def stage_gemm_small():
    import torch
    lib, h, ws = _handle()
    for bn in (128, 144, 192, 256):
        _gemm_case(lib, h, 128, bn, 64, bn, torch.bfloat16, save=f"gemm_small_bn{bn}_k64.pt")
        _gemm_case(lib, h, 256, 2 * bn, 128, bn, torch.bfloat16, save=f"gemm_small_bn{bn}.pt")
    _gemm_case(lib, h, 200, 288, 1152, 144, torch.bfloat16)        # M tail
    _gemm_case(lib, h, 128, 256, 96, 128, torch.bfloat16)          # K tail (zero fill)
    lib2, h2, ws2 = _handle("fp16")
    _gemm_case(lib2, h2, 256, 288, 128, 144, torch.float16)


def stage_gemm_shapes():
    import torch
    lib, h, ws = _handle()
    M = 16384
    for (N, K, bn) in [(3456, 1152, 144), (1152, 1152, 144), (1152, 1152, 128), (6144, 1152, 256), (1152, 3072, 144),
                       (1152, 3072, 192), (2304, 2304, 256), (6912, 2304, 192)]:
        err = _gemm_case(lib, h, M, N, K, bn, torch.bfloat16)
        # timing (plain epilogue, fp32 output) — indicative only
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
        w = torch.randn(N, K, generator=g, device="cuda").bfloat16()
        bias = torch.zeros(N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        from fitv2_b200 import _lib
        for _ in range(3):
            lib.fitv2_debug_gemm(h, 3, _p(a), _p(w), _p(bias), _p(out), M, N, K, bn, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.fitv2_debug_gemm(h, 3, _p(a), _p(w), _p(bias), _p(out), M, N, K, bn, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10):
            torch.matmul(a, w.t())
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        fl = 2.0 * M * N * K
        print(f"   time {ms*1e3:.1f} us = {fl/ms/1e9:.0f} TFLOP/s   (torch.matmul {ms_t*1e3:.1f} us = {fl/ms_t/1e9:.0f} TFLOP/s)", flush=True)


def _attn_ref(q, k, v, mask):
    import torch
    import torch.nn.functional as F
    am = mask[:, None, None, :]
    am = (am == am.transpose(-2, -1))
    o = F.scaled_dot_product_attention(q.float(), k.float(), v.float(), attn_mask=am)
    keep = (mask != 0).float()
    R, H, T, dh = q.shape
    return o.transpose(1, 2).reshape(R, T, H * dh) * keep[..., None]


def stage_attention():
    import torch
    from fitv2_b200 import _lib
    lib, h, ws = _handle()
    g = torch.Generator(device="cuda").manual_seed(0)
    H, dh = 16, 72
    for (R, T, masked) in [(1, 128, False), (2, 256, False), (2, 200, False), (2, 256, True), (1, 1024, False), (3, 100, True)]:
        q = torch.randn(R, H, T, dh, generator=g, device="cuda").bfloat16()
        k = torch.randn(R, H, T, dh, generator=g, device="cuda").bfloat16()
        v = torch.randn(R, H, T, dh, generator=g, device="cuda").bfloat16()
        tv = (T + 7) // 8 * 8
        vt = torch.zeros(R, H, dh, tv, device="cuda", dtype=torch.bfloat16)
        vt[..., :T] = v.transpose(-1, -2)
        mask = torch.ones(R, T, device="cuda")
        if masked:
            mask[0, T - 37:] = 0
            if R > 1:
                mask[1, : T // 3] = 2
        out = torch.full((R, T, H * dh), float("nan"), device="cuda", dtype=torch.bfloat16)
        dbg_s = torch.full((128, 128), float("nan"), device="cuda")
        dbg_o = torch.full((128, 80), float("nan"), device="cuda")
        _lib.check(lib.fitv2_debug_attention(h, _p(q), _p(k), _p(vt), _p(mask), _p(out), R, T, _p(dbg_s), _p(dbg_o), None))
        torch.cuda.synchronize()
        ref = _attn_ref(q, k, v, mask)
        n = min(T, 128)
        s_ref = q[0, 0, :n].float() @ k[0, 0, :n].float().t()
        print(f"attention R={R} T={T} masked={masked}: out max-rel-err {rel(out, ref):.3e} nans {int(torch.isnan(out.float()).sum())} "
              f"| S tile err {rel(dbg_s[:n, :n], s_ref):.3e}", flush=True)
        if rel(dbg_s[:n, :n], s_ref) > 1e-2:
            s_main = q[0, 0, :n, :64].float() @ k[0, 0, :n, :64].float().t()
            print(f"    S vs main-panel-only reference: {rel(dbg_s[:n, :n], s_main):.3e}")
            torch.save(dict(s=dbg_s.cpu(), s_ref=s_ref.cpu(), s_main=s_main.cpu()), os.path.join(OUT, f"attn_S_T{T}.pt"))
    # 3B head_dim 96
    from fitv2_b200 import _lib as L2
    cfg = L2.FitV2Config(2304, 1, 24, 96, 6144, 576, 16, 1001, 0, 1.0, 1.0)
    h3 = C.c_void_p()
    L2.check(lib.fitv2_create(C.byref(cfg), C.byref(h3)))
    L2.check(lib.fitv2_set_workspace(h3, _p(ws), ws.numel()))
    R, H, T, dh = 2, 24, 256, 96
    q = torch.randn(R, H, T, dh, generator=g, device="cuda").bfloat16()
    k = torch.randn(R, H, T, dh, generator=g, device="cuda").bfloat16()
    v = torch.randn(R, H, T, dh, generator=g, device="cuda").bfloat16()
    vt = v.transpose(-1, -2).contiguous()
    mask = torch.ones(R, T, device="cuda")
    out = torch.full((R, T, H * dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    dbg_s = torch.full((128, 128), float("nan"), device="cuda")
    L2.check(lib.fitv2_debug_attention(h3, _p(q), _p(k), _p(vt), _p(mask), _p(out), R, T, _p(dbg_s), None, None))
    torch.cuda.synchronize()
    s_ref = q[0, 0, :128].float() @ k[0, 0, :128].float().t()
    print(f"attention dh=96 R={R} T={T}: out max-rel-err {rel(out, _attn_ref(q, k, v, mask)):.3e} | S tile err {rel(dbg_s, s_ref):.3e}")


def _forward_compare(m, sd, kw, x, t, y, grid, mask, taps=True, quant=None):
    import torch
    from oracle import fitv2_oracle as O
    cfg = O.FiTConfig(**kw)
    tp = {} if taps else None
    ref = O.forward(cfg, {k: v.float() for k, v in sd.items()}, x, t, y, grid, mask, taps=tp)
    out = m(x.cuda(), t.cuda(), y.cuda(), grid.cuda(), mask.cuda())
    torch.cuda.synchronize()
    print(f"  OUT max-rel-err {rel(out.cpu(), ref):.3e}   (|ref|max {float(ref.abs().max()):.3f}) nans {int(torch.isnan(out).sum())}", flush=True)
    pad = (mask == 0)
    if pad.any():
        print(f"  pad rows exactly zero: {bool((out.cpu()[pad] == 0).all())}")
    if taps:
        R, N = x.shape[0], x.shape[1]
        H, dh = cfg.num_heads, cfg.head_dim
        for name, refv in [("c", tp["c"]), ("gmod", tp["global_adaln"]), ("q", tp["q"]), ("k", tp["k"]),
                           ("attn_out", tp["attn_out"]), ("x_res", tp["x1"])]:
            got = m.debug_tap(name).float().cpu()
            print(f"  tap {name:9s} max-rel-err {rel(got, refv):.3e}", flush=True)
        vt = m.debug_tap("vt").float().cpu()[..., :N]
        print(f"  tap vt        max-rel-err {rel(vt, tp['v'].transpose(-1, -2)):.3e}")
        rc = m.debug_tap("rope_cos").cpu()
        cos, sin = O.rope_cos_sin(cfg, grid)
        print(f"  tap rope_cos  max-abs-err {float((rc - cos[..., 0::2].permute(2, 0, 1)).abs().max()):.3e}")
    return out, ref


def _inputs(R, hp, wp, seed=3):
    import torch
    from fitv2_b200 import make_grid
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, hp * wp, 16, generator=g)
    t = torch.rand(R, generator=g)
    y = torch.randint(0, 1001, (R,), generator=g)
    return x, t, y, make_grid(R, hp, wp), torch.ones(R, hp * wp)


def stage_forward_depth1():
    m, sd, kw = _model(depth=1)
    print("depth-1 XL width, 4 rows x 256 tokens")
    _forward_compare(m, sd, kw, *_inputs(4, 16, 16))
    print("launches", m.kernel_launches())


def stage_forward_padded():
    import torch
    m, sd, kw = _model(depth=2, custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20, decouple=True, ori_max_pe_len=16)
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "xl_depth2_padded.pt"))
    print("golden fixture xl_depth2_padded (mixed 10x20/16x16/8x24/20x10 padded to 256)")
    out = m(fx["x"].cuda(), fx["t"].cuda(), fx["y"].cuda(), fx["grid"].cuda(), fx["mask"].cuda()).cpu()
    print(f"  vs REFERENCE golden: max-rel-err {rel(out, fx['out']):.3e}; pad rows zero {bool((out[fx['mask'] == 0] == 0).all())}")
    oc = m.forward_with_cfg(fx["x"].cuda(), fx["t"].cuda(), fx["y"].cuda(), fx["grid"].cuda(), fx["mask"].cuda(), None, 1.5).cpu()
    print(f"  forward_with_cfg vs golden: {rel(oc, fx['out_cfg']):.3e}")
    oc = m.forward_with_cfg(fx["x"].cuda(), fx["t"].cuda(), fx["y"].cuda(), fx["grid"].cuda(), fx["mask"].cuda(), None, 4.0, scale_pow=2.0).cpu()
    print(f"  forward_with_cfg(scale_pow=2) vs golden: {rel(oc, fx['out_cfg_pow']):.3e}")
    print("N=200 (10x20) unpadded, 4 rows")
    _forward_compare(m, sd, kw, *_inputs(4, 10, 20), taps=False)


def stage_forward_fp16():
    m, sd, kw = _model(depth=2, operand="fp16")
    print("depth-2 XL width fp16 operands")
    _forward_compare(m, sd, kw, *_inputs(4, 16, 16), taps=False)
    m, sd, kw = _model(depth=2, operand="bf16")
    print("depth-2 XL width bf16 operands")
    _forward_compare(m, sd, kw, *_inputs(4, 16, 16), taps=False)


def stage_forward_3b():
    m, sd, kw = _model(depth=1, width="3b")
    print("depth-1 3B width (D=2304, 24 heads of 96)")
    _forward_compare(m, sd, kw, *_inputs(2, 16, 16))


def stage_timing():
    import torch
    for operand in ("bf16",):
        m, sd, kw = _model(depth=36, operand=operand)
        R = 64
        x, t, y, grid, mask = [v.cuda() for v in _inputs(R, 16, 16)]
        for _ in range(2):
            m(x, t, y, grid, mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            m(x, t, y, grid, mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 304.62e9 * R
        print(f"XL/2 depth 36, 64 rows x 256 tokens, {operand}: {ms:.2f} ms / NFE = {fl/ms/1e9:.0f} TFLOP/s "
              f"-> {32/(250*ms/1e3):.2f} img/s/GPU", flush=True)
        # per-NFE parity at full depth vs oracle on 4 rows
        xs, ts, ys, gs, ms_ = _inputs(4, 16, 16, seed=9)
        _forward_compare(m, sd, kw, xs, ts, ys, gs, ms_, taps=False)


def main():
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage != "all":
        globals()["stage_" + stage]()
        return
    summary = {}
    for s in STAGES:
        t0 = time.time()
        log = os.path.join(OUT, f"probe_{s}.log")
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), s], stdout=f, stderr=subprocess.STDOUT,
                                   timeout=float(os.environ.get("PROBE_STAGE_TIMEOUT", "240")))
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        summary[s] = dict(rc=rc, seconds=round(time.time() - t0, 1))
        print(f"=== {s}: rc={rc} ({summary[s]['seconds']} s)", flush=True)
        with open(log) as f:
            txt = f.read()
        print(txt[-3000:], flush=True)
    with open(os.path.join(OUT, "probe_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
