"""GPU parity tests (run on a B200 with `-m gpu`): the CUDA path, called through the C ABI, against the CPU
oracle and the committed golden fixtures of the REAL reference.

Tolerances (stated, per north_star / SURVEY.md §8d):
  * per-NFE velocity: max_abs(ours - ref) / max_abs(ref) <= 1e-2 with 16-bit GEMM operands and fp32 accumulate;
  * fused CFG + Euler update, token masks, RoPE position tables: bit-exact / 1e-6;
  * latents after a trajectory prefix: <= 1e-2 (error of z is dsigma-weighted, far below the velocity error).
"""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fitv2_oracle as O          # checker only
from fitv2_b200 import FiT, EulerCFGSampler, euler_cfg_sample, make_grid, _lib

V_TOL = 1e-2
KW = dict(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora")
XL = dict(hidden_size=1152, num_heads=16, adaln_lora_dim=288)
B3 = dict(hidden_size=2304, num_heads=24, adaln_lora_dim=576)


def rel(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def _p(t):
    return C.c_void_p(t.data_ptr())


def build_model(depth, width=XL, operand="bf16", **extra):
    torch.manual_seed(0)
    m = FiT(**KW, depth=depth, operand_dtype=operand, **width, **extra).randomize_zero_init_(1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.cuda().eval(), sd, O.FiTConfig(depth=depth, **width, **extra)


def inputs(R, hp, wp, seed=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, hp * wp, 16, generator=g)
    t = torch.rand(R, generator=g)
    y = torch.randint(0, 1001, (R,), generator=g)
    return x, t, y, make_grid(R, hp, wp), torch.ones(R, hp * wp)


def run(m, *a):
    return m(*[v.cuda() for v in a]).cpu()


@pytest.fixture(scope="module")
def lib(built_lib):
    assert torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    return _lib.load()


@pytest.fixture(scope="module")
def xl2_padded(lib):
    return build_model(2, custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20, decouple=True, ori_max_pe_len=16)


# ------------------------------------------------------------------------------------------------
# elementwise: CFG + Euler (bit-exact) and channel-limited CFG combine
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N", [(2, 256), (32, 256), (3, 200), (1, 7), (32, 1024)])
def test_cfg_euler_bit_exact(lib, B, N):
    g = torch.Generator().manual_seed(B * 1000 + N)
    z, v2 = torch.randn(B, N, 16, generator=g), torch.randn(2 * B, N, 16, generator=g)
    sig = torch.linspace(0, 1, 251)
    vd = v2.cuda()
    for idx in (0, 17, 249):
        ref = O.cfg_euler_update(z, v2, 1.5, sig[idx], sig[idx + 1])
        zc = z.cuda().clone()
        _lib.check(lib.fitv2_cfg_euler(_p(zc), _p(vd), 1.5, float(sig[idx + 1] - sig[idx]), None, B, N, 16, None))
        assert torch.equal(zc.cpu(), ref)
        zc = z.cuda().clone()                      # step size from a device scalar (graph-replay form)
        ds = (sig[idx + 1] - sig[idx]).reshape(1).cuda()
        _lib.check(lib.fitv2_cfg_euler(_p(zc), _p(vd), 1.5, 0.0, _p(ds), B, N, 16, None))
        assert torch.equal(zc.cpu(), ref)


def test_cfg_combine_matches_forward_with_cfg_rule(lib):
    g = torch.Generator().manual_seed(0)
    out = torch.randn(6, 50, 16, generator=g)
    per = torch.tensor([1.5, 2.0, 4.0])
    ref = out.clone()
    c, u = out[:3, :, :12], out[3:, :, :12]
    gd = u + per.view(-1, 1, 1) * (c - u)
    ref[:3, :, :12], ref[3:, :, :12] = gd, gd
    oc = out.cuda()
    _lib.check(lib.fitv2_cfg_combine(_p(oc), _p(per.cuda()), 0.0, 3, 50, 16, 12, None))
    assert torch.equal(oc.cpu(), ref) and torch.equal(oc.cpu()[:, :, 12:], out[:, :, 12:])


# ------------------------------------------------------------------------------------------------
# components through the C ABI: tcgen05 GEMM and masked flash attention
# ------------------------------------------------------------------------------------------------
def _handle(lib, operand=0, D=1152, H=16, dh=72, Hm=3072, lora=288):
    cfg = _lib.FitV2Config(D, 1, H, dh, Hm, lora, 16, 1001, operand, 1.0, 1.0)
    h = C.c_void_p()
    _lib.check(lib.fitv2_create(C.byref(cfg), C.byref(h)))
    _lib.apply_env_options(h)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    _lib.check(lib.fitv2_set_workspace(h, _p(ws), ws.numel()))
    return h, ws


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (256, 288, 128, 144), (200, 288, 1152, 144), (128, 256, 96, 128),
                                      (4096, 1152, 3072, 192), (16384, 6144, 1152, 256), (1000, 3456, 1152, 144)])
@pytest.mark.parametrize("operand", [0, 1])
def test_gemm_matches_fp32_reference(lib, M, N, K, bn, operand):
    """C = A W^T + b with 16-bit operands and fp32 accumulation == fp32 matmul of the same (rounded) operands."""
    h, ws = _handle(lib, operand)
    dt = torch.float16 if operand else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g, device="cuda").to(dt)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(dt)
    bias = torch.randn(N, generator=g, device="cuda")
    ref = a.double() @ w.double().t() + bias.double()
    for epi in (3, 4):                           # 3: single-CTA tiles, 4: 2-CTA cluster with TMA-multicast weight tiles
        out = torch.full((M, N), float("nan"), device="cuda")
        _lib.check(lib.fitv2_debug_gemm(h, epi, _p(a), _p(w), _p(bias), _p(out), M, N, K, bn, None))
        assert rel(out, ref) < 2e-5, epi
    lib.fitv2_destroy(h)


def _attention_oracle(q, k, v, mask):
    """fit/model/modules.py:176-204 (checker, fp32 CPU)."""
    am = mask[:, None, None, :]
    am = (am == am.transpose(-2, -1))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=am)
    R, H, T, dh = q.shape
    return o.transpose(1, 2).reshape(R, T, H * dh) * (mask != 0).float()[..., None]


@pytest.mark.parametrize("R,T,dh,H,masked", [(1, 128, 72, 16, False), (2, 256, 72, 16, False), (2, 200, 72, 16, False), (2, 256, 72, 16, True),
                                             (1, 1024, 72, 16, False), (3, 100, 72, 16, True), (2, 256, 96, 24, True), (1, 1, 72, 16, False),
                                             (2, 77, 96, 24, True),
                                             # odd number of 128-row query tiles AND more work items than SMs: the second stream of the
                                             # pipelined kernel idles for whole items while its CTA keeps cycling the K / V rings
                                             (10, 128, 72, 16, False), (3, 777, 72, 16, True), (12, 384, 72, 16, False), (8, 100, 96, 24, True),
                                             (24, 256, 72, 16, True), (20, 200, 72, 16, True)])   # several masked items per CTA
@pytest.mark.parametrize("kernel", ["bounded", "online_max"])
def test_attention_matches_oracle(lib, R, T, dh, H, masked, kernel):
    """kernel 'bounded': attention_tm / attention_ws (constant logit bound; N(0,1) rows have the norm the QKV epilogue's
    LayerNorm produces); 'online_max': attention_general.cuh (running row maximum), fed logits far outside any bound."""
    h, ws = _handle(lib, 0, D=H * dh, H=H, dh=dh, Hm=3072 if dh == 72 else 6144, lora=288 if dh == 72 else 576)
    if kernel == "online_max":
        _lib.check(lib.fitv2_set_option(h, b"attn", 3))
    g = torch.Generator().manual_seed(R * 7 + T)
    q, k, v = [torch.randn(R, H, T, dh, generator=g).bfloat16() for _ in range(3)]
    if kernel == "online_max":                   # un-normalised q / k (FiTv1: no q/k norm): logits of several hundred
        q = (q.float() * torch.linspace(0.5, 6.0, T).view(1, 1, T, 1)).bfloat16()
        k = (k.float() * 3.0).bfloat16()
    mask = torch.ones(R, T)
    if masked:                                   # segment ids: 0 = padding, other ids = packed images
        mask[0, T - T // 4:] = 0
        if R > 1:
            mask[1, : T // 3] = 2
        for r in range(2, R):                    # larger batches: a different padding length / packing per sample
            if r % 3 == 0:
                mask[r, T - (r * 7) % (T // 2) - 1:] = 0
            elif r % 3 == 1:
                mask[r, : (r * 5) % (T // 2) + 1] = 3
    tv = (T + 7) // 8 * 8
    vt = torch.zeros(R, H, dh, tv, dtype=torch.bfloat16)
    vt[..., :T] = v.transpose(-1, -2)
    out = torch.full((R, T, H * dh), float("nan"), dtype=torch.bfloat16, device="cuda")
    qd, kd, vd, md = q.cuda(), k.cuda(), vt.cuda(), mask.cuda()          # keep the device tensors alive across the call
    _lib.check(lib.fitv2_debug_attention(h, _p(qd), _p(kd), _p(vd), _p(md), _p(out), R, T, None, None, None))
    torch.cuda.synchronize()
    ref = _attention_oracle(q.float(), k.float(), v.float(), mask)
    assert rel(out, ref) < 6e-3                  # P and the output are rounded to bf16 (2^-9)
    assert bool((out.float().cpu()[mask == 0] == 0).all())
    lib.fitv2_destroy(h)


def _key_length(r):
    """What seg_uniform_kernel must report for one row of segment ids: tokens when all ids are equal, n for n equal non-zero ids
    followed by zeros only (a padded sample), 0 for anything else (the kernels then compare ids per element)."""
    r = r.tolist()
    n = 0
    while n < len(r) and r[n] == r[0]:
        n += 1
    if n == len(r):
        return len(r)
    return n if r[0] != 0 and all(v == 0 for v in r[n:]) else 0


@pytest.mark.gpu
@pytest.mark.parametrize("T", [200, 256, 520])
def test_attention_key_length_paths(lib, T):
    """Padded samples take the key-length path of the bounded kernels (no per-element id compares): every boundary position of the
    valid length against the 64- / 128-key tiles, and the patterns that must NOT be classified as a prefix."""
    H, dh = 16, 72
    lens = [1, 63, 64, 65, 127, 128, 129, T - 1, T]
    rows = []
    for n in lens:
        r = torch.zeros(T); r[:n] = 1; rows.append(r)
    r = torch.zeros(T); r[:T // 2] = 7; rows.append(r)                       # non-unit id, still a prefix
    rows.append(torch.zeros(T))                                              # all padding: output zero
    r = torch.ones(T); r[T // 3: T // 2] = 0; rows.append(r)                 # hole in the middle: ids 1 | 0 | 1  -> compare path
    r = torch.ones(T); r[: T // 4] = 0; rows.append(r)                       # padding in FRONT                  -> compare path
    r = torch.ones(T); r[T // 2:] = 2; rows.append(r)                        # two packed images                 -> compare path
    r = torch.ones(T); r[T // 2: T - 5] = 2; r[T - 5:] = 0; rows.append(r)   # two images + padding              -> compare path
    mask = torch.stack(rows)
    R = mask.shape[0]
    h, ws = _handle(lib, 0, D=H * dh, H=H, dh=dh, Hm=3072, lora=288)
    g = torch.Generator().manual_seed(T)
    ln = lambda x: torch.nn.functional.layer_norm(x, (dh,))
    q, k = [ln(torch.randn(R, H, T, dh, generator=g)).bfloat16() for _ in range(2)]
    v = torch.randn(R, H, T, dh, generator=g).bfloat16()
    tv = (T + 7) // 8 * 8
    vt = torch.zeros(R, H, dh, tv, dtype=torch.bfloat16)
    vt[..., :T] = v.transpose(-1, -2)
    ref = _attention_oracle(q.float(), k.float(), v.float(), mask)
    for attn in (1, 2, 4, 3):                                                # P in TMEM, shared-memory P, one thread per row, online max
        _lib.check(lib.fitv2_set_option(h, b"attn", attn))
        out = torch.full((R, T, H * dh), float("nan"), dtype=torch.bfloat16, device="cuda")
        qd, kd, vd, md = q.cuda(), k.cuda(), vt.cuda(), mask.cuda()
        _lib.check(lib.fitv2_debug_attention(h, _p(qd), _p(kd), _p(vd), _p(md), _p(out), R, T, None, None, None))
        torch.cuda.synchronize()
        err = rel(out, ref)
        print(f"[key length] T={T} attn={attn}: {err:.2e}")
        assert err < 6e-3
        assert bool((out.float().cpu()[mask == 0] == 0).all())
    lib.fitv2_destroy(h)
    assert [_key_length(r) for r in mask][:len(lens)] == lens                # (the reference function of the tap test below)
    assert [_key_length(r) for r in mask][len(lens):] == [T // 2, T, 0, 0, 0, 0]


# ------------------------------------------------------------------------------------------------
# FiT.forward / forward_with_cfg against the reference goldens and the oracle
# ------------------------------------------------------------------------------------------------
def test_forward_golden_padded_mixed_aspect(xl2_padded, golden_dir):
    """BASELINE.json configs[2] flavour: dynntk + decouple, mixed 10x20 / 16x16 / 8x24 / 20x10 padded to 256."""
    m, sd, cfg = xl2_padded
    fx = torch.load(os.path.join(golden_dir, "xl_depth2_padded.pt"))
    a = (fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"])
    out = run(m, *a)
    assert rel(out, fx["out"]) < V_TOL
    assert bool((out[fx["mask"] == 0] == 0).all())                          # pad rows exactly zero
    ag = [v.cuda() for v in a]
    assert rel(m.forward_with_cfg(*ag, None, 1.5), fx["out_cfg"]) < V_TOL
    oc = m.forward_with_cfg(*ag, None, 4.0, scale_pow=2.0).cpu()
    assert rel(oc, fx["out_cfg_pow"]) < V_TOL
    assert torch.equal(oc[:2, :, :12], oc[2:, :, :12])                      # guided channels duplicated in both halves
    # token mask / RoPE position tables: bit-exact indices, 1e-6 tables
    cos, sin = O.rope_cos_sin(cfg, fx["grid"])
    assert float((m.debug_tap("rope_cos").cpu() - cos[..., 0::2].permute(2, 0, 1)).abs().max()) < 1e-6   # pair-major table
    assert float((m.debug_tap("rope_sin").cpu() - sin[..., 0::2].permute(2, 0, 1)).abs().max()) < 1e-6
    assert m.debug_tap("seg_uniform").cpu().tolist() == [_key_length(r) for r in fx["mask"]]


def test_forward_invariants(xl2_padded, golden_dir):
    m, sd, cfg = xl2_padded
    fx = torch.load(os.path.join(golden_dir, "xl_depth2_padded.pt"))
    a = (fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"])
    out = run(m, *a)
    assert torch.equal(run(m, *a), out)                                     # deterministic / idempotent
    out2 = run(m, fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"] * 2)
    assert torch.equal(out2, out * 2)                                       # raw-mask multiply (fit_model.py:230)
    assert torch.equal(run(m, fx["x"], fx["t"], fx["y"], fx["grid"], fx["mask"].bool()), out)
    solo = run(m, fx["x"][:1, :200], fx["t"][:1], fx["y"][:1], fx["grid"][:1, :, :200], fx["mask"][:1, :200])
    assert rel(solo, out[:1, :200]) < 2e-3                                  # padding invariance (different tiling, same math)
    ob = m(fx["x"].cuda().bfloat16(), *[v.cuda() for v in a[1:]])           # --mixed bf16 callers get bf16 back
    assert ob.dtype == torch.bfloat16 and rel(ob, out) < 3e-2
    # garbage in the padded positions of x must not leak into valid tokens
    xg = fx["x"].clone()
    xg[fx["mask"] == 0] = 1e3
    og = run(m, xg, *a[1:])
    valid = fx["mask"] != 0
    assert rel(og[valid], out[valid]) < 1e-6 and bool((og[~valid] == 0).all())


def test_forward_stage_taps_depth1(lib):
    """Per-stage parity on a depth-1 XL-width model: conditioning (fp32), QKV epilogue, attention, residual."""
    m, sd, cfg = build_model(1)
    x, t, y, grid, mask = inputs(4, 16, 16)
    taps = {}
    ref = O.forward(cfg, sd, x, t, y, grid, mask, taps=taps)
    out = run(m, x, t, y, grid, mask)
    assert rel(out, ref) < V_TOL
    # gmod comes from the tensor-pipe conditioning linears: tf32 weights (11 significant bits), fp32-exact activations
    assert rel(m.debug_tap("c"), taps["c"]) < 1e-5 and rel(m.debug_tap("gmod"), taps["global_adaln"]) < 1e-3
    for name, want in (("q", taps["q"]), ("k", taps["k"]), ("attn_out", taps["attn_out"])):
        assert rel(m.debug_tap(name), want) < 8e-3, name
    assert rel(m.debug_tap("vt")[..., :256], taps["v"].transpose(-1, -2)) < 8e-3
    assert rel(m.debug_tap("x_res"), taps["x1"]) < 5e-3
    assert 16 <= m.kernel_launches() <= 24       # 16 kernels + one finalize per split-K conditioning linear


def _tf32_rn(w):
    return ((w.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("width,R", [(XL, 4), (XL, 64), (XL, 130), (B3, 70)])
def test_conditioning_tensor_pipe_linears(lib, width, R):
    """csrc/cond_tc.cuh (tcgen05 kind::tf32, hi / lo split activations stacked along M, up to 3 row tiles incl. a ragged
    one): global / final adaLN and every block's LoRA modulation against the oracle evaluated with the SAME tf32-rounded
    adaLN weights -> only the fp32 summation order differs; against the unrounded fp32 oracle -> the stated tf32 weight
    rounding (2^-11 relative per weight)."""
    m, sd, cfg = build_model(2, width)
    x, t, y, grid, mask = inputs(R, 2, 3, seed=9)
    run(m, x, t, y, grid, mask)
    sd_r = dict(sd)
    for k in sd:
        if k.endswith("weight") and ("adaLN_modulation" in k):
            sd_r[k] = _tf32_rn(sd[k].float())
    for s_, tol in ((sd_r, 1e-5), (sd, 1e-3)):
        c = O.conditioning(cfg, s_, t, y)
        g = O._linear(torch.nn.functional.silu(c), s_, "global_adaLN_modulation.1")
        f = O._linear(torch.nn.functional.silu(c), s_, "final_layer.adaLN_modulation.1")
        mods = torch.stack([O.block_modulation(cfg, s_, c, i, g) for i in range(cfg.depth)])
        assert rel(m.debug_tap("gmod"), g) < tol and rel(m.debug_tap("fmod"), f) < tol
        assert rel(m.debug_tap("mod"), mods) < tol


@pytest.mark.parametrize("width,operand", [(B3, "bf16"), (XL, "fp16")])
def test_forward_other_widths_and_operands(lib, width, operand):
    m, sd, cfg = build_model(2 if width is XL else 1, width, operand)
    a = inputs(3, 10, 20, seed=5)                                            # 200 tokens: M tail + key tail
    assert rel(run(m, *a), O.forward(cfg, sd, *a)) < (2e-3 if operand == "fp16" else V_TOL)


def test_forward_swiglu_large_golden(lib, golden_dir):
    """use_swiglu_large=True (modules.py:248-249, SwiGLU hidden 4608) against the REAL reference's depth-1 forward."""
    fx = torch.load(os.path.join(golden_dir, "swiglu_large_xl_d1.pt"))
    m, sd, cfg = build_model(1, use_swiglu_large=True)
    R = fx["x"].shape[0]
    a = (fx["x"], fx["t"], fx["y"], make_grid(R, fx["hp"], fx["wp"]), torch.ones(R, fx["hp"] * fx["wp"]))
    out = run(m, *a)
    assert rel(out, fx["v_ref"]) < V_TOL and rel(out, O.forward(cfg, sd, *a)) < V_TOL
    big = inputs(40, 16, 16, seed=9)                                         # several token tiles per cluster, K = 4608 main loop
    assert rel(run(m, *big), O.forward(cfg, sd, *big)) < V_TOL


def test_xl_config1_full_depth_and_sampler(lib, golden_dir):
    """BASELINE.json configs[0]: XL/2 depth 36, CFG Euler steps at batch 2 against the REAL reference's outputs."""
    fx = torch.load(os.path.join(golden_dir, "xl_config1.pt"))
    m, sd, cfg = build_model(36)
    n, N = 2, 256
    grid, mask = make_grid(n, 16, 16), torch.ones(n, N)
    smp = EulerCFGSampler(m, fx["y"].cuda(), grid.cuda(), mask.cuda(), 250, 1.5)
    z1 = smp.sample(fx["z"].cuda(), first_steps=1).cpu()
    v2 = smp._v2.cpu()
    assert rel(v2, fx["v_step0"]) < V_TOL                                    # per-NFE velocity
    sig = torch.linspace(0, 1, 251)
    assert torch.equal(z1, O.cfg_euler_update(fx["z"], v2, 1.5, sig[0], sig[1]))   # fused update bit-exact on our velocity
    assert rel(z1, fx["z_step0"]) < 1e-4
    z2 = smp.sample(fx["z"].cuda(), first_steps=2).cpu()
    assert rel(z2, fx["z_step1"]) < 1e-4
    y2 = torch.cat([fx["y"], torch.full((n,), 1000)])
    vmid = run(m, torch.cat([fx["z"], fx["z"]]), torch.full((2 * n,), 0.5), y2, torch.cat([grid, grid]), torch.cat([mask, mask]))
    assert rel(vmid, fx["v_t05"]) < V_TOL
    # CUDA-graph replay of the step == eager launches, bit for bit
    zg = EulerCFGSampler(m, fx["y"].cuda(), grid.cuda(), mask.cuda(), 250, 1.5, use_cuda_graph=True).sample(fx["z"].cuda(), first_steps=2).cpu()
    assert torch.equal(zg, z2)
    # full-size batch (32 samples -> 64 rows, M = 16384): rows are independent, so the first two samples must
    # reproduce the batch-2 result (size-independent property at BASELINE.json's full size)
    g = torch.Generator().manual_seed(1)
    zb = torch.cat([fx["z"], torch.randn(30, N, 16, generator=g)])
    yb = torch.cat([fx["y"], torch.randint(0, 1000, (30,), generator=g)])
    zfull = euler_cfg_sample(m, zb.cuda(), yb.cuda(), make_grid(32, 16, 16).cuda(), torch.ones(32, N).cuda(), None, 250, 1.5, first_steps=2).cpu()
    # (1e-4, not bit-equal: the tile shapes follow M -- at 1 024 rows fc2 runs in the normal orientation, at 16 384 transposed with a
    # TMA reduce-add -- so the fp32 residual update associates differently and a few 16-bit operand roundings flip; measured 1.6e-5)
    d_full = rel(zfull[:2], z2)
    print(f"[config 1] batch-32 rows 0-1 vs batch-2 run after two steps: {d_full:.2e}")
    assert d_full < 1e-4 and bool(torch.isfinite(zfull).all())


def test_trajectory_prefix_matches_oracle(lib):
    """Latents after the first 6 Euler steps of a depth-4 XL-width model vs the oracle loop (error accumulates
    through the ODE; stated tolerance 1e-2 on max_abs(dz)/max_abs(z))."""
    m, sd, cfg = build_model(4)
    n, hp, wp = 2, 16, 16
    g = torch.Generator().manual_seed(21)
    z, y = torch.randn(n, hp * wp, 16, generator=g), torch.tensor([5, 321])
    grid, mask = make_grid(n, hp, wp), torch.ones(n, hp * wp)
    ref = O.euler_cfg_sample(cfg, sd, z, y, grid, mask, None, steps=250, cfg_scale=1.5, first_steps=6)
    got = euler_cfg_sample(m, z.cuda(), y.cuda(), grid.cuda(), mask.cuda(), None, 250, 1.5, first_steps=6).cpu()
    assert rel(got, ref) < 1e-2 and rel(got - z, ref - z) < 2e-2


def test_errors_are_loud(lib):
    m, sd, cfg = build_model(1)
    x, t, y, grid, mask = [v.cuda() for v in inputs(2, 4, 4)]
    with pytest.raises(ValueError):
        m(x, t[:1], y, grid, mask)
    with pytest.raises(_lib.FitV2Error):
        m(x.cpu(), t, y, grid, mask)
    h = C.c_void_p()
    cfgc = _lib.FitV2Config(1152, 1, 16, 72, 3072, 288, 16, 1001, 0, 1.0, 1.0)
    _lib.check(lib.fitv2_create(C.byref(cfgc), C.byref(h)))
    out = torch.empty(2, 16, 16, device="cuda")
    rc = lib.fitv2_forward(h, _p(x), 2, _p(t), _p(y), _p(grid), _p(mask.float()), _p(out), 2, 16, None)
    assert rc == -2 and b"not bound" in lib.fitv2_last_error()             # unbound weights: error, no fallback
    lib.fitv2_destroy(h)


# ------------------------------------------------------------------------------------------------
# after the trajectory: unpatchify + latent scaling, uint8 image pack (sample_fitv2_ddp.py:319-324)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,hp,wp", [(2, 16, 16), (3, 10, 20), (1, 32, 32), (2, 1, 3)])
def test_unpatchify_scale_bit_exact(lib, B, hp, wp):
    cfg = O.FiTConfig(hidden_size=1152, num_heads=16, depth=1)
    m = FiT(**KW, depth=1, **XL)
    g = torch.Generator().manual_seed(B * 100 + hp)
    z = torch.randn(B, hp * wp, 16, generator=g)
    ref = O.unpatchify(cfg, z, (hp * 2, wp * 2))
    assert torch.equal(m.unpatchify(z, (hp * 2, wp * 2)), ref)                      # CPU view path == oracle
    assert torch.equal(m.unpatchify(z.cuda(), (hp * 2, wp * 2)).cpu(), ref)          # CUDA kernel, bit-exact
    assert torch.equal(m.unpatchify(z.cuda(), (hp * 2, wp * 2), scaling_factor=0.18215).cpu(), ref / 0.18215)


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (1, 160, 320), (3, 8, 24)])
def test_pack_uint8_bit_exact(lib, B, H, W):
    from fitv2_b200 import pack_images_uint8
    g = torch.Generator().manual_seed(H + W)
    s = torch.randn(B, 3, H, W, generator=g) * 0.8
    s[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.99999994])
    ref = torch.clamp(127.5 * s.clamp(-1, 1) + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8).contiguous()
    assert torch.equal(pack_images_uint8(s.cuda()).cpu(), ref)


# ------------------------------------------------------------------------------------------------
# online_rope: per-sample dynamic NTK frequencies from `size` (fit_model.py:212-214, rope.py:234-274)
# ------------------------------------------------------------------------------------------------
def test_online_rope_forward_golden(lib, golden_dir):
    fx = torch.load(os.path.join(golden_dir, "xl_depth2_online.pt"))
    extra = dict(custom_freqs="ntk-aware", decouple=True, ori_max_pe_len=16, max_pe_len_h=16, max_pe_len_w=16, online_rope=True)
    m, sd, cfg = build_model(2, **extra)
    a = [fx[k].cuda() for k in ("x", "t", "y", "grid", "mask", "size")]
    out = m(*a).cpu()
    assert rel(out, fx["out"]) < V_TOL                                      # vs the REAL reference's online_rope forward
    assert bool((out[fx["mask"] == 0] == 0).all())
    cos, sin = O.rope_cos_sin_online(cfg, fx["grid"], fx["size"])
    assert float((m.debug_tap("rope_cos").cpu() - cos[..., 0::2].permute(2, 0, 1)).abs().max()) < 1e-6
    assert float((m.debug_tap("rope_sin").cpu() - sin[..., 0::2].permute(2, 0, 1)).abs().max()) < 1e-6
    # a different size tensor is picked up (different per-sample scale -> different output), the same one is cached
    size2 = fx["size"].clone(); size2[0, 0, 1] = 32
    assert rel(m(*a[:5], size2.cuda()).cpu(), out) > 1e-4
    with pytest.raises(ValueError):
        m(*a[:5], None)
    with pytest.raises(NotImplementedError):
        FiT(**KW, depth=1, **XL, custom_freqs="yarn", ori_max_pe_len=16, max_pe_len_h=16, max_pe_len_w=16, online_rope=True)
    with pytest.raises(NotImplementedError):
        FiT(**KW, depth=1, **XL, online_rope=True)                          # 'normal' has no online branch in the reference


def test_checkpoint_ingestion_repacks_kernel_weights(lib, tmp_path):
    """init_from_ckpt (eval_utils.py:12-71) after the model has already run: the packed kernel layouts follow the new weights."""
    from safetensors.torch import save_file
    from fitv2_b200 import init_from_ckpt
    m_a, _, _ = build_model(1)
    torch.manual_seed(123)
    m_b = FiT(**KW, depth=1, **XL).randomize_zero_init_(7).cuda().eval()
    a = inputs(2, 8, 8)
    out_a, out_b = run(m_a, *a), run(m_b, *a)
    assert rel(out_a, out_b) > 1e-2
    f = str(tmp_path / "model_ema.safetensors")
    save_file({f"_orig_mod.{k}": v.detach().cpu().contiguous() for k, v in m_b.state_dict().items()}, f)
    assert init_from_ckpt(m_a, f) == ([], [])
    assert torch.equal(run(m_a, *a), out_b)


@pytest.mark.parametrize("R,hp,wp", [(3, 10, 20), (5, 7, 9), (64, 16, 16)])
def test_transposed_residual_gemm_matches_normal(lib, monkeypatch, R, hp, wp):
    """EPI_RESID_T (weights as the M operand, 256-token-wide tiles, TMA reduce-add into the fp32 residual; the default for fc2
    at hidden 1152) against the normal orientation, forced both ways: token tail tiles (600 / 315 rows), samples that change
    inside a 16-token step (200 / 63 tokens per sample), and the headline shape."""
    a = inputs(R, hp, wp, seed=13)
    outs = []
    for mode in ("0", "1"):
        monkeypatch.setenv("FITV2_RESID_T", mode)
        m, sd, cfg = build_model(2)
        outs.append(run(m, *a))
        assert torch.equal(run(m, *a), outs[-1])                             # one add per element: deterministic
    # same products; the reduce-add rounds x + d once more, and a 1-ulp change of x can flip a 16-bit rounding of LN(x)
    assert rel(outs[1], outs[0]) < 5e-4
    if R <= 5:
        assert rel(outs[1], O.forward(cfg, sd, *a)) < V_TOL


@pytest.mark.parametrize("R,hp,wp", [(40, 8, 8), (20, 10, 20), (12, 12, 24), (7, 1, 1), (130, 2, 3), (50, 3, 7)])
def test_forward_many_rows_odd_tile_counts(lib, R, hp, wp):
    """Whole forward where every CTA of the attention kernel owns several work items and the number of 128-row query tiles
    is odd (64 / 288 tokens) or the last tile is ragged (200 tokens); 1 / 6 / 21 tokens: V^T rows padded to a multiple of 8
    (the workspace is poisoned with NaN by conftest, so a read of the never-written pad columns would show)."""
    m, sd, cfg = build_model(1)
    a = inputs(R, hp, wp, seed=11)
    out = run(m, *a)
    ref = O.forward(cfg, sd, *a)
    assert rel(out, ref) < V_TOL
    assert torch.equal(run(m, *a), out)


def test_plain_c_client_of_the_abi(built_lib, tmp_path):
    """The boundary is a C ABI: tools/cabi_client.c (gcc, no Python / torch / C++) creates a handle, binds weights, runs
    fitv2_forward twice on NaN-poisoned scratch memory and fitv2_cfg_euler, and checks finiteness, exact-zero pad rows and
    run-to-run identity itself (exit code)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cabi_client")
    libdir = os.path.join(root, "fitv2_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include", os.path.join(root, "tools", "cabi_client.c"),
                    "-o", exe, "-L", libdir, "-lfitv2_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "run-to-run identical 1" in r.stdout and "non-finite 0" in r.stdout


def test_sampler_without_cfg(lib):
    """cfg_scale = 1.0 (the script's default): no row duplication, z += dsigma * v (sample_fitv2_ddp.py:283-285,300-314)."""
    m, sd, cfg = build_model(2)
    n, hp, wp, steps = 3, 8, 8, 4
    g = torch.Generator().manual_seed(21)
    z0 = torch.randn(n, hp * wp, 16, generator=g)
    y = torch.randint(0, 1000, (n,), generator=g)
    grid, mask = make_grid(n, hp, wp), torch.ones(n, hp * wp)
    sig = torch.linspace(0, 1, 251)
    zr = z0.clone()
    for i in range(steps):
        v = O.forward(cfg, sd, zr, sig[i].expand(n), y, grid, mask)
        zr = zr + (sig[i + 1] - sig[i]) * v
    for graph in (False, True):
        z = euler_cfg_sample(m, z0.cuda(), y.cuda(), grid.cuda(), mask.cuda(), None, 250, 1.0, use_cuda_graph=graph, first_steps=steps).cpu()
        assert rel(z, zr) < 1e-4
    # the update itself is bit-exact given the velocity
    smp = EulerCFGSampler(m, y.cuda(), grid.cuda(), mask.cuda(), 250, 1.0)
    z1 = smp.sample(z0.cuda(), first_steps=1).cpu()
    assert torch.equal(z1, z0 + (sig[1] - sig[0]) * smp._v2.cpu())


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs 3 / 4 / 5 at their real depth, and the full 250-step trajectory, against outputs of the REAL
# reference (oracle/make_config_goldens.py).  Weights are regenerated from the seeds the fixtures were made with.
# ------------------------------------------------------------------------------------------------
def _check_weights(sd, fx):
    for k, (s_, a_) in fx["weight_checksum"].items():
        v = sd[k].double()
        assert abs(float(v.sum()) - s_) <= 1e-9 * max(1.0, abs(a_)) and abs(float(v.abs().sum()) - a_) <= 1e-9 * max(1.0, a_), k


def _cfg_nfe(m, z, y, grid, mask, t):
    n = z.shape[0]
    y2 = torch.cat([y, torch.full((n,), 1000)])
    return run(m, torch.cat([z, z]), torch.full((2 * n,), float(t)), y2, torch.cat([grid, grid]), torch.cat([mask, mask]))


DYN = dict(custom_freqs="ntk-aware", decouple=True, ori_max_pe_len=16)


def test_config3_xl_depth36_dynntk_golden(lib, golden_dir):
    """BASELINE.json configs[2]: XL/2 depth 36, 160x320 (10x20 tokens) ntk-aware decoupled, and the mixed-aspect batch padded to
    256 at full depth (M tails, key tails, segment masks through 36 blocks)."""
    fx = torch.load(os.path.join(golden_dir, "cfg3_xl_d36_dynntk.pt"))
    m, sd, cfg = build_model(36, max_pe_len_h=10, max_pe_len_w=20, **DYN)
    _check_weights(sd, fx)
    v = _cfg_nfe(m, fx["z"], fx["y"], make_grid(2, 10, 20), torch.ones(2, 200), 0.5)
    e1 = rel(v, fx["v_t05"])
    vp = run(m, fx["x_pad"], fx["t_pad"], fx["y_pad"], fx["grid_pad"], fx["mask_pad"])
    e2 = rel(vp, fx["v_pad"])
    print(f"[parity] config 3 (XL/2 d36 10x20 dynntk): CFG NFE {e1:.2e}, mixed padded batch {e2:.2e} (tolerance {V_TOL:.0e})")
    assert e1 < V_TOL and e2 < V_TOL
    assert bool((vp[fx["mask_pad"] == 0] == 0).all())


def test_config4_3b_depth40_golden(lib, golden_dir):
    """BASELINE.json configs[3]: FiTv2-3B/2 (hidden 2304, 24 heads of 96, depth 40), 256 tokens, one CFG NFE at batch 2 and the
    first Euler step: attention_ws at head_dim 96, the warp-pair LayerNorm, the 192-wide QKV tile, 128-wide adaLN-up tiles."""
    fx = torch.load(os.path.join(golden_dir, "cfg4_3b_d40.pt"))
    m, sd, cfg = build_model(40, B3)
    _check_weights(sd, fx)
    assert sum(v.numel() for v in sd.values()) == fx["n_params"]
    del sd
    grid, mask = make_grid(2, 16, 16), torch.ones(2, 256)
    e0 = rel(_cfg_nfe(m, fx["z"], fx["y"], grid, mask, 0.0), fx["v_t0"])
    e5 = rel(_cfg_nfe(m, fx["z"], fx["y"], grid, mask, 0.5), fx["v_t05"])
    z1 = EulerCFGSampler(m, fx["y"].cuda(), grid.cuda(), mask.cuda(), 250, 1.5).sample(fx["z"].cuda(), first_steps=1).cpu()
    ez = rel(z1, fx["z_step0"])
    print(f"[parity] config 4 (3B/2 d40 256 tokens): velocity t=0 {e0:.2e}, t=0.5 {e5:.2e}, z after step 0 {ez:.2e}")
    assert e0 < V_TOL and e5 < V_TOL and ez < 1e-4


def test_config5_xl_1024_tokens_golden(lib, golden_dir):
    """BASELINE.json configs[4]: XL/2 depth 36 at 512x512 -> 32x32 = 1024 tokens, ntk-aware decoupled (scale 2 on both axes):
    the long-sequence attention path (attention_ws, 8 key tiles) through 36 blocks."""
    fx = torch.load(os.path.join(golden_dir, "cfg5_xl_1024.pt"))
    m, sd, cfg = build_model(36, max_pe_len_h=32, max_pe_len_w=32, **DYN)
    _check_weights(sd, fx)
    grid, mask = make_grid(1, 32, 32), torch.ones(1, 1024)
    e0 = rel(_cfg_nfe(m, fx["z"], fx["y"], grid, mask, 0.0), fx["v_t0"])
    e5 = rel(_cfg_nfe(m, fx["z"], fx["y"], grid, mask, 0.5), fx["v_t05"])
    print(f"[parity] config 5 (XL/2 d36 1024 tokens): velocity t=0 {e0:.2e}, t=0.5 {e5:.2e}")
    assert e0 < V_TOL and e5 < V_TOL


Z_TOL = 5e-3          # stated tolerance on the final latents: max|z - z_ref| / max|z_ref| after the full 250-step trajectory


@pytest.mark.parametrize("operand", ["bf16", "fp16"])
def test_full_250_step_trajectory_final_latents(lib, golden_dir, operand):
    """north_star: "final latents after the full trajectory within a stated tolerance".  The script's complete 250-step CFG 1.5
    Euler loop at batch 1, XL/2 depth 36, against the latents the REAL reference model produced (326 s of CPU in the build
    container, tests/golden/xl_traj250.pt); checkpoints after 1 / 50 / 125 / 250 steps show how the error accumulates."""
    fx = torch.load(os.path.join(golden_dir, "xl_traj250.pt"))
    m, sd, cfg = build_model(36, operand=operand)
    _check_weights(sd, fx)
    grid, mask = make_grid(1, 16, 16), torch.ones(1, 256)
    smp = EulerCFGSampler(m, fx["y"].cuda(), grid.cuda(), mask.cuda(), 250, 1.5)
    errs = {}
    for steps in (1, 50, 125, 250):
        z = smp.sample(fx["z"].cuda(), first_steps=steps).cpu()
        ref = fx[f"z_step{steps}"]
        errs[steps] = (rel(z, ref), float((z - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))
        assert bool(torch.isfinite(z).all())
    print(f"[parity] 250-step trajectory, {operand} operands: max-rel / rms-rel after 1, 50, 125, 250 steps: "
          + ", ".join(f"{k}: {a:.2e}/{b:.2e}" for k, (a, b) in errs.items()) + f" (tolerance {Z_TOL:.0e})")
    assert errs[250][0] < Z_TOL and errs[1][0] < 1e-4
    zg = EulerCFGSampler(m, fx["y"].cuda(), grid.cuda(), mask.cuda(), 250, 1.5, use_cuda_graph=True).sample(fx["z"].cuda()).cpu()
    assert torch.equal(zg, z)                                              # graph replay of all 250 steps == eager launches


def test_tf32_conditioning_cost_at_depth36(lib, golden_dir, monkeypatch):
    """The adaLN linears run on the tensor pipe with tf32-rounded weights (csrc/cond_tc.cuh); FITV2_COND=simt keeps fp32 FMA and
    unrounded weights.  Same golden, both modes: what the rounding costs at full depth."""
    fx = torch.load(os.path.join(golden_dir, "xl_config1.pt"))
    grid, mask = make_grid(2, 16, 16), torch.ones(2, 256)
    errs = {}
    for mode in ("tc", "simt"):
        if mode == "simt":
            monkeypatch.setenv("FITV2_COND", "simt")
        m, sd, cfg = build_model(36)
        errs[mode] = (rel(_cfg_nfe(m, fx["z"], fx["y"], grid, mask, 0.0), fx["v_step0"]),
                      rel(_cfg_nfe(m, fx["z"], fx["y"], grid, mask, 0.5), fx["v_t05"]))
        del m
    print(f"[parity] XL/2 d36 velocity error vs the reference, t=0 / t=0.5: tf32 tensor-pipe conditioning {errs['tc'][0]:.2e} / {errs['tc'][1]:.2e}, "
          f"fp32 SIMT conditioning {errs['simt'][0]:.2e} / {errs['simt'][1]:.2e}")
    assert max(errs["tc"]) < V_TOL and max(errs["simt"]) < V_TOL


def test_final_layer_tensor_pipe_matches_fp32_kernel(lib, monkeypatch):
    """Final layer as LayerNorm+modulate (fp16 operand) + skinny tcgen05 GEMM against the fused fp32 SIMT kernel (FITV2_FINAL_TC=0)."""
    a = inputs(5, 10, 20, seed=17)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FITV2_FINAL_TC", mode)
        m, sd, cfg = build_model(2)
        outs[mode] = run(m, *a)
    ref = O.forward(cfg, sd, *a)
    e_tc, e_f32 = rel(outs["1"], ref), rel(outs["0"], ref)
    print(f"[parity] final layer: tensor pipe (fp16 operands) {e_tc:.2e}, fp32 SIMT {e_f32:.2e} vs oracle; between them {rel(outs['1'], outs['0']):.2e}")
    assert e_tc < V_TOL and e_f32 < V_TOL and rel(outs["1"], outs["0"]) < 2e-3


# ------------------------------------------------------------------------------------------------
# model variants (norms.py kinds) against the REAL reference (oracle/make_variant_goldens.py)
# ------------------------------------------------------------------------------------------------
def test_fitv1_config_golden(lib, golden_dir):
    """configs/fit/config_fit_xl.yaml params at depth 2: learn_sigma (32 output channels), (B, C, N) tensors, adaLN 'normal',
    no q / k norm (unbounded logits -> online-max attention), use_swiglu_large."""
    fx = torch.load(os.path.join(golden_dir, "fitv1_xl_d2.pt"))
    torch.manual_seed(0)
    m = FiT(**fx["kwargs"]).randomize_zero_init_(1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    _check_weights(sd, fx)
    m = m.cuda().eval()
    a = [fx[k].cuda() for k in ("x", "t", "y", "grid", "mask")]
    out = m(*a).cpu()
    assert out.shape == (4, 32, 256)
    e = rel(out, fx["out"])
    oc = m.forward_with_cfg(*a, None, 1.5).cpu()
    ec = rel(oc, fx["out_cfg"])
    print(f"[parity] FiTv1 config (depth 2): forward {e:.2e}, forward_with_cfg {ec:.2e}")
    assert e < V_TOL and ec < V_TOL
    assert bool((out.transpose(1, 2)[fx["mask"] == 0] == 0).all())
    assert torch.equal(oc[:2, :12], oc[2:, :12])                             # guided channels (dim 1 here) duplicated in both halves
    assert torch.equal(m.unpatchify(fx["unpatchify_in"].cuda(), (20, 40)).cpu(), fx["unpatchify_out"])
    assert torch.equal(m.unpatchify(fx["unpatchify_in"], (20, 40)), fx["unpatchify_out"])


def test_norm_variants_golden(lib, golden_dir):
    """norm_type 'rmsnorm' / 'w_layernorm', q/k norms rmsnorm / w_layernorm / none (fit/model/norms.py:35-77)."""
    for case in torch.load(os.path.join(golden_dir, "norm_variants_xl_d1.pt")):
        torch.manual_seed(0)
        m = FiT(**{**KW, **XL, **case["extra"]}, depth=1).randomize_zero_init_(1)
        m.load_state_dict(O.perturb_norm_weights({k: v.detach().clone() for k, v in m.state_dict().items()}, 2))
        _check_weights(m.state_dict(), case)
        m = m.cuda().eval()
        out = m(*[case[k].cuda() for k in ("x", "t", "y", "grid", "mask")]).cpu()
        e = rel(out, case["out"])
        print(f"[parity] norm variant {case['name']} {case['extra']}: {e:.2e}")
        assert e < V_TOL and bool((out[case["mask"] == 0] == 0).all())


def test_ctor_variant_goldens(lib, golden_dir):
    """The remaining constructor switches of fit_model.py:25-65 against outputs of the REAL reference classes
    (oracle/make_ctor_variant_goldens.py): the class defaults (timm Mlp with tanh-GELU, learn_sigma, (B, C, N), adaLN 'normal', no
    q / k norm), bias-free qkv / ffn with rel_pos_embed=None, add_rel_pe_to_v under 'XPOS', adaln_type 'swiglu'.  Every case is
    run and reported (with the conditioning / q / k / v taps against the oracle) before anything is asserted."""
    failures = []
    for c in torch.load(os.path.join(golden_dir, "ctor_variants_xl.pt")):
        name = c["name"]
        try:
            _run_ctor_variant(c, failures)
        except Exception as exc:                                            # keep going: one log shows every case
            print(f"[parity] ctor variant {name}: raised {type(exc).__name__}: {exc}")
            failures.append(name)
    assert not failures, failures


def _run_ctor_variant(c, failures):
    name = c["name"]
    torch.manual_seed(0)
    m = FiT(**c["kwargs"]).randomize_zero_init_(1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    _check_weights(sd, c)
    cfg = O.FiTConfig(**c["oracle_kwargs"])
    m = m.cuda().eval()
    a = [c[k].cuda() for k in ("x", "t", "y", "grid", "mask")]
    out = m(*a).cpu()
    # taps of this forward against the oracle: conditioning vector, per-block modulation, final modulation, q / k / v of block 0
    cvec = O.conditioning(cfg, sd, c["t"], c["y"])
    mods = torch.stack([O.block_modulation(cfg, sd, cvec, i, 0.0 if cfg.adaln_type != "lora" else
                                           O._linear(torch.nn.functional.silu(cvec), sd, "global_adaLN_modulation.1")) for i in range(cfg.depth)])
    info = dict(c=rel(m.debug_tap("c"), cvec), mod=rel(m.debug_tap("mod"), mods))
    if cfg.depth == 1:
        taps = {}
        O.forward(cfg, sd, c["x"], c["t"], c["y"], c["grid"], c["mask"], taps=taps)
        valid = (c["mask"] != 0)[:, None, :, None]
        info["q"] = rel(m.debug_tap("q").float().cpu() * valid, taps["q"] * valid)
        info["k"] = rel(m.debug_tap("k").float().cpu() * valid, taps["k"] * valid)
        info["v"] = rel(m.debug_tap("vt").float().cpu()[..., :256].transpose(2, 3) * valid, taps["v"] * valid)
    e = rel(out, c["out"])
    oc = m.forward_with_cfg(*a, None, 1.5).cpu()
    ec = rel(oc, c["out_cfg"])
    pad_ok = bool(((out if cfg.use_sit else out.transpose(1, 2))[c["mask"] == 0] == 0).all())
    print(f"[parity] ctor variant {name}: forward {e:.2e}, forward_with_cfg {ec:.2e}, pad rows zero {pad_ok}, taps "
          + ", ".join(f"{k} {v:.2e}" for k, v in info.items()))
    if not cfg.use_swiglu:
        # the two GELU epilogues (slab-staged EPI_GELU, the default, and the plain epilogue with act_gelu) are the same arithmetic
        m.set_option("gelu_epi", 1)
        same = torch.equal(m(*a).cpu(), out)
        m.set_option("gelu_epi", 0)
        print(f"[parity] ctor variant {name}: plain GELU epilogue bit-equal to the staged one: {same}")
        if not same:
            failures.append(name + " (gelu_epi)")
    if not (e < V_TOL and ec < V_TOL and pad_ok and out.shape == c["out"].shape):
        failures.append(name)


def test_ctor_variants_combined_at_3b_width_fp16(lib):
    """The switches of test_ctor_variant_goldens all at once, away from the golden shape: 3B/2 width (head_dim 96, Mlp hidden
    9216, SwiGLU-modulation hidden 1728 / 1152), fp16 operands, 3 rows x 200 tokens (ragged tiles) with a padded sample, CUDA
    graph capturable -- against the oracle (which the goldens pin to the real classes for each switch)."""
    kw = dict(hidden_size=2304, depth=1, num_heads=24, learn_sigma=False, use_sit=True, use_swiglu=False, qkv_bias=False,
              q_norm="layernorm", k_norm="layernorm", adaln_type="swiglu", add_rel_pe_to_v=True)
    torch.manual_seed(0)
    m = FiT(**kw, operand_dtype="fp16").randomize_zero_init_(1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    cfg = O.FiTConfig(hidden_size=2304, depth=1, num_heads=24, adaln_lora_dim=0, use_swiglu=False, qkv_bias=False,
                      adaln_type="swiglu", add_rel_pe_to_v=True)
    assert list(sd.keys()) == list(O.reference_init_state_dict(cfg, 0).keys())
    x, t, y, grid, mask = inputs(3, 10, 20)
    mask[2, 150:] = 0
    x[2, 150:] = 0
    ref = O.forward(cfg, sd, x, t, y, grid, mask)
    m = m.cuda().eval()
    out = run(m, x, t, y, grid, mask)
    e = rel(out, ref)
    print(f"[parity] constructor switches combined at 3B width, fp16 operands, 3 x 200 tokens: {e:.2e}")
    assert e < 2e-3 and bool((out[mask == 0] == 0).all())
    assert torch.equal(run(m, x, t, y, grid, mask), out)


# ------------------------------------------------------------------------------------------------
# the boundary as the reference script uses it (sample_fitv2_ddp.py:172-213): wrappers + jit trace
# ------------------------------------------------------------------------------------------------
def test_reference_script_wrappers_and_jit_trace(lib):
    """The script wraps the model in ModelWrapper / SamplingWrapper modules and traces them (fvcore FlopCountAnalysis ->
    torch.jit.trace) on rank 0 before sampling.  FiT.forward is the registered operator fitv2_b200::forward, so the trace
    records a node whose output depends on the latent input; the traced module reproduces the eager result."""
    m, sd, cfg = build_model(2)
    n, hp, wp = 2, 16, 16
    y = torch.tensor([3, 977]).cuda()
    y2 = torch.cat([y, torch.full((n,), 1000, device="cuda")])
    grid2 = make_grid(2 * n, hp, wp).cuda()
    mask2 = torch.ones(2 * n, hp * wp).cuda()
    size2 = torch.tensor([hp, wp]).repeat(2 * n, 1).unsqueeze(1).cuda()

    class ModelWrapper(torch.nn.Module):                                  # sample_fitv2_ddp.py:172-181
        def __init__(self, model):
            super().__init__()
            self.model = model

        def forward(self, z):
            t = torch.zeros(z.shape[0], device=z.device)
            return self.model(z, t, y=y2, grid=grid2, mask=mask2, size=size2)

    class SamplingWrapper(torch.nn.Module):                               # sample_fitv2_ddp.py:183-196: the Euler loop over the model
        def __init__(self, model, steps=3):
            super().__init__()
            self.model, self.steps = model, steps

        def forward(self, z):
            sig = torch.linspace(0, 1, 251)
            for i in range(self.steps):
                z_in = torch.cat([z, z], 0)
                v2 = self.model(z_in, sig[i].expand(z_in.shape[0]).to(z.device), y=y2, grid=grid2, mask=mask2, size=size2)
                c, u = v2.chunk(2, dim=0)
                z = z + (sig[i + 1] - sig[i]) * (u + 1.5 * (c - u))
            return z

    g = torch.Generator().manual_seed(2)
    z = torch.randn(n, hp * wp, 16, generator=g).cuda()
    z2 = torch.cat([z, z])
    eager = ModelWrapper(m)(z2)
    traced = torch.jit.trace(ModelWrapper(m), (z2,), check_trace=False)
    assert "fitv2_b200::forward" in str(traced.inlined_graph)
    assert torch.equal(traced(z2), eager)
    other = torch.randn(2 * n, hp * wp, 16, generator=g).cuda()
    assert torch.equal(traced(other), ModelWrapper(m)(other)) and not torch.equal(traced(other), eager)   # data dependence on z
    ts = torch.jit.trace(SamplingWrapper(m), (z,), check_trace=False)
    assert str(ts.inlined_graph).count("fitv2_b200::forward") == 3
    ref = euler_cfg_sample(m, z, y, make_grid(n, hp, wp).cuda(), torch.ones(n, hp * wp).cuda(), None, 250, 1.5, first_steps=3)
    assert rel(ts(z), ref) < 1e-6                                         # torch ops vs the fused update: same numbers up to fp32 contraction
    # meta / fake kernel: shape inference without running
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        fz = torch.empty(4, 256, 16, device="cuda")
        fo = torch.ops.fitv2_b200.forward(fz, torch.empty(4, device="cuda"), torch.empty(4, dtype=torch.long, device="cuda"),
                                          torch.empty(4, 2, 256, dtype=torch.long, device="cuda"), torch.empty(4, 256, device="cuda"), id(m), 4)
        assert fo.shape == (4, 256, 16) and fo.dtype == torch.float32


def test_out_of_range_label_is_reported(lib):
    """modules.py:101-106: the reference's embedding lookup raises on a label outside the table.  Here the kernel reads row 0
    instead of out-of-bounds memory and the handle reports the error at the next synchronisation point / call."""
    m, sd, cfg = build_model(1)
    x, t, y, grid, mask = [v.cuda() for v in inputs(2, 4, 4)]
    m(x, t, y, grid, mask)
    torch.cuda.synchronize()
    m.check_device_errors()                                                # clean
    bad = y.clone(); bad[1] = 1001
    out = m(x, t, bad, grid, mask)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    with pytest.raises(_lib.FitV2Error, match="class label"):
        m.check_device_errors()
    m.check_device_errors()                                                # the flag is sticky until reported, then cleared
    m(x, t, bad, grid, mask); torch.cuda.synchronize()
    with pytest.raises(_lib.FitV2Error, match="class label"):              # ... and fitv2_forward reports it on entry
        m(x, t, y, grid, mask)
    with pytest.raises(ValueError):                                        # CFG needs the null-class row
        torch.manual_seed(0)
        m0 = FiT(**KW, depth=1, **XL, class_dropout_prob=0.0).cuda()
        EulerCFGSampler(m0, y[:1], grid[:1], mask[:1], 10, 1.5)


def test_sampler_with_online_rope(lib, golden_dir):
    """the Euler loop of an online_rope model must use the per-sample frequencies of `size` (hr_xl / 3B-hr configs
    ship online_rope: true), and a NEW size tensor with other values must not hit a stale cache."""
    fx = torch.load(os.path.join(golden_dir, "xl_depth2_online.pt"))
    extra = dict(custom_freqs="ntk-aware", decouple=True, ori_max_pe_len=16, max_pe_len_h=16, max_pe_len_w=16, online_rope=True)
    m, sd, cfg = build_model(2, **extra)
    x, t, y, grid, mask, size = [fx[k] for k in ("x", "t", "y", "grid", "mask", "size")]
    n = x.shape[0]
    with pytest.raises(ValueError):
        EulerCFGSampler(m, y.cuda(), grid.cuda(), mask.cuda(), 250, 1.5)   # size is required, not silently ignored
    z = euler_cfg_sample(m, x.cuda(), y.cuda(), grid.cuda(), mask.cuda(), size.cuda(), 250, 1.5, first_steps=2).cpu()
    ref = O.euler_cfg_sample(cfg, sd, x, y, grid, mask, size, steps=250, cfg_scale=1.5, first_steps=2)
    assert rel(z, ref) < 1e-4
    zg = euler_cfg_sample(m, x.cuda(), y.cuda(), grid.cuda(), mask.cuda(), size.cuda(), 250, 1.5, use_cuda_graph=True, first_steps=2).cpu()
    assert torch.equal(zg, z)
    # value-keyed cache: fresh tensors of the same shape (same allocator block) with different contents
    a = [v.cuda() for v in (x, t, y, grid, mask)]
    outs = []
    for hw in ((16, 16), (32, 24), (16, 16)):
        s_new = torch.tensor(hw).repeat(n, 1).unsqueeze(1).cuda()
        outs.append(m(*a, s_new).cpu())
        del s_new
    assert torch.equal(outs[0], outs[2]) and rel(outs[1], outs[0]) > 1e-4
    for o_, hw in zip(outs[:2], ((16, 16), (32, 24))):
        assert rel(o_, O.forward(cfg, sd, x, t, y, grid, mask, torch.tensor(hw).repeat(n, 1).unsqueeze(1))) < V_TOL


def test_options_are_per_handle(lib, monkeypatch):
    """fitv2_set_option replaces the process-wide getenv latches: two handles in one process keep their own switches."""
    a = inputs(3, 10, 20, seed=5)
    m0, sd, cfg = build_model(1)
    monkeypatch.setenv("FITV2_QKV", "2"); monkeypatch.setenv("FITV2_RESID_T", "1"); monkeypatch.setenv("FITV2_ATTN", "ws")
    m1, _, _ = build_model(1)
    o1 = run(m1, *a)                                                                      # the handle is created (and reads the environment) at the first call
    monkeypatch.delenv("FITV2_QKV"); monkeypatch.delenv("FITV2_RESID_T"); monkeypatch.delenv("FITV2_ATTN")
    o0 = run(m0, *a)
    ref = O.forward(cfg, sd, *a)
    assert rel(o0, ref) < V_TOL and rel(o1, ref) < V_TOL and not torch.equal(o0, o1)     # different kernels, same function
    assert torch.equal(run(m0, *a), o0) and torch.equal(run(m1, *a), o1)                  # ... and each handle keeps its choice
    m0.set_option("attn", 3)                                                              # online-max kernel on a bounded model
    assert rel(run(m0, *a), ref) < V_TOL
    with pytest.raises(_lib.FitV2Error, match="unknown option"):
        m0.set_option("bogus", 1)


@pytest.mark.parametrize("width,depth,R,hp,wp,extra", [(XL, 2, 5, 10, 20, {}), (XL, 1, 3, 7, 9, {}), (B3, 1, 3, 10, 20, {}), (XL, 1, 2, 32, 32, {}),
                                                       (XL, 1, 66, 3, 5, {}), (XL, 1, 4, 16, 16, dict(q_norm=None, k_norm=None))])
def test_workspace_bounds_canary(lib, width, depth, R, hp, wp, extra):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer_unavailable.txt), so out-of-bounds writes are hunted
    with guard bands: every workspace buffer is followed by 4 KB of padding, the whole workspace (and a frame around the output
    tensor) is filled with a canary, and after the forward every byte of padding must still hold it -- for ragged shapes that
    exercise the M / key / V^T tails, the 1024-token path, 3B width, > 64 rows, and the online-max attention."""
    torch.manual_seed(0)
    m = FiT(**{**KW, **width, **extra}, depth=depth).randomize_zero_init_(1).cuda().eval()
    a = [v.cuda() for v in inputs(R, hp, wp, seed=23)]
    m.set_option("ws_guard", 4096)
    ref = m(*a).cpu()                                                        # sizes the workspace
    ws = m._workspace
    ws.fill_(0xA5)
    # the workspace belongs to the handle between calls (it caches the fp16 copy of the final weight there): re-announce it so that
    # the handle re-derives its cached contents after the fill above
    _lib.check(lib.fitv2_set_workspace(m._handle, _p(ws), ws.numel()))
    N = hp * wp
    frame = torch.full((R * N * 16 + 2048,), float("nan"), device="cuda")    # output in the middle of a NaN frame
    out = frame[1024:1024 + R * N * 16].view(R, N, 16)
    m._run(a[0].float().contiguous(), *a[1:], rows=R, out=out)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)
    assert bool(torch.isnan(frame[:1024]).all()) and bool(torch.isnan(frame[-1024:]).all())
    spans = m.workspace_layout()
    assert len(spans) >= 20
    host = ws.cpu()
    for i, (off, size) in enumerate(spans):
        end = spans[i + 1][0] if i + 1 < len(spans) else off + size + 4096
        pad = host[off + size:end]
        assert pad.numel() >= 4096 and bool((pad == 0xA5).all()), f"buffer {i} (offset {off}, {size} bytes): padding overwritten"
