"""Transport Sampler: the oracle against the golden vectors captured from the REAL reference
(`oracle/make_transport_golden.py`), the host logic of `fitv2_b200.transport`, and - on a GPU - the fused update
kernels through the C ABI, bit-exact against the reference's trajectories."""
import os

import pytest
import torch as th

from oracle import transport_oracle as T
from fitv2_b200 import Sampler, create_transport
from fitv2_b200 import transport as P


@pytest.fixture(scope="module")
def kat(golden_dir):
    return th.load(os.path.join(golden_dir, "transport_kat.pt"))


def same(a, b):
    return th.equal(th.nan_to_num(a, nan=123.0), th.nan_to_num(b, nan=123.0)) and th.equal(a.isnan(), b.isnan())


# ------------------------------------------------------------------------------------------------ CPU: oracle vs reference
def test_oracle_pointwise_matches_reference(kat):
    pw = kat["pointwise"]
    assert th.equal(T.score_from_velocity(pw["v"], pw["x"], pw["t"]), pw["score"])
    for form, ref in pw["diffusion"].items():
        got = T.compute_diffusion(pw["x"], pw["t"], form=form, norm=0.7)
        assert float((got - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), form     # cos / sin may differ in the last ulp across CPUs
        if form in ("constant", "SBDM", "sigma", "linear"):
            assert th.equal(got, ref), form


def test_oracle_sde_trajectories_match_reference(kat):
    init = kat["sde"]["init"]
    for c in kat["sde"]["cases"]:
        th.manual_seed(c["seed"])
        noises = [th.randn(init.size()) for _ in range(c["num_steps"] - 1)]
        xs = T.sample_sde(T.toy_velocity_model, init, sampling_method=c["method"], diffusion_form=c["form"], diffusion_norm=c["norm"],
                          last_step=c["last_step"], last_step_size=c["last_step_size"], num_steps=c["num_steps"], noises=noises)
        assert len(xs) == c["num_steps"]
        if c["form"] == "decreasing":
            assert float((xs[-1] - c["final"]).abs().max()) < 1e-5
        else:
            assert same(xs[-1], c["final"]) and same(xs[c["num_steps"] // 2], c["mid"]), c


def test_oracle_ode_fixed_grid(kat):
    o = kat["ode"]
    assert th.equal(T.sample_ode(T.toy_velocity_model, o["init"], sampling_method="euler", num_steps=o["num_steps"])[-1], o["euler_final"])
    assert th.equal(T.sample_ode(T.toy_velocity_model, o["init"], sampling_method="midpoint", num_steps=o["num_steps"])[-1], o["midpoint_final"])
    with pytest.raises(NotImplementedError):
        T.sample_ode(T.toy_velocity_model, o["init"], sampling_method="dopri8")


def test_oracle_dopri5_meets_its_tolerance():
    """The adaptive dopri5 restatement (torchdiffeq is not installed: unpinned) integrates to within its tolerance of a fine
    fixed-grid solution, at the requested grid times as well (dense output), forwards and on a decreasing grid."""
    x0 = th.linspace(-1.0, 1.0, 32, dtype=th.float64).reshape(2, 4, 4)
    model = lambda x, t, **kw: T.toy_velocity_model(x, t.to(x.dtype))
    th.set_default_dtype(th.float64)
    try:
        fine = T.sample_ode(model, x0, sampling_method="rk4", num_steps=2049)
        for rtol, bound in ((1e-3, 2e-4), (1e-6, 1e-5)):       # the dense output is 4th order: its error, not the step error, shows at 1e-6
            st = {}
            ys = T.sample_ode(model, x0, sampling_method="dopri5", num_steps=9, rtol=rtol, atol=1e-8, stats=st)
            assert len(ys) == 9 and st["nfe"] == 2 + 6 * st["steps"]
            for i in (4, 8):
                assert float((ys[i] - fine[i * 256]).abs().max()) < bound, (rtol, i, st)
        back = T.sample_ode(model, fine[-1], sampling_method="dopri5", num_steps=5, rtol=1e-6, atol=1e-8, reverse=True)
        ref_back = T.sample_ode(model, fine[-1], sampling_method="rk4", num_steps=2049, reverse=True)[-1]
        assert float((back[-1] - ref_back).abs().max()) < 1e-5
    finally:
        th.set_default_dtype(th.float32)


def test_dopri5_controller_host_logic_matches_oracle():
    """fitv2_b200.transport._dopri5 (the product's step-size controller) with torch stand-ins for its two kernels follows the
    oracle restatement step for step, and a solution that blows up raises instead of spinning."""
    from fitv2_b200 import transport as P
    init = th.linspace(-1.0, 1.0, 64).reshape(2, 2, 16) * 2.7
    for reverse, scale, kw in ((False, 1.0, {}), (False, 1.0, dict(rtol=1e-5, atol=1e-7)), (True, 0.2, {})):
        ts = th.linspace(1, 0, 7) if reverse else th.linspace(0, 1, 7)

        def f(tval, state):
            tv = th.ones(state.size(0)) * tval
            if reverse:
                tv = th.ones_like(tv) * (1 - tv)
            return T.toy_velocity_model(state, tv)
        st, st2 = {}, {}
        ys = P._dopri5(f, init * scale, ts, kw.get("rtol", 1e-3), kw.get("atol", 1e-6), st, ops=P._torch_ops(), max_steps=500)
        ref = T.sample_ode(T.toy_velocity_model, init * scale, sampling_method="dopri5", num_steps=7, reverse=reverse, stats=st2, **kw)
        # same number of steps / rejections; the error estimate of the tiny first step is rounding noise (~1e-8), so the next step
        # sizes differ in the third digit and the two solutions by a fraction of the local tolerance rtol * |y|
        assert {k: st[k] for k in ("nfe", "steps", "rejected")} == {k: st2[k] for k in ("nfe", "steps", "rejected")}
        bound = max(3e-4, 2.0 * kw.get("rtol", 1e-3) * float(ref[-1].abs().max()))
        assert max(float((a - b).abs().max()) for a, b in zip(ys, ref)) < bound
    with pytest.raises(RuntimeError, match="no progress"):                   # backwards from |y| = 2.7 the cubic term blows up
        ts = th.linspace(1, 0, 7)
        P._dopri5(lambda tv, s_: T.toy_velocity_model(s_, th.ones(s_.size(0)) * (1 - tv)), init, ts, 1e-3, 1e-6, {}, ops=P._torch_ops(), max_steps=300)


def test_oracle_rk_tableaus_have_their_published_order():
    """torchdiffeq is not installed, so the heun2 / heun3 / rk4 restatements cannot be pinned against it; what CAN be checked
    is the defining property of each Butcher tableau: the global error of y' = f(t, y) on a fixed grid falls as dt^order
    (euler 1, midpoint / heun2 2, heun3 3, rk4 4).  fp64 so that rounding does not mask the rate."""
    x0 = th.linspace(-1.0, 1.0, 32, dtype=th.float64).reshape(2, 4, 4)
    model = lambda x, t, **kw: T.toy_velocity_model(x, t.to(x.dtype))
    fine = T.sample_ode(model, x0, sampling_method="rk4", num_steps=2049)[-1]
    for method, order in (("euler", 1), ("midpoint", 2), ("heun2", 2), ("heun3", 3), ("rk4", 4)):
        errs = []
        for n in (9, 17, 33):
            ts_back = th.get_default_dtype()
            th.set_default_dtype(th.float64)                                 # linspace / ones inside the oracle
            try:
                errs.append(float((T.sample_ode(model, x0, sampling_method=method, num_steps=n)[-1] - fine).abs().max()))
            finally:
                th.set_default_dtype(ts_back)
        r1, r2 = errs[0] / errs[1], errs[1] / errs[2]
        assert 2 ** order * 0.7 < r1 < 2 ** order * 1.45 and 2 ** order * 0.7 < r2 < 2 ** order * 1.45, (method, errs)


# ------------------------------------------------------------------------------------------------ CPU: host logic of the product
def test_create_transport_and_intervals():
    tr = create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None, snr_type="lognorm")
    assert (tr.train_eps, tr.sample_eps) == (0, 0)
    for form in ("SBDM", "sigma"):
        for lss in (0.0, 0.04):
            assert tr.check_interval(0, 0, diffusion_form=form, sde=True, eval=True, last_step_size=lss) == \
                T.check_interval(0, diffusion_form=form, sde=True, last_step_size=lss)
    assert tr.check_interval(0, 0, sde=False, eval=True, reverse=True) == (1, 0)
    with pytest.raises(NotImplementedError):
        create_transport(path_type="VP")
    with pytest.raises(NotImplementedError):
        create_transport(prediction="noise")
    with pytest.raises(ValueError):
        create_transport(snr_type="bogus")
    s = Sampler(tr)
    with pytest.raises(NotImplementedError):
        s.sample_ode(sampling_method="dopri8")
    with pytest.raises(NotImplementedError):
        s.sample_sde(sampling_method="Midpoint")
    with pytest.raises(NotImplementedError):
        s.sample_sde(diffusion_form="bogus")
    with pytest.raises(NotImplementedError):
        s.sample_ode_likelihood()
    fn = s.sample_sde(diffusion_form="sigma", num_steps=4)
    with pytest.raises(P._lib.FitV2Error):                                   # no CPU path
        fn(th.zeros(2, 4, 16), T.toy_velocity_model)


def test_step_coefficients_follow_reference_expressions():
    """The host-side scalar table (fitv2_b200.transport._coef_row) == the reference's (B,1,1) tensors."""
    x = th.zeros(1, 1, 1)
    for form in P._DIFFUSION_FORMS:
        for tv in (0.0, 0.125, 0.5, 0.96):
            t = th.tensor([tv])
            dt = th.tensor(0.004)
            row = P._coef_row(t[0], dt, form, 0.9)
            d = T.compute_diffusion(x, t, form=form, norm=0.9).reshape(-1)[0]
            s, ds = T.sigma_t(t[0])
            var = s ** 2 - t[0] * ds * s
            assert same(row[1], var) and same(row[2], d) and same(row[4], th.sqrt(2 * d)) and same(row[5], th.sqrt(dt)), (form, tv)
            assert same(row[0], t[0]) and same(row[3], dt)


# ------------------------------------------------------------------------------------------------ GPU: kernels through the C ABI
def _gpu_toy(x, t, **kw):
    return T.toy_velocity_model(x, t)           # +, -, * only: the CUDA elementwise kernels round exactly like the CPU ones


@pytest.mark.gpu
def test_sde_kernels_bit_exact_against_reference_trajectories(kat, built_lib):
    s = Sampler(create_transport())
    init = kat["sde"]["init"]
    for c in kat["sde"]["cases"]:
        fn = s.sample_sde(sampling_method=c["method"], diffusion_form=c["form"], diffusion_norm=c["norm"], last_step=c["last_step"],
                          last_step_size=c["last_step_size"], num_steps=c["num_steps"], noise="reference")
        th.manual_seed(c["seed"])
        xs = fn(init.cuda(), _gpu_toy)
        assert len(xs) == c["num_steps"]
        if c["form"] == "decreasing":                                          # cos on the host: last-ulp differences allowed
            assert float((xs[-1].cpu() - c["final"]).abs().max()) < 1e-5
        else:
            assert same(xs[-1].cpu(), c["final"]) and same(xs[c["num_steps"] // 2].cpu(), c["mid"]), c


@pytest.mark.gpu
def test_ode_kernels_bit_exact(kat, built_lib):
    s = Sampler(create_transport())
    o = kat["ode"]
    for m in ("euler", "midpoint"):
        ys = s.sample_ode(sampling_method=m, num_steps=o["num_steps"])(o["init"].cuda(), _gpu_toy)
        assert len(ys) == o["num_steps"] and th.equal(ys[-1].cpu(), o[f"{m}_final"])
    for m in ("heun2", "heun3", "rk4"):                                       # fitv2_rk_stage: every stage bit-equal to the restated expressions
        ys = s.sample_ode(sampling_method=m, num_steps=o["num_steps"])(o["init"].cuda(), _gpu_toy)
        ref = T.sample_ode(T.toy_velocity_model, o["init"], sampling_method=m, num_steps=o["num_steps"])
        assert len(ys) == o["num_steps"] and all(th.equal(a.cpu(), b) for a, b in zip(ys, ref)), m
    # adaptive dopri5 (the reference default): same controller as the oracle restatement, same step sequence; the accepted states
    # agree to fp32 rounding, the interior grid points come from the dense-output polynomial whose coefficients cancel terms of
    # ~30 |y| in fp32 (18 y0 + 14 y1 - 32 y_mid): two evaluation orders differ by up to ~1e-4 there
    for kw in (dict(), dict(rtol=1e-5, atol=1e-7), dict(reverse=True)):
        st = {}
        init = o["init"] * (0.2 if kw.get("reverse") else 1.0)                # (backwards the toy ODE blows up in finite time from |y| ~ 2.7)
        ref = T.sample_ode(T.toy_velocity_model, init, sampling_method="dopri5", num_steps=7, stats=st, **kw)
        ys = s.sample_ode(sampling_method="dopri5", num_steps=7, **kw)(init.cuda(), _gpu_toy)
        assert len(ys) == 7 and all(s.last_ode_stats[k] == st[k] for k in ("nfe", "steps", "rejected")), (s.last_ode_stats, st)
        bound = max(3e-4, 2.0 * kw.get("rtol", 1e-3) * float(ref[-1].abs().max()))
        assert max(float((a.cpu() - b).abs().max()) for a, b in zip(ys, ref)) < bound
    yr = s.sample_ode(sampling_method="euler", num_steps=5, reverse=True)(o["init"].cuda(), _gpu_toy)
    assert th.equal(yr[-1].cpu(), T.sample_ode(T.toy_velocity_model, o["init"], sampling_method="euler", num_steps=5, reverse=True)[-1])


@pytest.mark.gpu
def test_sde_with_fit_forward_with_cfg(built_lib):
    """The upstream route: Sampler.sample_sde driving FiT.forward_with_cfg (sample_fitv2_ddp.py:138-146, 281)."""
    from oracle import fitv2_oracle as O
    from fitv2_b200 import FiT, make_grid
    kw = dict(hidden_size=1152, depth=2, num_heads=16, adaln_lora_dim=288)
    th.manual_seed(0)
    m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora", **kw).randomize_zero_init_(1)
    sd = {k: v.detach().clone().float() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    cfg = O.FiTConfig(**kw)
    n, hp, wp = 2, 8, 8
    g = th.Generator().manual_seed(3)
    z = th.randn(n, hp * wp, 16, generator=g)
    z2 = th.cat([z, z], 0)
    y2 = th.cat([th.tensor([5, 900]), th.full((n,), 1000)])
    grid2, mask2 = make_grid(2 * n, hp, wp), th.ones(2 * n, hp * wp)
    size2 = th.tensor((hp, wp)).repeat(2 * n, 1)[:, None, :]
    steps = 5
    th.manual_seed(77)
    noises = [th.randn(z2.size()) for _ in range(steps - 1)]
    ref_model = lambda x, t, **k: O.forward_with_cfg(cfg, sd, x, t, y2, grid2, mask2, size2, 1.5, 0.0)
    ref = T.sample_sde(ref_model, z2, sampling_method="Euler", diffusion_form="sigma", last_step="Mean", last_step_size=0.04,
                       num_steps=steps, noises=noises)
    fn = Sampler(create_transport()).sample_sde(sampling_method="Euler", diffusion_form="sigma", last_step="Mean", last_step_size=0.04,
                                                num_steps=steps, noise="reference")
    th.manual_seed(77)
    xs = fn(z2.cuda(), m.forward_with_cfg, y=y2.cuda(), grid=grid2.cuda(), mask=mask2.cuda(), size=size2.cuda(), cfg_scale=1.5, scale_pow=0.0)
    err = float((xs[-1].cpu() - ref[-1]).abs().max() / ref[-1].abs().max())
    assert err < 1e-2, err                                                     # bf16 GEMM operands inside the network; the update itself is exact


@pytest.mark.gpu
def test_ode_dopri5_and_rk4_with_fit_forward_with_cfg(built_lib):
    """The reference's ODE route with its default solver: Sampler.sample_ode(sampling_method="dopri5") driving FiT.forward_with_cfg
    (sample_fitv2_ddp.py:147-156, 281), and the fixed-grid rk4, against the oracle sampler driving the oracle model.  The network runs
    with bf16 operands, so the adaptive controller sees slightly different error norms than the fp32 oracle: the step counts may
    differ by a step, the solutions agree within the per-NFE tolerance."""
    from oracle import fitv2_oracle as O
    from fitv2_b200 import FiT, make_grid
    kw = dict(hidden_size=1152, depth=2, num_heads=16, adaln_lora_dim=288)
    th.manual_seed(0)
    m = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora", **kw).randomize_zero_init_(1)
    sd = {k: v.detach().clone().float() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    cfg = O.FiTConfig(**kw)
    n, hp, wp = 2, 8, 8
    g = th.Generator().manual_seed(3)
    z = th.randn(n, hp * wp, 16, generator=g)
    z2 = th.cat([z, z], 0)
    y2 = th.cat([th.tensor([5, 900]), th.full((n,), 1000)])
    grid2, mask2 = make_grid(2 * n, hp, wp), th.ones(2 * n, hp * wp)
    ref_model = lambda x, t, **k: O.forward_with_cfg(cfg, sd, x, t, y2, grid2, mask2, None, 1.5, 0.0)
    s = Sampler(create_transport())
    gpu_kw = dict(y=y2.cuda(), grid=grid2.cuda(), mask=mask2.cuda(), size=None, cfg_scale=1.5, scale_pow=0.0)
    for method, extra in (("dopri5", dict(atol=1e-6, rtol=1e-3)), ("rk4", {})):
        st = {}
        ref = T.sample_ode(ref_model, z2, sampling_method=method, num_steps=4, stats=st, **extra)
        ys = s.sample_ode(sampling_method=method, num_steps=4, **extra)(z2.cuda(), m.forward_with_cfg, **gpu_kw)
        err = max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(ys, ref))
        print(f"[parity] ODE {method} with FiT.forward_with_cfg: max-rel error over the grid {err:.2e}; oracle stats {st}, ours {s.last_ode_stats if method == 'dopri5' else {}}")
        assert len(ys) == 4 and err < 1e-2
        if method == "dopri5":
            assert abs(s.last_ode_stats["steps"] - st["steps"]) <= 1 and s.last_ode_stats["nfe"] >= 8
