"""Transport ``Sampler`` of the reference (SiT-style ODE / SDE integrators) on the B200 path.

Mirrors ``fit/scheduler/transport`` of the reference for the configuration FiTv2 uses
(``configs/fitv2/config_fitv2_xl.yaml:3-9``: path ``Linear``, prediction ``velocity``):

    transport = create_transport(path_type="Linear", prediction="velocity")      # __init__.py:5-71
    sampler = Sampler(transport)                                                  # transport.py:230-243
    sample_fn = sampler.sample_sde(sampling_method="Euler", diffusion_form="sigma", last_step="Mean",
                                   last_step_size=0.04, num_steps=250)            # transport.py:296-356
    xs = sample_fn(z, model.forward_with_cfg, **model_kwargs)                     # list of num_steps tensors

Same names, keyword arguments, return values (the list of intermediate samples) and errors as the reference.
What differs is the execution: every update is ONE fused elementwise CUDA kernel called through the C ABI
(``fitv2_sde_step`` & co, bit-exact with the reference's PyTorch expressions), and the SDE drift shares a single
network evaluation between the velocity and the score term (the reference calls the network twice per step,
transport.py:256-258).  The per-step scalars (score variance, diffusion coefficient, step sizes) are computed once
on the host in fp32 with the reference's expression order and kept in a device table.

ODE route: the fixed-grid solvers of ``torchdiffeq.odeint`` that the reference's ``--ode-sampling-method`` string can name —
``euler``, ``midpoint``, ``heun2``, ``heun3``, ``rk4`` (torchdiffeq's rk4 is the 3/8-rule variant) — with one fused update
kernel per stage (``fitv2_scaled_add`` / ``fitv2_rk_stage``), and the reference's DEFAULT, the adaptive ``dopri5``
(Dormand-Prince 5(4) with torchdiffeq's step-size controller and 4th-order dense output): stage arguments, solution, error
estimate and interpolation are ``fitv2_lincomb`` launches, the error / state norms ``fitv2_scaled_rms`` reductions, and the
accept / reject decision is taken on the host from three scalars per step (the number of network evaluations is data dependent
by construction).  ``torchdiffeq`` is an un-vendored dependency of the reference: its published ``fixed_grid.py`` /
``rk_common.py`` / ``dopri5.py`` / ``misc.py`` / ``interp.py`` are restated ("parity unpinned" for the integrator itself).

Not built: the other adaptive solvers (``dopri8``, ``bosh3``, ``fehlberg2``, ``adaptive_heun``), the Adams multistep family and
the likelihood ODE (autograd through the network); they raise ``NotImplementedError`` naming what exists.  Paths / predictions
other than Linear / velocity raise at ``create_transport``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Optional

import torch as th

from . import _lib

_ODE_FIXED = ("euler", "midpoint", "heun2", "heun3", "rk4")
# Dormand-Prince-Shampine tableau and dense-output mid-point weights (torchdiffeq/_impl/dopri5.py)
_DP_ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_DP_BETA = ((1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9), (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
            (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656), (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
_DP_C_ERROR = (35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
               11 / 84 - 649 / 6300, -1.0 / 60.0)
_DP_C_MID = (6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
             187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)
_DIFFUSION_FORMS = ("constant", "SBDM", "sigma", "linear", "decreasing", "increasing-decreasing")


class Transport:
    """Subset of transport.py:37-108 that sampling needs (Linear path, velocity prediction)."""

    def __init__(self, *, path_type="Linear", prediction="velocity", train_eps=0, sample_eps=0, snr_type="uniform",
                 loss_weight=None):
        self.path_type, self.prediction = path_type, prediction
        self.train_eps, self.sample_eps, self.snr_type, self.loss_weight = train_eps, sample_eps, snr_type, loss_weight

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        """transport.py:81-108 (ICPlan, ModelType.VELOCITY)."""
        t0, t1 = 0, 1
        eps = train_eps if not eval else sample_eps
        if sde:
            t0 = eps if (diffusion_form == "SBDM" and sde) else 0
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                     snr_type="uniform") -> Transport:
    """__init__.py:5-71.  Velocity on the Linear path is stable everywhere: both eps are 0 (lines 57-60)."""
    if path_type != "Linear" or prediction not in ("velocity", None):
        raise NotImplementedError(f"fitv2_b200 implements the FiTv2 transport (Linear path, velocity prediction), got {path_type}/{prediction}")
    if snr_type not in ("uniform", "lognorm"):
        raise ValueError(f"Invalid snr type {snr_type}")
    return Transport(path_type=path_type, prediction="velocity", train_eps=0, sample_eps=0, snr_type=snr_type, loss_weight=loss_weight)


def _f32(v) -> th.Tensor:
    return th.as_tensor(v, dtype=th.float32)


def _coef_row(t: th.Tensor, dt: th.Tensor, form: str, norm: float) -> th.Tensor:
    """The 8 per-step scalars of include/fitv2_b200.h, evaluated on 0-dim fp32 CPU tensors in the reference's own
    expression order (path.py:20-85), so that they carry the same roundings as its (B,1,1) tensors."""
    alpha, d_alpha = t, 1
    sigma, d_sigma = 1 - t, -1
    rar = alpha / d_alpha
    var = sigma ** 2 - rar * d_sigma * sigma
    if form == "constant":
        diff = th.tensor(norm)
    elif form == "SBDM":
        alpha_ratio = 1 / t
        diff = norm * (alpha_ratio * (sigma ** 2) - sigma * d_sigma)
    elif form == "sigma":
        diff = norm * sigma
    elif form == "linear":
        diff = norm * (1 - t)
    elif form == "decreasing":
        diff = 0.25 * (norm * th.cos(math.pi * t) + 1) ** 2
    elif form == "increasing-decreasing":
        diff = norm * th.sin(math.pi * t) ** 2
    else:
        raise NotImplementedError(f"Diffusion form {form} not implemented")
    return th.stack([_f32(rar), _f32(var), _f32(diff), _f32(dt), th.sqrt(2 * _f32(diff)), th.sqrt(_f32(dt)),
                     _f32(alpha), _f32((sigma ** 2) / alpha)])


def _p(t: Optional[th.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(th.cuda.current_stream(dev).cuda_stream)


def _check_state(x: th.Tensor, what: str) -> th.Tensor:
    if not isinstance(x, th.Tensor) or not x.is_cuda:
        raise _lib.FitV2Error(f"{what} must be a CUDA tensor: fitv2_b200 has no CPU path")
    if x.dtype != th.float32:
        raise _lib.FitV2Error(f"{what} must be float32 (the sampler state is integrated in fp32), got {x.dtype}")
    return x.contiguous()


def _device_ops(y0: th.Tensor):
    """The two device-side operations of the adaptive solver, bound to the C ABI for tensors shaped like ``y0``."""
    lib = _lib.load()
    dev, n = y0.device, y0.numel()
    st = _stream(dev)
    norms = th.empty(4, dtype=th.float32, device=dev)

    def lincomb(base: th.Tensor, cy, ks, cs) -> th.Tensor:
        out = th.empty_like(base)
        c = th.stack([_f32(cy)] + [_f32(v) for v in cs]).to(dev)
        ptrs = (C.c_void_p * max(len(ks), 1))(*[k.data_ptr() for k in ks])
        _lib.check(lib.fitv2_lincomb(_p(out), _p(base), ptrs, _p(c), len(ks), n, st), "fitv2_lincomb")
        return out

    def rms_norms(items) -> List[float]:
        """items: up to 4 tuples (a, b, s, atol, rtol) -> [sqrt(mean(((a - b) / (atol + rtol |s|))^2))]; ONE host read for all."""
        for slot, (a, b, s_, at, rt) in enumerate(items):
            _lib.check(lib.fitv2_scaled_rms(C.c_void_p(norms.data_ptr() + 4 * slot), _p(a), _p(b), _p(s_), float(at), float(rt), n, st),
                       "fitv2_scaled_rms")
        return norms[:len(items)].tolist()

    return lincomb, rms_norms


def _dopri5(f: Callable, y0: th.Tensor, ts: th.Tensor, rtol: float, atol: float, stats: dict, ops=None, max_steps: int = 100000) -> List[th.Tensor]:
    """torchdiffeq's adaptive dopri5 (RKAdaptiveStepsizeODESolver: _before_integrate, _advance, _adaptive_step, _runge_kutta_step;
    misc._select_initial_step / _compute_error_ratio / _optimal_step_size with safety 0.9, ifactor 10, dfactor 0.2, RMS norm;
    interp._interp_fit / _interp_evaluate): the solution at every time of ``ts``.  The step size and the times live on the host in
    fp32 like torchdiffeq's 0-dim tensors; per step the device reports three norms.  ``ops`` = (lincomb, rms_norms): the CUDA kernels
    by default; the CPU tests inject torch implementations to exercise this controller without a GPU."""
    lincomb, rms_norms = ops if ops is not None else _device_ops(y0)
    sign = 1.0 if float(ts[-1]) >= float(ts[0]) else -1.0                    # decreasing grids: integrate -f(-t, y) forwards
    tt = ts * sign
    nfe = 0

    def func(t: th.Tensor, state: th.Tensor) -> th.Tensor:
        nonlocal nfe
        nfe += 1
        out = f(t * sign, state)
        return out if sign > 0 else lincomb(out, -1.0, [], [])

    t0 = tt[0]
    f0 = func(t0, y0)
    # ---- _select_initial_step (order 4) ----
    d0, d1 = [th.tensor(v) for v in rms_norms([(y0, None, y0, atol, rtol), (f0, None, y0, atol, rtol)])]
    h0 = th.tensor(1e-6) if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(t0 + h0, lincomb(y0, 1.0, [f0], [h0]))
    d2 = th.tensor(rms_norms([(f1, f0, y0, atol, rtol)])[0]) / h0
    h1 = th.max(th.tensor(1e-6), h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = th.min(100 * h0, h1).to(th.float32)
    y, f_cur, t_lo, t_hi = y0, f0, t0, t0
    coeff = None                                                             # [e, d, c, b, a] of the last accepted step
    out = [y0.clone()]
    steps = rejected = 0
    hist = []                                                                # (t, dt, error ratio) of every attempted step
    for t_next in tt[1:]:
        while t_next > t_hi:
            if steps >= max_steps or not bool(th.isfinite(dt)) or float(dt) <= 0.0:
                raise RuntimeError(f"dopri5: no progress after {steps} steps (dt = {float(dt)}, t = {float(t_hi) * sign}): the model output is "
                                   "not finite or the tolerance cannot be met")
            ks = [f_cur]                                                     # _runge_kutta_step
            for alpha_i, beta_i in zip(_DP_ALPHA, _DP_BETA):
                ti = t_hi + dt if alpha_i == 1.0 else t_hi + alpha_i * dt
                yi = lincomb(y, 1.0, ks, [b * dt for b in beta_i])
                ks.append(func(ti, yi))
            y1, f_new = yi, ks[-1]                                           # c_sol == beta[-1] (first-same-as-last)
            err = lincomb(y, 0.0, ks, [c * dt for c in _DP_C_ERROR])
            ny, ny1, nerr = rms_norms([(y, None, None, 0.0, 0.0), (y1, None, None, 0.0, 0.0), (err, None, None, 0.0, 0.0)])
            ratio = nerr / (atol + rtol * max(ny, ny1))                      # the one host synchronisation of the step
            steps += 1
            hist.append((float(t_hi) * sign, float(dt), float(ratio)))
            if ratio <= 1:
                y_mid = lincomb(y, 1.0, ks, [c * dt for c in _DP_C_MID])
                fa, fb = ks[0], ks[-1]                                       # _interp_fit
                pa = lincomb(y, -8.0, [y1, y_mid, fa, fb], [-8.0, 16.0, -2 * dt, 2 * dt])
                pb = lincomb(y, 18.0, [y1, y_mid, fa, fb], [14.0, -32.0, 5 * dt, -3 * dt])
                pc = lincomb(y, -11.0, [y1, y_mid, fa, fb], [-5.0, 16.0, -4 * dt, dt])
                pd = lincomb(fa, dt, [], [])
                coeff = [y, pd, pc, pb, pa]
                t_lo, t_hi = t_hi, t_hi + dt
                y, f_cur = y1, f_new
            else:
                rejected += 1
            if ratio == 0:                                                   # _optimal_step_size
                dt = dt * 10.0
            elif ratio != ratio:                                             # NaN error norm: shrink like a rejected step
                dt = dt * 0.2
            else:
                dt = dt * min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2))
        x = (t_next - t_lo) / (t_hi - t_lo)                                  # _interp_evaluate
        out.append(lincomb(coeff[0], 1.0, coeff[1:], [x, x * x, x * x * x, x * x * x * x]))
    stats.clear()
    stats.update(nfe=nfe, steps=steps, rejected=rejected, history=hist)
    return out


def _torch_ops():
    """CPU stand-ins for the two kernels (tests of the controller logic only; the product path never uses them)."""
    def lincomb(base, cy, ks, cs):
        acc = _f32(cy) * base
        for k, c in zip(ks, cs):
            acc = acc + _f32(c) * k
        return acc

    def rms_norms(items):
        out = []
        for a, b, s_, at, rt in items:
            v = a - b if b is not None else a
            if s_ is not None:
                v = v / (at + rt * s_.abs())
            out.append(float(v.double().pow(2).mean().sqrt()))
        return out
    return lincomb, rms_norms


class Sampler:
    """transport.py:230-401."""

    def __init__(self, transport: Transport):
        self.transport = transport
        self.last_ode_stats = {}                       # dopri5: network evaluations / steps / rejected steps of the last call

    # ------------------------------------------------------------------ SDE
    def sample_sde(self, *, sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0, last_step="Mean",
                   last_step_size=0.04, num_steps=250, noise: str = "device",
                   generator: Optional[th.Generator] = None) -> Callable[..., List[th.Tensor]]:
        """transport.py:296-356.  ``noise``: "device" draws the Wiener increments on the GPU (``generator`` optional);
        "reference" draws them exactly like the reference, ``th.randn(x.size())`` from the default CPU generator, and
        copies them over (bit-parity runs)."""
        if sampling_method not in ("Euler", "Heun"):
            raise NotImplementedError("Smapler type not implemented.")           # integrators.py:58-59 (sic)
        if diffusion_form not in _DIFFUSION_FORMS:
            raise NotImplementedError(f"Diffusion form {diffusion_form} not implemented")
        if last_step not in (None, "Mean", "Tweedie", "Euler"):
            raise NotImplementedError()
        if noise not in ("device", "reference"):
            raise ValueError("noise must be 'device' or 'reference'")
        if last_step is None:
            last_step_size = 0.0
        tr = self.transport
        t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, diffusion_form=diffusion_form, sde=True, eval=True,
                                   reverse=False, last_step_size=last_step_size)
        ts = th.linspace(t0, t1, num_steps)                                      # integrators.py:22-23 (CPU fp32)
        dt = ts[1] - ts[0]
        norm = diffusion_norm
        rows = [_coef_row(ti, dt, diffusion_form, norm) for ti in ts[:-1]]
        rows2 = [_coef_row(ti + dt, dt, diffusion_form, norm) for ti in ts[:-1]] if sampling_method == "Heun" else []
        t1f = th.ones(1) * t1                                                    # transport.py:347
        last = _coef_row(t1f[0], _f32(last_step_size), diffusion_form, norm)
        half_dt = 0.5 * dt
        lib = _lib.load()

        def _sample(init: th.Tensor, model: Callable, **model_kwargs) -> List[th.Tensor]:
            x = _check_state(init, "init").clone()
            dev, n, B = x.device, x.numel(), x.size(0)
            coef = th.stack(rows).to(dev)
            coef2 = th.stack(rows2).to(dev) if rows2 else None
            coef_last = last.to(dev)
            sc = th.stack([th.stack([th.sqrt(dt), r[4]]) for r in rows]).to(dev)             # (w * sqrt(dt)) * sqrt(2D)
            sc_dt = th.stack([_f32(1.0), dt]).to(dev)
            sc_last = th.stack([_f32(1.0), _f32(last_step_size)]).to(dev)
            c_half = half_dt.reshape(1).to(dev)
            xs: List[th.Tensor] = []
            with th.cuda.device(dev), th.no_grad():
                st = _stream(dev)
                for i, ti in enumerate(ts[:-1]):
                    if noise == "reference":
                        w = th.randn(x.size()).to(x)                             # integrators.py:30 / :40
                    else:
                        w = th.randn(x.shape, device=dev, dtype=th.float32, generator=generator)
                    t = th.full((B,), float(ti), device=dev, dtype=th.float32)
                    if sampling_method == "Euler":
                        v = _check_state(model(x, t, **model_kwargs), "model output")
                        _lib.check(lib.fitv2_sde_step(_p(x), _p(v), _p(w), _p(coef[i]), n, st), "fitv2_sde_step")
                    else:
                        xhat = th.empty_like(x)
                        _lib.check(lib.fitv2_scaled_add(_p(xhat), _p(x), _p(w), _p(sc[i]), n, st), "fitv2_scaled_add")
                        v1 = _check_state(model(xhat, t, **model_kwargs), "model output")
                        k1 = th.empty_like(x)
                        _lib.check(lib.fitv2_sde_drift(_p(k1), _p(xhat), _p(v1), _p(coef[i]), n, st), "fitv2_sde_drift")
                        xp = th.empty_like(x)
                        _lib.check(lib.fitv2_scaled_add(_p(xp), _p(xhat), _p(k1), _p(sc_dt), n, st), "fitv2_scaled_add")
                        t2 = th.full((B,), float(ti + dt), device=dev, dtype=th.float32)
                        v2 = _check_state(model(xp, t2, **model_kwargs), "model output")
                        k2 = th.empty_like(x)
                        _lib.check(lib.fitv2_sde_drift(_p(k2), _p(xp), _p(v2), _p(coef2[i]), n, st), "fitv2_sde_drift")
                        _lib.check(lib.fitv2_heun_combine(_p(x), _p(xhat), _p(k1), _p(k2), _p(c_half), n, st), "fitv2_heun_combine")
                    xs.append(x.clone())
                # ---- last step (transport.py:268-292, 346-349) ----
                if last_step is not None:
                    tl = th.full((B,), float(t1f[0]), device=dev, dtype=th.float32)
                    v = _check_state(model(x, tl, **model_kwargs), "model output")
                    if last_step == "Mean":
                        _lib.check(lib.fitv2_sde_step(_p(x), _p(v), None, _p(coef_last), n, st), "fitv2_sde_step")
                    elif last_step == "Euler":
                        _lib.check(lib.fitv2_scaled_add(_p(x), _p(x), _p(v), _p(sc_last), n, st), "fitv2_scaled_add")
                    else:
                        _lib.check(lib.fitv2_tweedie(_p(x), _p(x), _p(v), _p(coef_last), n, st), "fitv2_tweedie")
                xs.append(x)
            assert len(xs) == num_steps, "Samples does not match the number of steps"
            return xs

        return _sample

    # ------------------------------------------------------------------ ODE
    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False
                   ) -> Callable[..., List[th.Tensor]]:
        """transport.py:358-401 with the fixed-grid methods of ``torchdiffeq.odeint`` on ``linspace(t0, t1, num_steps)``
        (integrators.py:95-116); returns the solution at every grid point.  ``atol`` / ``rtol`` are accepted for signature
        compatibility (fixed-grid solvers ignore them, as in torchdiffeq)."""
        if sampling_method not in _ODE_FIXED + ("dopri5",):
            raise NotImplementedError(f"ODE method {sampling_method!r}: built are the fixed-grid solvers {_ODE_FIXED} and the adaptive 'dopri5'")
        tr = self.transport
        t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, sde=False, eval=True, reverse=reverse, last_step_size=0.0)
        ts = th.linspace(t0, t1, num_steps)
        lib = _lib.load()
        third, two_thirds = 1.0 / 3.0, 2.0 / 3.0                                 # rk_common.py: _one_third, _two_thirds

        def _sample(x: th.Tensor, model: Callable, **model_kwargs) -> List[th.Tensor]:
            y = _check_state(x, "x").clone()
            dev, n, B = y.device, y.numel(), y.size(0)
            ys = [y.clone()]

            def f(tval: th.Tensor, state: th.Tensor) -> th.Tensor:
                tv = th.ones(B) * tval                                           # integrators.py:102
                if reverse:
                    tv = th.ones_like(tv) * (1 - tv)                             # transport.py:376-377
                return _check_state(model(state, tv.to(dev), **model_kwargs), "model output")

            if sampling_method == "dopri5":
                with th.cuda.device(dev), th.no_grad():
                    return _dopri5(f, y, ts, float(rtol), float(atol), self.last_ode_stats)

            with th.cuda.device(dev), th.no_grad():
                st = _stream(dev)

                def stage(out, k1, k2, k3, k4, dt, mode, s1=0.0, s2=0.0, s3=0.0):
                    sc = th.stack([_f32(dt), _f32(s1), _f32(s2), _f32(s3)]).to(dev)
                    _lib.check(lib.fitv2_rk_stage(_p(out), _p(y), _p(k1), _p(k2), _p(k3), _p(k4), _p(sc), mode, n, st), "fitv2_rk_stage")
                    return out

                for i in range(num_steps - 1):
                    ta, tb = ts[i], ts[i + 1]
                    dt = tb - ta
                    if sampling_method == "euler":
                        s = th.stack([_f32(1.0), dt]).to(dev)
                        _lib.check(lib.fitv2_scaled_add(_p(y), _p(y), _p(f(ta, y)), _p(s), n, st), "fitv2_scaled_add")
                    elif sampling_method == "midpoint":
                        half = 0.5 * dt
                        y_mid = th.empty_like(y)
                        s1 = th.stack([_f32(1.0), half]).to(dev)
                        _lib.check(lib.fitv2_scaled_add(_p(y_mid), _p(y), _p(f(ta, y)), _p(s1), n, st), "fitv2_scaled_add")
                        s2 = th.stack([_f32(1.0), dt]).to(dev)
                        _lib.check(lib.fitv2_scaled_add(_p(y), _p(y), _p(f(ta + half, y_mid)), _p(s2), n, st), "fitv2_scaled_add")
                    elif sampling_method == "heun2":                             # tableau [[0,0,0],[1,1,0],[0,1/2,1/2]] (rk2_step_func)
                        k1 = f(ta, y)
                        k2 = f(ta + dt * 1.0, stage(th.empty_like(y), k1, None, None, None, dt, 0, 1.0))
                        stage(y, k1, k2, None, None, dt, 1, 0.5, 0.5)
                    elif sampling_method == "heun3":                             # tableau c = (0, 1/3, 2/3), b = (1/4, 0, 3/4) (rk3_step_func)
                        k1 = f(ta, y)
                        k2 = f(ta + dt * third, stage(th.empty_like(y), k1, None, None, None, dt, 0, third))
                        k3 = f(ta + dt * two_thirds, stage(th.empty_like(y), k1, k2, None, None, dt, 1, 0.0, two_thirds))
                        stage(y, k1, k2, k3, None, dt, 2, 0.25, 0.0, 0.75)
                    else:                                                        # rk4: torchdiffeq's 3/8 rule (rk4_alt_step_func)
                        k1 = f(ta, y)
                        k2 = f(ta + dt * third, stage(th.empty_like(y), k1, None, None, None, dt, 0, third))
                        k3 = f(ta + dt * two_thirds, stage(th.empty_like(y), k1, k2, None, None, dt, 3, third))
                        k4 = f(tb, stage(th.empty_like(y), k1, k2, k3, None, dt, 4))
                        stage(y, k1, k2, k3, k4, dt, 5)
                    ys.append(y.clone())
            return ys

        return _sample

    def sample_ode_likelihood(self, **kw):
        raise NotImplementedError("the likelihood ODE differentiates through the network (autograd); fitv2_b200 is inference only")
