"""CPU restatement of the reference's transport ``Sampler`` (SiT-style ODE / SDE integrators) for the
configuration FiTv2 uses (``configs/fitv2/config_fitv2_xl.yaml:3-9``: path Linear, prediction velocity).

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and ``oracle/make_transport_golden.py``), never by the
product package ``fitv2_b200``.  Plain PyTorch fp32 on the CPU, same expression order as the reference so that
the results are bit-equal to it (``oracle/make_transport_golden.py`` asserts that against the real classes of
/root/reference and writes ``tests/golden/transport_kat.pt``).

Reference lines restated (paths relative to the reference repo):
  fit/scheduler/transport/__init__.py:5-71     create_transport (eps defaults)
  fit/scheduler/transport/transport.py:81-108  Transport.check_interval
  fit/scheduler/transport/transport.py:194-228 get_drift (velocity_ode) / get_score (velocity -> score)
  fit/scheduler/transport/transport.py:245-356 Sampler: SDE drift/diffusion, last step, sample_sde
  fit/scheduler/transport/transport.py:358-401 Sampler.sample_ode
  fit/scheduler/transport/path.py:20-100       ICPlan: alpha/sigma, compute_drift, compute_diffusion, score from velocity
  fit/scheduler/transport/integrators.py:8-75  sde: Euler-Maruyama / Heun steps and the forward loop
  fit/scheduler/transport/integrators.py:77-116 ode: torchdiffeq.odeint on linspace(t0, t1, num_steps)

``torchdiffeq`` (requirements.txt, unpinned, NOT vendored and not installed here) supplies ``odeint``; for the
fixed-grid methods its published algorithm is  y_{i+1} = y_i + (t_{i+1} - t_i) * f(t_i, y_i)  ("euler") and
y_{i+1} = y_i + dt * f(t_i + dt/2, y_i + dt/2 * f(t_i, y_i))  ("midpoint"), evaluated on exactly the grid passed in
(no ``step_size`` option is given by the reference).  The other fixed-grid solvers the method string can name are restated
from ``torchdiffeq/_impl/rk_common.py`` (rk2_step_func / rk3_step_func / rk4_alt_step_func) and ``fixed_grid.py``:
  "heun2":  k2 = f(t + dt*1, y + dt*k1*1);                      y' = y + dt*(k1*1/2 + k2*1/2)
  "heun3":  k2 = f(t + dt/3, y + dt*k1*(1/3)); k3 = f(t + 2dt/3, y + dt*(k1*0 + k2*(2/3)));  y' = y + dt*(k1/4 + k2*0 + k3*3/4)
  "rk4":    the 3/8 rule: k2 = f(t + dt/3, y + dt*k1/3); k3 = f(t + 2dt/3, y + dt*(k2 - k1/3)); k4 = f(t1, y + dt*(k1 - k2 + k3));
            y' = y + (k1 + 3*(k2 + k3) + k4)*dt*0.125
Parity of the ODE integrator itself is therefore anchored on the reference's call site only ("parity unpinned" for
torchdiffeq.odeint), while everything on the SDE route is pinned against the reference's own code.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch as th


def expand_t_like_x(t: th.Tensor, x: th.Tensor) -> th.Tensor:
    """path.py:5-13."""
    return t.view(t.size(0), *([1] * (x.dim() - 1)))


# ------------------------------------------------------------------------------------------------
# ICPlan (path.py:17-100), linear coupling  x_t = t * x1 + (1 - t) * x0
# ------------------------------------------------------------------------------------------------
def alpha_t(t):
    return t, 1


def sigma_t(t):
    return 1 - t, -1


def compute_drift(x, t):
    """path.py:35-43 (t already broadcastable): returns (-drift, diffusion)."""
    alpha_ratio = 1 / t
    s, ds = sigma_t(t)
    drift = alpha_ratio * x
    diffusion = alpha_ratio * (s ** 2) - s * ds
    return -drift, diffusion


def compute_diffusion(x, t, form="constant", norm=1.0):
    """path.py:45-69; t is the (B,) time vector."""
    t = expand_t_like_x(t, x)
    if form == "constant":
        return th.tensor(norm).to(x)
    if form == "SBDM":
        return norm * compute_drift(x, t)[1]
    if form == "sigma":
        return norm * sigma_t(t)[0]
    if form == "linear":
        return norm * (1 - t)
    if form == "decreasing":
        return 0.25 * (norm * th.cos(math.pi * t) + 1) ** 2
    if form == "increasing-decreasing":
        return norm * th.sin(math.pi * t) ** 2
    raise NotImplementedError(f"Diffusion form {form} not implemented")


def score_from_velocity(velocity, x, t):
    """path.py:71-85."""
    t = expand_t_like_x(t, x)
    a, da = alpha_t(t)
    s, ds = sigma_t(t)
    reverse_alpha_ratio = a / da
    var = s ** 2 - reverse_alpha_ratio * ds * s
    return (reverse_alpha_ratio * velocity - x) / var


def create_transport_eps(path_type="Linear", prediction="velocity", train_eps=None, sample_eps=None):
    """__init__.py:50-60: velocity on a Linear / GVP path is stable everywhere -> both eps are 0."""
    if path_type != "Linear" or prediction != "velocity":
        raise NotImplementedError("oracle restates the FiTv2 configuration: Linear path, velocity prediction")
    return 0, 0


def check_interval(sample_eps=0, *, diffusion_form="SBDM", sde=False, reverse=False, last_step_size=0.0):
    """transport.py:81-108 for ICPlan + ModelType.VELOCITY, eval=True."""
    t0, t1 = 0, 1
    eps = sample_eps
    if sde:                                       # `self.model_type != VELOCITY or sde`
        t0 = eps if (diffusion_form == "SBDM" and sde) else 0
        t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
    if reverse:
        t0, t1 = 1 - t0, 1 - t1
    return t0, t1


# ------------------------------------------------------------------------------------------------
# SDE (transport.py:245-356, integrators.py:8-75)
# ------------------------------------------------------------------------------------------------
def sde_drift(model, x, t, diffusion_form, diffusion_norm, **kw):
    """transport.py:256-258.  The reference evaluates the model twice (drift and score); the outputs are identical."""
    v = model(x, t, **kw)
    return v + compute_diffusion(x, t, form=diffusion_form, norm=diffusion_norm) * score_from_velocity(v, x, t)


def sample_sde(model: Callable, init: th.Tensor, *, sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0,
               last_step: Optional[str] = "Mean", last_step_size=0.04, num_steps=250,
               noises: Optional[Sequence[th.Tensor]] = None, sample_eps=0, **model_kwargs) -> List[th.Tensor]:
    """transport.py:296-356 + integrators.py:64-75.  `noises` (one tensor per step) replaces the reference's
    ``th.randn(x.size())`` draws from the default CPU generator, in the same order."""
    if last_step is None:
        last_step_size = 0.0
    t0, t1 = check_interval(sample_eps, diffusion_form=diffusion_form, sde=True, last_step_size=last_step_size)
    ts = th.linspace(t0, t1, num_steps)
    dt = ts[1] - ts[0]
    drift = lambda x, t: sde_drift(model, x, t, diffusion_form, diffusion_norm, **model_kwargs)
    x = init
    xs = []
    for i, ti in enumerate(ts[:-1]):
        w_cur = (noises[i] if noises is not None else th.randn(x.size())).to(x)
        dw = w_cur * th.sqrt(dt)
        if sampling_method == "Euler":                                  # integrators.py:29-37
            t = th.ones(x.size(0)).to(x) * ti
            d = drift(x, t)
            diffusion = compute_diffusion(x, t, form=diffusion_form, norm=diffusion_norm)
            mean_x = x + d * dt
            x = mean_x + th.sqrt(2 * diffusion) * dw
        elif sampling_method == "Heun":                                 # integrators.py:39-48
            t_cur = th.ones(x.size(0)).to(x) * ti
            diffusion = compute_diffusion(x, t_cur, form=diffusion_form, norm=diffusion_norm)
            xhat = x + th.sqrt(2 * diffusion) * dw
            k1 = drift(xhat, t_cur)
            xp = xhat + dt * k1
            k2 = drift(xp, t_cur + dt)
            x = xhat + 0.5 * dt * (k1 + k2)
        else:
            raise NotImplementedError("Smapler type not implemented.")
        xs.append(x)
    t_last = th.ones(init.size(0), device=init.device) * t1
    if last_step is None:                                               # transport.py:268-292
        x = xs[-1]
    elif last_step == "Mean":
        x = xs[-1] + drift(xs[-1], t_last) * last_step_size
    elif last_step == "Tweedie":
        v = model(xs[-1], t_last, **model_kwargs)
        a = alpha_t(t_last)[0][0]
        s = sigma_t(t_last)[0][0]
        x = xs[-1] / a + (s ** 2) / a * score_from_velocity(v, xs[-1], t_last)
    elif last_step == "Euler":
        x = xs[-1] + model(xs[-1], t_last, **model_kwargs) * last_step_size
    else:
        raise NotImplementedError()
    xs.append(x)
    assert len(xs) == num_steps, "Samples does not match the number of steps"
    return xs


# ------------------------------------------------------------------------------------------------
# ODE (transport.py:358-401, integrators.py:77-116): fixed-grid methods of torchdiffeq.odeint
# ------------------------------------------------------------------------------------------------
def sample_ode(model: Callable, x: th.Tensor, *, sampling_method="euler", num_steps=50, reverse=False, atol=1e-6, rtol=1e-3,
               stats: Optional[dict] = None, **model_kwargs) -> List[th.Tensor]:
    """Returns the solution at every grid point (odeint returns a (num_steps, ...) tensor; the script takes [-1])."""
    t0, t1 = check_interval(0, sde=False, reverse=reverse)
    ts = th.linspace(t0, t1, num_steps)

    def f(t, y):
        tv = th.ones(y.size(0)).to(y) * t
        if reverse:
            tv = th.ones_like(tv) * (1 - tv)
        return model(y, tv, **model_kwargs)

    if sampling_method == "dopri5":
        return dopri5_integrate(f, x, ts, rtol=rtol, atol=atol, stats=stats)
    ys = [x]
    y = x
    for i in range(num_steps - 1):
        ta, tb = ts[i], ts[i + 1]
        dt = tb - ta
        if sampling_method == "euler":
            y = y + dt * f(ta, y)
        elif sampling_method == "midpoint":
            half_dt = 0.5 * dt
            y_mid = y + f(ta, y) * half_dt
            y = y + dt * f(ta + half_dt, y_mid)
        elif sampling_method == "heun2":
            k1 = f(ta, y)
            k2 = f(ta + dt * 1.0, y + dt * k1 * 1.0)
            y = y + dt * (k1 * (1 / 2) + k2 * (1 / 2))
        elif sampling_method == "heun3":
            k1 = f(ta, y)
            k2 = f(ta + dt * (1 / 3), y + dt * k1 * (1 / 3))
            k3 = f(ta + dt * (2 / 3), y + dt * (k1 * 0.0 + k2 * (2 / 3)))
            y = y + dt * (k1 * (1 / 4) + k2 * 0.0 + k3 * (3 / 4))
        elif sampling_method == "rk4":
            one_third, two_thirds = 1 / 3, 2 / 3
            k1 = f(ta, y)
            k2 = f(ta + dt * one_third, y + dt * k1 * one_third)
            k3 = f(ta + dt * two_thirds, y + dt * (k2 - k1 * one_third))
            k4 = f(tb, y + dt * (k1 - k2 + k3))
            y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        else:
            raise NotImplementedError(f"fixed-grid ODE method {sampling_method!r} not restated (adaptive dopri5 has data-dependent NFE)")
        ys.append(y)
    return ys


def toy_velocity_model(x, t, **kw):
    """Polynomial stand-in for the network (+, -, * only: bit-reproducible on any CPU); used by the golden vectors."""
    tt = expand_t_like_x(t, x)
    return 0.5 * x * (1 + tt) - 0.125 * x * x * x + 0.3 * tt


# ------------------------------------------------------------------------------------------------
# adaptive dopri5 (the reference's DEFAULT ODE method, fit/utils/sit_eval_utils.py:20 -> torchdiffeq.odeint(method="dopri5"))
# ------------------------------------------------------------------------------------------------
# torchdiffeq is an un-vendored, uninstalled dependency ("parity unpinned"): the constants and the control flow below restate its
# published _impl/dopri5.py (Dormand-Prince-Shampine tableau, DPS_C_MID), _impl/rk_common.py (_runge_kutta_step, _adaptive_step,
# RKAdaptiveStepsizeODESolver._before_integrate / _advance), _impl/misc.py (_select_initial_step, _compute_error_ratio,
# _optimal_step_size with safety 0.9, ifactor 10, dfactor 0.2, _rms_norm) and _impl/interp.py (_interp_fit, _interp_evaluate).
DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
DP_C_ERROR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720, -2187 / 6784 - -12231 / 42400,
              11 / 84 - 649 / 6300, -1.0 / 60.0]
DP_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
            187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def _rms(x):
    return x.abs().pow(2).mean().sqrt()


def dopri5_integrate(f: Callable, y0: th.Tensor, ts: th.Tensor, rtol=1e-3, atol=1e-6, stats: Optional[dict] = None,
                     max_steps: int = 100000) -> List[th.Tensor]:
    """Solution at every time of ``ts`` (increasing or decreasing grid), torchdiffeq's adaptive dopri5 semantics: the solver
    walks its own step sequence from ts[0] and evaluates a 4th-order dense-output polynomial at the requested times."""
    sign = 1.0 if float(ts[-1]) >= float(ts[0]) else -1.0                 # torchdiffeq flips time for decreasing grids
    tt = ts * sign
    func = (lambda t, y: f(t, y)) if sign > 0 else (lambda t, y: -f(-t, y))
    nfe = 0
    t0 = tt[0]
    f0 = func(t0, y0); nfe += 1
    # _select_initial_step (order = 4)
    scale = atol + y0.abs() * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    h0 = th.tensor(1e-6, dtype=tt.dtype) if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(t0 + h0, y0 + h0 * f0); nfe += 1
    d2 = _rms((f1 - f0) / scale) / h0
    h1 = th.max(th.tensor(1e-6, dtype=tt.dtype), h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = th.min(100 * h0, h1).to(tt.dtype)
    y, f_cur, t_lo, t_hi = y0, f0, t0, t0
    coeff = [y0] * 5
    out = [y0]
    steps = rejected = 0
    hist = []
    for t_next in tt[1:]:
        while t_next > t_hi:
            if steps >= max_steps or not bool(th.isfinite(dt)) or float(dt) <= 0.0:   # (torchdiffeq's own guard is max_num_steps = 2**31 - 1)
                raise RuntimeError(f"dopri5: no progress after {steps} steps (dt = {float(dt)}): solution blow-up or unreachable tolerance")
            # _runge_kutta_step
            ks = [f_cur]
            for alpha_i, beta_i in zip(DP_ALPHA, DP_BETA):
                ti = t_hi + dt if alpha_i == 1.0 else t_hi + alpha_i * dt
                yi = y + sum(k * (b * dt) for k, b in zip(ks, beta_i))
                ks.append(func(ti, yi)); nfe += 1
            y1 = yi                                                          # c_sol[:-1] == beta[-1] and c_sol[-1] == 0
            f_new = ks[-1]
            err = sum(k * (c * dt) for k, c in zip(ks, DP_C_ERROR))
            tol = atol + rtol * th.max(_rms(y), _rms(y1))
            ratio = _rms(err) / tol
            steps += 1
            hist.append((float(t_hi) * sign, float(dt), float(ratio)))
            if ratio <= 1:
                y_mid = y + sum(k * (c * dt) for k, c in zip(ks, DP_C_MID))
                fa, fb = ks[0], ks[-1]
                a = 2 * dt * (fb - fa) - 8 * (y1 + y) + 16 * y_mid          # _interp_fit
                b = dt * (5 * fa - 3 * fb) + 18 * y + 14 * y1 - 32 * y_mid
                c = dt * (fb - 4 * fa) - 11 * y - 5 * y1 + 16 * y_mid
                coeff = [y, dt * fa, c, b, a]
                t_lo, t_hi = t_hi, t_hi + dt
                y, f_cur = y1, f_new
            else:
                rejected += 1
            # _optimal_step_size
            if ratio == 0:
                dt = dt * 10.0
            else:
                dfactor = 1.0 if ratio < 1 else 0.2
                factor = min(10.0, max(0.9 / float(ratio) ** 0.2, dfactor))
                dt = dt * factor
        x = (t_next - t_lo) / (t_hi - t_lo)                                 # _interp_evaluate
        total = coeff[0] + x * coeff[1]
        xp = x
        for cf in coeff[2:]:
            xp = xp * x
            total = total + xp * cf
        out.append(total)
    if stats is not None:
        stats.update(nfe=nfe, steps=steps, rejected=rejected, history=hist)
    return out
