"""Pin oracle/transport_oracle.py against the REAL reference transport code (build container only; needs
/root/reference).  torchdiffeq is not installed, so it is stubbed: the SDE route, the score / diffusion / interval
functions run from the reference's own code; `odeint` cannot (see transport_oracle.py header).

    python oracle/make_transport_golden.py        -> tests/golden/transport_kat.pt
"""
from __future__ import annotations

import os
import sys
import types

import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("FITV2_REFERENCE", "/root/reference")

if "torchdiffeq" not in sys.modules:
    stub = types.ModuleType("torchdiffeq")
    def _no_odeint(*a, **k):
        raise RuntimeError("torchdiffeq is not installed in this image")
    stub.odeint = _no_odeint
    sys.modules["torchdiffeq"] = stub
if "tqdm" not in sys.modules:
    try:
        import tqdm  # noqa: F401
    except ImportError:
        t = types.ModuleType("tqdm"); t.tqdm = lambda x, **k: x; sys.modules["tqdm"] = t

# import the transport package without executing fit/__init__ side effects
import importlib.util
def _load(name, path, pkg=None):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)] if pkg else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
tp = _load("ref_transport", os.path.join(REF, "fit/scheduler/transport/__init__.py"), pkg=True)

from oracle import transport_oracle as T


def main():
    out = {}
    notes = []
    tr = tp.create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                             snr_type="lognorm")
    assert (tr.train_eps, tr.sample_eps) == T.create_transport_eps()
    sampler = tp.Sampler(tr)
    # ---- intervals ----
    for form in ("SBDM", "sigma", "constant"):
        for lss in (0.0, 0.04):
            ref = tr.check_interval(tr.train_eps, tr.sample_eps, diffusion_form=form, sde=True, eval=True, reverse=False, last_step_size=lss)
            assert ref == T.check_interval(0, diffusion_form=form, sde=True, last_step_size=lss), (form, lss)
    assert tr.check_interval(0, 0, sde=False, eval=True, reverse=False, last_step_size=0.0) == T.check_interval(0, sde=False)
    assert tr.check_interval(0, 0, sde=False, eval=True, reverse=True, last_step_size=0.0) == T.check_interval(0, sde=False, reverse=True)
    notes.append("check_interval: 8 cases equal")
    # ---- pointwise functions ----
    g = th.Generator().manual_seed(11)
    x = th.randn(3, 5, 16, generator=g)
    v = th.randn(3, 5, 16, generator=g)
    t = th.tensor([0.1, 0.5, 0.96])
    ps = tr.path_sampler
    assert th.equal(ps.get_score_from_velocity(v, x, t), T.score_from_velocity(v, x, t))
    forms = ("constant", "SBDM", "sigma", "linear", "decreasing", "increasing-decreasing")
    for form in forms:
        assert th.equal(ps.compute_diffusion(x, t, form=form, norm=0.7), T.compute_diffusion(x, t, form=form, norm=0.7)), form
    notes.append("score_from_velocity + 6 diffusion forms bit-equal")
    out["pointwise"] = dict(x=x, v=v, t=t, score=ps.get_score_from_velocity(v, x, t),
                            diffusion={f: ps.compute_diffusion(x, t, form=f, norm=0.7) for f in forms})
    # ---- SDE trajectories with the polynomial toy model ----
    init = th.randn(4, 6, 16, generator=g)
    cases = []
    for method in ("Euler", "Heun"):
        for form, last, lss, steps in (("sigma", "Mean", 0.04, 12), ("constant", "Euler", 0.04, 9), ("linear", None, 0.04, 7),
                                        ("sigma", "Tweedie", 0.04, 8), ("SBDM", "Mean", 0.04, 6), ("decreasing", "Mean", 0.02, 10)):
            fn = sampler.sample_sde(sampling_method=method, diffusion_form=form, diffusion_norm=0.9, last_step=last,
                                    last_step_size=lss, num_steps=steps)
            th.manual_seed(100 + steps)
            xs_ref = fn(init, T.toy_velocity_model)
            th.manual_seed(100 + steps)
            noises = [th.randn(init.size()) for _ in range(steps - 1)]
            xs = T.sample_sde(T.toy_velocity_model, init, sampling_method=method, diffusion_form=form, diffusion_norm=0.9,
                              last_step=last, last_step_size=lss, num_steps=steps, noises=noises)
            assert len(xs) == len(xs_ref) == steps
            for a, b in zip(xs, xs_ref):
                assert th.equal(th.nan_to_num(a, nan=123.0), th.nan_to_num(b, nan=123.0)) and th.equal(a.isnan(), b.isnan()), (method, form, last)
            cases.append(dict(method=method, form=form, norm=0.9, last_step=last, last_step_size=lss, num_steps=steps,
                              seed=100 + steps, final=xs_ref[-1], mid=xs_ref[steps // 2]))
    out["sde"] = dict(init=init, cases=cases)
    notes.append(f"sample_sde: {len(cases)} trajectories (Euler-Maruyama / Heun x 6 diffusion-form / last-step settings) bit-equal, NaN pattern included")
    # ---- ODE: the drift the reference hands to odeint is the bare model output (velocity_ode) ----
    drift = tr.get_drift()
    tt = th.ones(4) * 0.3
    assert th.equal(drift(init, tt, T.toy_velocity_model), T.toy_velocity_model(init, tt))
    ys = T.sample_ode(T.toy_velocity_model, init, sampling_method="euler", num_steps=11)
    ym = T.sample_ode(T.toy_velocity_model, init, sampling_method="midpoint", num_steps=11)
    out["ode"] = dict(init=init, euler_final=ys[-1], midpoint_final=ym[-1], num_steps=11)
    notes.append("ODE: reference drift == model output (velocity_ode); euler / midpoint grids restated from torchdiffeq's published fixed-grid algorithm (torchdiffeq not installed: integrator itself unpinned)")
    dst = os.path.join(ROOT, "tests", "golden", "transport_kat.pt")
    th.save(out, dst)
    with open(os.path.join(ROOT, "tests", "golden", "README.md"), "a") as f:
        f.write("\n## transport_kat.pt (oracle/make_transport_golden.py, real reference fit/scheduler/transport)\n\n")
        for n in notes:
            f.write(f"* {n}\n")
    print("\n".join(notes))
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
