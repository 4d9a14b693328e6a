"""Golden fixtures for BASELINE.json configs 3 / 4 / 5 at their REAL depth and for the full 250-step trajectory,
generated from the REAL reference classes (run in the build container only; TEST INFRASTRUCTURE).

State dicts are regenerated from seeds on the test side (reference init under ``torch.manual_seed(0)`` + the
zero-parameter redraw of SURVEY.md §8d), so only inputs, reference outputs and weight checksums are stored.

  cfg4_3b_d40.pt        FiTv2-3B/2 (depth 40, hidden 2304, 24 heads of 96), 256 tokens, one CFG NFE at batch 2 (4 rows)
                        at t = 0 and t = 0.5, and the latents after the first Euler step
                        (configs/fitv2/config_fitv2_3B.yaml:30-47, sample_fitv2_ddp.py:287-314)
  cfg5_xl_1024.pt       FiTv2-XL/2 depth 36 at 512x512 -> 32x32 = 1024 tokens, ntk-aware + decouple, ori_max_pe_len 16
                        (sample_fitv2_ddp.py:75-99), one CFG NFE at batch 1 (2 rows) at t = 0 and t = 0.5
  cfg3_xl_d36_dynntk.pt FiTv2-XL/2 depth 36, 160x320 -> 10x20 = 200 tokens dynntk; (i) one CFG NFE at batch 2, t = 0.5,
                        (ii) a mixed-aspect batch padded to 256 (fit/data/in1k_latent_dataset.py:54-69 layout)
  xl_traj250.pt         FiTv2-XL/2 depth 36, 256 tokens, batch 1: latents after 50 / 125 / 250 steps of the script's
                        250-step CFG 1.5 Euler loop

Usage:  python oracle/make_config_goldens.py [--only cfg4,cfg5,cfg3,traj] [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fitv2_oracle as O                       # noqa: E402
from oracle.make_golden import install_reference, ref_kwargs, mixed_padded_batch   # noqa: E402


def build_reference_inplace(FiT, cfg, init_seed=0, redraw_seed=1, std=0.02):
    """Reference model with the synthetic weights, without the two extra state_dict copies of make_golden.build_reference
    (3B/2 is 12 GB in fp32)."""
    torch.manual_seed(init_seed)
    m = FiT(**ref_kwargs(cfg)).eval()
    g = torch.Generator().manual_seed(redraw_seed)
    for _, p in m.state_dict().items():                    # O.redraw_zero_params, in place
        if p.is_floating_point() and not bool(p.any()):
            p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float32) * std)
    return m


def checksum_some(sd, every=37):
    keys = list(sd.keys())
    pick = keys[::every] + keys[-3:]
    return {k: (float(sd[k].double().sum()), float(sd[k].double().abs().sum())) for k in pick}


def script_inputs(n, hp, wp):
    """sample_fitv2_ddp.py:54-55,257-268 with global_seed 0, rank 0 (labels from the CPU generator)."""
    torch.manual_seed(0)
    N = hp * wp
    z = torch.randn(n, N, 16)
    y = torch.randint(0, 1000, (n,))
    return z, y, O.make_grid(n, hp, wp), torch.ones(n, N)


def cfg_nfe(m, z, y, grid, mask, t):
    n = z.shape[0]
    y2 = torch.cat([y, torch.full((n,), 1000)], 0)
    return m(torch.cat([z, z], 0), torch.full((2 * n,), float(t)), y=y2, grid=torch.cat([grid, grid], 0),
             mask=torch.cat([mask, mask], 0), size=None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--only", default="cfg3,cfg5,traj,cfg4")
    args = ap.parse_args()
    only = set(args.only.split(","))
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count() or 1)
    FiT = install_reference()
    report = []
    dyn = dict(custom_freqs="ntk-aware", decouple=True, ori_max_pe_len=16)

    if "cfg3" in only:
        t0 = time.time()
        cfg = O.FiTConfig(**O.XL2, **dyn, max_pe_len_h=10, max_pe_len_w=20)
        m = build_reference_inplace(FiT, cfg)
        sd = m.state_dict()
        z, y, grid, mask = script_inputs(2, 10, 20)
        v05 = cfg_nfe(m, z, y, grid, mask, 0.5)
        y2 = torch.cat([y, torch.full((2,), 1000)], 0)
        o05 = O.forward(cfg, sd, torch.cat([z, z]), torch.full((4,), 0.5), y2, torch.cat([grid, grid]), torch.cat([mask, mask]))
        assert torch.equal(v05, o05), (v05 - o05).abs().max()
        xp, gp, mp = mixed_padded_batch(cfg, [(10, 20), (16, 16), (8, 24), (20, 10)], 256, seed=11)
        tp = torch.tensor([0.1, 0.5, 0.9, 0.3])
        yp = torch.tensor([7, 1000, 999, 1000])
        vp = m(xp, tp, y=yp, grid=gp, mask=mp, size=None)
        assert torch.equal(vp, O.forward(cfg, sd, xp, tp, yp, gp, mp))
        torch.save(dict(z=z, y=y, v_t05=v05, x_pad=xp, t_pad=tp, y_pad=yp, grid_pad=gp, mask_pad=mp, v_pad=vp,
                        weight_checksum=checksum_some(sd)), os.path.join(args.out, "cfg3_xl_d36_dynntk.pt"))
        report.append(f"cfg3_xl_d36_dynntk: XL/2 depth 36, 10x20 dynntk CFG NFE (t=0.5) + mixed padded batch; oracle bit-equal "
                      f"(|v|max {float(v05.abs().max()):.4f} / {float(vp.abs().max()):.4f}, {time.time() - t0:.0f} s)")
        del m, sd

    if "cfg5" in only:
        t0 = time.time()
        cfg = O.FiTConfig(**O.XL2, **dyn, max_pe_len_h=32, max_pe_len_w=32)
        m = build_reference_inplace(FiT, cfg)
        sd = m.state_dict()
        z, y, grid, mask = script_inputs(1, 32, 32)
        v0 = cfg_nfe(m, z, y, grid, mask, 0.0)
        v05 = cfg_nfe(m, z, y, grid, mask, 0.5)
        y2 = torch.cat([y, torch.full((1,), 1000)], 0)
        o0 = O.forward(cfg, sd, torch.cat([z, z]), torch.zeros(2), y2, torch.cat([grid, grid]), torch.cat([mask, mask]))
        assert torch.equal(v0, o0), (v0 - o0).abs().max()
        torch.save(dict(z=z, y=y, v_t0=v0, v_t05=v05, weight_checksum=checksum_some(sd)), os.path.join(args.out, "cfg5_xl_1024.pt"))
        report.append(f"cfg5_xl_1024: XL/2 depth 36, 32x32 = 1024 tokens ntk-aware decouple ori 16, CFG NFE at t=0 / 0.5; oracle "
                      f"bit-equal at t=0 (|v|max {float(v0.abs().max()):.4f}, {time.time() - t0:.0f} s)")
        del m, sd

    if "traj" in only:
        t0 = time.time()
        cfg = O.FiTConfig(**O.XL2)
        m = build_reference_inplace(FiT, cfg)
        z, y, grid, mask = script_inputs(1, 16, 16)
        sig = torch.linspace(0, 1, 251)
        y2 = torch.cat([y, torch.full((1,), 1000)], 0)
        grid2, mask2 = torch.cat([grid, grid], 0), torch.cat([mask, mask], 0)
        zz, keep = z, {}
        for idx in range(250):                             # sample_fitv2_ddp.py:297-314 verbatim semantics
            z_in = torch.cat([zz, zz], 0)
            v2 = m(z_in, sig[idx].expand(2), y=y2, grid=grid2, mask=mask2, size=None)
            cond, uncond = v2.chunk(2, dim=0)
            v = uncond + 1.5 * (cond - uncond)
            zz = zz + (sig[idx + 1] - sig[idx]) * v
            if idx + 1 in (1, 50, 125, 250):
                keep[idx + 1] = zz.clone()
        torch.save(dict(z=z, y=y, z_step1=keep[1], z_step50=keep[50], z_step125=keep[125], z_step250=keep[250],
                        weight_checksum=checksum_some(m.state_dict())), os.path.join(args.out, "xl_traj250.pt"))
        report.append(f"xl_traj250: XL/2 depth 36, batch 1, the script's 250-step CFG 1.5 Euler loop on the REAL reference model "
                      f"(|z250|max {float(keep[250].abs().max()):.4f}, rms {float(keep[250].pow(2).mean().sqrt()):.4f}, {time.time() - t0:.0f} s)")
        del m

    if "cfg4" in only:
        t0 = time.time()
        cfg = O.FiTConfig(**O.B3_2)
        m = build_reference_inplace(FiT, cfg)
        sd = m.state_dict()
        z, y, grid, mask = script_inputs(2, 16, 16)
        v0 = cfg_nfe(m, z, y, grid, mask, 0.0)
        v05 = cfg_nfe(m, z, y, grid, mask, 0.5)
        y2 = torch.cat([y, torch.full((2,), 1000)], 0)
        o0 = O.forward(cfg, sd, torch.cat([z, z]), torch.zeros(4), y2, torch.cat([grid, grid]), torch.cat([mask, mask]))
        assert torch.equal(v0, o0), (v0 - o0).abs().max()
        sig = torch.linspace(0, 1, 251)
        z1 = O.cfg_euler_update(z, v0, 1.5, sig[0], sig[1])
        torch.save(dict(z=z, y=y, v_t0=v0, v_t05=v05, z_step0=z1, weight_checksum=checksum_some(sd),
                        n_params=sum(v.numel() for v in sd.values())), os.path.join(args.out, "cfg4_3b_d40.pt"))
        report.append(f"cfg4_3b_d40: 3B/2 depth 40 ({sum(v.numel() for v in sd.values())} params), 256 tokens, CFG NFE at "
                      f"t=0 / 0.5; oracle bit-equal at t=0 (|v|max {float(v0.abs().max()):.4f}, {time.time() - t0:.0f} s)")
        del m, sd

    with open(os.path.join(args.out, "README.md"), "a") as f:
        f.write("\n## BASELINE configs 3 / 4 / 5 at full depth + the 250-step trajectory (oracle/make_config_goldens.py, real reference)\n\n"
                + "\n".join(f"* {r}" for r in report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
