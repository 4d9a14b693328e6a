"""Golden fixtures for the non-default model variants (SURVEY.md §8 f4 and the norm kinds of fit/model/norms.py),
generated from the REAL reference classes (build container only; TEST INFRASTRUCTURE).

  fitv1_xl_d2.pt         the `params:` block of configs/fit/config_fit_xl.yaml:20-36 at depth 2: learn_sigma (32 output
                         channels), use_sit=False ((B, C, N) tensors), adaln_type 'normal', no q / k norm, use_swiglu_large;
                         forward + forward_with_cfg on a mixed-aspect padded batch
  norm_variants_xl_d1.pt two depth-1 XL-width FiTv2-style models with the other create_norm kinds:
                         (a) norm_type 'rmsnorm', q_norm 'rmsnorm', k_norm 'layernorm' + qk_norm_weight (-> w_layernorm)
                         (b) norm_type 'w_layernorm', q_norm None, k_norm 'rmsnorm'
                         norm weights perturbed by oracle.perturb_norm_weights (they initialise to ones)

State dicts are regenerated from seeds on the test side; inputs, reference outputs and weight checksums are stored.
Usage:  python oracle/make_variant_goldens.py
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fitv2_oracle as O                                            # noqa: E402
from oracle.make_golden import install_reference, mixed_padded_batch, checksum  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FITV1 = dict(context_size=256, patch_size=2, in_channels=4, hidden_size=1152, depth=2, num_heads=16, mlp_ratio=4.0,
             class_dropout_prob=0.1, num_classes=1000, learn_sigma=True, use_swiglu=True, use_swiglu_large=True,
             rel_pos_embed="rope")                                              # config_fit_xl.yaml:22-36, depth 28 -> 2


def build(FiT, kwargs, cfg, perturb=False):
    torch.manual_seed(0)
    m = FiT(**kwargs).eval()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    osd0 = O.reference_init_state_dict(cfg, 0)
    assert list(osd0.keys()) == list(sd0.keys()), (list(osd0.keys())[:20], list(sd0.keys())[:20])
    for k in sd0:
        assert torch.equal(sd0[k], osd0[k]), k
    sd = O.redraw_zero_params(sd0, 1)
    if perturb:
        sd = O.perturb_norm_weights(sd, 2)
    m.load_state_dict(sd)
    return m, sd


def main():
    torch.set_grad_enabled(False)
    FiT = install_reference()
    report = []

    # ---- FiTv1 ----
    cfg = O.FiTConfig(hidden_size=1152, depth=2, num_heads=16, adaln_lora_dim=0, learn_sigma=True, use_sit=False,
                      adaln_type="normal", q_norm=None, k_norm=None, use_swiglu_large=True)
    m, sd = build(FiT, FITV1, cfg)
    x, grid, mask = mixed_padded_batch(cfg, [(10, 20), (16, 16), (8, 24), (20, 10)], 256, seed=11)
    xc = x.transpose(1, 2).contiguous()                                         # (B, C, N)
    t = torch.tensor([0.1, 0.5, 0.9, 0.3])
    y = torch.tensor([7, 1000, 999, 1000])
    ref = m(xc, t, y=y, grid=grid, mask=mask, size=None)
    got = O.forward(cfg, sd, xc, t, y, grid, mask)
    assert ref.shape == (4, 32, 256) and torch.equal(ref, got), (ref - got).abs().max()
    refc = m.forward_with_cfg(xc, t, y, grid, mask, None, 1.5)
    assert torch.equal(refc, O.forward_with_cfg(cfg, sd, xc, t, y, grid, mask, None, 1.5))
    lat = torch.randn(2, 16, 200, generator=torch.Generator().manual_seed(3))
    unp = m.unpatchify(lat, (20, 40))
    torch.save(dict(kwargs=FITV1, x=xc, t=t, y=y, grid=grid, mask=mask, out=ref, out_cfg=refc, unpatchify_in=lat, unpatchify_out=unp,
                    weight_checksum=checksum(sd)), os.path.join(OUT, "fitv1_xl_d2.pt"))
    report.append(f"fitv1_xl_d2: config_fit_xl.yaml params at depth 2 (learn_sigma, (B,C,N) layout, adaLN 'normal', no q/k norm, "
                  f"swiglu_large): init + forward + forward_with_cfg + unpatchify, oracle bit-equal (|out|max {float(ref.abs().max()):.4f})")

    # ---- norm variants ----
    base = dict(context_size=256, patch_size=2, in_channels=4, hidden_size=1152, depth=1, num_heads=16, mlp_ratio=4.0,
                class_dropout_prob=0.1, num_classes=1000, learn_sigma=False, use_sit=True, use_swiglu=True, rel_pos_embed="rope",
                adaln_type="lora", adaln_lora_dim=288)
    cases = []
    for name, extra in (("a", dict(norm_type="rmsnorm", q_norm="rmsnorm", k_norm="layernorm", qk_norm_weight=True)),
                        ("b", dict(norm_type="w_layernorm", q_norm=None, k_norm="rmsnorm"))):
        cfg = O.FiTConfig(hidden_size=1152, depth=1, num_heads=16, adaln_lora_dim=288, **extra)
        m, sd = build(FiT, {**base, **extra}, cfg, perturb=True)
        x, grid, mask = mixed_padded_batch(cfg, [(10, 20), (16, 16), (8, 24)], 256, seed=13)
        t = torch.tensor([0.2, 0.6, 1.0])
        y = torch.tensor([1, 1000, 500])
        ref = m(x, t, y=y, grid=grid, mask=mask, size=None)
        got = O.forward(cfg, sd, x, t, y, grid, mask)
        assert torch.equal(ref, got), (name, (ref - got).abs().max())
        cases.append(dict(name=name, extra=extra, x=x, t=t, y=y, grid=grid, mask=mask, out=ref, weight_checksum=checksum(sd)))
        report.append(f"norm_variants_xl_d1[{name}] {extra}: init + forward, oracle bit-equal (|out|max {float(ref.abs().max()):.4f})")
    torch.save(cases, os.path.join(OUT, "norm_variants_xl_d1.pt"))

    with open(os.path.join(OUT, "README.md"), "a") as f:
        f.write("\n## model variants (oracle/make_variant_goldens.py, real reference)\n\n" + "\n".join(f"* {r}" for r in report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
