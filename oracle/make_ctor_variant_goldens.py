"""Golden fixtures for the remaining constructor switches of fit.model.fit_model.FiT (fit_model.py:25-65), generated from
the REAL reference classes (build container only; TEST INFRASTRUCTURE).  No shipped config sets these switches, but the class
DEFAULTS do (`FiT(hidden_size=..., depth=..., num_heads=...)` is a GELU-Mlp, learn_sigma, (B, C, N), adaLN 'normal' model).

  ctor_variants_xl.pt   four XL-width models, each on a mixed-aspect padded batch:
      defaults      every keyword at its class default except hidden_size / depth (2) / num_heads: timm Mlp with tanh-GELU
                    (use_swiglu=False, modules.py:253), learn_sigma, use_sit=False, adaln_type 'normal', no q / k norm
      nobias_norope FiTv2 layout with qkv_bias=False, ffn_bias=False (modules.py:140,248-253) and rel_pos_embed=None
                    (modules.py:153,170: q / k are not rotated)
      rope_v        FiTv2 layout with add_rel_pe_to_v=True (modules.py:171-172) and rel_pos_embed='XPOS' (lower-cased, same
                    rotation as 'rope')
      adaln_swiglu  FiTv2 layout with adaln_type='swiglu' (modules.py:265-268,284-285: SwiGLU modulation MLPs on c, no global term)

`adaln_bias=False` cannot be constructed in the reference (initialize_weights calls nn.init.constant_ on the missing bias,
fit_model.py:141-153), so there is nothing to pin and the drop-in keeps rejecting it.

State dicts are regenerated from seeds on the test side; inputs, reference outputs and weight checksums are stored.
Usage:  python oracle/make_ctor_variant_goldens.py
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fitv2_oracle as O                                            # noqa: E402
from oracle.make_golden import install_reference, mixed_padded_batch, checksum  # noqa: E402
from oracle.make_variant_goldens import build                                   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
V2 = dict(context_size=256, patch_size=2, in_channels=4, hidden_size=1152, depth=1, num_heads=16, mlp_ratio=4.0,
          class_dropout_prob=0.1, num_classes=1000, learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm",
          k_norm="layernorm", rel_pos_embed="rope", adaln_type="lora", adaln_lora_dim=288)
O2 = dict(hidden_size=1152, depth=1, num_heads=16, adaln_lora_dim=288)

# name -> (reference kwargs, oracle config kwargs)
CASES = {
    "defaults": (dict(hidden_size=1152, depth=2, num_heads=16),
                 dict(hidden_size=1152, depth=2, num_heads=16, adaln_lora_dim=0, learn_sigma=True, use_sit=False, adaln_type="normal",
                      q_norm=None, k_norm=None, use_swiglu=False)),
    "nobias_norope": ({**V2, "qkv_bias": False, "ffn_bias": False, "rel_pos_embed": None},
                      {**O2, "qkv_bias": False, "ffn_bias": False, "rel_pos_embed": None}),
    "rope_v": ({**V2, "add_rel_pe_to_v": True, "rel_pos_embed": "XPOS"},
               {**O2, "add_rel_pe_to_v": True, "rel_pos_embed": "XPOS"}),
    "adaln_swiglu": ({**V2, "adaln_type": "swiglu", "adaln_lora_dim": None},
                     {**O2, "adaln_type": "swiglu", "adaln_lora_dim": 0}),
}


def main():
    torch.set_grad_enabled(False)
    FiT = install_reference()
    report, cases = [], []
    for name, (kw, okw) in CASES.items():
        cfg = O.FiTConfig(**okw)
        m, sd = build(FiT, kw, cfg)
        x, grid, mask = mixed_padded_batch(cfg, [(10, 20), (16, 16), (8, 24), (20, 10)], 256, seed=17)
        if not cfg.use_sit:
            x = x.transpose(1, 2).contiguous()                                  # (B, C, N)
        t = torch.tensor([0.1, 0.5, 0.9, 0.3])
        y = torch.tensor([7, 1000, 999, 1000])
        ref = m(x, t, y=y, grid=grid, mask=mask, size=None)
        got = O.forward(cfg, sd, x, t, y, grid, mask)
        assert torch.equal(ref, got), (name, (ref - got).abs().max())
        refc = m.forward_with_cfg(x, t, y, grid, mask, None, 1.5)
        assert torch.equal(refc, O.forward_with_cfg(cfg, sd, x, t, y, grid, mask, None, 1.5)), name
        cases.append(dict(name=name, kwargs=kw, oracle_kwargs=okw, x=x, t=t, y=y, grid=grid, mask=mask, out=ref, out_cfg=refc,
                          keys=list(sd.keys()), weight_checksum=checksum(sd)))
        report.append(f"ctor_variants_xl[{name}]: init + forward + forward_with_cfg, oracle bit-equal "
                      f"({len(sd)} tensors, |out|max {float(ref.abs().max()):.4f})")
    torch.save(cases, os.path.join(OUT, "ctor_variants_xl.pt"))
    with open(os.path.join(OUT, "README.md"), "a") as f:
        f.write("\n## remaining constructor switches (oracle/make_ctor_variant_goldens.py, real reference)\n\n"
                + "\n".join(f"* {r}" for r in report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
