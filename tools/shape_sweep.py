"""Shape sweep of the whole forward (depth 1) against the CPU oracle: many rows / few tokens, one token, ragged and odd
tile counts, padded mixed-aspect batches, both widths.  One-off robustness run (the oracle needs CPU time);
    python tools/shape_sweep.py  -> gpurun_out/shape_sweep.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FITV2_POISON_WORKSPACE", "1")
import torch
from oracle import fitv2_oracle as O                      # checker
from fitv2_b200 import FiT, make_grid

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 1)
KW = dict(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora")
WIDTHS = {"xl": dict(hidden_size=1152, num_heads=16, adaln_lora_dim=288), "3b": dict(hidden_size=2304, num_heads=24, adaln_lora_dim=576)}
OPERAND = sys.argv[1] if len(sys.argv) > 1 else "bf16"
CASES = [("xl", 2, 64, 64, "pad"), ("xl", 3, 40, 50, "mixed"), ("3b", 3, 32, 32, "pad"), ("xl", 64, 16, 16, "none"), ("xl", 33, 16, 16, "pad"), ("xl", 64, 4, 4, "none"), ("xl", 7, 1, 1, "none"), ("xl", 50, 3, 7, "pad"),
         ("xl", 5, 32, 32, "pad"), ("xl", 17, 12, 24, "mixed"), ("xl", 64, 10, 20, "none"), ("xl", 96, 8, 8, "mixed"), ("xl", 130, 2, 3, "none"),
         ("3b", 20, 8, 8, "pad"), ("3b", 6, 16, 16, "mixed"), ("3b", 30, 5, 5, "none")]
models = {}
res = []
for width, R, hp, wp, mk in CASES:
    if width not in models:
        torch.manual_seed(0)
        m = FiT(**KW, depth=1, operand_dtype=OPERAND, **WIDTHS[width]).randomize_zero_init_(1)
        sd = {k: v.detach().clone().float() for k, v in m.state_dict().items()}
        models[width] = (m.cuda().eval(), sd, O.FiTConfig(depth=1, **WIDTHS[width]))
    m, sd, cfg = models[width]
    N = hp * wp
    g = torch.Generator().manual_seed(R * 1000 + N)
    x = torch.randn(R, N, 16, generator=g); t = torch.rand(R, generator=g); y = torch.randint(0, 1001, (R,), generator=g)
    grid, mask = make_grid(R, hp, wp), torch.ones(R, N)
    if mk != "none":
        for r in range(R):
            if mk == "pad" and r % 2:
                mask[r, N - (r * 3) % max(N // 2, 1) - 1:] = 0
            elif mk == "mixed":
                if r % 3 == 0:
                    mask[r, N - (r * 5) % max(N // 2, 1) - 1:] = 0
                elif r % 3 == 1:
                    mask[r, : (r * 7) % max(N // 2, 1) + 1] = 2       # two packed images
        x = x * (mask != 0).float()[..., None]
    t0 = time.time()
    out = m(x.cuda(), t.cuda(), y.cuda(), grid.cuda(), mask.cuda())
    torch.cuda.synchronize()
    out2 = m(x.cuda(), t.cuda(), y.cuda(), grid.cuda(), mask.cuda()).cpu()
    out = out.cpu()
    ref = O.forward(cfg, sd, x, t, y, grid, mask)
    err = float((out - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    row = dict(width=width, rows=R, tokens=N, mask=mk, err=err, deterministic=bool(torch.equal(out, out2)),
               pad_zero=bool((out[mask == 0] == 0).all()), finite=bool(torch.isfinite(out).all()), seconds=round(time.time() - t0, 1))
    res.append(row)
    print(row, flush=True)
bad = [r for r in res if not (r["err"] < 1e-2 and r["deterministic"] and r["pad_zero"] and r["finite"])]
print("FAILED" if bad else "all ok", len(res), "cases; worst err", max(r["err"] for r in res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"shape_sweep_{OPERAND}.json"), "w") as f:
    json.dump(res, f, indent=1)
