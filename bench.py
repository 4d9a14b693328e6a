"""Benchmark of the FiTv2 denoising hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload xl256|xl160x320|xl512|3b256]

A "step" is one CFG Euler/ODE denoising step of the sampler loop of sample_fitv2_ddp.py:297-314 over one
batch of 32 samples per GPU: one FiT forward over 64 rows (cond + uncond) + the CFG combine + Euler update.
250 steps make one batch of images, so   images/sec = 32 * n_gpus / (250 * seconds_per_step).

Our arm:  `value`  = device-timed (CUDA events), latents resident in HBM;
          `e2e`    = the same loop driven through the public API with HOST (pinned) buffers: every step
                     copies that step's inputs host->device and reads the updated latents back device->host;
          `roofline` = the dominant kernel (fused gate/up SwiGLU GEMM) timed live with CUDA events on the
                     launching stream during the timed region, against the measured bf16 tensor peak;
          `cpu_baseline` = the CPU oracle (bit-equal restatement of the reference's PyTorch path) on the
                     host cores, bounded sample (N=1, rank 0 only).
Reference arm (`--impl reference`): the reference's CPU PyTorch path (oracle port; the reference needs a timm
shim + a constructor fix to run at all and does not travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    #            model kw                                                         grid (h, w) patches, rope kw
    "xl256":     (dict(hidden_size=1152, depth=36, num_heads=16, adaln_lora_dim=288), (16, 16), {}),
    "xl160x320": (dict(hidden_size=1152, depth=36, num_heads=16, adaln_lora_dim=288), (10, 20),
                  dict(custom_freqs="ntk-aware", max_pe_len_h=10, max_pe_len_w=20, decouple=True, ori_max_pe_len=16)),
    "xl512":     (dict(hidden_size=1152, depth=36, num_heads=16, adaln_lora_dim=288), (32, 32),
                  dict(custom_freqs="ntk-aware", max_pe_len_h=32, max_pe_len_w=32, decouple=True, ori_max_pe_len=16)),
    "3b256":     (dict(hidden_size=2304, depth=40, num_heads=24, adaln_lora_dim=576), (16, 16), {}),
}
NUM_SAMPLING_STEPS = 250
CFG_SCALE = 1.5
METRIC = "images/sec FiTv2-XL/2 256^2 250-step ODE CFG 1.5"
METRIC_BY_WORKLOAD = {
    "xl256": METRIC,                                                     # BASELINE.json metric (the headline, default)
    "xl160x320": "images/sec FiTv2-XL/2 160x320 (dynntk, decouple) 250-step ODE CFG 1.5",
    "xl512": "images/sec FiTv2-XL/2 512^2 (1024 tokens) 250-step ODE CFG 1.5",
    "3b256": "images/sec FiTv2-3B/2 256^2 250-step ODE CFG 1.5",
}


def flops_per_forward_row(kw, n_tokens):
    """2*MAC, attention included, per sample row."""
    D, L, lora, N = kw["hidden_size"], kw["depth"], kw["adaln_lora_dim"], n_tokens
    Hm = (int(D * 4.0) * 2) // 3
    mac = N * (L * (4 * D * D + 3 * D * Hm + 2 * N * D) + 32 * D) + L * (D * lora + 6 * D * lora) + 256 * D + D * D + 6 * D * D + 2 * D * D
    return 2.0 * mac


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"], hbm_gbs=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full summary."""
    p = os.path.join(ROOT, "profiles", "r2_gateup_ncu_full.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_gateup_ncu_full.json")
    if workload != "xl256" or not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get("traffic_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, pw = [], 0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm),
                    power_w=statistics.median(pw) if pw else None)


# --------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) — `cpu_baseline` leg and `--impl reference`
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(workload: str, steps: int, warmup: int, n_samples: int, budget_s: float):
    """Times CFG Euler steps of the reference's PyTorch path on the host cores (fp32, all threads)."""
    import torch
    from oracle import fitv2_oracle as O        # bench.py cpu legs are allowed to execute the oracle (task ④)
    kw, (hp, wp), rope_kw = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    cfg = O.FiTConfig(**kw, **rope_kw)
    sd = O.synthetic_state_dict(cfg)
    N = hp * wp
    torch.manual_seed(0)
    z = torch.randn(n_samples, N, 16)
    y = torch.randint(0, 1000, (n_samples,))
    grid, mask = O.make_grid(n_samples, hp, wp), torch.ones(n_samples, N)
    y2 = torch.cat([y, torch.full((n_samples,), 1000)])
    grid2, mask2 = torch.cat([grid, grid]), torch.cat([mask, mask])
    sig = torch.linspace(0, 1, NUM_SAMPLING_STEPS + 1)

    def one(idx, zz):
        v2 = O.forward(cfg, sd, torch.cat([zz, zz]), sig[idx].expand(2 * n_samples), y2, grid2, mask2)
        return O.cfg_euler_update(zz, v2, CFG_SCALE, sig[idx], sig[idx + 1])
    idx = 0
    for _ in range(warmup):
        z = one(idx % NUM_SAMPLING_STEPS, z); idx += 1
    times, t_start = [], time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        z = one(idx % NUM_SAMPLING_STEPS, z); idx += 1
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    sec = sum(times) / len(times)
    return dict(images_per_sec=n_samples / (NUM_SAMPLING_STEPS * sec), sec_per_step=sec, steps_timed=len(times),
                cores=torch.get_num_threads(), n_samples=n_samples)


def run_reference(args):
    from fitv2_b200.distributed import dist_env
    rank, _, world = dist_env()
    if rank != 0:
        return                      # rank 0 alone runs and prints the CPU arm
    n_samples = 1
    r = cpu_reference_run(args.workload, steps=args.steps, warmup=min(args.warmup, 1), n_samples=n_samples, budget_s=args.cpu_budget)
    kw, (hp, wp), _ = WORKLOADS[args.workload]
    sample = (f"{r['steps_timed']} CFG Euler steps of the {NUM_SAMPLING_STEPS}-step trajectory at batch {n_samples} "
              f"({2 * n_samples} model rows x {hp * wp} tokens), fp32, {r['cores']} threads; images/s = batch / (250 * s_per_step)")
    line = dict(impl="reference", metric=METRIC_BY_WORKLOAD[args.workload], value=r["images_per_sec"], unit="images/sec", n_gpus=args.gpus, steps=r["steps_timed"],
                steps_requested=args.steps, warmup=min(args.warmup, 1), ms_per_step=r["sec_per_step"] * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="fp32", data="synthetic",
                config=dict(workload=f"FiTv2-{'3B' if args.workload == '3b256' else 'XL'}/2 {args.workload}, {NUM_SAMPLING_STEPS}-step ODE, CFG {CFG_SCALE}",
                            per_gpu_batch=n_samples, rows=2 * n_samples, tokens=hp * wp, device="host CPU"),
                cpu_baseline=dict(value=r["images_per_sec"], unit="images/sec", cores=r["cores"], kind="port", sample=sample),
                e2e=dict(value=r["images_per_sec"], unit="images/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                nfe_per_sec=1.0 / r["sec_per_step"], gpu_launches=0)
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fitv2_b200 import FiT, EulerCFGSampler, make_grid, _lib
    from fitv2_b200.distributed import init_process_group, draw_rank_inputs, gather_latents, max_over_ranks

    rank, local_rank, world = init_process_group("nccl")
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device: fitv2_b200 has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    kw, (hp, wp), rope_kw = WORKLOADS[args.workload]
    N, n = hp * wp, args.batch
    torch.manual_seed(0)                                  # reference init seed
    model = FiT(learn_sigma=False, use_sit=True, use_swiglu=True, q_norm="layernorm", k_norm="layernorm", adaln_type="lora",
                operand_dtype=args.operand, **kw, **rope_kw).randomize_zero_init_(seed=1).to(dev).eval()
    z_host, y_host = draw_rank_inputs(0, world, rank, n, N, 16, 1000)
    grid, mask = make_grid(n, hp, wp), torch.ones(n, N)
    smp = EulerCFGSampler(model, y_host.to(dev), grid.to(dev), mask.to(dev), NUM_SAMPLING_STEPS, CFG_SCALE, use_cuda_graph=args.cuda_graph)
    z = z_host.to(dev).contiguous()
    lib = _lib.load()

    if args.cuda_graph:
        smp.capture(z)                                    # the timed loop replays the captured step (scalars copied per step)

    def step(i):
        if args.cuda_graph:
            smp.replay_step(i % NUM_SAMPLING_STEPS)
        else:
            smp._step(z, smp.t_table[i % NUM_SAMPLING_STEPS], smp.dsig[(i % NUM_SAMPLING_STEPS):(i % NUM_SAMPLING_STEPS) + 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident timing (`value`) ----------------
    for i in range(args.warmup):
        step(i)
    if not args.cuda_graph:
        model.profile(["gateup_gemm"])                    # dominant kernel, event-timed during the timed region (eager launches only)
    l0 = model.kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for i in range(args.steps):
            step(args.warmup + i)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev)
    launches = model.kernel_launches() - l0 + args.steps      # + one cfg_euler kernel per step
    if args.cuda_graph:                                       # replayed launches are not counted by the handle: the captured step's count
        launches = args.steps * (smp.launches_per_step + 1)
        prof = {"gateup_gemm": (0.0, 0)}
    else:
        prof = model.profile_read()
        model.profile(None)
    ms_step = ms_total / args.steps
    img_s = n * world / (NUM_SAMPLING_STEPS * ms_step / 1e3)

    # ---------------- end-to-end through the public API with host buffers (`e2e`) ----------------
    pin = lambda t: t.contiguous().pin_memory()
    hz, ht = pin(z_host.clone()), pin(torch.zeros(2 * n))
    hy, hgrid, hmask = pin(smp.y2.cpu()), pin(smp.grid2.cpu()), pin(smp.mask2.cpu())
    hout = pin(torch.empty(n, N, 16))
    sig = smp.sigmas
    h2d = sum(t.numel() * t.element_size() for t in (hz, ht, hy, hgrid, hmask))
    d2h = hout.numel() * hout.element_size()

    def e2e_step(i):
        idx = i % NUM_SAMPLING_STEPS
        ht.fill_(float(sig[idx]))
        dz = hz.to(dev, non_blocking=True)
        dt, dy = ht.to(dev, non_blocking=True), hy.to(dev, non_blocking=True)
        dg, dm = hgrid.to(dev, non_blocking=True), hmask.to(dev, non_blocking=True)
        v2 = model(torch.cat([dz, dz], 0), dt, dy, dg, dm)                     # the reference-facing call (FiT.forward)
        _lib.check(lib.fitv2_cfg_euler(dz.data_ptr(), v2.data_ptr(), CFG_SCALE, float(sig[idx + 1] - sig[idx]), None, n, N, 16,
                                       torch.cuda.current_stream(dev).cuda_stream))
        hout.copy_(dz, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()                          # the host consumes the result every step
        hz.copy_(hout)
    e2e_steps = max(3, min(args.steps, args.e2e_steps or args.steps))
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps, dev)
    e2e_img_s = n * world / (NUM_SAMPLING_STEPS * e2e_ms / 1e3)

    # ---------------- the one collective of the path: final latents all-gather (outside the step loop) ----------------
    gathered = gather_latents(z)
    torch.cuda.synchronize(dev)

    if rank != 0:
        return
    peaks = measured_peaks()
    fl_row = flops_per_forward_row(kw, N)
    fl_step = fl_row * 2 * n
    D, Hm = kw["hidden_size"], (int(kw["hidden_size"] * 4.0) * 2) // 3
    gu_ms, gu_cnt = prof["gateup_gemm"]
    gu_flops = 2.0 * (2 * n * N) * D * (2 * Hm)           # algorithmic FLOPs of one fused gate/up GEMM launch
    gu_tflops = gu_flops / (gu_ms / max(gu_cnt, 1) * 1e-3) / 1e12 if gu_cnt else None
    line = dict(
        metric=METRIC_BY_WORKLOAD[args.workload], value=img_s, unit="images/sec", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.operand, data="synthetic",
        config=dict(workload=f"FiTv2-{'3B' if args.workload == '3b256' else 'XL'}/2 {args.workload}: {NUM_SAMPLING_STEPS}-step ODE, CFG {CFG_SCALE}, "
                             f"{n} samples/GPU ({2 * n} model rows x {N} tokens), random-init weights, synthetic noise",
                    per_gpu_batch=n, rows=2 * n, tokens=N, parallelism=f"dp{world} (independent trajectories per GPU, no collective in the step loop)",
                    l2="working set per step (1.3 GB weights + activations) far exceeds the 126 MB L2; no flush needed",
                    accumulate="fp32 (TMEM), fp32 residual/LayerNorm/softmax/conditioning", cuda_graph=bool(args.cuda_graph)),
        nfe_per_sec=world * 1e3 / ms_step, model_tflops=fl_step / (ms_step * 1e-3) / 1e12,
        frac_of_tensor_peak_sustained=fl_step / (ms_step * 1e-3) / 1e12 / peaks["tflops_sustained"],
        e2e=dict(value=e2e_img_s, unit="images/sec", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms, steps=e2e_steps),
        gpu_launches=int(launches),
        roofline=dict(bound="tensor", kernel="gemm_tc_kernel<256, EPI_SWIGLU> (fused gate/up GEMM + SwiGLU epilogue)",
                      achieved=gu_tflops, peak=peaks["tflops_sustained"], unit="TFLOP/s",
                      frac=(gu_tflops / peaks["tflops_sustained"]) if gu_tflops else None, traffic=ncu_traffic(args.workload),
                      peak_source=peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                      launches_timed=gu_cnt, us_per_launch=(gu_ms / gu_cnt * 1e3) if gu_cnt else None),
        clocks=clocks.summary(),
        gathered_latents=list(gathered.shape),
    )
    # the step runs power-capped: time ~ energy / cap, so J/step is the quantity a kernel change has to lower
    pw = line["clocks"].get("power_w")
    line["energy_j_per_step"] = (pw * ms_step / 1e3) if pw else None
    line["energy_j_per_image"] = (pw * ms_step / 1e3 * NUM_SAMPLING_STEPS / n) if pw else None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.workload, steps=3, warmup=1, n_samples=1, budget_s=args.cpu_budget)
        line["cpu_baseline"] = dict(value=r["images_per_sec"], unit="images/sec", cores=r["cores"], kind="port",
                                    sample=f"{r['steps_timed']} CFG Euler steps at batch 1 (2 rows x {N} tokens) of the same model, fp32 oracle "
                                           f"(bit-equal port of the reference PyTorch path), {r['sec_per_step']:.2f} s/step")
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="xl256", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=32, help="samples per GPU (the script's --per-proc-batch-size)")
    ap.add_argument("--operand", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--cuda-graph", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the end-to-end leg (0 = as many as --steps, the same power/thermal regime)")
    ap.add_argument("--cpu-budget", type=float, default=120.0, help="seconds of CPU work allowed for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
