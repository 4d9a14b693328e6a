# Round-end evidence run on one B200 (gpurun): tests, smoke, bench (all named workloads + the batch-2 latency line), per-class
# breakdowns, the ncu launch list of the bench command and one `ncu --set full` capture of every kernel family.  Outputs: gpurun_out/f_*.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -m gpu -q -s --tb=short 2>&1 | tail -40 > gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.log 2> gpurun_out/f_bench_reference.err
for w in xl160x320 3b256 xl512; do timeout 300 python bench.py --workload $w --steps 40 --no-cpu-baseline > gpurun_out/f_bench_$w.log 2> gpurun_out/f_bench_$w.err; done
timeout 200 python bench.py --batch 2 --steps 100 --no-cpu-baseline > gpurun_out/f_bench_config1_eager.log 2>&1
timeout 200 python bench.py --batch 2 --steps 100 --no-cpu-baseline --cuda-graph > gpurun_out/f_bench_config1_graph.log 2>&1
for w in xl256 xl160x320 3b256 xl512; do timeout 200 python tools/profile_breakdown.py $w > gpurun_out/f_bd_$w.log 2>&1; done
timeout 200 python tools/attn_bench.py > gpurun_out/f_attn_bench.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 810 -c 540 --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 2 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
# (the full-set capture is a separate call, tools/final_ncu.sh: gpurun merges at most 64 MiB back)
tail -3 gpurun_out/f_pytest.log; tail -1 gpurun_out/f_smoke.log; cut -c1-300 gpurun_out/f_bench.log
