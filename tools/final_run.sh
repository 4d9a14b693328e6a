set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err
for w in xl160x320 3b256 xl512; do timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/f_bench_$w.log 2> gpurun_out/f_bench_$w.err; done
for w in xl256 xl160x320 3b256 xl512; do timeout 200 python tools/profile_breakdown.py $w > gpurun_out/f_bd_$w.log 2>&1; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 810 -c 540 --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 2 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attention|ln_modulate|cond_tc" -s 16 -c 9 -o gpurun_out/prof_r1_final python tools/ncu_target.py xl256 2 > gpurun_out/f_ncu_full.log 2>&1
tail -2 gpurun_out/f_pytest.log; tail -1 gpurun_out/f_smoke.log; cut -c1-300 gpurun_out/f_bench.log
