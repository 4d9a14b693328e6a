# One `ncu --set full` capture of every kernel family on the depth-2 target (tools/ncu_target.py): the first forward's launches
# (conditioning, patch-embed, one block, final layer) and the variant kernels at the end.  Two reports of <= 32 MiB each.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attention|ln_modulate|cond_tc|patch_embed|f32_to_f16" -s 13 -c 13 -o gpurun_out/prof_r2_block -f python tools/ncu_target.py xl256 1 > gpurun_out/f_ncu_full_block.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"attention_general|transpose_inner|lincomb|scaled_rms|cfg_euler|gemm_tc_kernel<144, 5|gemm_tc_kernel<32" -c 8 -o gpurun_out/prof_r2_variants -f python tools/ncu_target.py xl256 1 > gpurun_out/f_ncu_full_variants.log 2>&1
ls -la gpurun_out/*.ncu-rep
