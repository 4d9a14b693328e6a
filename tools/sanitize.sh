#!/bin/bash
# compute-sanitizer over a reduced -m gpu selection: one small case per kernel family (SURVEY.md §5 "race detection / sanitizers").
#   tools/sanitize.sh [memcheck|initcheck|racecheck|synccheck ...]      -> gpurun_out/r2_sanitizer_<tool>.txt
# initcheck runs with FITV2_POISON_WORKSPACE=empty so that the scratch memory really is uninitialised until a kernel writes it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='test_cfg_euler_bit_exact and 3-200 or test_cfg_combine or test_gemm_matches_fp32_reference and 256-288-128 or test_attention_matches_oracle and (3-100-72 or 2-77-96 or 2-256-72-16-True) or test_forward_golden_padded or test_conditioning_tensor_pipe_linears and width0-4 or test_forward_other_widths and width0 or test_transposed_residual_gemm and 3-10-20 or test_forward_many_rows and (7-1-1 or 5-7-9 or 12-12-24) or test_unpatchify_scale_bit_exact and 2-1-3 or test_pack_uint8 and 3-8-24 or test_online_rope_forward or test_fitv1_config_golden or test_norm_variants or test_final_layer_tensor_pipe or test_out_of_range_label or test_sde_kernels or test_ode_kernels'
for tool in "${@:-memcheck initcheck racecheck}"; do
  extra=""; poison=1
  [ "$tool" = initcheck ] && { extra="--track-unused-memory no"; poison=empty; }
  out=gpurun_out/r2_sanitizer_$tool.txt
  echo "# compute-sanitizer --tool $tool $extra python -m pytest tests -m gpu -q -x --tb=line -k \"$SEL\"   (FITV2_POISON_WORKSPACE=$poison)" > $out
  FITV2_POISON_WORKSPACE=$poison timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 30 --error-exitcode 0 \
      python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "$SEL" 2>&1 | grep -v "^$" | tail -120 >> $out
  echo "# exit $?" >> $out
done
