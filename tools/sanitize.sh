#!/bin/bash
# compute-sanitizer over the mbarrier / TMEM / TMA / cluster kernels (SURVEY.md §5 race detection): memcheck,
# racecheck and synccheck on a small test selection that launches every kernel family once
# (1-CTA and 2-CTA implicit GEMM, fused resblock pair incl. its CTA-pair variant, tensor-core attention, single-output-channel conv, AdaIN ring / TMA / cluster,
# attention, BiLSTM H=128 / H=256 cluster, MAS).  Logs -> gpurun_out/r02_sanitizer_<tool>.log
set -u
mkdir -p gpurun_out
SEL='test_conv_igemm_1d and shape1 or test_resblock_pair and shape0 and mid or test_resblock_pair and shape3 and stage_end or test_resblock_pair and shape1 and branch_end or test_resblock_pair and shape2 and mid or test_conv_igemm_2cta_pairs and shape0 or test_conv_single_output_channel and shape2 or test_adain_norm_fused and shape0 or test_adain_norm_fused and shape5 and False or test_relpos_attention and 37 or test_conformer_attention and 50 or test_bilstm and 5-60 or test_repeat_and_length_regulate or test_mas_edge_cases or test_round_durations_kernel'
for tool in memcheck racecheck synccheck; do
  timeout 480 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py tests/test_serving_gpu.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$tool.log | tail -3
done
