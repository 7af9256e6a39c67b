#!/bin/bash
# compute-sanitizer over a reduced GPU-test subset: one case (or a few) per kernel family.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# (round 2, second run: includes the kernels written this round -- conv_rows with the TMEM ring, convt_rows, wgrad_tc, the cluster kernel se_tail, diff_ops)
# memcheck: out-of-bounds / misaligned global + shared accesses; racecheck: shared-memory hazards (the mbarrier / TMEM pipelines of
# conv_tc.cu and conv_rows.cu, the cp.async rings of plane_fir.cu, the TMA rings of blur_rows.cu); synccheck: barrier misuse.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_r02.txt
: > $OUT
SUBSET='test_upfirdn2d_golden_cases or test_upfirdn2d_asymmetric or test_fused_bias_act_golden or test_layout_roundtrip or test_modulation_and_demod or test_conv3x3_fused_epilogue or (test_conv3x3_tcgen05_raw and case0) or (test_stride2_pad1_conv and case0) or (test_grouped_stride2_conv and case0) or test_conv1x1_and_tap_sum or (test_blur_act and bfloat16) or (test_torgb and 34) or (test_alignnet_tail and 32) or (test_field_step and 32) or test_warp_mix or (test_mask_blend and 128) or (test_imgio_vs_oracle and shape0) or test_bicubic_up_add or test_alignnet_norm_kernels or test_conv3x3_prelu_epilogue or (test_conv3x3_row_sliding_kernel and case0) or (test_conv3x3_fused_torgb_epilogue and case0) or (test_se_gate_and_residual and case0) or (test_conv3x3_seeded_accumulator and case0) or (test_alignnet_split_front and 64) or (test_tap_sum_tiled and 70) or (test_conv3x3_fused_output_statistics and case0) or (test_conv_transposed_fused_phases and case0) or (test_conv3x3_encoder_epilogues and case0) or (test_strided_gather_conv and case0) or (test_blur_adjoint and hw0) or test_act_bwd_dot_torgb_bwd or (test_warp_mix_bwd and shape0) or (test_mask_blend_bwd and 64) or (test_field_step_bwd and 32) or (test_conv_transposed_row_streaming_kernel and case0) or (test_conv_transposed_split_into_exact_tiles and case2) or (test_se_tail_cluster_kernel and case0) or test_encoder_glue_kernels or (test_thumbnail_nhwc and hw1) or (test_conv1x1_stride2 and case0 and 0-) or (test_conv3x3_f16_storage and case0) or (test_conv_wgrad_tcgen05 and case0) or test_diff_ops_against_autograd or (test_spm_warp_gradient_vs_oracle_autograd and 32-12)'
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool" >> $OUT
  OOD_ROWS_MIN_STRIPS=1 timeout 1200 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py tests/test_zz_samm_bwd_gpu.py tests/test_alignment_grad_gpu.py -q -x -m gpu -k "$SUBSET" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|Race reported|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -30 >> $OUT
done
cat $OUT
