#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over CI-sized launches of every tensor-core kernel.
# usage (on the GPU box): bash tools/sanitize.sh   -> logs under gpurun_out/sanitize_*.log
TESTS="tests/test_gpu_bf16.py::test_bf16_batched_equals_looped \
tests/test_gpu_bf16.py::test_decoder_group_mode_bit_identical[S-groups0] \
tests/test_gpu_bf16.py::test_decoder_pair_mode_bit_identical[S-9] \
tests/test_gpu_bf16.py::test_fused_postnet_stack_matches_layer_by_layer \
tests/test_gpu_scale.py::test_decoder_two_super_tiles_in_flight[S-1-1152] \
tests/test_gpu_conv_img.py::test_conv_img_image_epilogue[700-256-256-5] \
tests/test_gpu_conv_img.py::test_conv_img_layernorm_epilogues[900-384-384-True] \
tests/test_gpu_conv_img.py::test_conv_img_blocked_epilogues[900-256-1024-True] \
tests/test_teacher_forced.py::test_forward_teacher_forced_against_reference_golden[S_tf_n40_drop-fp16]"
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest $TESTS -x -q -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " gpurun_out/sanitize_$tool.log | tail -5
done
