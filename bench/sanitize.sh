#!/bin/bash
# compute-sanitizer targets for every kernel family (run on the GPU box; SURVEY.md §5 "race detection").
# memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards; synccheck: barrier misuse.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_kernels.py \
      -k "operand_layouts and 256-256-784 or fused_ce or softmax_ce or gather_batch or fused_update_matches or maxpool or im2col" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
