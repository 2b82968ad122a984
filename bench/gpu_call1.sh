#!/bin/bash
# round-2 first GPU call: sanity + baseline + sanitizer logs + pending test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/nvsmi_r2.txt 2>&1
echo "=== gpu tests"; timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
echo "=== pending test"; SPARKTORCH_PENDING_TESTS=1 timeout 200 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_engine.py -k "hogwild_fused_cnn" 2>&1 | tail -3
echo "=== bench n1"; timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2_base_n1.json 2> gpurun_out/r2_base_n1.err; tail -c 600 gpurun_out/r2_base_n1.json
echo "=== sanitizer"
for tool in memcheck racecheck synccheck; do
  echo "--- $tool"
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_kernels.py \
      -k "operand_layouts and 256-256-784 or fused_ce or softmax_ce or gather_batch or fused_update_matches or maxpool or im2col" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
