mkdir -p gpurun_out
timeout 300 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_norm.py tests/test_gpu_resnet.py tests/test_gpu_engine.py -k "epilogue or resnet or hogwild_fused or predictor or persistent" > gpurun_out/ci_b.log 2>&1; echo rc=$?
grep -E "^E  |passed|failed|FAILED" gpurun_out/ci_b.log | head -30
echo "=== resnet50 fused b32"; timeout 300 python bench/resnet50_bench.py --gpus 1 --steps 10 --warmup 3 --batch 32 --engine fused 2>&1 | tail -1 | cut -c1-200
echo "=== resnet50 fused b128"; timeout 300 python bench/resnet50_bench.py --gpus 1 --steps 10 --warmup 3 --batch 128 --engine fused 2>&1 | tail -1 | cut -c1-200
