mkdir -p gpurun_out
timeout 600 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_resnet.py -s > gpurun_out/ci_resnet.log 2>&1; echo rc=$?; grep -E "passed|failed|cos |Error" gpurun_out/ci_resnet.log | head -120
for b in 32 128; do echo "=== resnet50 fused batch $b"; timeout 600 python bench/resnet50_bench.py --gpus 1 --steps 10 --warmup 3 --batch $b --engine fused 2>&1 | tail -1 | cut -c1-400; done
