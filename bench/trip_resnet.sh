mkdir -p gpurun_out
timeout 600 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_gpu_resnet.py -s > gpurun_out/ci_resnet.log 2>&1; echo rc=$?; tail -40 gpurun_out/ci_resnet.log
timeout 300 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_engine.py 2>&1 | tail -3
for e in fused eager; do echo "=== resnet50 $e"; timeout 600 python bench/resnet50_bench.py --gpus 1 --steps 10 --warmup 3 --batch 32 --engine $e 2>&1 | tail -4 | cut -c1-800; done
