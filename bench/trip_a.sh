mkdir -p gpurun_out
bash bench/gpu_ci.sh 2>&1 | tail -40
grep -E "cos |worst" gpurun_out/ci_resnet.log | head -8
echo "=== resnet50 fused b32"; timeout 300 python bench/resnet50_bench.py --gpus 1 --steps 10 --warmup 3 --batch 32 --engine fused 2>&1 | tail -1 | cut -c1-300
echo "=== dnn n1"; timeout 200 python bench.py --gpus 1 --steps 2000 --warmup 50 --no-e2e 2>/dev/null | tail -1 | cut -c1-260
