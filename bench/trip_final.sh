mkdir -p gpurun_out
echo "=== resnet50 reference arm b32"; timeout 240 python bench/resnet50_bench.py --impl reference --batch 32 --steps 5 --warmup 2 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/resnet_ref_n1.json | cut -c1-300
echo "=== resnet50 fused b32"; timeout 240 python bench/resnet50_bench.py --batch 32 --steps 10 --warmup 3 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/resnet_fused_n1.json | cut -c1-300
echo "=== inference ours cuda"; timeout 200 python bench/inference_bench.py --impl ours --rows 500000 --device cuda 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/inference_ours_cuda.json | cut -c1-300
echo "=== bench.py default"; timeout 300 python bench.py 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ours_dnn_n1_final.json | cut -c1-400
