# 8-GPU validation + scaling numbers (kept short: every second is charged 8x)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
tr() { port=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
js() { grep '^{' | tail -1 | tee gpurun_out/$1 | cut -c1-330; }
echo "=== multi_gpu_check N=$N"
tr 29611 tests/multi_gpu_check.py > gpurun_out/mgc_n$N.log 2>&1; echo rc=$?; grep -E "OK|Error|assert" gpurun_out/mgc_n$N.log | head -8
echo "=== ours dnn N=$N"
tr 29613 bench.py --gpus $N --steps 2000 --warmup 50 2>gpurun_out/stderr_last.log | js ours_dnn_n$N.json
echo "=== reference dnn N=$N"
tr 29612 bench.py --impl reference --gpus $N --steps 300 --warmup 20 2>gpurun_out/stderr_last.log | js ref_dnn_n$N.json
echo "=== ours cnn N=$N"
tr 29615 bench.py --gpus $N --steps 500 --warmup 20 --model cnn 2>gpurun_out/stderr_last.log | js ours_cnn_n$N.json
echo "=== reference cnn N=$N"
tr 29614 bench.py --impl reference --model cnn --gpus $N --steps 200 --warmup 20 2>gpurun_out/stderr_last.log | js ref_cnn_n$N.json
echo "=== ours cnn hogwild N=$N"
tr 29616 bench.py --gpus $N --steps 500 --warmup 20 --model cnn --mode hogwild 2>gpurun_out/stderr_last.log | js ours_cnn_hw_n$N.json
echo "=== allreduce sweep N=$N"
SWEEP_MAX_BYTES=$((1<<28)) tr 29618 bench/allreduce_sweep.py > gpurun_out/ar_sweep_n$N.log 2>&1; echo rc=$?; grep "bytes" gpurun_out/ar_sweep_n$N.log | cut -c1-330
echo "=== kar profile N=$N"
tr 29619 bench/kar_profile.py > gpurun_out/kar_n$N.log 2>&1; echo rc=$?; grep "per K-AR" gpurun_out/kar_n$N.log
echo "=== resnet50 fused N=$N"
tr 29620 bench/resnet50_bench.py --gpus $N --steps 10 --warmup 3 --batch 32 --engine fused 2>gpurun_out/stderr_last.log | js resnet_fused_n$N.json
