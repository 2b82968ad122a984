# usage: bash bench/trip.sh N   – multi-GPU validation + scaling numbers on one box
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 | sort | uniq -c
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_n$N.txt
tr() { port=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
echo "=== multi_gpu_check N=$N"
tr 29611 tests/multi_gpu_check.py > gpurun_out/mgc_n$N.log 2>&1; echo rc=$?; grep -E "OK|Error|assert" gpurun_out/mgc_n$N.log | head -8
echo "=== pytest test_gpu_multi"
timeout 600 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_multi.py > gpurun_out/ci_multi_n$N.log 2>&1; echo rc=$?; tail -3 gpurun_out/ci_multi_n$N.log
echo "=== reference dnn N=$N"
tr 29612 bench.py --impl reference --gpus $N --steps 300 --warmup 20 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ref_dnn_n$N.json | cut -c1-400
echo "=== ours dnn N=$N"
tr 29613 bench.py --gpus $N --steps 2000 --warmup 50 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ours_dnn_n$N.json | cut -c1-400
echo "=== ours dnn N=1 (same box)"
timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 50 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ours_dnn_n1_box$N.json | cut -c1-300
echo "=== reference cnn N=$N"
tr 29614 bench.py --impl reference --model cnn --gpus $N --steps 200 --warmup 20 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ref_cnn_n$N.json | cut -c1-400
echo "=== ours cnn N=$N"
tr 29615 bench.py --gpus $N --steps 500 --warmup 20 --model cnn 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ours_cnn_n$N.json | cut -c1-400
echo "=== ours cnn hogwild N=$N"
tr 29616 bench.py --gpus $N --steps 500 --warmup 20 --model cnn --mode hogwild 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ours_cnn_hw_n$N.json | cut -c1-400
echo "=== ours dnn hogwild N=$N"
tr 29617 bench.py --gpus $N --steps 1000 --warmup 50 --mode hogwild 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/ours_dnn_hw_n$N.json | cut -c1-400
echo "=== allreduce sweep N=$N"
tr 29618 bench/allreduce_sweep.py > gpurun_out/ar_sweep_n$N.log 2>&1; echo rc=$?; tail -24 gpurun_out/ar_sweep_n$N.log
echo "=== kar profile N=$N"
tr 29619 bench/kar_profile.py > gpurun_out/kar_n$N.log 2>&1; echo rc=$?; tail -12 gpurun_out/kar_n$N.log
echo "=== resnet50 fused N=$N"
tr 29620 bench/resnet50_bench.py --gpus $N --steps 10 --warmup 3 --batch 32 --engine fused 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/resnet_fused_n$N.json | cut -c1-400
echo "=== resnet50 eager N=$N"
tr 29621 bench/resnet50_bench.py --gpus $N --steps 10 --warmup 3 --batch 32 --engine eager 2>gpurun_out/stderr_last.log | grep '^{' | tail -1 | tee gpurun_out/resnet_eager_n$N.json | cut -c1-400
