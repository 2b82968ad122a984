mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -2
bash bench/gpu_ci.sh 2>&1 | tail -30
echo "=== ours n1 x2"
for i in 1 2; do timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 50 2>&1 | tail -1; done
echo "=== multi_gpu_check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > gpurun_out/mgc.log 2>&1; echo rc=$?; tail -25 gpurun_out/mgc.log
echo "=== ours n2 push"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2000 --warmup 50 2>&1 | tail -1
echo "=== ours n2 twoshot"
SPARKTORCH_EXCHANGE=twoshot timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 2000 --warmup 50 2>&1 | tail -1
echo "=== cnn n1"
timeout 300 python bench.py --gpus 1 --steps 500 --warmup 20 --model cnn 2>&1 | tail -1
echo "=== cnn n2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 500 --warmup 20 --model cnn 2>&1 | tail -1
echo "=== hogwild n2 dnn"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 2 --steps 1000 --warmup 50 --mode hogwild 2>&1 | tail -1
