#!/bin/bash
# First GPU minutes of the next round (things written or left open after round 1's GPU budget ran out).
#   gpurun --gpus 8 --timeout 400 -- 'bash bench/next_round_checks.sh 8'
N=${1:-8}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout ${T:-120} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
echo "=== 1. open bug: synchronous CNN at $N GPUs - a stall now aborts with a thread dump after 45 s"
SPARKTORCH_STALL_TIMEOUT_S=45 T=150 tr 29701 bench.py --gpus $N --model cnn --steps 100 --warmup 10 --no-e2e > gpurun_out/cnn_sync_n$N.log 2>&1
echo "rc=$?"; grep -E "^\{|no progress|Error|error" gpurun_out/cnn_sync_n$N.log | head -20 | cut -c1-300
echo "    same without PDL / without NVLS (one of them passing localises the problem)"
SPARKTORCH_PDL=0 SPARKTORCH_STALL_TIMEOUT_S=45 T=150 tr 29702 bench.py --gpus $N --model cnn --steps 100 --warmup 10 --no-e2e 2>&1 | grep -E "^\{|no progress" | cut -c1-200
SPARKTORCH_NVLS=0 SPARKTORCH_STALL_TIMEOUT_S=45 T=150 tr 29703 bench.py --gpus $N --model cnn --steps 100 --warmup 10 --no-e2e 2>&1 | grep -E "^\{|no progress" | cut -c1-200
echo "=== 2. multi-GPU correctness script (thresholds for 8 asynchronous workers were set without a run)"
T=300 tr 29704 tests/multi_gpu_check.py > gpurun_out/mgc_next_n$N.log 2>&1; echo "rc=$?"; grep -E "OK|Error|assert" gpurun_out/mgc_next_n$N.log | head
echo "=== 3. pending single-GPU tests"
SPARKTORCH_PENDING_TESTS=1 timeout 200 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_engine.py -k "hogwild_fused_cnn" 2>&1 | tail -3
echo "=== 4. compute-sanitizer over the kernel tests (single GPU)"
timeout 600 bash bench/sanitize.sh 2>&1 | tail -15
