#!/bin/bash
# What the end of round 2 left unrun (GPU budget spent) - run these first, in this order.
#   gpurun --timeout 900 -- 'bash bench/next_round_checks.sh 1'        (one GPU)
#   gpurun --gpus 8 --timeout 400 -- 'bash bench/next_round_checks.sh 8'
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  echo "=== implicit-GEMM convolution (TMA im2col): numerics with convention diagnostics, then timing vs explicit im2col+GEMM and cuDNN"
  timeout 300 python bench/conv_implicit_check.py 2>&1 | tail -25
  echo "=== its GPU test (opt-in until the line above is green)"
  SPARKTORCH_IMPLICIT_CONV=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k implicit 2>&1 | tail -3
  echo "=== full GPU suite (last full run predates the CTA-pair dispatch in gemm_bf16)"
  timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
  echo "=== headline bench, driver arguments"
  timeout 200 python bench.py --steps 20 --warmup 5 | tail -1 | cut -c1-400
  echo "=== ResNet-50 bf16 step with the CTA-pair GEMM in the forward convolutions"
  timeout 400 python bench.py --model resnet50 --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
else
  tr() { port=$1; shift; timeout ${T:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
  echo "=== start-up of the Spark stand-in at $N ranks, 3 times (the round-1/2 hang: profiles/hang_8gpu_startup.md)"
  for rep in 1 2 3; do T=120 tr $((29800+rep)) bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>gpurun_out/start_$rep.err | tail -1 | cut -c1-200; done
  echo "=== headline bench at $N GPUs with the final K-STEP build"
  T=150 tr 29810 bench.py --gpus $N --steps 200 --warmup 20 | tail -1 | cut -c1-600
fi
