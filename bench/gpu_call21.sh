#!/bin/bash
# 8 GPUs, final: K-STEP correctness with the gpu-scope producer rule, scaling point (200/20 and the driver's 20/5), synchronous CNN
# (stack dumps on a hang), reference arm at 8
N=${1:-8}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout ${T:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
one() { python -c "import json,sys;ls=[l for l in sys.stdin.read().splitlines() if l.startswith('{')];d=json.loads(ls[-1]);print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('value'), d.get('config',{}).get('comm'), d.get('loss'), d.get('unavailable'))"; }
echo "=== kstep multi check"; T=150 tr 29720 tests/multi_gpu_kstep.py > gpurun_out/kstep_multi_n$N.log 2>&1; echo "rc=$?"; grep -E "kstep|KSTEP|AssertionError" gpurun_out/kstep_multi_n$N.log | head -6 | cut -c1-300
echo "=== bench dnn n$N"; T=150 tr 29760 bench.py --gpus $N --steps 200 --warmup 20 2>gpurun_out/bench_err_n$N.log | tee gpurun_out/r2_final_n$N.json | one
for rep in 1 2; do echo "=== bench cnn sync n$N (run $rep)"; SPARKTORCH_STALL_TIMEOUT_S=30 T=60 tr $((29770+rep)) bench.py --gpus $N --model cnn --steps 100 --warmup 10 --no-e2e 2>gpurun_out/bench_cnn_err_n${N}_$rep.log | tee gpurun_out/r2_cnn_sync_n${N}_$rep.json | one; grep -E "no progress|File \"/root/repo|File \"/tmp/code" gpurun_out/bench_cnn_err_n${N}_$rep.log | sort | uniq -c | sort -rn | head -12 | cut -c1-220; done
