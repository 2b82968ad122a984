#!/bin/bash
# 8 GPUs: K-STEP correctness + scaling point, hogwild correctness at 8, synchronous CNN (round-1 hang), hogwild CNN both arms
N=${1:-8}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout ${T:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
one() { python -c "import json,sys;ls=[l for l in sys.stdin.read().splitlines() if l.startswith('{')];d=json.loads(ls[-1]);print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('value'), d.get('config',{}).get('comm'), d.get('loss'), d.get('unavailable'))"; }
echo "=== kstep multi check"; T=300 tr 29720 tests/multi_gpu_kstep.py > gpurun_out/kstep_multi_n$N.log 2>&1; echo "rc=$?"; grep -E "kstep|KSTEP|AssertionError" gpurun_out/kstep_multi_n$N.log | head -6 | cut -c1-300
echo "=== bench dnn n$N"; T=240 tr 29760 bench.py --gpus $N --steps 200 --warmup 20 2>gpurun_out/bench_err_n$N.log | tee gpurun_out/r2_mega_n$N.json | one
echo "=== bench dnn n$N prof"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_n$N.json T=240 tr 29765 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>/dev/null | one; python bench/mega_timeline.py gpurun_out/mega_prof_n$N.json | sed -n 3,14p
echo "=== multi_gpu_check (hogwild included)"; T=400 tr 29704 tests/multi_gpu_check.py > gpurun_out/mgc_r2_n$N.log 2>&1; echo "rc=$?"; grep -E " ok|OK|AssertionError" gpurun_out/mgc_r2_n$N.log | head -14 | cut -c1-300
for rep in 1 2 3; do echo "=== bench cnn sync n$N (run $rep)"; SPARKTORCH_STALL_TIMEOUT_S=45 T=150 tr $((29770+rep)) bench.py --gpus $N --model cnn --steps 100 --warmup 10 --no-e2e 2>gpurun_out/bench_cnn_err_n${N}_$rep.log | tee gpurun_out/r2_cnn_sync_n${N}_$rep.json | one; grep -E "no progress|Error" gpurun_out/bench_cnn_err_n${N}_$rep.log | head -3 | cut -c1-200; done
echo "=== bench cnn hogwild n$N ours"; T=200 tr 29780 bench.py --gpus $N --model cnn --mode hogwild --steps 100 --warmup 10 2>gpurun_out/bench_cnnhw_err_n$N.log | tee gpurun_out/r2_cnn_hw_n$N.json | one
echo "=== bench cnn hogwild n$N reference"; T=200 tr 29790 bench.py --impl reference --gpus $N --model cnn --mode hogwild --steps 100 --warmup 10 2>gpurun_out/bench_cnnhw_ref_err_n$N.log | tee gpurun_out/r2_cnn_hw_ref_n$N.json | one
echo "=== reference dnn n$N"; T=200 tr 29795 bench.py --impl reference --gpus $N --steps 200 --warmup 20 2>/dev/null | tee gpurun_out/r2_ref_n$N.json | one
