#!/bin/bash
# N GPUs: K-STEP with the chained head at world > 1, epoch flags vs cross-GPU counters (SPARKTORCH_MEGA_XCNT)
N=${1:-2}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout ${T:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
one() { python -c "import json,sys;ls=[l for l in sys.stdin.read().splitlines() if l.startswith('{')];d=json.loads(ls[-1]);print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('value'), d.get('config',{}).get('comm'), d.get('loss'))"; }
echo "=== kstep multi check (flags)"; T=300 tr 29720 tests/multi_gpu_kstep.py > gpurun_out/kstep_multi_n$N.log 2>&1; echo "rc=$?"; grep -E "kstep|KSTEP|AssertionError" gpurun_out/kstep_multi_n$N.log | head -6 | cut -c1-300
echo "=== kstep multi check (xcnt)"; SPARKTORCH_MEGA_XCNT=1 T=300 tr 29725 tests/multi_gpu_kstep.py > gpurun_out/kstep_multi_xcnt_n$N.log 2>&1; echo "rc=$?"; grep -E "kstep|KSTEP|AssertionError|Error" gpurun_out/kstep_multi_xcnt_n$N.log | head -6 | cut -c1-300
echo "=== bench dnn n$N (flags)"; T=240 tr 29760 bench.py --gpus $N --steps 200 --warmup 20 2>gpurun_out/bench_err_n$N.log | tee gpurun_out/r2_chain_n$N.json | one
echo "=== bench dnn n$N (xcnt)"; SPARKTORCH_MEGA_XCNT=1 T=240 tr 29762 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>>gpurun_out/bench_err_n$N.log | tee gpurun_out/r2_chain_xcnt_n$N.json | one
echo "=== bench dnn n$N (xcnt, 20/5)"; SPARKTORCH_MEGA_XCNT=1 T=240 tr 29764 bench.py --gpus $N --steps 20 --warmup 5 2>>gpurun_out/bench_err_n$N.log | one
echo "=== timeline (xcnt)"; SPARKTORCH_MEGA_XCNT=1 SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_xcnt_n$N.json T=240 tr 29766 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>/dev/null | one; python bench/mega_timeline.py gpurun_out/mega_prof_xcnt_n$N.json | sed -n 3,22p
