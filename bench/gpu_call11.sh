#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout ${T:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
one() { python -c "import json,sys;ls=[l for l in sys.stdin.read().splitlines() if l.startswith('{')];d=json.loads(ls[-1]);print(d['ms_per_step'], d.get('e2e',{}).get('ms_per_step'), d['config']['comm'])"; }
echo "=== kstep multi check"; T=300 tr 29720 tests/multi_gpu_kstep.py > gpurun_out/kstep_multi_n$N.log 2>&1; echo "rc=$?"; grep -E "kstep|KSTEP|Error|assert" gpurun_out/kstep_multi_n$N.log | head -12 | cut -c1-400
echo "=== bench n$N (prof, both ranks)"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_n$N.json T=240 tr 29730 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>/dev/null | one
python bench/mega_timeline.py gpurun_out/mega_prof_n$N.json | head -14; echo "--- rank 1"; python bench/mega_timeline.py gpurun_out/mega_prof_n$N.json.rank1 | sed -n 3,14p
echo "=== bench n$N NVLS off"; SPARKTORCH_NVLS=0 T=240 tr 29740 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>/dev/null | one
echo "=== bench n$N steps per launch 1"; SPARKTORCH_MEGA_STEPS=1 T=240 tr 29750 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>/dev/null | one
echo "=== bench n$N full (with e2e)"; T=240 tr 29760 bench.py --gpus $N --steps 200 --warmup 20 2>/dev/null | tee gpurun_out/r2_mega_n$N.json | one
