#!/bin/bash
# 2 GPUs: K-STEP multi-GPU correctness + timing
N=${1:-2}
mkdir -p gpurun_out
tr() { port=$1; shift; timeout ${T:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
echo "=== kstep multi check"; T=300 tr 29720 tests/multi_gpu_kstep.py > gpurun_out/kstep_multi_n$N.log 2>&1; echo "rc=$?"; grep -E "kstep|KSTEP|Error|error|assert" gpurun_out/kstep_multi_n$N.log | head -20 | cut -c1-400
echo "=== bench n$N"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_n$N.json T=240 tr 29730 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r2_mega_n$N.json 2> gpurun_out/r2_mega_n$N.err; python -c "import json;d=json.load(open(\"gpurun_out/r2_mega_n$N.json\"));print(d[\"ms_per_step\"], d[\"e2e\"][\"ms_per_step\"], d[\"value\"], d[\"config\"][\"comm\"])"; tail -3 gpurun_out/r2_mega_n$N.err | cut -c1-300; python bench/mega_timeline.py gpurun_out/mega_prof_n$N.json | head -22
echo "=== bench n$N legacy"; SPARKTORCH_MEGA=0 T=240 tr 29740 bench.py --gpus $N --steps 200 --warmup 20 --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d[\"ms_per_step\"])"
