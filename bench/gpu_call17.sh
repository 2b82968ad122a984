#!/bin/bash
# 1 GPU: quiet chain heads + early-A first layer + fewer epilogue syncs
mkdir -p gpurun_out
one() { python -c "import json,sys;ls=[l for l in sys.stdin.read().splitlines() if l.startswith('{')];d=json.loads(ls[-1]);print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('value'), d.get('loss'))"; }
echo "=== mega tests"; timeout 300 python -m pytest tests/test_gpu_mega.py -x -q -m gpu 2>&1 | tail -4
echo "=== bench n1"; timeout 200 python bench.py --steps 200 --warmup 20 2>gpurun_out/bench_err_n1.log | tee gpurun_out/r2_chain2_n1.json | one
echo "=== bench n1 driver-like (20/5)"; timeout 200 python bench.py --steps 20 --warmup 5 2>>gpurun_out/bench_err_n1.log | tee gpurun_out/r2_chain2_n1_s20.json | one
echo "=== no early A"; SPARKTORCH_MEGA_EARLY_A=0 timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e 2>>gpurun_out/bench_err_n1.log | one
echo "=== no chain"; SPARKTORCH_MEGA_CHAIN=0 timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e 2>>gpurun_out/bench_err_n1.log | one
echo "=== timeline"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_chain2_n1.json timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e 2>/dev/null | one; python bench/mega_timeline.py gpurun_out/mega_prof_chain2_n1.json > gpurun_out/kstep_timeline_chain2_n1.md; head -24 gpurun_out/kstep_timeline_chain2_n1.md
echo "=== engine+swap tests"; timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_swap.py -q -m gpu -x 2>&1 | tail -4
