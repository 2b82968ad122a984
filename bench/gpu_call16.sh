#!/bin/bash
# 1 GPU: validate the chained head / window split / warp-uniform GEMM / multi-step streaming, then profile
mkdir -p gpurun_out
one() { python -c "import json,sys;ls=[l for l in sys.stdin.read().splitlines() if l.startswith('{')];d=json.loads(ls[-1]);print(d.get('ms_per_step'), d.get('e2e',{}).get('ms_per_step'), d.get('value'), d.get('loss'), (d.get('step_kernel') or {}).get('groups'))"; }
echo "=== mega tests"; timeout 300 python -m pytest tests/test_gpu_mega.py -x -q -m gpu 2>&1 | tail -4
echo "=== bench n1 (chain + split)"; timeout 200 python bench.py --steps 200 --warmup 20 2>gpurun_out/bench_err_n1.log | tee gpurun_out/r2_chain_n1.json | one
echo "=== bench n1 driver-like (20/5)"; timeout 200 python bench.py --steps 20 --warmup 5 2>>gpurun_out/bench_err_n1.log | tee gpurun_out/r2_chain_n1_s20.json | one
echo "=== bench n1 no chain"; SPARKTORCH_MEGA_CHAIN=0 timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e 2>>gpurun_out/bench_err_n1.log | tee gpurun_out/r2_nochain_n1.json | one
echo "=== bench n1 no chain, no split"; SPARKTORCH_MEGA_CHAIN=0 SPARKTORCH_MEGA_EPI_SPLIT=1 timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e 2>>gpurun_out/bench_err_n1.log | one
echo "=== timeline"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_chain_n1.json timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e 2>/dev/null | one; python bench/mega_timeline.py gpurun_out/mega_prof_chain_n1.json > gpurun_out/kstep_timeline_chain_n1.md; head -24 gpurun_out/kstep_timeline_chain_n1.md
echo "=== gpu tests (all)"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "=== gemm bench"; timeout 300 python bench/gemm_bench.py > gpurun_out/gemm_bench_r2.log 2>&1; tail -25 gpurun_out/gemm_bench_r2.log
echo "=== ncu k-step"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_step_kernel -c 1 -f -o gpurun_out/kstep_r2 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/ncu_kstep.log 2>&1; tail -3 gpurun_out/ncu_kstep.log; ls -la gpurun_out/kstep_r2.ncu-rep
