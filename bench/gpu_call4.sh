#!/bin/bash
mkdir -p gpurun_out
echo "=== swap tests"; timeout 300 python -m pytest -m gpu -x -q -p no:cacheprovider tests/test_gpu_swap.py 2>&1 | tail -25 | cut -c1-500
echo "=== bench n1"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_n1.json timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 --no-e2e > gpurun_out/r2_mega_n1.json 2> gpurun_out/r2_mega_n1.err; python -c "import json;d=json.load(open(\"gpurun_out/r2_mega_n1.json\"));print(d[\"ms_per_step\"])"; tail -3 gpurun_out/r2_mega_n1.err; python bench/mega_timeline.py gpurun_out/mega_prof_n1.json
