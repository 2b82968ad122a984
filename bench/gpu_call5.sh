#!/bin/bash
mkdir -p gpurun_out
echo "=== mega tests"; timeout 300 python -m pytest -m gpu -x -q -p no:cacheprovider tests/test_gpu_mega.py 2>&1 | tail -25 | cut -c1-600
echo "=== engine tests"; timeout 400 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_engine.py 2>&1 | tail -8 | cut -c1-300
echo "=== bench n1 (prof)"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_n1.json timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 --no-e2e > gpurun_out/r2_mega_n1p.json 2> gpurun_out/r2_mega_n1p.err; python -c "import json;d=json.load(open(\"gpurun_out/r2_mega_n1p.json\"));print(d[\"ms_per_step\"])"; tail -3 gpurun_out/r2_mega_n1p.err; python bench/mega_timeline.py gpurun_out/mega_prof_n1.json
echo "=== bench n1"; timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2_mega_n1.json 2> gpurun_out/r2_mega_n1.err; python -c "import json;d=json.load(open(\"gpurun_out/r2_mega_n1.json\"));print(d[\"ms_per_step\"], d[\"e2e\"][\"ms_per_step\"], d[\"value\"])"; tail -3 gpurun_out/r2_mega_n1.err
echo "=== bench n1 flags protocol (what N>1 runs), no fused opt"; SPARKTORCH_MEGA_FLAGS=1 timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d[\"ms_per_step\"])"
