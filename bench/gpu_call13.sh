#!/bin/bash
mkdir -p gpurun_out
echo "=== all gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -12 | cut -c1-400
echo "=== bench n1"; timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2_mega_n1.json 2> gpurun_out/r2_mega_n1.err; python -c "import json;d=json.load(open(\"gpurun_out/r2_mega_n1.json\"));print(d[\"ms_per_step\"], d[\"e2e\"][\"ms_per_step\"], d[\"value\"], d[\"loss\"])"; tail -3 gpurun_out/r2_mega_n1.err
echo "=== reference n1"; timeout 200 python bench.py --impl reference --gpus 1 --steps 200 --warmup 20 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d.get('ms_per_step'), d.get('loss'), d.get('unavailable'))"
echo "=== reference hogwild cnn n1"; timeout 200 python bench.py --impl reference --model cnn --mode hogwild --gpus 1 --steps 50 --warmup 10 2>&1 | tail -2 | cut -c1-600
echo "=== ours hogwild cnn n1"; timeout 200 python bench.py --model cnn --mode hogwild --gpus 1 --steps 50 --warmup 10 2>&1 | tail -2 | cut -c1-600
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
