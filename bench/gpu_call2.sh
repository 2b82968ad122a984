#!/bin/bash
# K-STEP first light: new tests (each under its own timeout), then the engine tests, then bench + task timeline
mkdir -p gpurun_out
echo "=== mega tests"
for t in test_one_step_matches_the_multi_kernel_graph test_many_steps_adam_match_the_multi_kernel_graph test_tracks_eager_fp32_with_odd_shapes test_early_stop_random_sampling_and_streaming_paths; do
  timeout 150 python -m pytest -m gpu -x -q -p no:cacheprovider tests/test_gpu_mega.py -k "$t" 2>&1 | tail -25 | cut -c1-400
done
echo "=== engine tests"; timeout 400 python -m pytest -m gpu -q -p no:cacheprovider tests/test_gpu_engine.py 2>&1 | tail -15 | cut -c1-300
echo "=== bench n1"; SPARKTORCH_MEGA_PROF=gpurun_out/mega_prof_n1.json timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2_mega_n1.json 2> gpurun_out/r2_mega_n1.err; python -c "import json;d=json.load(open(\"gpurun_out/r2_mega_n1.json\"));print(d[\"ms_per_step\"], d[\"e2e\"][\"ms_per_step\"])"; tail -3 gpurun_out/r2_mega_n1.err; python bench/mega_timeline.py gpurun_out/mega_prof_n1.json
