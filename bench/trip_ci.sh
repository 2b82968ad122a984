mkdir -p gpurun_out
bash bench/gpu_ci.sh 2>&1 | tail -40
for f in gpurun_out/ci_*.log; do grep -E "^E  |^FAILED" $f | head -6; done
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
