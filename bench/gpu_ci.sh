#!/bin/bash
# Run on the GPU box: every test group in its own process with its own timeout so a
# hung kernel in one group cannot hide the results of the others.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 420 python -m pytest -m gpu -q -p no:cacheprovider "$@" > gpurun_out/ci_$name.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/ci_$name.log | grep -E "passed|failed|error|Error|rc=" | tail -5; }
run gemm_k tests/test_gpu_kernels.py -k "gemm and (False-False or block_n or deep_k or forward_epilogue or k_ag) and not persistent"
run gemm_mn tests/test_gpu_kernels.py -k "gemm and not (False-False or block_n or deep_k or forward_epilogue or k_ag) and not persistent"
run misc tests/test_gpu_kernels.py -k "not gemm"
run persist tests/test_gpu_kernels.py -k "persistent"
run conv tests/test_gpu_kernels.py -k "im2col or conv or maxpool"
run engine tests/test_gpu_engine.py
run resnet tests/test_gpu_resnet.py tests/test_gpu_norm.py
