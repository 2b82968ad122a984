"""tcgen05 GEMM kernel vs cuBLAS (torch.matmul) – device-timed, inputs larger than L2 are
rotated between iterations.  Writes gpurun_out/gemm_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparktorch_b200.ops import native  # noqa: E402

C = native.load(required=True)
C.set_pair_mode(0)         # this script measures the single-CTA kernels; bench/gemm2cta_check.py measures the CTA-pair kernel
dev = torch.device("cuda", 0)
PEAK = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops": 1590.0}
out = []
SHAPES = [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 4096, 1024), (2048, 2048, 2048), (42000, 256, 784), (256, 256, 784)]
for (M, N, K) in SHAPES:
    nbuf = max(2, int(300e6 // ((M * K + N * K) * 2)) + 1)
    As = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    Bs = [torch.randn(N, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    o = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    rec = {"M": M, "N": N, "K": K}
    for bn, mode in ((128, 0), (256, 0), (128, 2), (256, 2)):
        C.set_persistent_mode(mode)
        tag = f"ours_bn{bn}" + ("_persistent" if mode else "")

        def ours(i):
            C.gemm_bf16(As[i % nbuf], False, Bs[i % nbuf], False, M, N, K, 1.0, None, False, None, o, None, None, None, False, None, bn, 1)
        for i in range(3):
            ours(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for i in range(iters):
            ours(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rec[f"{tag}_ms"] = ms
        rec[f"{tag}_tflops"] = 2 * M * N * K / ms / 1e9
    C.set_persistent_mode(1)
    for i in range(3):
        torch.matmul(As[i % nbuf], Bs[i % nbuf].t(), out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        torch.matmul(As[i % nbuf], Bs[i % nbuf].t(), out=o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    rec["cublas_ms"] = ms
    rec["cublas_tflops"] = 2 * M * N * K / ms / 1e9
    rec["best_ours_tflops"] = max(v for k, v in rec.items() if k.startswith("ours") and k.endswith("tflops"))
    rec["frac_of_measured_peak"] = rec["best_ours_tflops"] / PEAK["bf16_tflops"]
    out.append(rec)
    print(rec, flush=True)
    del As, Bs
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_bench.json", "w"), indent=1)
