"""CTA-pair GEMM (tcgen05.mma.cta_group::2, csrc/gemm2cta_sm100.cu): numerics against an fp32 reference, then device-timed
throughput against the single-CTA kernels and cuBLAS (torch.matmul).  Writes gpurun_out/gemm2cta.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparktorch_b200.ops import native  # noqa: E402

C = native.load(required=True)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
res = {"numerics": [], "perf": []}
for (M, N, K, bias, relu) in [(256, 256, 64, False, False), (512, 512, 256, False, False), (1000, 640, 200, True, True), (2048, 1024, 4096 + 64, True, False)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bv = torch.randn(N, device=dev) if bias else None
    o = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    rc = C.gemm_bf16_2cta(a, b, o, bv, 0.5, relu)
    torch.cuda.synchronize()
    ref = 0.5 * (a.float() @ b.float().t())
    if bias:
        ref = ref + bv
    if relu:
        ref = ref.clamp_min(0)
    err = (o.float() - ref).abs().max().item()
    tol = 2e-2 * ref.abs().max().item() + 1e-2
    ok = rc == 0 and bool(torch.isfinite(o.float()).all()) and err <= tol
    res["numerics"].append({"M": M, "N": N, "K": K, "rc": rc, "max_err": err, "tol": tol, "ok": ok})
    print("numerics", M, N, K, "rc", rc, "max_err %.4f tol %.4f" % (err, tol), "OK" if ok else "FAIL", flush=True)
if not all(r["ok"] for r in res["numerics"]):
    json.dump(res, open("gpurun_out/gemm2cta.json", "w"), indent=1)
    sys.exit(1)


def timed(fn, iters=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (M, N, K) in [(8192, 8192, 8192), (4096, 4096, 4096), (16384, 4096, 1024), (2048, 2048, 2048), (42000, 256, 784)]:
    nbuf = max(2, int(300e6 // ((M * K + N * K) * 2)) + 1)          # operands rotate through > L2-sized storage
    As = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    Bs = [torch.randn(N, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    o = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
    fl = 2 * M * N * K / 1e9
    rec = {"M": M, "N": N, "K": K}
    rec["pair_ms"] = timed(lambda i: C.gemm_bf16_2cta(As[i % nbuf], Bs[i % nbuf], o, None, 1.0, False))
    C.set_persistent_mode(1)
    if hasattr(C, "set_pair_mode"):
        C.set_pair_mode(0)
    rec["single_ms"] = timed(lambda i: C.gemm_bf16(As[i % nbuf], False, Bs[i % nbuf], False, M, N, K, 1.0, None, False, None, o, None, None, None, False, None, 256, 1))
    if hasattr(C, "set_pair_mode"):
        C.set_pair_mode(1)
    rec["cublas_ms"] = timed(lambda i: torch.matmul(As[i % nbuf], Bs[i % nbuf].t(), out=o))
    for k in ("pair", "single", "cublas"):
        rec[k + "_tflops"] = fl / rec[k + "_ms"]
    rec["pair_vs_cublas"] = rec["cublas_ms"] / rec["pair_ms"]
    res["perf"].append(rec)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}, flush=True)
    del As, Bs
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm2cta.json", "w"), indent=1)
