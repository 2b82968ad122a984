"""Where does a tiny GEMM spend its time?  Per-CTA SM-clock timestamps of the kernel phases
(L2-warm, as inside the training step).  Writes gpurun_out/gemm_phases.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparktorch_b200.ops import native  # noqa: E402

C = native.load(required=True)
dev = torch.device("cuda", 0)
out = {}
CASES = {"fwd1_256x256x784": (256, 256, 784, False, False), "fwd2_256x256x256": (256, 256, 256, False, False),
         "head_256x10x256": (256, 10, 256, False, False), "dgrad2_mnB": (256, 256, 256, False, True),
         "wgrad1_mnAB": (256, 784, 256, True, True)}
for name, (M, N, K, a_mn, b_mn) in CASES.items():
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).to(torch.bfloat16)
    o = torch.zeros(M, (N + 15) // 16 * 16, dtype=torch.bfloat16, device=dev)[:, :N]
    bias = torch.zeros(N, device=dev)
    prof = torch.zeros(4096, 8, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(5):
        prof.zero_()
        torch.cuda.synchronize()
        e0.record()
        C.gemm_bf16(A, a_mn, B, b_mn, M, N, K, bias=bias, relu=True, out_bf16=o, prof=prof)
        e1.record()
        torch.cuda.synchronize()
    p = prof[(prof[:, 0] != 0)].cpu()
    t0 = p[:, 0].min()
    rel = (p - p[:, :1]).float()
    names = ["start", "prologue_done", "deps_met", "first_tile_landed", "accum_complete", "epilogue_done", "dealloc_done"]
    rec = {"ctas": int(p.shape[0]), "event_us": e0.elapsed_time(e1) * 1e3,
           "cta_start_spread_cycles": float((p[:, 0].max() - t0)),
           "span_cycles_first_start_to_last_end": float(p[:, 6].max() - t0)}
    for i, nme in enumerate(names):
        rec[nme + "_median_cycles"] = float(rel[:, i].median())
    out[name] = rec
    print(name, rec, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_phases.json", "w"), indent=1)
