"""Gradient all-reduce bandwidth sweep: our peer-memory kernel vs torch.distributed (NCCL)
(BASELINE.md B4).  Run under torch.distributed.run with one rank per GPU:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29650 bench/allreduce_sweep.py

Device-timed with CUDA events, max over ranks; bus bandwidth = 2*(N-1)/N * bytes / t.
Writes gpurun_out/allreduce_sweep_N.json on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{int(os.environ['MASTER_PORT']) + 23}", rank=rank,
                            world_size=world, device_id=dev)
    from sparktorch_b200.ops import native
    from sparktorch_b200.parallel.p2p import SymmBlock

    C = native.load(required=True)
    max_bytes = int(os.environ.get("SWEEP_MAX_BYTES", 1 << 30))
    blk = SymmBlock(dev, world, rank)
    blk.reserve("buf", max_bytes)
    blk.reserve("flags", C.flag_words() * 4)
    blk.rendezvous()
    state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev)
    full = blk.view("buf", torch.float32, max_bytes // 4)
    nccl_buf = torch.zeros(max_bytes // 4, device=dev)
    out = []
    size = 1024
    while size <= max_bytes:
        n = size // 4
        iters = 200 if size <= (1 << 20) else (40 if size <= (1 << 26) else 10)
        p2p = timed(lambda: C.allreduce_sum(blk.peer_ptrs("buf"), blk.peer_ptrs("flags"), state, world, rank, n), iters, dev)
        mc = blk.multicast_ptr("buf")
        nvls = timed(lambda: C.allreduce_sum(blk.peer_ptrs("buf"), blk.peer_ptrs("flags"), state, world, rank, n, mc), iters,
                     dev) if mc else float("inf")
        ours = min(p2p, nvls)
        nccl = timed(lambda: dist.all_reduce(nccl_buf[:n]), iters, dev)
        bus = 2 * (world - 1) / world * size
        rec = {"bytes": size, "ours_us": ours * 1e3, "nccl_us": nccl * 1e3, "ours_p2p_us": p2p * 1e3,
               "ours_nvls_us": nvls * 1e3 if mc else None, "ours_busbw_GBs": bus / ours / 1e6,
               "nccl_busbw_GBs": bus / nccl / 1e6}
        out.append(rec)
        if rank == 0:
            print(rec, flush=True)
        size *= 4
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"world": world, "multicast": bool(blk.multicast_base), "sweep": out},
                  open(f"gpurun_out/allreduce_sweep_{world}.json", "w"), indent=1)
    del full
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
