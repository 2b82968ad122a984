"""Cost of the fused all-reduce+optimizer kernel (K-AR) in isolation, multi-GPU (run under torch.distributed.run).
Back-to-back launches, device-timed, max over ranks; arena = MNIST-DNN (269 322 params) and a 25.6 M one."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{int(os.environ['MASTER_PORT']) + 29}", rank=rank,
                            world_size=world, device_id=dev)
    from sparktorch_b200.parallel.p2p import FusedUpdate
    from sparktorch_b200.runtime.optim import parse_optimizer

    spec = parse_optimizer(torch.optim.Adam, {"lr": 1e-3})
    out = {}
    for name, numel in (("mnist_dnn", 269376), ("resnet50", 25557056)):
        for nvls in (True, False):
            fu = FusedUpdate(numel, spec, world, rank, dev, patience=-1, max_iters=8, wbf16_elems=numel, use_multicast=nvls)
            fu.build([[0, 1, numel, 0, numel, -1, 0, 1]])
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(5):
                fu.updater.step()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 200 if numel < 1e6 else 30
            e0.record()
            for _ in range(iters):
                fu.updater.step()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            key = f"{name}_nvls{int(fu.use_multicast)}"
            out[key] = float(t[0])
            if rank == 0:
                print(key, f"{float(t[0]):.2f} us per K-AR launch (world {world})", flush=True)
            del fu
            dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/kar_profile_{world}.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
