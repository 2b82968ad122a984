"""Multi-GPU correctness script (run under torch.distributed.run, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 tests/multi_gpu_check.py

Checks, against NCCL / single-process PyTorch oracles:
  1. peer-memory all-reduce kernel == NCCL all_reduce
  2. K-AR (fused reduce-scatter + Adam + publish) == average-then-torch.optim
  3. fused MLP fit: replicas stay bit-identical, loss decreases, in-band early stop agrees
Prints 'MULTI_GPU_OK' on rank 0 on success.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{int(os.environ['MASTER_PORT']) + 17}", rank=rank,
                            world_size=world, device_id=dev)
    from sparktorch_b200.ops import native
    from sparktorch_b200.parallel.p2p import FusedUpdate, SymmBlock
    from sparktorch_b200.runtime.flat import TAIL
    from sparktorch_b200.runtime.optim import parse_optimizer

    C = native.load(required=True)

    # ---- 1. plain all-reduce over peer memory ---------------------------------------------------
    for n in (1024, 1 << 20, (1 << 22) + 4096):
        blk = SymmBlock(dev, world, rank)
        blk.reserve("buf", n * 4)
        blk.reserve("flags", C.flag_words() * 4)
        blk.rendezvous()
        buf = blk.view("buf", torch.float32, n)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        src = torch.randn(n, generator=g).to(dev)
        buf.copy_(src)
        ref = src.clone()
        dist.all_reduce(ref)
        state = torch.zeros(C.step_state_bytes() // 4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        for it in range(4):                     # repeated launches exercise the epoch protocol; odd ones go through NVLS
            buf.copy_(src)
            torch.cuda.synchronize()
            dist.barrier()
            mc = blk.multicast_ptr("buf") if it % 2 else 0
            C.allreduce_sum(blk.peer_ptrs("buf"), blk.peer_ptrs("flags"), state, world, rank, n, mc)
            torch.cuda.synchronize()
            torch.testing.assert_close(buf, ref, rtol=1e-5, atol=1e-5)
            dist.barrier()
        if rank == 0:
            print(f"allreduce n={n} ok; multicast={'yes' if blk.multicast_base else 'no'}", flush=True)

    # ---- 2. K-AR vs oracle --------------------------------------------------------------------------
    for cls, kw in ((torch.optim.Adam, dict(lr=1e-2)), (torch.optim.SGD, dict(lr=0.1, momentum=0.9))):
        spec = parse_optimizer(cls, kw)
        rows, cols = 37, 24
        numel = 960                               # 888 padded to a multiple of 64
        fu = FusedUpdate(numel, spec, world, rank, dev, patience=-1, max_iters=16, wbf16_elems=rows * cols)
        fu.build([[0, rows, cols, 0, cols, -1, 0, 1]])
        torch.manual_seed(0)
        p_ref = torch.nn.Parameter(torch.randn(rows, cols, device=dev))
        fu.param[: rows * cols].copy_(p_ref.detach().reshape(-1))
        opt = cls([p_ref], **kw)
        torch.cuda.synchronize()
        dist.barrier()
        for t in range(5):
            g = torch.randn(rows, cols, generator=torch.Generator().manual_seed(1000 * t + rank)).to(dev)
            fu.grad[: rows * cols].copy_(g.reshape(-1))
            fu.grad[numel] = float(rank + t)       # loss slot
            gsum = g.clone()
            dist.all_reduce(gsum)
            p_ref.grad = gsum / world
            opt.step()
            fu.updater.step()
            torch.cuda.synchronize()
            torch.testing.assert_close(fu.param[: rows * cols].view(rows, cols), p_ref.detach(), rtol=3e-5, atol=3e-6)
            torch.testing.assert_close(fu.wbf16.view(rows, cols).float(), p_ref.detach().to(torch.bfloat16).float(), rtol=0, atol=0)
            assert abs(fu.last_loss() - (sum(range(world)) / world + t)) < 1e-5, fu.last_loss()
            assert float(fu.grad.abs().max()) == 0.0
            dist.barrier()
        if rank == 0:
            print(f"K-AR {cls.__name__} ok (nvls={'on' if fu.use_multicast else 'off'})", flush=True)

    # ---- 3. fused MLP fit across ranks --------------------------------------------------------------
    from sparktorch_b200.models import mnist_dnn
    from sparktorch_b200.runtime.engine import TrainConfig, train_partition

    rng = np.random.default_rng(7 + rank)
    X = (rng.integers(0, 256, size=(2048, 784)) * (rng.random((2048, 784)) < 0.2)).astype(np.float32) / 64
    Wt = np.random.default_rng(99).normal(size=(784, 10)).astype(np.float32)
    Y = (X @ Wt).argmax(1).astype(np.float32).reshape(-1, 1)
    torch.manual_seed(1234 + rank)                # deliberately different init per rank: rank 0 must win
    cfg = TrainConfig(iters=120, mini_batch=256, device="cuda", seed=5, engine="fused")
    res = train_partition(rank, world, mnist_dnn(), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 2e-3}, X, Y, cfg, dev)
    h = res.history
    assert h["engine"] == "fused_mlp" and h["iters_run"] == 120, h
    assert np.mean(h["train_loss"][-10:]) < 0.7 * np.mean(h["train_loss"][:3]), h["train_loss"][::10]
    losses = torch.tensor(h["train_loss"], device=dev)
    gathered = [torch.zeros_like(losses) for _ in range(world)]
    dist.all_gather(gathered, losses)
    for gl in gathered:
        assert torch.equal(gl, gathered[0]), "ranks disagree on the distributed loss history"
    # one-shot push exchange (opt-in): replicas agree to rounding, training converges the same way
    os.environ["SPARKTORCH_EXCHANGE"] = "push"
    try:
        torch.manual_seed(1234 + rank)
        resp = train_partition(rank, world, mnist_dnn(), nn.CrossEntropyLoss(), torch.optim.Adam, {"lr": 2e-3}, X, Y,
                               TrainConfig(iters=120, mini_batch=256, device="cuda", seed=5, engine="fused"), dev)
    finally:
        os.environ.pop("SPARKTORCH_EXCHANGE", None)
    assert resp.history["comm"] == "P2PPush", resp.history["comm"]
    lp = torch.tensor(resp.history["train_loss"], device=dev)
    gp = [torch.zeros_like(lp) for _ in range(world)]
    dist.all_gather(gp, lp)
    for g in gp:
        assert torch.allclose(g, gp[0], rtol=1e-3, atol=1e-4), "push exchange: ranks drifted apart"
    assert torch.allclose(lp, losses, rtol=5e-2, atol=5e-3), (lp[-5:], losses[-5:])
    cfg = TrainConfig(iters=40, mini_batch=256, device="cuda", seed=5, engine="fused", early_stop_patience=2)
    res = train_partition(rank, world, mnist_dnn(), nn.CrossEntropyLoss(), torch.optim.SGD, {"lr": 0.0}, X, Y, cfg, dev)
    stops = torch.tensor([res.history["stopped_at"] if res.history["stopped_at"] is not None else -1], device=dev)
    allst = [torch.zeros_like(stops) for _ in range(world)]
    dist.all_gather(allst, stops)
    assert all(int(s) == int(allst[0]) and int(s) >= 2 for s in allst), allst
    # ---- 4. hogwild over peer memory: no barriers between workers, remote optimizer steps -------------
    from sparktorch_b200.parallel.hogwild_p2p import train_partition_hogwild

    for lock in (True, False):
        cfg = TrainConfig(iters=120, mini_batch=128, device="cuda", seed=9)
        # W asynchronous workers apply W optimizer steps per "iteration": keep the effective step size constant
        hw = train_partition_hogwild(rank, world, mnist_dnn(), nn.CrossEntropyLoss(), torch.optim.Adam,
                                     {"lr": 2e-3 / world}, X, Y, cfg, dev, acquire_lock=lock, window_len=world)
        hh = hw.history
        # asynchronous workers must LEARN at every world size (no per-world thresholds): final loss < 0.7 x initial
        assert np.isfinite(hh["train_loss"]).all() and np.mean(hh["train_loss"][-8:]) < 0.7 * np.mean(hh["train_loss"][:3]), hh["train_loss"][::10]
        if rank == 0:
            assert hh["pushes_total"] == 120 * world, hh["pushes_total"]
            assert all(torch.isfinite(v).all() for v in hw.state_dict.values())
            print(f"hogwild lock={lock} ok: pushes={hh['pushes_total']} stripes={hh['stripes']} loss {hh['train_loss'][0]:.3f}->{hh['train_loss'][-1]:.3f}", flush=True)
        dist.barrier()
    from sparktorch_b200.parallel.hogwild_p2p import train_partition_hogwild_fused

    for lock in (True, False):           # acquireLock=True is the reference's default: it runs the fused K-AG worker as well
        cfg = TrainConfig(iters=200, mini_batch=256, device="cuda", seed=9, sampling="sweep")
        hf = train_partition_hogwild_fused(rank, world, mnist_dnn(), nn.CrossEntropyLoss(), torch.optim.Adam,
                                           {"lr": 2e-3 / world}, X, Y, cfg, dev, window_len=world, acquire_lock=lock)
        assert hf is not None and hf.history["iters_run"] == 200
        tl = hf.history["train_loss"]
        assert np.isfinite(tl).all() and np.mean(tl[-10:]) < 0.7 * np.mean(tl[:3]), tl[::20]
        if rank == 0:
            assert hf.history["pushes_total"] == 200 * world
            print(f"fused hogwild (K-AG, lock={lock}) ok: {tl[0]:.3f}->{tl[-1]:.3f} launches/step {hf.history['gpu_launches_per_step']} "
                  f"{1e6 * hf.history['loop_seconds'] / 200:.1f} us/iter comm {hf.history['comm']}", flush=True)
        dist.barrier()
    if rank == 0:
        print("fused fit ok:", [round(v, 3) for v in h["train_loss"][::20]], "comm", h["comm"], flush=True)
        print("MULTI_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
