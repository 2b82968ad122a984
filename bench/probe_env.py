"""Environment probe for the GPU box (written to gpurun_out/)."""
import ctypes
import os
import subprocess

import torch

print("torch", torch.__version__, "cuda", torch.version.cuda, "ndev", torch.cuda.device_count())
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    print(i, p.name, p.major, p.minor, p.multi_processor_count, p.total_memory >> 20, "MiB")
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
try:
    lib = ctypes.CDLL("libcuda.so.1")
    print("libcuda ok", lib)
except OSError as e:
    print("libcuda missing", e)
try:
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29555")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    t = symm.empty(1024, dtype=torch.float32, device="cuda:0")
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print("symm_mem world=1 ok: ptrs", h.buffer_ptrs, "signal", h.signal_pad_ptrs, "mc", getattr(h, "multicast_ptr", None))
    dist.destroy_process_group()
except Exception as e:
    print("symm_mem probe failed:", repr(e))
print("cpu count", os.cpu_count())
print(subprocess.run("free -g | head -2; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv", shell=True, capture_output=True, text=True).stdout)
