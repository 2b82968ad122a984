"""In-tree build of the native extension (``sparktorch_b200/_C*.so``).

Kernels are plain CUDA translation units (no torch headers → seconds per file)
compiled for sm_100a only:

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 ...

``bindings.cpp`` is the only file that sees torch/pybind11.  Objects land in
``csrc/_build`` (git-ignored); the shared object sits next to the Python
package so it travels with the source tree to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["gemm_sm100.cu", "gemm2cta_sm100.cu", "loss.cu", "data.cu", "comm_ar.cu", "comm_hw.cu", "conv.cu", "norm.cu", "mlp_step.cu"]
CPP_SOURCES = ["bindings.cpp"]
MODULE = "_C"


def so_path() -> str:
    return os.path.join(PKG, MODULE + sysconfig.get_config_var("EXT_SUFFIX"))


def _digest(paths: List[str], extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _run(cmd: List[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if cmd[0] == NVCC and "-o" in cmd:       # keep ptxas' resource report next to the object (registers, spills, stack)
        with open(cmd[cmd.index("-o") + 1] + ".ptxas.txt", "w") as fh:
            fh.write(r.stdout + r.stderr)
    if verbose and (r.stdout.strip() or r.stderr.strip()):
        print(r.stdout + r.stderr)


def build_all(verbose: bool = False, force: bool = False) -> str:
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(BUILD, exist_ok=True)
    cu = [s for s in CU_SOURCES if os.path.exists(os.path.join(HERE, s))]
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    all_src = [os.path.join(HERE, s) for s in cu + CPP_SOURCES] + headers
    stamp = os.path.join(BUILD, "stamp.txt")
    digest = _digest(all_src, torch.__version__ + sys.version)
    out = so_path()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    nvcc_flags = ARCH + ["-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                         "-Xptxas", "-v"]
    objs = []
    jobs = []
    for s in cu:
        o = os.path.join(BUILD, s + ".o")
        objs.append(o)
        jobs.append([NVCC] + nvcc_flags + ["-I", HERE, "-c", os.path.join(HERE, s), "-o", o])
    inc = []
    for p in ce.include_paths() + [os.path.join(CUDA_HOME, "include"), sysconfig.get_paths()["include"], HERE]:
        inc += ["-I", p]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    for s in CPP_SOURCES:
        o = os.path.join(BUILD, s + ".o")
        objs.append(o)
        jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=" + MODULE,
                     "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-Wno-deprecated-declarations"]
                    + inc + ["-c", os.path.join(HERE, s), "-o", o])
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
        list(pool.map(lambda c: _run(c, verbose), jobs))
    libs = []
    for p in ce.library_paths() + [os.path.join(CUDA_HOME, "lib64")]:
        libs += ["-L", p, "-Wl,-rpath," + p]
    link = ["g++", "-shared", "-o", out] + objs + libs + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                                                           "-ltorch_python", "-lcudart"]
    _run(link, verbose)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return out


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
