"""In-tree build of libsdpnet_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "lib", "libsdpnet_b200.so")
TORCH_LIB = os.path.join(HERE, "lib", "libsdpnet_b200_torch.so")     # TORCH_LIBRARY wrappers over the C-ABI (csrc/torch_ops.cpp)
SOURCES = ["gemm_tc.cu", "gemm_simt.cu", "norm.cu", "dwconv.cu", "dwconv_slab.cu", "attention.cu", "attention_tc.cu", "preprocess.cu", "forward.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "sdpnet_b200.h")]
    nvcc = _nvcc()

    def compile_one(src: str):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            r = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
            with open(o + ".log", "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(f"compiled {src}")
        return o

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"linked {LIB}")
    build_torch_ops(force, verbose)
    return LIB


def build_torch_ops(force: bool = False, verbose: bool = False) -> str:
    """g++ build of the thin `torch.ops.sdpnet_b200.*` wrappers against the installed torch headers; links the C-ABI
    library next to it (rpath $ORIGIN)."""
    src = os.path.join(CSRC, "torch_ops.cpp")
    if not (force or _stale(TORCH_LIB, [src, LIB, os.path.join(ROOT, "include", "sdpnet_b200.h")])):
        return TORCH_LIB
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + ["-I/usr/local/cuda/include"]
    tlib = ce.library_paths()[0]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + inc + \
          [src, "-o", TORCH_LIB, f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
           f"-L{os.path.dirname(LIB)}", "-lsdpnet_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on torch_ops.cpp:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"linked {TORCH_LIB}")
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
