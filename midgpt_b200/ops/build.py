"""In-tree build of the native extensions (no JIT cache outside the repo, so the .so files travel with it).

    python -m midgpt_b200.ops.build            # build everything that is out of date
    python -m midgpt_b200.ops.build --force

* ``_ext*.so``            — CUDA kernels (csrc/*.cu, sm_100a only) + torch bindings (csrc/bindings.cpp)
* ``_native_loader*.so``  — C++ data-loader helper (csrc/dataloader.cpp, pybind11, no torch)

The .cu files do not include torch headers, so nvcc cross-compiles them in seconds on a box without a
GPU; only ``bindings.cpp`` pays for ``torch/extension.h``.  A content hash of the sources + flags is
stored next to each product; an up-to-date product is left alone.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
CU_SOURCES = ["gemm.cu", "gemm_mx.cu", "elementwise.cu", "attention.cu", "comm.cu"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    return "nvcc"


def _hash(paths, extra="") -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode() + b"\0" + f.read())      # (basename: the digest must not depend on where the repo is checked out)
    return h.hexdigest()


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(cmd)}\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
    return r


def ext_path() -> str:
    return os.path.join(HERE, "_ext" + EXT_SUFFIX)


def loader_path() -> str:
    return os.path.join(HERE, "_native_loader" + EXT_SUFFIX)


def build_ext(force: bool = False, verbose: bool = True) -> str:
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OBJ, exist_ok=True)
    cus = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    srcs = cus + [os.path.join(CSRC, "bindings.cpp"), os.path.join(CSRC, "common.cuh")]
    stamp = os.path.join(HERE, "_ext.stamp")
    digest = _hash(srcs, " ".join(NVCC_FLAGS) + torch.__version__ + ",".join(os.path.basename(c) for c in cus))
    out = ext_path()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    if verbose:
        print(f"[build] compiling {len(cus)} CUDA sources for sm_100a + bindings …", flush=True)
    inc = ce.include_paths("cuda") if hasattr(ce, "include_paths") else []
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:
        inc = ce.include_paths(True)
    py_inc = sysconfig.get_paths()["include"]
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    defines = ["-DTORCH_EXTENSION_NAME=_ext", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={cxx11}",
               "-DHAS_COMM_CU=" + ("1" if os.path.exists(os.path.join(CSRC, "comm.cu")) else "0")]

    def compile_cu(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        _run([_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-c", src, "-o", obj], log=obj + ".log")
        return obj

    def compile_cpp(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-c", src, "-o", obj, "-I", CSRC, "-I", py_inc, "-I", "/usr/local/cuda/include"]
        for i in inc:
            cmd += ["-isystem", i]
        _run(cmd + defines, log=obj + ".log")
        return obj

    with ThreadPoolExecutor(max_workers=6) as ex:
        futs = [ex.submit(compile_cu, c) for c in cus] + [ex.submit(compile_cpp, os.path.join(CSRC, "bindings.cpp"))]
        objs = [f.result() for f in futs]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["g++", "-shared", "-o", out] + objs + [f"-L{torch_lib}", "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu",
                                                   "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                                                   f"-Wl,-rpath,{torch_lib}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    _run(link)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[build] wrote {out}", flush=True)
    return out


def build_loader(force: bool = False, verbose: bool = True) -> str:
    import pybind11
    src = os.path.join(CSRC, "dataloader.cpp")
    stamp = os.path.join(HERE, "_native_loader.stamp")
    digest = _hash([src])
    out = loader_path()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-pthread", src, "-o", out,
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"]]
    _run(cmd)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"[build] wrote {out}", flush=True)
    return out


def build_all(force: bool = False, verbose: bool = True):
    return build_loader(force, verbose), build_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
