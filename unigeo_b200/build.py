"""In-tree build of libunigeo_b200.so (nvcc, sm_100a only).

``python -m unigeo_b200.build`` or ``__graft_entry__.build()``.  Objects go to
``unigeo_b200/csrc/build/`` and the library to ``unigeo_b200/libunigeo_b200.so``; both
are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libunigeo_b200.so")
SOURCES = ["tapgemm.cu", "fmha.cu", "kernels.cu", "ops.cu", "unet.cu", "unet2d.cu", "vae.cu", "clip.cu", "post.cu", "metrics.cu", "stitch.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return nv


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(tag: str, defines) -> str:
    """A differently-compiled copy of the whole library (e.g. tag "trace", defines ["UG_TAPGEMM_TRACE"]) as
    ``libunigeo_b200_<tag>.so`` next to the product library; objects under ``csrc/build/<tag>/``.  Development
    tooling (tools/trace_tapgemm.py): the product build and its stamp are not touched."""
    obj_dir = os.path.join(OBJ, tag)
    os.makedirs(obj_dir, exist_ok=True)
    lib = os.path.join(HERE, f"libunigeo_b200_{tag}.so")
    nv = _nvcc()
    flags = [*NVCC_FLAGS, *[f"-D{d}" for d in defines]]

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        r = subprocess.run([nv, *flags, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nv, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "unigeo_b200.h"))
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nv = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nv, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([nv, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
