"""Builds libsdf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "libsdf_b200.so")

SOURCES = ["capi.cu", "lif.cu", "bn.cu", "window.cu", "attn_qkgate.cu", "attn_qktv.cu", "conv_small.cu", "spike_gemm.cu", "spike_wgrad.cu", "voxel_input.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src, obj, verbose):
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ and link libsdf_b200.so.  Incremental per source file."""
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "sdf_b200.h"))
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        stamp = obj + ".sha"
        dig = _digest([src, *headers])
        objs.append(obj)
        old = open(stamp).read() if os.path.exists(stamp) else ""
        if force or old != dig or not os.path.exists(obj):
            jobs.append((src, obj, stamp, dig))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            futs = [(ex.submit(_compile_one, src, obj, verbose), stamp, dig) for src, obj, stamp, dig in jobs]
            for fut, stamp, dig in futs:
                fut.result()
                with open(stamp, "w") as f:
                    f.write(dig)
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


def build_library_locked(force=False, verbose=False):
    """build_library() under an exclusive file lock, so N ranks starting together compile once."""
    import fcntl
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            return build_library(force=force, verbose=verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
