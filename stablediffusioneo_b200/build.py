"""Builds libsdeo.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

Usage: python -m stablediffusioneo_b200.build [--force]
The .so lands in stablediffusioneo_b200/_C/ so that it travels with the source tree.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libsdeo.so")
SOURCES = ["host_util.cu", "gemm_conv.cu", "attention.cu", "norm.cu", "groupnorm_stream.cu", "elementwise.cu", "precise.cu", "canny.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libsdeo.so")


def _source_hash():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/sdeo.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.stamp")
    digest = _source_hash()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)

    def unit_hash(src):
        # one translation unit = its .cu + every header it may include (all of csrc's and the public one) + the flags
        h = hashlib.sha256()
        deps = [src] + sorted(n for n in os.listdir(CSRC) if n.endswith((".cuh", ".h"))) + ["../../include/sdeo.h"]
        for name in deps:
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
        h.update(" ".join(NVCC_FLAGS).encode())
        return h.hexdigest()

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        tag, want = obj + ".hash", unit_hash(src)
        if not force and os.path.exists(obj) and os.path.exists(tag):
            with open(tag) as f:
                if f.read().strip() == want:
                    return obj  # unchanged unit: keep its object (gemm_conv.cu alone takes minutes)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        with open(tag, "w") as f:
            f.write(want)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + ".tmp"  # link aside, then rename: a reader (or a repo snapshot) never sees a half-written library
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(path)
