"""In-tree build of libseal_embedded_b200.so: hand-written sm_100a CUDA kernels (csrc/*.cu, nvcc) plus
the C host layer that carries the reference's se_* API (host/seal_embedded.c, gcc).

The .so is git-ignored but travels to the GPU box with the snapshot; nvcc cross-compiles for
sm_100a without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libseal_embedded_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + ARCH
# per-file extras: the FP64 encode must never contract a*b+c into an FMA (bit-exactness)
EXTRA = {"seb_encode.cu": ["-fmad=false"]}
CU_SOURCES = ["seb_api.cu", "seb_sample.cu", "seb_encode.cu", "seb_encrypt.cu", "seb_verify.cu"]
C_SOURCES = ["seal_embedded.c"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _deps() -> list[str]:
    out = [os.path.abspath(__file__), os.path.join(HERE, "..", "include", "seal_embedded_b200.h")]
    for d in (CSRC, HOST):
        out += [os.path.join(d, f) for f in os.listdir(d)]
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    deps = _deps()
    if not force and _newer(LIB, deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src in CU_SOURCES:
        obj = os.path.join(BUILD, src + ".o")
        jobs.append((obj, [nvcc] + NVCC_FLAGS + EXTRA.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]))
    for src in C_SOURCES:
        obj = os.path.join(BUILD, src + ".o")
        jobs.append((obj, ["gcc", "-O2", "-std=gnu11", "-fPIC", "-Wall", "-c", os.path.join(HOST, src), "-o", obj]))

    def run(job):
        obj, cmd = job
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
        if p.returncode != 0:
            raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{p.stdout}")
        if verbose:
            print(p.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        objs = list(ex.map(run, jobs))
    link = [nvcc, "-shared", "-o", LIB] + ARCH + objs + ["-lm"]
    p = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed: {' '.join(link)}\n{p.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
