"""Shared fixtures.

`-m "not gpu"`: oracle vs golden vectors / reference KATs / compiled reference, host logic (g++
emulation of the per-thread device functions), C-ABI symbol check, gloo sharding tests.
`-m gpu`: parity tests proper — the CUDA path through the C ABI against the oracle.
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O

    O.build()  # compiles the C restatement; also the reference .so when /root/reference is mounted
    return O


@pytest.fixture(scope="session")
def orc(oracle_mod):
    return oracle_mod.Oracle()


@pytest.fixture(scope="session")
def seb():
    """The product package (hyphenated directory name)."""
    return importlib.import_module("seal-embedded_b200")


@pytest.fixture(scope="session")
def emul():
    """g++ build of tests/host_emul/emul.cpp (sequential emulation of device functions)."""
    import ctypes

    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    out = os.path.join(ROOT, "tests", "host_emul", "libemul.so")
    csrc = os.path.join(ROOT, "seal-embedded_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        # SEB_ENC_E=16 compiles the emulation for the encode's 16-values-per-thread variant (A/B builds)
        extra = [f"-DENC_E={os.environ['SEB_ENC_E']}"] if os.environ.get("SEB_ENC_E") else []
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", csrc, "-I", cuda_inc] + extra +
                       [src, "-o", out], check=True)
    return ctypes.CDLL(out)


@pytest.fixture(scope="session")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


def keys_for(O, orc, n: int, np_: int):
    """Deterministic sk and the matching pk (oracle gen_pk; identical to the reference's gen_pk)."""
    sk = O.make_sk(n)
    pk0, pk1 = orc.gen_pk(n, np_, sk)
    return sk, pk0, pk1


def to_dev(torch, a: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()
