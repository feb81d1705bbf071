#!/usr/bin/env python
"""Latency of ONE se_encrypt_seeded call (the reference's own usage pattern: one message per call, send callback
per component) through the drop-in API, against the reference library's same call on one host core.
  python tools/latency_single.py   (needs a GPU; run from the repo root)"""
import importlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")
from oracle import oracle as O  # test infrastructure: key files + the reference arm of this measurement

res = []
for n, np_, asym in ((4096, 3, True), (4096, 3, False), (1024, 1, False), (16384, 6, False)):
    orc = O.Oracle()
    sk = O.make_sk(n)
    pk0, pk1 = orc.gen_pk(n, np_, sk)
    primes = orc.primes(n, np_)
    tmp = tempfile.mkdtemp()
    O.write_key_files(tmp, n, primes, sk, pk0, pk1)
    cwd = os.getcwd(); os.chdir(tmp)
    se = seb.SealEmbedded()
    se.se_setup(n, np_, 0.0, seb.api.SE_ASYM_ENCR if asym else seb.api.SE_SYM_ENCR)
    vals = O.make_values(1, n // 2, seed=1)[0]
    seed = O.make_seeds(1, b"lat")[0]; sseed = O.make_seeds(1, b"lat-share")[0]
    sink = []
    def send(data): sink.append(len(data)); return len(data)
    for _ in range(5): se.se_encrypt_seeded(sseed, seed, send, vals)
    t0 = time.perf_counter(); reps = 50
    for _ in range(reps): se.se_encrypt_seeded(sseed, seed, send, vals)
    gpu_us = (time.perf_counter() - t0) / reps * 1e6
    se.se_cleanup(); os.chdir(cwd)
    ref_us = None
    if O.have_reference():
        ref = O.ReferenceLib(); ref.setup(n, np_, asym, sk=sk, pk0=pk0, pk1=pk1, primes=primes)
        try:
            ref.encrypt_seeded(sseed, seed, vals)
            t0 = time.perf_counter(); r2 = 10
            for _ in range(r2): ref.encrypt_seeded(sseed, seed, vals)
            ref_us = (time.perf_counter() - t0) / r2 * 1e6
        finally:
            ref.close(); os.chdir(cwd)
    res.append({"n": n, "nprimes": np_, "asym": asym, "b200_us_per_call": round(gpu_us, 1),
                "reference_us_per_call_one_core": None if ref_us is None else round(ref_us, 1)})
    print(json.dumps(res[-1]), flush=True)
