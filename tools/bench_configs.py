#!/usr/bin/env python
"""Per-GPU throughput of every BASELINE.json configuration plus the NTT-only sweep (config E), one GPU.

  python tools/bench_configs.py [--out profiles/rNN_configs.json] [--quick]

Config B is what bench.py reports; the others are parity-test configurations, measured here so their
kernels are not flying blind.  Batches are the per-GPU shards of the 8-GPU configurations.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


PRIMES30 = [1053818881, 1054015489, 1054212097, 1055260673, 1056178177, 1056440321, 1058209793, 1060175873,
            1060700161, 1060765697, 1061093377, 1062469633, 1062535169]  # device/lib/parameters.c:129-174


def timed(fn, stream, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def full_path(name, n, np_, asym, batch, reps=3):
    ctx = seb.Context(n, np_, asym=asym, device=0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1)
    if asym:
        pk = [np.stack([rng.integers(0, q, n, dtype=np.uint32) for q in ctx.primes]) for _ in range(2)]
        ctx.set_public_key(*pk)
    else:
        t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
        ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
    vlen = n // 2
    gen = torch.Generator(device="cuda").manual_seed(3)
    d_vals = torch.rand((batch, vlen), generator=gen, device="cuda", dtype=torch.float32) * 32 - 16
    d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")

    def step():
        if asym:
            ctx.encrypt_asym_device(d_vals, vlen, d_seeds, batch, d_out)
        else:
            ctx.encrypt_sym_device(d_vals, vlen, d_ss, d_seeds, batch, d_out, False)

    ms = timed(step, stream, reps)
    assert ctx.encode_failures() == 0
    ctx.profile_begin(2)
    step()
    step()
    k = ctx.profile_end().mean(axis=0)
    names = ctx.PROFILE_SEGMENTS[asym]
    res = {"config": name, "n": n, "nprimes": np_, "asym": asym, "batch": batch, "ms_per_step": ms,
           "ciphertexts_per_s": batch / (ms * 1e-3), "out_GB_per_step": d_out.numel() * 4 / 1e9,
           "kernels_ms": {nm: float(v) for nm, v in zip(names, k)}}
    ctx.close()
    del d_out, d_vals
    torch.cuda.empty_cache()
    return res


def ntt_sweep(quick):
    out = []
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    # config E: n x primes x batch.  Prime counts beyond the reference's parameter sets (8 primes at n <= 8192)
    # use the first primes of its 30-bit list with the library's minimal 2n-th roots (seb_minimal_psi).
    for n, nps in ((1024, (1, 8)), (2048, (1,)), (4096, (1, 3, 8)), (8192, (1, 4, 8)), (16384, (1, 6, 8, 13))):
        for np_ in nps:
            ctx = seb.Context(n, np_, asym=True, device=0, primes=PRIMES30[:np_])
            ctx.set_stream(stream.cuda_stream)
            batches = [1, 64, 4096, (1 << 30) // (4 * n * np_)] if not quick else [(1 << 30) // (4 * n * np_)]
            for batch in batches:
                polys = torch.randint(0, 1 << 27, (batch, np_, n), device="cuda", dtype=torch.int32)
                ms = timed(lambda: ctx.ntt_device(polys, batch), stream, 5)
                gbs = 8.0 * n * np_ * batch / (ms * 1e-3) / 1e9
                out.append({"n": n, "nprimes": np_, "batch": batch, "ms": ms, "ntt_per_s": batch * np_ / (ms * 1e-3),
                            "GBps": gbs, "frac_of_hbm_peak": gbs / PEAK})
                del polys
            ctx.close()
    return out


def e2e_sym(name, n, np_, batch, reps=2):
    """Host-pointer symmetric path with pinned buffers, full ciphertexts vs the seed-compressed form (c0 only):
    the ciphertext D2H is the bound, so halving it should nearly double the end-to-end rate."""
    import time
    ctx = seb.Context(n, np_, asym=False, device=0)
    rng = np.random.default_rng(2)
    t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
    ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
    vlen = n // 2
    vals = torch.empty((batch, vlen), dtype=torch.float32).uniform_(-16, 16).pin_memory()
    seeds = torch.randint(0, 256, (batch, 64), dtype=torch.uint8).pin_memory()
    ss = torch.randint(0, 256, (batch, 64), dtype=torch.uint8).pin_memory()
    res = {"config": name, "n": n, "nprimes": np_, "batch": batch}
    for mode, words in (("full", 2 * np_ * n), ("seedct", np_ * n)):
        out = torch.empty((batch, words), dtype=torch.int32).pin_memory()
        fn = (lambda: ctx.encrypt_sym_host(vals.numpy(), ss.numpy(), seeds.numpy(), out.numpy().view(np.uint32))) \
            if mode == "full" else \
            (lambda: ctx.encrypt_sym_seedct_host(vals.numpy(), ss.numpy(), seeds.numpy(), out.numpy().view(np.uint32)))
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        res[mode] = {"ms_per_step": dt * 1e3, "ciphertexts_per_s": batch / dt, "d2h_GB_per_step": batch * words * 4 / 1e9,
                     "d2h_GBps": batch * words * 4 / 1e9 / dt}
        del out
    ctx.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    res = {"hbm_peak_GBps": PEAK, "full_path": [], "e2e_sym": [], "ntt_only": []}
    cfgs = [("A: n=1024 1 prime sym, batch 65536 (the reference's CPU case, batched)", 1024, 1, False, 65536),
            ("B: n=4096 3 primes asym, batch 65536", 4096, 3, True, 65536),
            ("B-sym: n=4096 3 primes sym, batch 65536", 4096, 3, False, 65536),
            ("C: n=8192 4 primes asym, batch 32768 (1/8 of 262144)", 8192, 4, True, 32768),
            ("D: n=16384 6 primes sym, batch 16384 (1/8 of 131072)", 16384, 6, False, 16384)]
    for c in cfgs:
        r = full_path(*c)
        res["full_path"].append(r)
        print(json.dumps(r), flush=True)
    res["e2e_sym"] = []
    for c in (("B-sym e2e: n=4096 3 primes sym, batch 16384 (host buffers)", 4096, 3, 16384),
              ("D e2e: n=16384 6 primes sym, batch 4096 (host buffers)", 16384, 6, 4096)):
        r = e2e_sym(*c)
        res["e2e_sym"].append(r)
        print(json.dumps(r), flush=True)
    res["ntt_only"] = ntt_sweep(a.quick)
    for r in res["ntt_only"]:
        print(json.dumps(r), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
