"""Worker processes that run the reference's own CPU path over slices of a batch (test infrastructure).

The reference keeps static state (device/lib/seal_embedded.c:18-22), so it runs one process per host core; each
worker sets the reference up once (`oracle/_ref/libseref.so` = the unmodified reference compiled by oracle/Makefile,
or the oracle port when that file is absent), encrypts its chunks through se_encrypt_seeded and returns the per-item
64-bit digests of the byte streams (oracle/ref_shim.c: ref_encrypt_digests = the function of seb_digest_device).

Inputs are a pure function of (tag, chunk index), so the GPU side regenerates exactly the same items:
  values of chunk c = make_values(CHUNK, n/2, seed = value_seed + c)
  seeds of item i   = SHAKE256(tag || LE64(i))[0:64]   (shareable seeds: tag + b"-share")
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = 256  # items per unit of work


def chunk_inputs(O, n: int, tag: bytes, value_seed: int, c: int, asym: bool):
    vals = O.make_values(CHUNK, n // 2, seed=value_seed + c)
    seeds = O.make_seeds(CHUNK, tag, start=c * CHUNK)
    sseeds = None if asym else O.make_seeds(CHUNK, tag + b"-share", start=c * CHUNK)
    return vals, seeds, sseeds


def _worker(args):
    n, np_, asym, tag, value_seed, chunks = args
    from oracle import oracle as O

    orc = O.Oracle()
    sk = O.make_sk(n)
    pk0 = pk1 = None
    if asym:
        pk0, pk1 = orc.gen_pk(n, np_, sk)
    ref = None
    if O.have_reference():
        ref = O.ReferenceLib()
        ref.setup(n, np_, asym, sk=sk, pk0=pk0, pk1=pk1, primes=orc.primes(n, np_))
    out = {}
    try:
        for c in chunks:
            vals, seeds, sseeds = chunk_inputs(O, n, tag, value_seed, c, asym)
            if ref is not None:
                out[c] = ref.encrypt_digests(sseeds, seeds, vals)
            else:  # the oracle port (same bytes: tests/test_oracle.py pins it against the reference)
                cts = np.empty((CHUNK, np_, 2, n), np.uint32)
                for b in range(CHUNK):
                    if asym:
                        ok, cts[b] = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
                    else:
                        ok, cts[b] = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk, ref_quirk=True)
                    assert ok
                out[c] = O.digest_words(cts.reshape(CHUNK, -1))
    finally:
        if ref is not None:
            ref.close()
    return out


def reference_digests(n: int, np_: int, asym: bool, tag: bytes, value_seed: int, nchunks: int, procs: int | None = None):
    """digests [nchunks * CHUNK] of the reference's ciphertext streams, computed in `procs` processes."""
    from oracle import oracle as O

    procs = max(1, min(procs or (os.cpu_count() or 1), nchunks))
    jobs = [(n, np_, asym, tag, value_seed, list(range(w, nchunks, procs))) for w in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        parts = pool.map(_worker, jobs)
    merged = {}
    for p in parts:
        merged.update(p)
    kind = "reference" if O.have_reference() else "port"
    return np.concatenate([merged[c] for c in range(nchunks)]), kind, procs
