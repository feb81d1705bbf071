"""Generates tests/golden/*.npz from the UNMODIFIED reference library (oracle/_ref/libseref.so, built
from /root/reference/device/lib by oracle/Makefile).  Run in the build container only
(`python tests/golden/make_golden.py`); the fixtures are committed so the GPU box, which has no
/root/reference, can check against them.

encrypt_golden.npz, per config key k:
  k_cfg      (n, nprimes, asym)
  k_values   [batch][vlen] fp32      k_seeds / k_sseeds [batch][64]
  k_digest   [batch][32]  sha256 of the byte stream se_encrypt_seeded sent (c0,c1 per prime)
  k_pkdigest sha256(pk0 || pk1) of the public key the stream was produced with (asym only)
  k_ct0      full stream of item 0 (small configs only)
stage_golden.npz (n = 1024 and 4096): encode output, u, e0+pt, e1, PRNG counter, ntt of a ramp.
kat.npz: scalar known answers of device/test/modulo_tests.c and uintmodarith_tests.c re-evaluated
  through the reference's own functions.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = [(1024, 1, 0), (4096, 3, 1), (4096, 3, 0), (8192, 4, 1), (16384, 6, 0)]


def one_config(args):
    n, np_, asym = args
    ref = O.ReferenceLib()
    sk = O.make_sk(n)
    pk0, pk1 = ref.gen_pk(n, np_, sk)
    primes = O.Oracle().primes(n, np_)
    ref.setup(n, np_, bool(asym), sk=sk, pk0=pk0, pk1=pk1, primes=primes)
    batch = 2
    vlens = [n // 2, n // 2]
    vals = O.make_values(batch, n // 2, seed=1000 + n + asym)
    seeds = O.make_seeds(batch, b"golden-%d-%d" % (n, asym))
    sseeds = O.make_seeds(batch, b"golden-share-%d-%d" % (n, asym))
    digests = np.zeros((batch, 32), np.uint8)
    ct0 = None
    for b in range(batch):
        ok, ct = ref.encrypt_seeded(sseeds[b], seeds[b], vals[b][: vlens[b]])
        assert ok
        digests[b] = np.frombuffer(hashlib.sha256(ct.tobytes()).digest(), np.uint8)
        if b == 0:
            ct0 = ct.copy()
    ref.close()
    key = "n%d_p%d_%s" % (n, np_, "asym" if asym else "sym")
    out = {key + "_cfg": np.array([n, np_, asym], np.int64), key + "_values": vals, key + "_seeds": seeds,
           key + "_sseeds": sseeds, key + "_digest": digests}
    if asym:
        out[key + "_pkdigest"] = np.frombuffer(hashlib.sha256(pk0.tobytes() + pk1.tobytes()).digest(), np.uint8)
    if n <= 1024:
        out[key + "_ct0"] = ct0
    return out


def stage_vectors():
    ref = O.ReferenceLib()
    out = {}
    for n in (1024, 4096):
        v = O.make_values(1, n // 2, seed=7 + n)[0]
        seed = O.make_seeds(1, b"stage-%d" % n)[0]
        ok, pt = ref.encode(n, v)
        assert ok
        u, pte, e1, ctr = ref.asym_init(n, seed, pt)
        np_ = 1 if n == 1024 else 3
        ramp = (np.arange(n, dtype=np.uint64) * 2654435761 % 134012929).astype(np.uint32)
        out.update({f"n{n}_values": v, f"n{n}_seed": seed, f"n{n}_pt": pt, f"n{n}_u": u, f"n{n}_pte": pte,
                    f"n{n}_e1": e1, f"n{n}_ctr": np.array([ctr], np.uint64), f"n{n}_ramp": ramp,
                    f"n{n}_ntt_ramp": np.stack([ref.ntt(n, np_, p, ramp) for p in range(np_)]),
                    f"n{n}_index_map": ref.index_map(n)})
        a, c = ref.sample_uniform(n, np_, 0, seed, 0)
        out[f"n{n}_uniform_p0"] = a
        out[f"n{n}_uniform_ctr"] = np.array([c], np.uint64)
    return out


def kats():
    ref = O.ReferenceLib()
    L = ref.lib
    MAX = 0xFFFFFFFF
    rows32, rows64, rowsmul = [], [], []
    for q in (134012929, 1053818881):
        for x in (0, 1, q - 1, q, q + 1, (q << 1) & MAX, (q << 2) & MAX, 0x36934613, MAX):
            rows32.append((x, q, L.ref_barrett32(x, q)))
        for hi, lo in ((0, 0), (0, 1), (0, q - 1), (0, q), (0, q + 1), (0, MAX), (0x33345624, 0x47193658), (MAX, MAX)):
            rows64.append((lo, hi, q, L.ref_barrett64(lo, hi, q)))
        for a, b in ((0, 0), (1, 1), (q - 1, q - 1), (0x38573475 % q, 0x83748563 % q), (0x38573475, 0x83748563)):
            rowsmul.append((a, b, q, L.ref_mul_mod(a, b, q)))
    return {"barrett32": np.array(rows32, np.uint64), "barrett64": np.array(rows64, np.uint64),
            "mul_mod": np.array(rowsmul, np.uint64)}


if __name__ == "__main__":
    import multiprocessing as mp

    O.build(ref=True)
    with mp.Pool(len(CONFIGS)) as pool:  # one process per config: the reference keeps static state
        parts = pool.map(one_config, CONFIGS)
    merged = {}
    for p in parts:
        merged.update(p)
    np.savez_compressed(os.path.join(HERE, "encrypt_golden.npz"), **merged)
    np.savez_compressed(os.path.join(HERE, "stage_golden.npz"), **stage_vectors())
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kats())
    for f in ("encrypt_golden.npz", "stage_golden.npz", "kat.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
