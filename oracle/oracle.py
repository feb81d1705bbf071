"""ctypes front-end to the TEST ORACLE (oracle/liboracle.so) and, when it has been built, to the
unmodified reference library (oracle/_ref/libseref.so).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product (seal-embedded_b200/) never does.

Parity status: PINNED — see oracle/se_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libseref.so")
REFERENCE_SRC = "/root/reference/device/lib"
REF_DEMO = os.path.join(HERE, "_ref", "se_reference_api_demo")  # built by `make refdemo` (reference headers)

SEED_BYTES = 64

_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(ref: bool | None = None) -> None:
    """Compile the oracle (always) and the reference .so (when /root/reference is mounted)."""
    targets = ["oracle"]
    if ref is None:
        ref = os.path.isdir(REFERENCE_SRC)
    if ref:
        targets.append("ref")
        # the reference-header build of the reference-API-only demo, linked against the product library
        if os.path.exists(os.path.join(HERE, "..", "seal-embedded_b200", "libseal_embedded_b200.so")):
            targets.append("refdemo")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _ptr(a: np.ndarray, typ):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


def _seed(seed) -> np.ndarray:
    a = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    assert a.size == SEED_BYTES
    return a


# --------------------------------------------------------------------------------------------
# deterministic synthetic inputs shared by tests, fixtures and the bench
# --------------------------------------------------------------------------------------------
def make_seeds(batch: int, tag: bytes = b"se-b200", start: int = 0) -> np.ndarray:
    """seed[b] = SHAKE256(tag || LE64(start + b))[0:64]"""
    out = np.empty((batch, SEED_BYTES), dtype=np.uint8)
    for b in range(batch):
        out[b] = np.frombuffer(hashlib.shake_256(tag + struct.pack("<Q", start + b)).digest(SEED_BYTES), dtype=np.uint8)
    return out


def make_sk(n: int, tag: bytes = b"se-b200-sk") -> np.ndarray:
    """n/4 bytes, four 2-bit fields in {0,1,2} per byte (sk_<n>.dat format, fileops.c:140-170)."""
    raw = np.frombuffer(hashlib.shake_256(tag + struct.pack("<Q", n)).digest(n), dtype=np.uint8)
    t = (raw % 3).astype(np.uint8).reshape(n // 4, 4)
    return ((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8)


def make_values(batch: int, vlen: int, seed: int = 0) -> np.ndarray:
    """fp32 messages, i.i.d. uniform in [-16, 16)"""
    rng = np.random.Generator(np.random.Philox(seed))
    return (rng.random((batch, vlen), dtype=np.float32) * np.float32(32.0) - np.float32(16.0)).astype(np.float32)


# --------------------------------------------------------------------------------------------
# the restatement
# --------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_shake256.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t]
        L.orc_prng_fill.argtypes = [_u8p, C.c_uint64, C.c_size_t, _u8p]
        L.orc_default_primes.argtypes = [C.c_size_t, C.c_size_t, _u32p]
        L.orc_default_primes.restype = C.c_int
        L.orc_default_scale.argtypes = [C.c_size_t]
        L.orc_default_scale.restype = C.c_double
        L.orc_ntt_root.argtypes = [C.c_size_t, C.c_uint32]
        L.orc_ntt_root.restype = C.c_uint32
        L.orc_const_ratio.argtypes = [C.c_uint32, _u32p]
        for name, nargs in (("orc_barrett32", 2), ("orc_barrett64", 3), ("orc_add_mod", 3), ("orc_neg_mod", 2),
                            ("orc_sub_mod", 3), ("orc_mul_mod", 3)):
            f = getattr(L, name)
            f.argtypes = [C.c_uint32] * nargs
            f.restype = C.c_uint32
        L.orc_pow_mod.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32]
        L.orc_pow_mod.restype = C.c_uint32
        L.orc_index_map.argtypes = [C.c_size_t, _u16p]
        L.orc_ifft_twiddles.argtypes = [C.c_size_t, _f64p]
        L.orc_encode.argtypes = [C.c_size_t, C.c_double, _f32p, C.c_size_t, _i64p]
        L.orc_encode.restype = C.c_int
        L.orc_decode.argtypes = [C.c_size_t, C.c_double, C.c_uint32, _u32p, C.c_size_t, _f32p]
        L.orc_sample_ternary_small.argtypes = [C.c_size_t, _u8p, C.POINTER(C.c_uint64), _u8p]
        L.orc_sample_cbd.argtypes = [C.c_size_t, _u8p, C.POINTER(C.c_uint64), _i8p]
        L.orc_sample_uniform.argtypes = [C.c_size_t, C.c_uint32, _u8p, C.POINTER(C.c_uint64), _u32p]
        L.orc_expand_ternary.argtypes = [C.c_size_t, C.c_uint32, _u8p, _u32p]
        L.orc_reduce_small.argtypes = [C.c_size_t, C.c_uint32, _i8p, _u32p]
        L.orc_reduce_pte.argtypes = [C.c_size_t, C.c_uint32, _i64p, _u32p]
        L.orc_ntt_roots.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, _u32p]
        L.orc_ntt.argtypes = [C.c_size_t, C.c_uint32, _u32p, _u32p]
        L.orc_intt.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, _u32p]
        L.orc_ntt_default.argtypes = [C.c_size_t, C.c_uint32, _u32p]
        L.orc_negacyclic_mul.argtypes = [C.c_size_t, C.c_uint32, _u32p, _u32p, _u32p]
        L.orc_encrypt_asym.argtypes = [C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _u8p, _u32p, _u32p, _u32p]
        L.orc_encrypt_asym.restype = C.c_int
        L.orc_encrypt_sym.argtypes = [C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _u8p, _u8p, _u8p, C.c_int, _u32p]
        L.orc_encrypt_sym.restype = C.c_int
        L.orc_gen_pk_prime.argtypes = [C.c_size_t, C.c_uint32, _u8p, _u8p, _i8p, _u32p, _u32p]
        L.orc_decrypt_ntt.argtypes = [C.c_size_t, C.c_uint32, _u32p, _u32p, _u8p, _u32p]
        L.orc_encrypt_asym_batch.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _u8p, _u32p,
                                             _u32p, _u32p]
        L.orc_encrypt_asym_batch.restype = C.c_int
        L.orc_encrypt_asym_ex.argtypes = [C.c_size_t, C.c_size_t, _u32p, _u32p, C.c_double, _f32p, C.c_size_t, _u8p,
                                          _u32p, _u32p, _u32p]
        L.orc_encrypt_asym_ex.restype = C.c_int
        L.orc_encrypt_sym_ex.argtypes = [C.c_size_t, C.c_size_t, _u32p, _u32p, C.c_double, _f32p, C.c_size_t, _u8p,
                                         _u8p, _u8p, C.c_int, _u32p]
        L.orc_encrypt_sym_ex.restype = C.c_int
        L.orc_gen_pk_prime_ex.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, _u8p, _u8p, _i8p, _u32p, _u32p]
        L.orc_decrypt_ntt_ex.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, _u32p, _u32p, _u8p, _u32p]
        L.orc_ntt_psi.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, _u32p]

    # -- hashing / prng
    def shake256(self, data: bytes, outlen: int) -> bytes:
        inp = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, np.uint8)
        out = np.empty(max(outlen, 1), np.uint8)
        self.lib.orc_shake256(_ptr(out, _u8p), outlen, _ptr(inp, _u8p), len(data))
        return out[:outlen].tobytes()

    def prng_fill(self, seed, counter: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, np.uint8)
        self.lib.orc_prng_fill(_ptr(_seed(seed), _u8p), counter, nbytes, _ptr(out, _u8p))
        return out

    # -- parameters
    def primes(self, n: int, np_: int) -> list[int]:
        buf = np.zeros(16, np.uint32)
        if not self.lib.orc_default_primes(n, np_, _ptr(buf, _u32p)):
            raise ValueError(f"illegal parameter set n={n} nprimes={np_}")
        return [int(x) for x in buf[:np_]]

    def scale(self, n: int) -> float:
        return float(self.lib.orc_default_scale(n))

    def ntt_root(self, n: int, q: int) -> int:
        return int(self.lib.orc_ntt_root(n, q))

    def const_ratio(self, q: int) -> tuple[int, int]:
        buf = np.zeros(2, np.uint32)
        self.lib.orc_const_ratio(q, _ptr(buf, _u32p))
        return int(buf[0]), int(buf[1])

    # -- encode
    def index_map(self, n: int) -> np.ndarray:
        out = np.empty(n, np.uint16)
        self.lib.orc_index_map(n, _ptr(out, _u16p))
        return out

    def ifft_twiddles(self, n: int) -> np.ndarray:
        out = np.empty(2 * n, np.float64)
        self.lib.orc_ifft_twiddles(n, _ptr(out, _f64p))
        return out

    def encode(self, n: int, values: np.ndarray, scale: float | None = None):
        v = np.ascontiguousarray(values, dtype=np.float32)
        out = np.empty(n, np.int64)
        ok = self.lib.orc_encode(n, self.scale(n) if scale is None else scale, _ptr(v, _f32p), v.size,
                                 _ptr(out, _i64p))
        return bool(ok), out

    def decode(self, n: int, q: int, pt: np.ndarray, vlen: int, scale: float | None = None) -> np.ndarray:
        p = np.ascontiguousarray(pt, dtype=np.uint32)
        out = np.empty(vlen, np.float32)
        self.lib.orc_decode(n, self.scale(n) if scale is None else scale, q, _ptr(p, _u32p), vlen, _ptr(out, _f32p))
        return out

    # -- samplers (each returns (array, counter_after))
    def sample_ternary_small(self, n: int, seed, counter: int = 0):
        out = np.empty(n // 4, np.uint8)
        ctr = C.c_uint64(counter)
        self.lib.orc_sample_ternary_small(n, _ptr(_seed(seed), _u8p), C.byref(ctr), _ptr(out, _u8p))
        return out, ctr.value

    def sample_cbd(self, n: int, seed, counter: int = 0):
        out = np.empty(n, np.int8)
        ctr = C.c_uint64(counter)
        self.lib.orc_sample_cbd(n, _ptr(_seed(seed), _u8p), C.byref(ctr), _ptr(out, _i8p))
        return out, ctr.value

    def sample_uniform(self, n: int, q: int, seed, counter: int = 0):
        out = np.empty(n, np.uint32)
        ctr = C.c_uint64(counter)
        self.lib.orc_sample_uniform(n, q, _ptr(_seed(seed), _u8p), C.byref(ctr), _ptr(out, _u32p))
        return out, ctr.value

    def expand_ternary(self, n: int, q: int, packed: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.empty(n, np.uint32)
        self.lib.orc_expand_ternary(n, q, _ptr(p, _u8p), _ptr(out, _u32p))
        return out

    def reduce_small(self, n: int, q: int, e: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(e, dtype=np.int8)
        out = np.empty(n, np.uint32)
        self.lib.orc_reduce_small(n, q, _ptr(a, _i8p), _ptr(out, _u32p))
        return out

    def reduce_pte(self, n: int, q: int, pte: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(pte, dtype=np.int64)
        out = np.empty(n, np.uint32)
        self.lib.orc_reduce_pte(n, q, _ptr(a, _i64p), _ptr(out, _u32p))
        return out

    # -- ntt
    def ntt_roots(self, n: int, q: int, psi: int | None = None) -> np.ndarray:
        out = np.empty(n, np.uint32)
        self.lib.orc_ntt_roots(n, q, self.ntt_root(n, q) if psi is None else psi, _ptr(out, _u32p))
        return out

    def ntt(self, n: int, q: int, vec: np.ndarray, psi: int | None = None) -> np.ndarray:
        v = np.array(vec, dtype=np.uint32, copy=True)
        roots = self.ntt_roots(n, q, psi)
        self.lib.orc_ntt(n, q, _ptr(roots, _u32p), _ptr(v, _u32p))
        return v

    def intt(self, n: int, q: int, vec: np.ndarray, psi: int | None = None) -> np.ndarray:
        v = np.array(vec, dtype=np.uint32, copy=True)
        self.lib.orc_intt(n, q, self.ntt_root(n, q) if psi is None else psi, _ptr(v, _u32p))
        return v

    def negacyclic_mul(self, n: int, q: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.empty(n, np.uint32)
        self.lib.orc_negacyclic_mul(n, q, _ptr(a, _u32p), _ptr(b, _u32p), _ptr(out, _u32p))
        return out

    # -- full path
    def encrypt_asym(self, n: int, np_: int, values: np.ndarray, seed, pk0: np.ndarray, pk1: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float32)
        pk0 = np.ascontiguousarray(pk0, dtype=np.uint32)
        pk1 = np.ascontiguousarray(pk1, dtype=np.uint32)
        out = np.zeros((np_, 2, n), np.uint32)
        ok = self.lib.orc_encrypt_asym(n, np_, _ptr(v, _f32p), v.size, _ptr(_seed(seed), _u8p), _ptr(pk0, _u32p),
                                       _ptr(pk1, _u32p), _ptr(out, _u32p))
        return bool(ok), out

    def encrypt_sym(self, n: int, np_: int, values: np.ndarray, share_seed, seed, sk: np.ndarray,
                    ref_quirk: bool = False):
        v = np.ascontiguousarray(values, dtype=np.float32)
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        out = np.zeros((np_, 2, n), np.uint32)
        ok = self.lib.orc_encrypt_sym(n, np_, _ptr(v, _f32p), v.size, _ptr(_seed(share_seed), _u8p),
                                      _ptr(_seed(seed), _u8p), _ptr(sk, _u8p), int(ref_quirk), _ptr(out, _u32p))
        return bool(ok), out

    def encrypt_asym_batch(self, n: int, np_: int, values: np.ndarray, seeds: np.ndarray, pk0, pk1) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float32)
        s = np.ascontiguousarray(seeds, dtype=np.uint8)
        pk0 = np.ascontiguousarray(pk0, dtype=np.uint32)
        pk1 = np.ascontiguousarray(pk1, dtype=np.uint32)
        batch, vlen = v.shape
        out = np.zeros((batch, np_, 2, n), np.uint32)
        ok = self.lib.orc_encrypt_asym_batch(n, np_, batch, _ptr(v, _f32p), vlen, _ptr(s, _u8p), _ptr(pk0, _u32p),
                                             _ptr(pk1, _u32p), _ptr(out, _u32p))
        assert ok
        return out

    def gen_pk(self, n: int, np_: int, sk: np.ndarray, ep_seed=bytes([7]) * 64, seed_base=bytes([9]) * 64):
        """Same recipe as ReferenceLib.gen_pk: ep = CBD(PRNG(ep_seed)); a from PRNG(seed_base, byte0=p)."""
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        ep, _ = self.sample_cbd(n, ep_seed)
        pk0 = np.zeros((np_, n), np.uint32)
        pk1 = np.zeros((np_, n), np.uint32)
        for p, q in enumerate(self.primes(n, np_)):
            sd = bytearray(seed_base)
            sd[0] = p
            self.lib.orc_gen_pk_prime(n, q, _ptr(_seed(sd), _u8p), _ptr(sk, _u8p), _ptr(ep, _i8p),
                                      _ptr(pk0[p], _u32p), _ptr(pk1[p], _u32p))
        return pk0, pk1

    # -- the same with an explicit chain (custom primes: "parity unpinned" against the reference, se_oracle.h)
    def encrypt_asym_ex(self, n, primes, psis, scale, values, seed, pk0, pk1):
        v = np.ascontiguousarray(values, dtype=np.float32)
        pr = np.ascontiguousarray(primes, dtype=np.uint32)
        ps = np.ascontiguousarray(psis, dtype=np.uint32)
        pk0 = np.ascontiguousarray(pk0, dtype=np.uint32)
        pk1 = np.ascontiguousarray(pk1, dtype=np.uint32)
        out = np.zeros((len(pr), 2, n), np.uint32)
        ok = self.lib.orc_encrypt_asym_ex(n, len(pr), _ptr(pr, _u32p), _ptr(ps, _u32p), float(scale), _ptr(v, _f32p),
                                          v.size, _ptr(_seed(seed), _u8p), _ptr(pk0, _u32p), _ptr(pk1, _u32p),
                                          _ptr(out, _u32p))
        return bool(ok), out

    def encrypt_sym_ex(self, n, primes, psis, scale, values, share_seed, seed, sk, ref_quirk=False):
        v = np.ascontiguousarray(values, dtype=np.float32)
        pr = np.ascontiguousarray(primes, dtype=np.uint32)
        ps = np.ascontiguousarray(psis, dtype=np.uint32)
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        out = np.zeros((len(pr), 2, n), np.uint32)
        ok = self.lib.orc_encrypt_sym_ex(n, len(pr), _ptr(pr, _u32p), _ptr(ps, _u32p), float(scale), _ptr(v, _f32p),
                                         v.size, _ptr(_seed(share_seed), _u8p), _ptr(_seed(seed), _u8p), _ptr(sk, _u8p),
                                         int(ref_quirk), _ptr(out, _u32p))
        return bool(ok), out

    def gen_pk_ex(self, n, primes, psis, sk, ep_seed=bytes([7]) * 64, seed_base=bytes([9]) * 64):
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        ep, _ = self.sample_cbd(n, ep_seed)
        pk0 = np.zeros((len(primes), n), np.uint32)
        pk1 = np.zeros((len(primes), n), np.uint32)
        for p, (q, psi) in enumerate(zip(primes, psis)):
            sd = bytearray(seed_base)
            sd[0] = p
            self.lib.orc_gen_pk_prime_ex(n, int(q), int(psi), _ptr(_seed(sd), _u8p), _ptr(sk, _u8p), _ptr(ep, _i8p),
                                         _ptr(pk0[p], _u32p), _ptr(pk1[p], _u32p))
        return pk0, pk1

    def decrypt_decode_ex(self, n, primes, psis, scale, ct, sk, vlen, prime_idx=0):
        q, psi = int(primes[prime_idx]), int(psis[prime_idx])
        c0 = np.ascontiguousarray(ct[prime_idx, 0], dtype=np.uint32)
        c1 = np.ascontiguousarray(ct[prime_idx, 1], dtype=np.uint32)
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        ptn = np.empty(n, np.uint32)
        self.lib.orc_decrypt_ntt_ex(n, q, psi, _ptr(c0, _u32p), _ptr(c1, _u32p), _ptr(sk, _u8p), _ptr(ptn, _u32p))
        return self.decode(n, q, self.intt(n, q, ptn, psi), vlen, scale)

    def decrypt_ntt(self, n: int, q: int, c0, c1, sk) -> np.ndarray:
        c0 = np.ascontiguousarray(c0, dtype=np.uint32)
        c1 = np.ascontiguousarray(c1, dtype=np.uint32)
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        out = np.empty(n, np.uint32)
        self.lib.orc_decrypt_ntt(n, q, _ptr(c0, _u32p), _ptr(c1, _u32p), _ptr(sk, _u8p), _ptr(out, _u32p))
        return out

    def decrypt_decode(self, n: int, np_: int, ct: np.ndarray, sk: np.ndarray, vlen: int, prime_idx: int = 0):
        """ct [np][2][n] -> decoded floats using one prime (device/test/ckks_tests_common.c:173-231)."""
        q = self.primes(n, np_)[prime_idx]
        ptn = self.decrypt_ntt(n, q, ct[prime_idx, 0], ct[prime_idx, 1], sk)
        return self.decode(n, q, self.intt(n, q, ptn), vlen)


def digest_words(words: np.ndarray) -> np.ndarray:
    """Per-row digest of a [items][W] u32 array: sum_i mix64((i << 32) | w_i) mod 2^64 (splitmix64 finaliser) — the
    function of seb_digest_device (csrc/seb_verify.cu) and ref_encrypt_digests (ref_shim.c), in numpy."""
    w = np.ascontiguousarray(words, dtype=np.uint32)
    w = w.reshape(w.shape[0], -1)
    with np.errstate(over="ignore"):
        z = (np.arange(w.shape[1], dtype=np.uint64) << np.uint64(32))[None, :] | w.astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return z.sum(axis=1, dtype=np.uint64)


def write_key_files(workdir: str, n: int, primes: list[int], sk: np.ndarray | None, pk0=None, pk1=None) -> str:
    """Lay out adapter_output_data/ the way the reference reads it (fileops.c:140-204)."""
    d = os.path.join(workdir, "adapter_output_data")
    os.makedirs(d, exist_ok=True)
    if sk is not None:
        np.ascontiguousarray(sk, dtype=np.uint8).tofile(os.path.join(d, f"sk_{n}.dat"))
    if pk0 is not None:
        for p, q in enumerate(primes):
            np.ascontiguousarray(pk0[p], dtype="<u4").tofile(os.path.join(d, f"pk0_ntt_{n}_{q}.dat"))
            np.ascontiguousarray(pk1[p], dtype="<u4").tofile(os.path.join(d, f"pk1_ntt_{n}_{q}.dat"))
    return d


# --------------------------------------------------------------------------------------------
# the real reference, when its .so is present (built here from /root/reference; travels to the
# GPU box as a prebuilt file)
# --------------------------------------------------------------------------------------------
def have_reference() -> bool:
    return os.path.exists(REF_SO)


class ReferenceLib:
    """One context per process (the reference keeps static state, seal_embedded.c:18-22)."""

    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        self._cwd = os.getcwd()
        self._tmp = None
        self.n = self.np_ = 0
        L.ref_setup.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.c_char_p]
        L.ref_setup.restype = C.c_int
        L.ref_scale.restype = C.c_double
        L.ref_nprimes.restype = C.c_size_t
        L.ref_prime.argtypes = [C.c_size_t]
        L.ref_prime.restype = C.c_uint32
        L.ref_ratio.argtypes = [C.c_size_t, C.c_size_t]
        L.ref_ratio.restype = C.c_uint32
        L.ref_encrypt_seeded.argtypes = [_u8p, _u8p, _f32p, C.c_size_t, _u32p]
        L.ref_encrypt_seeded.restype = C.c_int
        L.ref_encrypt_loop.argtypes = [C.c_size_t, _u8p, _u8p, _f32p, C.c_size_t, _u32p]
        L.ref_encrypt_loop.restype = C.c_double
        if hasattr(L, "ref_encrypt_digests"):
            L.ref_encrypt_digests.argtypes = [C.c_size_t, _u8p, _u8p, _f32p, C.c_size_t, C.c_size_t, _u32p,
                                              C.POINTER(C.c_uint64)]
            L.ref_encrypt_digests.restype = C.c_size_t
            L.ref_digest_words.argtypes = [_u32p, C.c_size_t]
            L.ref_digest_words.restype = C.c_uint64
        L.ref_index_map.argtypes = [C.c_size_t, _u16p]
        L.ref_encode.argtypes = [C.c_size_t, _f32p, C.c_size_t, _i64p]
        L.ref_encode.restype = C.c_int
        L.ref_prng_fill.argtypes = [_u8p, C.c_uint64, C.c_size_t, _u8p]
        L.ref_asym_init.argtypes = [C.c_size_t, _u8p, _i64p, _u8p, _i8p]
        L.ref_asym_init.restype = C.c_uint64
        L.ref_sample_ternary_small.argtypes = [C.c_size_t, _u8p, C.c_uint64, _u8p]
        L.ref_sample_ternary_small.restype = C.c_uint64
        L.ref_sample_cbd.argtypes = [C.c_size_t, _u8p, C.c_uint64, _i8p]
        L.ref_sample_cbd.restype = C.c_uint64
        L.ref_sample_uniform.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, _u8p, C.c_uint64, _u32p]
        L.ref_sample_uniform.restype = C.c_uint64
        L.ref_ntt.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, _u32p]
        if hasattr(L, "ref_ntt_loop"):
            L.ref_ntt_loop.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, _u32p, C.c_size_t]
            L.ref_ntt_loop.restype = C.c_double
        L.ref_reduce_pte.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, _i64p, _u32p]
        L.ref_encrypt_sym_c1a.argtypes = [C.c_size_t, C.c_size_t, _u8p, _u8p, _u8p, _f32p, C.c_size_t, _u32p]
        L.ref_encrypt_sym_c1a.restype = C.c_int
        L.ref_gen_pk.argtypes = [C.c_size_t, C.c_size_t, _u8p, _i8p, _u8p, _u32p, _u32p]
        for name, nargs in (("ref_barrett32", 2), ("ref_barrett64", 3), ("ref_mul_mod", 3), ("ref_add_mod", 3),
                            ("ref_sub_mod", 3)):
            f = getattr(L, name)
            f.argtypes = [C.c_uint32] * nargs
            f.restype = C.c_uint32

    # -- API level
    def setup(self, n: int, np_: int, asym: bool, sk=None, pk0=None, pk1=None, primes=None) -> None:
        """Writes the key files into a private temp dir, chdir()s there (the reference opens
        CWD-relative paths) and calls se_setup."""
        self._tmp = tempfile.TemporaryDirectory(prefix="seref_")
        write_key_files(self._tmp.name, n, primes or [], sk, pk0, pk1)
        ok = self.lib.ref_setup(n, np_, int(asym), self._tmp.name.encode())
        assert ok
        self.n, self.np_ = n, np_

    def close(self) -> None:
        if self.n:
            self.lib.ref_cleanup()
            self.n = 0
        os.chdir(self._cwd)
        if self._tmp is not None:
            self._tmp.cleanup()
            self._tmp = None

    def primes(self) -> list[int]:
        return [int(self.lib.ref_prime(i)) for i in range(self.lib.ref_nprimes())]

    def ratios(self) -> list[tuple[int, int]]:
        return [(int(self.lib.ref_ratio(i, 0)), int(self.lib.ref_ratio(i, 1))) for i in range(self.lib.ref_nprimes())]

    def scale(self) -> float:
        return float(self.lib.ref_scale())

    def encrypt_seeded(self, share_seed, seed, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float32)
        out = np.zeros((self.np_, 2, self.n), np.uint32)
        ss = _ptr(_seed(share_seed), _u8p) if share_seed is not None else None
        ok = self.lib.ref_encrypt_seeded(ss, _ptr(_seed(seed), _u8p), _ptr(v, _f32p), v.size * 4, _ptr(out, _u32p))
        return bool(ok), out

    def encrypt_loop(self, share_seeds, seeds: np.ndarray, values: np.ndarray) -> float:
        v = np.ascontiguousarray(values, dtype=np.float32)
        s = np.ascontiguousarray(seeds, dtype=np.uint8)
        ss = None
        if share_seeds is not None:
            share_seeds = np.ascontiguousarray(share_seeds, dtype=np.uint8)
            ss = _ptr(share_seeds, _u8p)
        out = np.zeros((self.np_, 2, self.n), np.uint32)
        return float(self.lib.ref_encrypt_loop(v.shape[0], ss, _ptr(s, _u8p), _ptr(v, _f32p), v.shape[1],
                                               _ptr(out, _u32p)))

    def encrypt_digests(self, share_seeds, seeds: np.ndarray, values: np.ndarray) -> np.ndarray:
        """se_encrypt_seeded over consecutive items; returns the per-item 64-bit digest of each byte stream
        (the function seb_digest_device computes on the GPU)."""
        v = np.ascontiguousarray(values, dtype=np.float32)
        s = np.ascontiguousarray(seeds, dtype=np.uint8)
        ss = None
        if share_seeds is not None:
            share_seeds = np.ascontiguousarray(share_seeds, dtype=np.uint8)
            ss = _ptr(share_seeds, _u8p)
        words = 2 * self.np_ * self.n
        scratch = np.zeros(words, np.uint32)
        out = np.zeros(v.shape[0], np.uint64)
        bad = self.lib.ref_encrypt_digests(v.shape[0], ss, _ptr(s, _u8p), _ptr(v, _f32p), v.shape[1], words,
                                           _ptr(scratch, _u32p), out.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert bad == 0, f"{bad} reference calls failed"
        return out

    # -- stage level
    def index_map(self, n: int) -> np.ndarray:
        out = np.empty(n, np.uint16)
        self.lib.ref_index_map(n, _ptr(out, _u16p))
        return out

    def encode(self, n: int, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float32)
        out = np.zeros(n, np.int64)
        ok = self.lib.ref_encode(n, _ptr(v, _f32p), v.size, _ptr(out, _i64p))
        return bool(ok), out

    def prng_fill(self, seed, counter: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, np.uint8)
        self.lib.ref_prng_fill(_ptr(_seed(seed), _u8p), counter, nbytes, _ptr(out, _u8p))
        return out

    def asym_init(self, n: int, seed, pt: np.ndarray):
        pte = np.array(pt, dtype=np.int64, copy=True)
        u = np.zeros(n // 4, np.uint8)
        e1 = np.zeros(n, np.int8)
        ctr = self.lib.ref_asym_init(n, _ptr(_seed(seed), _u8p), _ptr(pte, _i64p), _ptr(u, _u8p), _ptr(e1, _i8p))
        return u, pte, e1, int(ctr)

    def sample_ternary_small(self, n: int, seed, counter: int = 0):
        out = np.zeros(n // 4, np.uint8)
        c = self.lib.ref_sample_ternary_small(n, _ptr(_seed(seed), _u8p), counter, _ptr(out, _u8p))
        return out, int(c)

    def sample_cbd(self, n: int, seed, counter: int = 0):
        out = np.zeros(n, np.int8)
        c = self.lib.ref_sample_cbd(n, _ptr(_seed(seed), _u8p), counter, _ptr(out, _i8p))
        return out, int(c)

    def sample_uniform(self, n: int, np_: int, prime_idx: int, seed, counter: int = 0):
        out = np.zeros(n, np.uint32)
        c = self.lib.ref_sample_uniform(n, np_, prime_idx, _ptr(_seed(seed), _u8p), counter, _ptr(out, _u32p))
        return out, int(c)

    def ntt(self, n: int, np_: int, prime_idx: int, vec: np.ndarray) -> np.ndarray:
        v = np.array(vec, dtype=np.uint32, copy=True)
        self.lib.ref_ntt(n, np_, prime_idx, _ptr(v, _u32p))
        return v

    def ntt_loop_seconds(self, n: int, np_: int, prime_idx: int, reps: int) -> float:
        """Seconds for `reps` calls of the reference's ntt_inpl on one core (roots built once, outside)."""
        v = np.arange(n, dtype=np.uint32)
        return float(self.lib.ref_ntt_loop(n, np_, prime_idx, _ptr(v, _u32p), reps))

    def reduce_pte(self, n: int, np_: int, prime_idx: int, pte: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(pte, dtype=np.int64)
        out = np.zeros(n, np.uint32)
        self.lib.ref_reduce_pte(n, np_, prime_idx, _ptr(a, _i64p), _ptr(out, _u32p))
        return out

    def encrypt_sym_c1a(self, n: int, np_: int, share_seed, seed, sk: np.ndarray, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float32)
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        out = np.zeros((np_, 2, n), np.uint32)
        ok = self.lib.ref_encrypt_sym_c1a(n, np_, _ptr(_seed(share_seed), _u8p), _ptr(_seed(seed), _u8p),
                                          _ptr(sk, _u8p), _ptr(v, _f32p), v.size, _ptr(out, _u32p))
        return bool(ok), out

    def gen_pk(self, n: int, np_: int, sk: np.ndarray, ep_seed=bytes([7]) * 64, seed_base=bytes([9]) * 64):
        sk = np.ascontiguousarray(sk, dtype=np.uint8)
        ep, _ = self.sample_cbd(n, ep_seed)
        pk0 = np.zeros((np_, n), np.uint32)
        pk1 = np.zeros((np_, n), np.uint32)
        self.lib.ref_gen_pk(n, np_, _ptr(sk, _u8p), _ptr(ep, _i8p), _ptr(_seed(seed_base), _u8p), _ptr(pk0, _u32p),
                            _ptr(pk1, _u32p))
        return pk0, pk1
