"""CPU tests of the kernels' per-thread logic through tests/host_emul (the .cuh device functions
compiled with g++ and run one "thread" after another): index math, shared-memory layout, twiddle and
key-table orderings, Keccak and the samplers' bit tricks — all against the oracle or hashlib.
No product result is computed on the CPU here; the product path stays CUDA-only.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import struct

import numpy as np
import pytest

LOGNS = [10, 11, 12, 13, 14]
# NTT plan keys (seb_ntt.cuh NttCfg): log2(n) = 16 coefficients per thread; 16 + log2(n) = 32 per thread (n >= 8192)
KEYS = [10, 11, 12, 13, 14, 29, 30]


def _key_logn(key):
    return key & 15


def _key_e(key):
    return 32 if key & 16 else 16


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.fixture(scope="module")
def E(emul):
    emul.emul_ntt.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                              C.c_uint32, C.POINTER(C.c_uint32)]
    emul.emul_pad.argtypes = [C.c_int, C.c_uint32]
    emul.emul_pad.restype = C.c_uint32
    emul.emul_smem_words.argtypes = [C.c_int]
    emul.emul_smem_words.restype = C.c_uint32
    emul.emul_plan.argtypes = [C.c_int, C.POINTER(C.c_int)]
    emul.emul_ntt_elem.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    emul.emul_ntt_elem.restype = C.c_int64
    emul.emul_ntt_sync_scope.argtypes = [C.c_int, C.c_int]
    emul.emul_barrett64.argtypes = [C.c_uint64, C.c_uint32]
    emul.emul_barrett64.restype = C.c_uint32
    emul.emul_barrett32.argtypes = [C.c_uint32, C.c_uint32]
    emul.emul_barrett32.restype = C.c_uint32
    emul.emul_shoup_lazy.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    emul.emul_shoup_lazy.restype = C.c_uint32
    emul.emul_encode.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint16), C.POINTER(C.c_double),
                                 C.c_double, C.POINTER(C.c_int64)]
    emul.emul_encode_mag.restype = C.c_uint32
    emul.emul_enc_vskew.argtypes = [C.c_int, C.c_uint32]
    emul.emul_enc_vskew.restype = C.c_uint32
    emul.emul_enc_vwords.argtypes = [C.c_int]
    emul.emul_enc_vwords.restype = C.c_uint32
    emul.emul_enc_pos.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    emul.emul_enc_pos.restype = C.c_int64
    emul.emul_enc_sync_width.argtypes = [C.c_int, C.c_int]
    emul.emul_enc_e.restype = C.c_int
    emul.emul_keccak.argtypes = [C.POINTER(C.c_uint64)]
    emul.emul_prng_block.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64)]
    emul.emul_ternary_block.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32)]
    emul.emul_cbd_block.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint32)]
    emul.emul_cbd_block_plain.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint32)]
    emul.emul_keccak_il.argtypes = [C.POINTER(C.c_uint64)]
    emul.emul_prng_word_il.argtypes = [C.POINTER(C.c_uint8), C.c_uint64]
    emul.emul_prng_word_il.restype = C.c_uint32
    emul.emul_ternary_block_raw.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    emul.emul_mod3_bytes.argtypes = [C.c_uint32]
    emul.emul_mod3_bytes.restype = C.c_uint32
    return emul


def _plan(E, logn):
    r = (C.c_int * 4)()
    k = E.emul_plan(logn, r)
    return [r[i] for i in range(k)]


# ------------------------------------------------------------------------------------------------
# modular arithmetic helpers
# ------------------------------------------------------------------------------------------------
def test_barrett_and_shoup(E):
    rng = np.random.default_rng(1)
    for q in (134012929, 1053818881, 1062535169):
        xs = [0, 1, q - 1, q, q + 1, 2 * q, 0xFFFFFFFF, (1 << 63) - 1, (1 << 64) - 1] + \
             [int(x) for x in rng.integers(0, 1 << 63, 200, dtype=np.uint64)]
        for x in xs:
            assert E.emul_barrett64(x, q) == x % q
            assert E.emul_barrett32(x & 0xFFFFFFFF, q) == (x & 0xFFFFFFFF) % q
        for x, w in rng.integers(0, 1 << 32, (300, 2), dtype=np.uint64):
            x, w = int(x), int(w) % q
            r = E.emul_shoup_lazy(x, w, q)  # uintmodarith.h:308-331: [0, 2q), congruent to x*w
            assert r < 2 * q and r % q == (x * w) % q


# ------------------------------------------------------------------------------------------------
# NTT: plan, layout, transform
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", KEYS)
def test_ntt_plan(key, E):
    plan = _plan(E, key)
    logn, e = _key_logn(key), _key_e(key)
    assert sum(plan) == logn and all(r in (3, 4, 5) for r in plan)
    assert (1 << plan[0]) == e  # the first pass reads E strided coefficients per thread: one group
    assert all((1 << r) <= e for r in plan)


@pytest.mark.parametrize("key", KEYS)
def test_smem_layout_conflict_free(key, E):
    """seb_pad: (1) injective and inside NttSmem::WORDS; (2) additive over bit-disjoint fields, which is
    what turns every per-element address into base + immediate; (3) for every pass, the 32 lanes of
    every warp hit 32 distinct banks on each scalar access, and the last pass' 128-bit reads are
    conflict free per quarter-warp phase."""
    logn, epl = _key_logn(key), _key_e(key)
    n = 1 << logn
    T = n // epl
    pad = np.array([E.emul_pad(key, a) for a in range(n)], dtype=np.int64)
    words = E.emul_smem_words(key)
    assert len(set(pad.tolist())) == n and pad.max() < words and words % 4 == 0
    rng = np.random.default_rng(logn)
    for _ in range(2000):
        x = int(rng.integers(0, n))
        y = int(rng.integers(0, n)) & ~x
        assert E.emul_pad(key, x | y) == E.emul_pad(key, x) + E.emul_pad(key, y)
    plan = _plan(E, key)
    s0 = 0
    for pi, R in enumerate(plan):
        LS = logn - s0 - R
        GP = epl >> R
        last = pi == len(plan) - 1
        if pi > 0 or not last:
            for i in range(GP):
                for w0 in range(0, T, 32):
                    # element 0 of slot i of the warp's 32 threads, from the kernels' own mapping
                    base = np.array([E.emul_ntt_elem(key, pi, t, i, 0) for t in range(w0, w0 + 32)], dtype=np.int64)
                    if last and pi > 0:
                        # 128-bit reads: 8 lanes per phase, each 4 consecutive words
                        for k in range((1 << R) // 4):
                            addr = pad[base] + 4 * k
                            assert (addr % 4 == 0).all()
                            for ph in range(4):
                                banks = (addr[8 * ph: 8 * ph + 8] // 4) % 8
                                assert len(set(banks.tolist())) == 8, (key, pi, i, w0, k)
                    else:
                        for j in range(1 << R):
                            banks = pad[base | (j << LS)] % 32
                            assert len(set(banks.tolist())) == 32, (key, pi, i, j, w0)
        s0 += R


@pytest.mark.parametrize("key", KEYS)
def test_ntt_barrier_scopes(key, E):
    """Each pass covers every coefficient exactly once, and wherever the kernels replace the CTA barrier
    between two passes by a narrower one (NttSync: __syncwarp() or a named barrier over an aligned group of
    64 / 128 threads), every coefficient a thread reads in the later pass was written in the earlier one by
    a thread of the same warp / group."""
    logn, epl = _key_logn(key), _key_e(key)
    n = 1 << logn
    T = n // epl
    plan = _plan(E, key)
    owner = []
    for p, R in enumerate(plan):
        own = np.full(n, -1, np.int64)
        for t in range(T):
            for i in range(epl >> R):
                for j in range(1 << R):
                    e = E.emul_ntt_elem(key, p, t, i, j)
                    assert 0 <= e < n and own[e] == -1
                    own[e] = t
        assert (own >= 0).all()
        owner.append(own)
    scopes = []
    for p in range(len(plan) - 1):
        width = E.emul_ntt_sync_scope(key, p)  # 0: whole CTA, 32: warp, else threads per named barrier
        if width:
            shift = width.bit_length() - 1
            assert 1 << shift == width
            assert bool(((owner[p] >> shift) == (owner[p + 1] >> shift)).all()), (key, p, width)
            if width > 32:
                # hardware barriers 1..15 only: barrier 0 is __syncthreads()
                assert T % width == 0 and T // width <= 15
        scopes.append(width)
    # the 32-coefficient plans keep ONE CTA-wide barrier per transform
    expect = {10: [0, 32], 11: [0, 0], 12: [0, 32], 13: [0, 64, 32], 14: [0, 128, 32], 29: [0, 32], 30: [0, 32]}[key]
    assert scopes == expect
    assert E.emul_ntt_sync_scope(key, len(plan) - 1) == 0


@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("npoly", [1, 3])
def test_ntt_emulation_matches_oracle(key, npoly, E, orc):
    """The register-blocked multi-pass transform (seb_ntt_pass + per-pass twiddle tables + the
    epilogue-order key-table indexing) equals ntt_inpl (ntt.c:124-189) bit for bit."""
    logn = _key_logn(key)
    n = 1 << logn
    nprimes = {10: 1, 11: 1, 12: 3, 13: 6, 14: 13}[logn]
    rng = np.random.default_rng(100 + logn)
    for q in (orc.primes(n, nprimes)[0], orc.primes(n, nprimes)[-1]):
        roots = orc.ntt_roots(n, q)
        wq = ((roots.astype(np.uint64) << np.uint64(32)) // np.uint64(q)).astype(np.uint32)
        x = rng.integers(0, q, (npoly, n), dtype=np.uint32)
        x[0, :4] = (0, 1, q - 1, q - 2)
        out = np.zeros_like(x)
        rc = E.emul_ntt(key, npoly, _p(x, C.c_uint32), _p(roots, C.c_uint32), _p(wq, C.c_uint32), q,
                        _p(out, C.c_uint32))
        assert rc == 0
        for p in range(npoly):
            assert np.array_equal(out[p], orc.ntt(n, q, x[p])), (key, q, p)


def test_ntt_split_form_matches_oracle(E, orc):
    """The split form of the symmetric kernel at n = 16384 (k_encrypt_sym_split, seb_host_build_tw_sym): stage 0 pairs x[i]
    with x[i + n/2] under roots[1]; half r is then an independent 8192-point transform on the plan of n = 8192 whose root
    at local stage s, group j is roots[2^(s+1) + r 2^s + j] of the full table.  Here: stage 0 in numpy, the two halves through
    the emulated 32-coefficient plan with those tables, against ntt_inpl on the whole polynomial."""
    n, nh = 16384, 8192
    rng = np.random.default_rng(77)
    for q in (orc.primes(n, 6)[0], orc.primes(n, 13)[-1]):
        roots = orc.ntt_roots(n, q)
        x = rng.integers(0, q, n, dtype=np.uint32)
        x[:4] = (0, 1, q - 1, q - 2)
        a, b = x[:nh].astype(np.uint64), x[nh:].astype(np.uint64)
        t = (b * np.uint64(roots[1])) % np.uint64(q)
        halves = [((a + t) % np.uint64(q)).astype(np.uint32), ((a + np.uint64(q) - t) % np.uint64(q)).astype(np.uint32)]
        got = np.zeros(n, np.uint32)
        for r in range(2):
            rh = np.zeros(nh, np.uint32)
            rh[0] = roots[0]
            for s in range(13):
                rh[(1 << s):(2 << s)] = roots[(2 << s) + (r << s):(2 << s) + (r << s) + (1 << s)]
            wq = ((rh.astype(np.uint64) << np.uint64(32)) // np.uint64(q)).astype(np.uint32)
            out = np.zeros((1, nh), np.uint32)
            xin = halves[r].reshape(1, nh).copy()
            assert E.emul_ntt(29, 1, _p(xin, C.c_uint32), _p(rh, C.c_uint32), _p(wq, C.c_uint32), q, _p(out, C.c_uint32)) == 0
            got[r * nh:(r + 1) * nh] = out[0]
        assert np.array_equal(got, orc.ntt(n, q, x)), q


def test_ntt_emulation_lazy_inputs(E, orc):
    """The fused encrypt feeds on-load values anywhere below 4q (e.g. q - 1 + small): inputs in [0,4q)
    must still give the canonical transform of their residues."""
    logn, n, q = 12, 4096, 1053818881
    rng = np.random.default_rng(9)
    x = rng.integers(0, 4 * q, (1, n), dtype=np.uint64).astype(np.uint32)
    roots = orc.ntt_roots(n, q)
    wq = ((roots.astype(np.uint64) << np.uint64(32)) // np.uint64(q)).astype(np.uint32)
    out = np.zeros_like(x)
    assert E.emul_ntt(logn, 1, _p(x, C.c_uint32), _p(roots, C.c_uint32), _p(wq, C.c_uint32), q, _p(out, C.c_uint32)) == 0
    assert np.array_equal(out[0], orc.ntt(n, q, (x[0] % np.uint32(q))))


# ------------------------------------------------------------------------------------------------
# encode
# ------------------------------------------------------------------------------------------------
def _src_map(orc, n):
    im = orc.index_map(n)
    src = np.zeros(n, np.uint16)
    src[im[: n // 2]] = np.arange(n // 2, dtype=np.uint16)
    src[im[n // 2:]] = np.arange(n // 2, dtype=np.uint16)
    return src


@pytest.mark.parametrize("lognl", [10, 11, 12, 13])
def test_encode_barrier_scopes(lognl, E):
    """Every encode pass touches every position of the CTA exactly once, and wherever the kernel synchronises
    less than the CTA between two passes (enc_sync_width: a warp after pass 0, a named barrier over 64 or 128
    threads after pass 1), every position a thread reads in the later pass was written in the earlier one by
    a thread of the same unit.  Named barriers stay within hardware barriers 1..15."""
    nl = 1 << lognl
    epl = E.emul_enc_e()  # complex values per thread the library is compiled for: 8 or 16
    lr = epl.bit_length() - 1
    T = nl // epl
    npass = (lognl + lr - 1) // lr
    owner = []
    for p in range(npass):
        R = lr if lognl - lr * p >= lr else lognl - lr * p
        own = np.full(nl, -1, np.int64)
        for t in range(T):
            for i in range(epl >> R):
                for j in range(1 << R):
                    pos = E.emul_enc_pos(lognl, p, t, i, j)
                    assert 0 <= pos < nl and own[pos] == -1
                    own[pos] = t
        assert (own >= 0).all()
        owner.append(own)
    widths = []
    for p in range(npass - 1):
        w = E.emul_enc_sync_width(lognl, p)
        if w:
            shift = w.bit_length() - 1
            assert 1 << shift == w
            assert bool(((owner[p] >> shift) == (owner[p + 1] >> shift)).all()), (lognl, p, w)
            if w > 32:
                assert T % w == 0 and T // w <= 15
        widths.append(w)
    if epl == 8:
        assert widths[:2] == [32, 128 if T > 960 else 64] and all(w == 0 for w in widths[2:])
    else:
        assert widths[:2] == [32, 256 if T > 256 else 0] and all(w == 0 for w in widths[2:])


@pytest.mark.parametrize("logn", LOGNS)
def test_encode_gather_conflict_free(logn, E, orc):
    """Pass 0 of the encode gathers values[src_map[pos]] for 8 consecutive positions per thread from the
    message staged in shared memory.  enc_vskew is injective, stays inside EncVals::WORDS, and makes the 32
    lanes of every warp hit 32 distinct banks in every gather instruction (a linear layout collides up to
    32-way at n = 16384); the staging stores stay at most 2-way."""
    n = 1 << logn
    src = _src_map(orc, n).astype(np.int64)
    phys = np.array([E.emul_enc_vskew(logn, s) for s in range(n // 2)], dtype=np.int64)
    assert len(set(phys.tolist())) == n // 2 and phys.max() < E.emul_enc_vwords(logn)
    nl = n // 2 if logn == 14 else n  # positions per CTA (n = 16384 runs as a 2-CTA cluster)
    epl = E.emul_enc_e()
    T = nl // epl
    worst_linear = 0
    allowed = 2 if (epl == 16 and logn == 10) else 1  # 16 positions per thread at n = 1024: two-way at worst
    for cta0 in range(0, n, nl):
        for w0 in range(0, T, 32):
            g = np.arange(w0, min(w0 + 32, T))
            for j in range(epl):
                slots = src[cta0 + epl * g + j]
                assert np.bincount(phys[slots] % 32, minlength=32).max() <= allowed, (logn, cta0, w0, j)
                worst_linear = max(worst_linear, int(np.bincount(slots % 32, minlength=32).max()))
    if epl == 8:
        assert worst_linear == {10: 2, 11: 4, 12: 8, 13: 16, 14: 32}[logn]
    for i0 in range(0, n // 2, 32):  # staging: 32 consecutive slots per store instruction
        assert np.bincount(phys[i0:i0 + 32] % 32, minlength=32).max() <= 2


@pytest.mark.parametrize("logn", LOGNS)
def test_encode_emulation_matches_oracle(logn, E, orc, oracle_mod):
    """Radix-8 passes of fused radix-2 stages, swizzled shared memory, the 2-CTA split for
    n = 16384: same int64 coefficients as ckks_encode_base (ckks_common.c:105-215), bit for bit."""
    n = 1 << logn
    tw = orc.ifft_twiddles(n)
    src = _src_map(orc, n)
    n_inv = orc.scale(n) / n
    for vlen, seed in ((n // 2, 1), (n // 2, 2), (5, 3), (0, 4)):
        v = oracle_mod.make_values(1, n // 2, seed=seed * 100 + logn)[0]
        out = np.zeros(n, np.int64)
        bad = E.emul_encode(logn, _p(v, C.c_float), vlen, _p(src, C.c_uint16), _p(tw, C.c_double), n_inv,
                            _p(out, C.c_int64))
        ok, exp = orc.encode(n, v[:vlen])
        assert bad == 0 and ok and np.array_equal(out, exp), (logn, vlen)
        # per-item magnitude handed to the encrypt kernels: max |coefficient| clipped to 32 bits
        assert E.emul_encode_mag() == min(int(np.abs(exp).max()), 0xFFFFFFFF)
    # overflow flag (ckks_common.c:195-204)
    v = np.full(n // 2, 3.0e38, np.float32)
    out = np.zeros(n, np.int64)
    assert E.emul_encode(logn, _p(v, C.c_float), n // 2, _p(src, C.c_uint16), _p(tw, C.c_double), n_inv,
                         _p(out, C.c_int64)) == 1


# ------------------------------------------------------------------------------------------------
# Keccak / samplers
# ------------------------------------------------------------------------------------------------
def test_keccak_permutation_and_prng_block(E, oracle_mod):
    st = (C.c_uint64 * 25)()
    E.emul_keccak(st)  # Keccak-f[1600] of the zero state: first lane is the well-known 0xF1258F7940E1DDE7
    assert st[0] == 0xF1258F7940E1DDE7 and st[24] == 0xEAF1FF7B5CECA249
    seeds = oracle_mod.make_seeds(4, b"kk")
    for i, ctr in enumerate((0, 1, 0xFFFFFFFF, 1 << 50)):
        out = np.zeros(17, np.uint64)
        E.emul_prng_block(_p(seeds[i], C.c_uint8), ctr, _p(out, C.c_uint64))
        assert out.tobytes() == hashlib.shake_256(seeds[i].tobytes() + struct.pack("<Q", ctr)).digest(136)


def test_keccak_bit_interleaved(E):
    """The bit-interleaved round function (even/odd halves, 174 operations) is the same permutation."""
    st = (C.c_uint64 * 25)()
    E.emul_keccak_il(st)
    assert st[0] == 0xF1258F7940E1DDE7 and st[24] == 0xEAF1FF7B5CECA249
    rng = np.random.default_rng(11)
    for _ in range(8):
        v = rng.integers(0, 1 << 64, 25, dtype=np.uint64)
        a = (C.c_uint64 * 25)(*[int(x) for x in v])
        b = (C.c_uint64 * 25)(*[int(x) for x in v])
        E.emul_keccak(a)
        E.emul_keccak_il(b)
        assert list(a) == list(b)


def test_prng_word_interleaved(E, oracle_mod):
    """A 4-byte PRNG call through the interleaved sponge with round 0 folded and the last round pruned to one word."""
    seeds = oracle_mod.make_seeds(16, b"w0")
    for i in range(16):
        for ctr in (0, 1, 77 + i, 0xFFFF, 0x12345678, 0xFFFFFFFF, (1 << 32) + 5, (1 << 63) + 12345):
            want = struct.unpack("<I", hashlib.shake_256(seeds[i].tobytes() + struct.pack("<Q", ctr)).digest(4))[0]
            assert E.emul_prng_word_il(_p(seeds[i], C.c_uint8), ctr) == want


def test_mod3_bytes_exhaustive(E):
    for b in range(256):
        x = b | ((255 - b) << 8) | (((b * 7) & 0xFF) << 16) | (((b * 13 + 5) & 0xFF) << 24)
        r = E.emul_mod3_bytes(x)
        for k in range(4):
            assert (r >> (8 * k)) & 0xFF == ((x >> (8 * k)) & 0xFF) % 3


def test_ternary_block_every_byte_value(E):
    """The multiply-based mod 3 / flag gathering of seb_ternary_block on every byte value at every position of a word,
    beside neighbours that make the 16-bit partial products overflow into each other (0xFF) or not (0x00)."""
    rng = np.random.default_rng(5)
    for fill in (0x00, 0xFF, 0xFD, None):
        for lane in range(4):
            for base in range(0, 256, 24):
                blk = np.full(96, fill, np.uint8) if fill is not None else rng.integers(0, 256, 96, dtype=np.uint8)
                vals = [(base + i) & 0xFF for i in range(24)]
                for i, v in enumerate(vals):
                    blk[4 * i + lane] = v
                packed = np.zeros(6, np.uint32)
                mask = np.zeros(3, np.uint32)
                E.emul_ternary_block_raw(_p(blk, C.c_uint8), _p(packed, C.c_uint32), _p(mask, C.c_uint32))
                pb = packed.tobytes()
                for pos in range(96):
                    rej = int(blk[pos]) >= 0xFE
                    assert ((int(mask[pos // 32]) >> (pos % 32)) & 1) == int(rej)
                    if not rej:
                        assert (pb[pos // 4] >> (6 - 2 * (pos % 4))) & 3 == int(blk[pos]) % 3


def test_ternary_and_cbd_blocks(E, oracle_mod):
    """One 96-byte PRNG call as a ternary block (packed fields + rejection masks, sample.c:223-241) and as
    16 CBD samples (sample.c:263-321)."""
    seeds = oracle_mod.make_seeds(64, b"blk")
    seen_reject = 0
    for i in range(64):
        ctr = i * 3
        buf = hashlib.shake_256(seeds[i].tobytes() + struct.pack("<Q", ctr)).digest(96)
        packed = np.zeros(6, np.uint32)
        mask = np.zeros(3, np.uint32)
        E.emul_ternary_block(_p(seeds[i], C.c_uint8), ctr, _p(packed, C.c_uint32), _p(mask, C.c_uint32))
        pb = packed.tobytes()
        for pos in range(96):
            rej = buf[pos] >= 0xFE
            assert ((int(mask[pos // 32]) >> (pos % 32)) & 1) == int(rej)
            field = (pb[pos // 4] >> (6 - 2 * (pos % 4))) & 3
            assert rej or field == buf[pos] % 3  # a rejected field is a don't-care: the walk overwrites it
            seen_reject += rej
        o = np.zeros(4, np.uint32)
        E.emul_cbd_block(_p(seeds[i], C.c_uint8), ctr, _p(o, C.c_uint32))
        got = np.frombuffer(o.tobytes(), np.int8)
        hw = lambda v: bin(v).count("1")  # noqa: E731
        for s in range(16):
            x = buf[6 * s: 6 * s + 6]
            assert got[s] == hw(x[0]) + hw(x[1]) + hw(x[2] & 0x1F) - hw(x[3]) - hw(x[4]) - hw(x[5] & 0x1F)
        o2 = np.zeros(4, np.uint32)
        E.emul_cbd_block_plain(_p(seeds[i], C.c_uint8), ctr, _p(o2, C.c_uint32))
        assert np.array_equal(o, o2)
    assert seen_reject > 0


# ------------------------------------------------------------------------------------------------
# ternary sampler: the pointer-doubling resolution of a wave against the in-order walk it replaces
# ------------------------------------------------------------------------------------------------
def _walk_sequential(st, lo, hi, acc, need_full, need_last, nblocks):
    """sample.c:223-241 as the kernel's first version walked it: counter i is a redraw of the current block while it still
    owes one (consumed whether or not its byte is acceptable), else the next block, else the walk is over."""
    j, need, cur, served, consumed = st
    is_blk = is_red = 0
    owner = -1
    for i in range(lo, hi):
        if need > 0:
            if (acc >> i) & 1:
                is_red |= 1 << i
                served += 1
                need -= 1
        elif j < nblocks:
            need = need_last[i] if j == nblocks - 1 else need_full[i]
            is_blk |= 1 << i
            owner, cur, served = i, j, 0
            j += 1
        else:
            break
        consumed += 1
    return (j, need, cur, served, consumed), is_blk, is_red, owner


def _walk_doubling(st, lo, hi, acc, need_full, need_last, nblocks):
    """seb_tern_walk (seb_sample.cu) restated: carried block first, "next block" links per lane, five doubling rounds."""
    j, need, cur, served, consumed = st
    below = lambda p: (1 << min(p, 32)) - 1  # noqa: E731
    A = acc & below(hi) & ~below(lo)

    def nth_after(mask, k):  # lane after the k-th (1-based) set bit of mask
        for _ in range(k - 1):
            mask &= mask - 1
        return (mask & -mask).bit_length()

    p0, need_c = lo, 0
    if need > 0:
        have = bin(A).count("1")
        if have < need:
            p0, need_c = 32, need - have
        else:
            p0 = nth_after(A, need)
    is_blk = 0
    avail = nblocks - j
    if p0 < hi and avail > 0:
        nxt = []
        for lane in range(32):
            above = A & ~below(lane + 1)
            if need_full[lane] == 0:
                e = lane + 1
            elif bin(above).count("1") < need_full[lane]:
                e = 32
            else:
                e = nth_after(above, need_full[lane])
            nxt.append(32 if e >= hi else e)
        is_blk = 1 << p0
        for _ in range(5):
            add = 0
            for lane in range(32):
                if (is_blk >> lane) & 1 and nxt[lane] < 32:
                    add |= 1 << nxt[lane]
            is_blk |= add
            nxt = [nxt[nxt[lane]] if nxt[lane] < 32 else 32 for lane in range(32)]
        if bin(is_blk).count("1") > avail:
            m, keep = is_blk, 0
            for _ in range(avail):
                keep |= m & -m
                m &= m - 1
            is_blk = keep
    nb = bin(is_blk).count("1")
    owner, end = -1, hi
    j += nb
    if nb > 0:
        owner = is_blk.bit_length() - 1
        n_b = need_last[owner] if j == nblocks else need_full[owner]
        rest = A & ~below(owner + 1)
        got = bin(rest).count("1")
        cur = j - 1
        if got >= n_b:
            need, served = 0, n_b
            if j == nblocks:
                end = owner + 1 if n_b == 0 else nth_after(rest, n_b)
        else:
            need, served = n_b - got, got
    else:
        served += bin(A & below(p0)).count("1")
        need = need_c
        if need_c == 0 and j == nblocks:
            end = p0
    consumed += end - lo
    return (j, need, cur, served, consumed), is_blk, A & ~is_blk & below(end), owner


def test_ternary_walk_doubling_equals_sequential():
    """Random waves (dense and sparse acceptable bytes, blocks with 0..4 rejections, carried needs, split waves, the
    ciphertext ending inside the wave) chained over several waves per state: same roles, same state, same counter."""
    rng = np.random.default_rng(2024)
    for trial in range(3000):
        nblocks = int(rng.integers(1, 60))
        st_a = st_b = (0, 0, 0, 0, 0)
        if rng.random() < 0.3:  # start in the middle of a ciphertext, possibly owing redraws
            j0 = int(rng.integers(0, nblocks + 1))
            st_a = st_b = (j0, int(rng.integers(0, 5)) if j0 > 0 else 0, max(j0 - 1, 0), int(rng.integers(0, 3)), 17)
        for wave in range(4):
            p_acc = rng.choice([0.992, 0.9, 0.5])
            acc = int(sum(1 << i for i in range(32) if rng.random() < p_acc))
            need_full = [int(v) for v in rng.choice([0, 0, 0, 1, 1, 2, 3, 4], 32)]
            need_last = [min(v, int(rng.integers(0, 3))) for v in need_full]
            lo = int(rng.choice([0, 0, 0, 5, 16]))
            hi = int(rng.choice([32, 32, 32, 11, 20]))
            if hi <= lo:
                lo, hi = 0, 32
            ra = _walk_sequential(st_a, lo, hi, acc, need_full, need_last, nblocks)
            rb = _walk_doubling(st_b, lo, hi, acc, need_full, need_last, nblocks)
            assert ra[1] == rb[1] and ra[2] == rb[2], (trial, wave, "roles")
            sa, sb = ra[0], rb[0]
            assert (sa[0], sa[1], sa[4]) == (sb[0], sb[1], sb[4]), (trial, wave, sa, sb)
            if sa[1] > 0:  # the carried block and what it has been served only matter while it still owes something
                assert (sa[2], sa[3]) == (sb[2], sb[3]) and (ra[3] >= 0) == (rb[3] >= 0), (trial, wave, sa, sb)
            st_a, st_b = sa, sb


# ------------------------------------------------------------------------------------------------
# uniform sampler: the streamed fix-up's lane assignment against the per-ciphertext order of sample.c:39-57
# ------------------------------------------------------------------------------------------------
def test_uniform_fix_stream_assignment():
    """k_uniform_fix_stream restated: a warp serves K ciphertexts in turn; a wave gives the current one as many lanes as it
    still has rejected words and the lanes behind them to the next one.  For every ciphertext the candidates must be
    examined in counter order without gaps, the i-th accepted one must land in its i-th rejected word, and the counter
    must end behind the last candidate consumed - exactly what one ciphertext alone would do."""
    rng = np.random.default_rng(99)
    for trial in range(300):
        K = int(rng.choice([2, 4, 8]))
        cnt = [int(rng.choice([0, 1, 5, 31, 32, 33, 76, 150])) for _ in range(K)]
        c0 = [int(rng.integers(0, 1000)) for _ in range(K)]
        p_acc = float(rng.choice([0.98, 0.75]))
        ok = [rng.random(4 * max(cnt) + 200) < p_acc for _ in range(K)]  # ok[k][t]: candidate c0 + 1 + t acceptable
        # reference: each ciphertext by itself
        want_fill, want_ctr = [], []
        for k in range(K):
            idx = np.flatnonzero(ok[k])[:cnt[k]]
            want_fill.append(list(idx))
            want_ctr.append(c0[k] + 1 + (int(idx[-1]) + 1 if cnt[k] else 0))
        # the stream
        fill = [[] for _ in range(K)]
        ctr = [None] * K
        done = [0] * K
        base = [c + 1 for c in c0]
        last = list(c0)
        a = 0
        guard = 0
        while a < K:
            guard += 1
            assert guard < 10000
            if done[a] == cnt[a]:
                ctr[a] = last[a] + 1
                a += 1
                continue
            s = min(32, cnt[a] - done[a])
            b = a + 1
            nb = min(32 - s, cnt[b] - done[b]) if b < K else 0
            for k, lanes in ((a, s), (b, nb)):
                if lanes == 0:
                    continue
                t0 = base[k] - (c0[k] + 1)
                for lane in range(lanes):
                    if ok[k][t0 + lane]:
                        fill[k].append(t0 + lane)
                        last[k] = base[k] + lane
                done[k] = len(fill[k])
                base[k] += lanes
        assert fill == want_fill and ctr == want_ctr, trial
