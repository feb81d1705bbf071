"""CPU tests of the ORACLE (oracle/se_oracle.c): it must reproduce every vector the reference pins
for this path before the GPU parity tests may trust it.

  * scalar known answers of device/test/modulo_tests.c:78-179 and uintmodarith_tests.c:179-195
    (tests/golden/kat.npz, re-evaluated through the reference's own functions);
  * stage dumps and full-ciphertext digests produced by the UNMODIFIED reference library
    (tests/golden/{stage,encrypt}_golden.npz, generator: tests/golden/make_golden.py);
  * live differential tests against oracle/_ref/libseref.so where that prebuilt file exists;
  * the reference's own property tests (ntt.ntt vs schoolbook, encrypt -> decrypt -> decode within
    0.1: device/test/ntt_tests.c, ckks_tests_common.c:173-231).
"""
from __future__ import annotations

import hashlib
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN

ALL_CONFIGS = [(1024, 1), (2048, 1), (4096, 3), (8192, 6), (16384, 13)]


# ------------------------------------------------------------------------------------------------
# scalar KATs
# ------------------------------------------------------------------------------------------------
def test_kat_barrett_mulmod(orc):
    k = np.load(os.path.join(GOLDEN, "kat.npz"))
    for x, q, r in k["barrett32"]:
        assert orc.lib.orc_barrett32(int(x), int(q)) == int(r) == int(x) % int(q)
    for lo, hi, q, r in k["barrett64"]:
        assert orc.lib.orc_barrett64(int(lo), int(hi), int(q)) == int(r) == ((int(hi) << 32) | int(lo)) % int(q)
    for a, b, q, r in k["mul_mod"]:
        assert orc.lib.orc_mul_mod(int(a), int(b), int(q)) == int(r)


def test_reference_pinned_constants(orc):
    """device/test/modulo_tests.c:100 pins MAX_ZZ mod q; ntt.c:213-289 tabulates psi; each psi must be a
    primitive 2n-th root (psi^n = -1) of its prime."""
    assert 0xFFFFFFFF % 1053818881 == 79691771
    assert orc.lib.orc_barrett32(0xFFFFFFFF, 1053818881) == 79691771
    for n, np_ in ALL_CONFIGS:
        for q in orc.primes(n, np_):
            psi = orc.ntt_root(n, q)
            assert psi and orc.lib.orc_pow_mod(psi, n, q) == q - 1, (n, q)
            lo, hi = orc.const_ratio(q)
            assert ((hi << 32) | lo) == (1 << 64) // q
    assert orc.scale(1024) == 2.0 ** 20 and orc.scale(4096) == 2.0 ** 25  # parameters.c:197-225
    with pytest.raises(ValueError):
        orc.primes(4096, 4)  # parameters.c:204-213: n=4096 takes at most 3 primes


def test_add_sub_neg_mod(orc):
    """uintmodarith.h:26-88"""
    q = 1053818881
    L = orc.lib
    assert L.orc_add_mod(q - 1, q - 1, q) == q - 2
    assert L.orc_add_mod(0, 0, q) == 0
    assert L.orc_neg_mod(0, q) == 0 and L.orc_neg_mod(1, q) == q - 1
    assert L.orc_sub_mod(0, 1, q) == q - 1 and L.orc_sub_mod(5, 5, q) == 0
    rng = np.random.default_rng(0)
    for a, b in rng.integers(0, q, (200, 2)):
        a, b = int(a), int(b)
        assert L.orc_add_mod(a, b, q) == (a + b) % q
        assert L.orc_sub_mod(a, b, q) == (a - b) % q
        assert L.orc_mul_mod(a, b, q) == (a * b) % q


# ------------------------------------------------------------------------------------------------
# PRNG
# ------------------------------------------------------------------------------------------------
def test_prng_is_shake256(orc):
    """rng.h:78-91: every fill is SHAKE256(seed || LE64(counter)); Appendix A pins two outputs."""
    seed = bytes([2]) * 64
    out = orc.prng_fill(seed, 0, 200).tobytes()
    assert out == hashlib.shake_256(seed + struct.pack("<Q", 0)).digest(200)
    assert out[:8].hex() == "497823cc8f417a86" and out[-4:].hex() == "c4bbfb7f"
    assert orc.prng_fill(seed, 1, 8).tobytes().hex() == "820fa868f5eecf76"
    for ctr, nbytes in ((0, 1), (7, 96), (1 << 40, 136), (3, 137), (9, 4096 * 4)):
        assert orc.prng_fill(seed, ctr, nbytes).tobytes() == \
            hashlib.shake_256(seed + struct.pack("<Q", ctr)).digest(nbytes)
    for data in (b"", b"abc", bytes(range(200))):
        assert orc.shake256(data, 64) == hashlib.shake_256(data).digest(64)


# ------------------------------------------------------------------------------------------------
# stage vectors from the compiled reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1024, 4096])
def test_stage_golden(n, orc):
    g = np.load(os.path.join(GOLDEN, "stage_golden.npz"))
    v, seed = g[f"n{n}_values"], g[f"n{n}_seed"]
    assert np.array_equal(orc.index_map(n), g[f"n{n}_index_map"])
    ok, pt = orc.encode(n, v)
    assert ok and np.array_equal(pt, g[f"n{n}_pt"])
    u, ctr = orc.sample_ternary_small(n, seed, 0)
    assert np.array_equal(u, g[f"n{n}_u"])
    e0, ctr = orc.sample_cbd(n, seed, ctr)
    assert np.array_equal(pt + e0.astype(np.int64), g[f"n{n}_pte"])
    e1, ctr = orc.sample_cbd(n, seed, ctr)
    assert np.array_equal(e1, g[f"n{n}_e1"])
    assert ctr == int(g[f"n{n}_ctr"][0])
    np_ = g[f"n{n}_ntt_ramp"].shape[0]
    for p, q in enumerate(orc.primes(n, np_)):
        assert np.array_equal(orc.ntt(n, q, g[f"n{n}_ramp"]), g[f"n{n}_ntt_ramp"][p])
    a, c = orc.sample_uniform(n, orc.primes(n, np_)[0], seed, 0)
    assert np.array_equal(a, g[f"n{n}_uniform_p0"]) and c == int(g[f"n{n}_uniform_ctr"][0])


def _golden_keys():
    g = np.load(os.path.join(GOLDEN, "encrypt_golden.npz"))
    return [k[: -len("_digest")] for k in g.files if k.endswith("_digest")]


@pytest.mark.parametrize("key", _golden_keys())
def test_encrypt_golden(key, orc, oracle_mod):
    """sha256 of the byte stream se_encrypt_seeded sent (seal_embedded.c:145-213), per item."""
    g = np.load(os.path.join(GOLDEN, "encrypt_golden.npz"))
    n, np_, asym = (int(x) for x in g[key + "_cfg"])
    vals, seeds, sseeds = g[key + "_values"], g[key + "_seeds"], g[key + "_sseeds"]
    sk = oracle_mod.make_sk(n)
    if asym:
        pk0, pk1 = orc.gen_pk(n, np_, sk)
        assert hashlib.sha256(pk0.tobytes() + pk1.tobytes()).digest() == g[key + "_pkdigest"].tobytes()
    for b in range(vals.shape[0]):
        if asym:
            ok, ct = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
        else:
            # the reference's se_encrypt stream carries ntt(m+e) in the c1 slot
            ok, ct = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk, ref_quirk=True)
        assert ok
        assert hashlib.sha256(ct.tobytes()).digest() == g[key + "_digest"][b].tobytes(), (key, b)
        if b == 0 and key + "_ct0" in g.files:
            assert np.array_equal(ct, g[key + "_ct0"].reshape(ct.shape))


# ------------------------------------------------------------------------------------------------
# live differential tests against the compiled reference (prebuilt file; skipped when absent)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libseref.so not built")
    return oracle_mod.ReferenceLib()


def test_vs_reference_stages(ref, orc, oracle_mod):
    rng = np.random.default_rng(5)
    for n, np_ in ((1024, 1), (4096, 3), (8192, 4)):
        seed = oracle_mod.make_seeds(1, b"diff-%d" % n)[0]
        v = oracle_mod.make_values(1, n // 2, seed=n)[0]
        assert np.array_equal(ref.index_map(n), orc.index_map(n))
        okr, ptr_ = ref.encode(n, v)
        oko, pto = orc.encode(n, v)
        assert okr and oko and np.array_equal(ptr_, pto)
        # partial inputs are zero padded (fresh pool)
        okr, ptr2 = ref.encode(n, v[:37])
        oko, pto2 = orc.encode(n, v[:37])
        assert okr and oko and np.array_equal(ptr2, pto2)
        ur, cr = ref.sample_ternary_small(n, seed, 0)
        uo, co = orc.sample_ternary_small(n, seed, 0)
        assert np.array_equal(ur, uo) and cr == co
        er, cr2 = ref.sample_cbd(n, seed, cr)
        eo, co2 = orc.sample_cbd(n, seed, co)
        assert np.array_equal(er, eo) and cr2 == co2
        assert ref.prng_fill(seed, 11, 333).tobytes() == orc.prng_fill(seed, 11, 333).tobytes()
        ctr = 0
        for p, q in enumerate(orc.primes(n, np_)):
            x = rng.integers(0, q, n, dtype=np.uint32)
            assert np.array_equal(ref.ntt(n, np_, p, x), orc.ntt(n, q, x))
            pte = rng.integers(-(1 << 40), 1 << 40, n, dtype=np.int64)
            assert np.array_equal(ref.reduce_pte(n, np_, p, pte), orc.reduce_pte(n, q, pte))
            ar, cra = ref.sample_uniform(n, np_, p, seed, ctr)
            ao, coa = orc.sample_uniform(n, q, seed, ctr)
            assert np.array_equal(ar, ao) and cra == coa
            ctr = coa


@pytest.mark.parametrize("n,np_,asym", [(1024, 1, False), (4096, 3, True), (4096, 3, False), (8192, 4, True)])
def test_vs_reference_encrypt(n, np_, asym, oracle_mod, orc):
    """se_encrypt_seeded of the unmodified reference vs the restatement, several items, partial vlen."""
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libseref.so not built")
    import multiprocessing as mp

    # the reference keeps static state (seal_embedded.c:18-22): one process per configuration
    with mp.get_context("spawn").Pool(1) as pool:
        digests = pool.apply(_ref_encrypt_digests, ((n, np_, asym),))
    sk = oracle_mod.make_sk(n)
    vals = oracle_mod.make_values(3, n // 2, seed=77 + n)
    seeds = oracle_mod.make_seeds(3, b"diffenc")
    sseeds = oracle_mod.make_seeds(3, b"diffenc-share")
    pk0 = pk1 = None
    if asym:
        pk0, pk1 = orc.gen_pk(n, np_, sk)
    vlens = [n // 2, n // 2, 100]
    for b in range(3):
        v = vals[b][: vlens[b]]
        if asym:
            ok, ct = orc.encrypt_asym(n, np_, v, seeds[b], pk0, pk1)
        else:
            ok, ct = orc.encrypt_sym(n, np_, v, sseeds[b], seeds[b], sk, ref_quirk=True)
        assert ok and hashlib.sha256(ct.tobytes()).hexdigest() == digests[b], (n, np_, asym, b)


def _ref_encrypt_digests(cfg):
    from oracle import oracle as O

    n, np_, asym = cfg
    ref = O.ReferenceLib()
    sk = O.make_sk(n)
    pk0, pk1 = ref.gen_pk(n, np_, sk)
    ref.setup(n, np_, bool(asym), sk=sk, pk0=pk0, pk1=pk1, primes=O.Oracle().primes(n, np_))
    vals = O.make_values(3, n // 2, seed=77 + n)
    seeds = O.make_seeds(3, b"diffenc")
    sseeds = O.make_seeds(3, b"diffenc-share")
    vlens = [n // 2, n // 2, 100]
    out = []
    # the 100-value item comes LAST so no stale slots from a longer earlier call matter... they do:
    # the reference keeps slots past the input from the previous call, so re-setup first
    for b in range(3):
        if vlens[b] != n // 2:
            ref.close()
            ref.setup(n, np_, bool(asym), sk=sk, pk0=pk0, pk1=pk1, primes=O.Oracle().primes(n, np_))
        ok, ct = ref.encrypt_seeded(sseeds[b], seeds[b], vals[b][: vlens[b]])
        assert ok
        out.append(hashlib.sha256(ct.tobytes()).hexdigest())
    ref.close()
    return out


def test_gen_pk_matches_reference(oracle_mod, orc):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libseref.so not built")
    ref = oracle_mod.ReferenceLib()
    for n, np_ in ((1024, 1), (4096, 3)):
        sk = oracle_mod.make_sk(n)
        r0, r1 = ref.gen_pk(n, np_, sk)
        o0, o1 = orc.gen_pk(n, np_, sk)
        assert np.array_equal(r0, o0) and np.array_equal(r1, o1)


# ------------------------------------------------------------------------------------------------
# the reference's property tests, on the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,np_", [(1024, 1), (2048, 1), (4096, 3)])
def test_ntt_properties(n, np_, orc):
    """device/test/ntt_tests.c: intt(ntt(a) . ntt(b)) equals the schoolbook negacyclic product;
    intt(ntt(a)) = a; outputs canonical."""
    rng = np.random.default_rng(n)
    for q in orc.primes(n, np_):
        a = rng.integers(0, q, n, dtype=np.uint32)
        b = rng.integers(0, q, n, dtype=np.uint32)
        A, B = orc.ntt(n, q, a), orc.ntt(n, q, b)
        assert A.max() < q
        assert np.array_equal(orc.intt(n, q, A), a)
        prod = ((A.astype(np.uint64) * B.astype(np.uint64)) % np.uint64(q)).astype(np.uint32)
        assert np.array_equal(orc.intt(n, q, prod), orc.negacyclic_mul(n, q, a, b))
    # edge vectors: zero, one, all q-1
    q = orc.primes(n, np_)[0]
    assert not orc.ntt(n, q, np.zeros(n, np.uint32)).any()
    one = np.zeros(n, np.uint32)
    one[0] = 1
    assert np.array_equal(orc.ntt(n, q, one), np.ones(n, np.uint32))
    top = np.full(n, q - 1, np.uint32)
    assert np.array_equal(orc.intt(n, q, orc.ntt(n, q, top)), top)


def test_sampler_edges(orc, oracle_mod):
    """Rejection paths: ternary bytes >= 0xFE redraw one byte from a new PRNG call (sample.c:223-241);
    uniform words >= max_multiple redraw 4 bytes (sample.c:45-56).  Checked against a pure-Python
    restatement driven by hashlib."""
    def X(seed, c, k):
        return hashlib.shake_256(bytes(seed) + struct.pack("<Q", c)).digest(k)

    for tag in (b"e1", b"e2", b"e3"):
        seed = oracle_mod.make_seeds(1, tag)[0]
        n = 1024
        u, ctr = orc.sample_ternary_small(n, seed, 0)
        c, vals = 0, []
        for j in range((n + 95) // 96):
            buf = X(seed, c, 96)
            c += 1
            for i in range(min(96, n - 96 * j)):
                r = buf[i]
                while r >= 0xFE:
                    r = X(seed, c, 1)[0]
                    c += 1
                vals.append(r % 3)
        assert c == ctr
        t = np.array(vals, np.uint8).reshape(-1, 4)
        assert np.array_equal(u, (t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3])

        q = 1053818881
        a, ctr = orc.sample_uniform(n, q, seed, 5)
        M = 0xFFFFFFFF - (0xFFFFFFFF % q) - 1
        c = 5
        bulk = np.frombuffer(X(seed, c, 4 * n), "<u4")
        c += 1
        exp = []
        for r in bulk:
            r = int(r)
            while r >= M:
                r = struct.unpack("<I", X(seed, c, 4))[0]
                c += 1
            exp.append(r % q)
        assert c == ctr and np.array_equal(a, np.array(exp, np.uint32))
        assert ctr > 6  # ~1.86 % of 1024 words are rejected: the redraw path ran


def test_cbd_definition(orc, oracle_mod):
    """sample.c:263-284: 6 bytes per sample, x2 and x5 masked to 5 bits, range [-21, 21]."""
    seed = oracle_mod.make_seeds(1, b"cbd")[0]
    n = 1024
    e, ctr = orc.sample_cbd(n, seed, 3)
    assert ctr == 3 + n // 16
    exp = []
    for t in range(n // 16):
        buf = hashlib.shake_256(bytes(seed) + struct.pack("<Q", 3 + t)).digest(96)
        for i in range(16):
            x = buf[6 * i: 6 * i + 6]
            hw = lambda v: bin(v).count("1")  # noqa: E731
            exp.append(hw(x[0]) + hw(x[1]) + hw(x[2] & 0x1F) - hw(x[3]) - hw(x[4]) - hw(x[5] & 0x1F))
    assert np.array_equal(e, np.array(exp, np.int8))
    assert e.min() >= -21 and e.max() <= 21


@pytest.mark.parametrize("n,np_,asym", [(1024, 1, False), (4096, 3, True), (4096, 3, False)])
def test_encrypt_decrypt_roundtrip(n, np_, asym, orc, oracle_mod):
    """device/test/ckks_tests_common.c:173-231: decrypt + decode returns the message within 0.1."""
    sk = oracle_mod.make_sk(n)
    v = oracle_mod.make_values(1, n // 2, seed=3)[0]
    seed = oracle_mod.make_seeds(1, b"rt")[0]
    sseed = oracle_mod.make_seeds(1, b"rt-share")[0]
    if asym:
        pk0, pk1 = orc.gen_pk(n, np_, sk)
        ok, ct = orc.encrypt_asym(n, np_, v, seed, pk0, pk1)
    else:
        ok, ct = orc.encrypt_sym(n, np_, v, sseed, seed, sk)
    assert ok
    for p, q in enumerate(orc.primes(n, np_)):
        assert ct[p].max() < q
        dec = orc.decrypt_decode(n, np_, ct, sk, n // 2, prime_idx=p)
        assert np.abs(dec - v).max() < 0.1
    if not asym:
        # c1 = a is reproducible from the shareable seed alone; the quirk stream differs only in c1
        ok, ctq = orc.encrypt_sym(n, np_, v, sseed, seed, sk, ref_quirk=True)
        assert ok and np.array_equal(ctq[:, 0], ct[:, 0]) and not np.array_equal(ctq[:, 1], ct[:, 1])
        a0, _ = orc.sample_uniform(n, orc.primes(n, np_)[0], sseed, 0)
        assert np.array_equal(ct[0, 1], a0)


def test_encode_edges(orc):
    """ckks_common.c:105-215: empty input encodes to zero; a value whose coefficient passes 2^63
    makes the encode fail (the only `false` se_encrypt can return)."""
    n = 1024
    ok, pt = orc.encode(n, np.zeros(0, np.float32))
    assert ok and not pt.any()
    ok, _ = orc.encode(n, np.full(n // 2, 3.0e38, np.float32))
    assert not ok
    ok, pt = orc.encode(n, np.full(n // 2, 1.0, np.float32))
    assert ok and pt[0] == 1 << 20 and not pt[1:].any()  # constant message -> constant polynomial
