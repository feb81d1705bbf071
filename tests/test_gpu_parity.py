"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on identical seeded
inputs — bit-exact (integer/byte work; the FP64 encode is required to be bit-exact too, so the
tolerance is zero everywhere).  Reference behaviour being pinned is cited per test.
"""
from __future__ import annotations

import hashlib
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, keys_for

pytestmark = pytest.mark.gpu

CONFIGS = [(1024, 1), (2048, 1), (4096, 3), (8192, 4), (16384, 6)]


def dev(torch, a: np.ndarray):
    a = np.ascontiguousarray(a)
    view = {np.dtype(np.uint32): np.int32, np.dtype(np.uint16): np.int16, np.dtype(np.uint64): np.int64}.get(a.dtype)
    return torch.from_numpy(a.view(view) if view else a).cuda()


def host(t, dtype):
    return t.cpu().numpy().view(dtype)


@pytest.fixture(scope="module")
def ctxs(seb, torch_cuda):
    cache = {}

    def get(n, np_, asym):
        key = (n, np_, asym)
        if key not in cache:
            cache[key] = seb.Context(n, np_, asym, device=0)
        return cache[key]

    yield get
    for c in cache.values():
        c.close()


def test_prng_blocks(seb, torch_cuda, oracle_mod, ctxs):
    """rng.h:78-91 / fips202.c:105-128 — SHAKE256(seed || LE64(ctr)), against hashlib."""
    torch = torch_cuda
    ctx = ctxs(1024, 1, True)
    count = 300
    seeds = oracle_mod.make_seeds(count)
    ctrs = np.arange(count, dtype=np.uint64) * np.uint64(0x0101010101) + np.uint64(3)
    d_out = torch.zeros(count * 17, dtype=torch.int64, device="cuda")
    ctx.prng_blocks_device(dev(torch, seeds), dev(torch, ctrs), count, d_out)
    torch.cuda.synchronize()
    out = host(d_out, np.uint8).reshape(count, 136)
    for i in range(count):
        exp = hashlib.shake_256(seeds[i].tobytes() + struct.pack("<Q", int(ctrs[i]))).digest(136)
        assert out[i].tobytes() == exp, i


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_encode(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """ckks_common.c:105-215 + fft.c:69-144, incl. short inputs (zero padded) and the 9 message
    patterns of device/test/ckks_tests_common.c:25-57."""
    torch = torch_cuda
    ctx = ctxs(n, np_, True)
    vlen = n // 2
    rows = [oracle_mod.make_values(1, vlen, seed=n + k)[0] for k in range(4)]
    pat = np.zeros((7, vlen), np.float32)
    pat[0, 0] = 1
    pat[1, 0] = 2
    pat[2, :] = 1
    pat[3, :] = 2
    pat[4, :] = 1.1
    pat[5, :] = -2.1
    pat[6, 1::2] = 1
    big = (oracle_mod.make_values(1, vlen, seed=5)[0] * np.float32(1e6)).astype(np.float32)
    vals = np.stack(rows + list(pat) + [big])
    batch = vals.shape[0]
    d_pt = torch.zeros(batch * n, dtype=torch.int64, device="cuda")
    ctx.encode_device(dev(torch, vals), vlen, batch, d_pt)
    assert ctx.encode_failures() == 0
    got = d_pt.cpu().numpy().reshape(batch, n)
    for b in range(batch):
        ok, exp = orc.encode(n, vals[b])
        assert ok and np.array_equal(got[b], exp), (n, b, np.nonzero(got[b] != exp)[0][:8])
    # short rows: vlen < n/2 is zero padded
    for short in (1, 7, n // 4):
        sv = np.ascontiguousarray(vals[:3, :short])
        ctx.encode_device(dev(torch, sv), short, 3, d_pt)
        assert ctx.encode_failures() == 0
        got = d_pt.cpu().numpy().reshape(batch, n)
        for b in range(3):
            ok, exp = orc.encode(n, sv[b])
            assert ok and np.array_equal(got[b], exp), (n, short, b)


def test_encode_overflow_flag(seb, torch_cuda, orc, ctxs):
    """ckks_common.c:195-204: |coeff| > 2^63 makes the encode fail (returns false)."""
    torch = torch_cuda
    n = 4096
    ctx = ctxs(n, 3, True)
    vals = np.zeros((3, n // 2), np.float32)
    vals[0, :] = 1.0
    vals[1, :] = 3e37  # scale 2^25 * 3e37 >> 2^63
    vals[2, 0] = 0.5
    assert orc.encode(n, vals[1])[0] is False
    d_pt = torch.zeros(3 * n, dtype=torch.int64, device="cuda")
    ctx.encode_device(dev(torch, vals), n // 2, 3, d_pt)
    assert ctx.encode_failures() == 1


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_samplers_asym(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """sample.c:218-242 (ternary with interleaved single-byte redraws) and :311-356 (CBD), PRNG
    counter order of ckks_asym.c:184-201."""
    torch = torch_cuda
    ctx = ctxs(n, np_, True)
    batch = 48
    seeds = oracle_mod.make_seeds(batch, b"sampler-%d" % n)
    d_u = torch.zeros(batch * n // 4, dtype=torch.uint8, device="cuda")
    d_e = torch.zeros(batch * 2 * n, dtype=torch.int8, device="cuda")
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ctx.sample_asym_device(dev(torch, seeds), batch, d_u, d_e, d_ctr)
    torch.cuda.synchronize()
    u = d_u.cpu().numpy().reshape(batch, n // 4)
    e = d_e.cpu().numpy().reshape(batch, 2, n)
    ctr = host(d_ctr, np.uint32)
    for b in range(batch):
        eu, c = orc.sample_ternary_small(n, seeds[b])
        assert ctr[b] == c, (b, ctr[b], c)
        assert np.array_equal(u[b], eu), (n, b)
        e0, c = orc.sample_cbd(n, seeds[b], c)
        e1, c = orc.sample_cbd(n, seeds[b], c)
        assert np.array_equal(e[b, 0], e0) and np.array_equal(e[b, 1], e1), (n, b)


@pytest.mark.parametrize("batch", [1, 2, 7, 601])
def test_sampler_ternary_paired_warps(batch, seb, torch_cuda, oracle_mod, orc, ctxs):
    """n = 4096 runs the ternary sampler with TWO ciphertexts per warp sharing their third wave of PRNG counters
    (k_sample_ternary_pair): lone / odd batches leave a warp with one ciphertext, and over 601 items every branch is taken
    — ciphertexts done inside the shared wave (counter < 80), ciphertexts that need further waves of their own (the
    counter after u is checked to cover both).  Packed u and the PRNG counter equal sample.c:218-242 item by item."""
    torch = torch_cuda
    n, np_ = 4096, 3
    ctx = ctxs(n, np_, True)
    seeds = oracle_mod.make_seeds(batch, b"ternary-pair-%d" % batch)
    d_u = torch.zeros(batch * n // 4, dtype=torch.uint8, device="cuda")
    d_e = torch.zeros(batch * 2 * n, dtype=torch.int8, device="cuda")
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ctx.sample_asym_device(dev(torch, seeds), batch, d_u, d_e, d_ctr)
    torch.cuda.synchronize()
    u = d_u.cpu().numpy().reshape(batch, n // 4)
    ctr = host(d_ctr, np.uint32)
    for b in range(batch):
        eu, c = orc.sample_ternary_small(n, seeds[b])
        assert ctr[b] == c and np.array_equal(u[b], eu), (batch, b, ctr[b], c)
    if batch > 500:
        assert ctr.min() < 80 <= ctr.max()  # both kinds of ciphertext occurred


@pytest.mark.parametrize("wide", ["0", "1"])
@pytest.mark.parametrize("coop", ["0", "1"])
@pytest.mark.parametrize("n,np_", CONFIGS)
def test_sampler_uniform_both_kernels(n, np_, coop, wide, seb, torch_cuda, oracle_mod, orc, ctxs, monkeypatch):
    """sample_poly_uniform (sample.c:39-57) through each of the two bulk kernels — one sequential sponge per thread
    (large batches) and the warp-cooperative sponge spread over 25 lanes (small batches), forced with
    the "uniform_coop" option — and each of the two fix-up kernels — a warp per ciphertext with 32-candidate waves and a CTA
    per ciphertext with one round of candidates, forced with "uniform_fix_wide": same polynomials, same
    counters as the oracle, for a batch that is not a multiple of the warps per CTA."""
    torch = torch_cuda
    ctx = ctxs(n, np_, False)
    ctx.set_option("uniform_coop", int(coop))  # switches are context state (read from the environment only in seb_create)
    ctx.set_option("uniform_fix_wide", int(wide))
    batch = 9
    seeds = oracle_mod.make_seeds(batch, b"uniform-k-%d" % n)
    d_seeds = dev(torch, seeds)
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    d_out = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
    try:
        for p in range(np_):
            ctx.sample_uniform_device(d_seeds, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
        torch.cuda.synchronize()
    finally:
        ctx.set_option("uniform_coop", -1)
        ctx.set_option("uniform_fix_wide", -1)
    out = host(d_out, np.uint32).reshape(batch, np_, n)
    ctr = host(d_ctr, np.uint32)
    for b in range(batch):
        c = 0
        for p, q in enumerate(ctx.primes):
            exp, c = orc.sample_uniform(n, q, seeds[b], c)
            assert np.array_equal(out[b, p], exp), (n, coop, wide, b, p)
        assert ctr[b] == c


@pytest.mark.parametrize("batch", [1, 9, 16, 37])
@pytest.mark.parametrize("n,np_", CONFIGS)
def test_sampler_uniform_pair_kernel(n, np_, batch, seb, torch_cuda, oracle_mod, orc, ctxs):
    """The third bulk kernel (k_uniform_bulk_pair, mid-size batches): two lanes per sponge with the Keccak state
    bit-interleaved (even / odd bits), forced with the "uniform_pair" option.  Same polynomials, reject lists (through
    the fix-up) and counters as sample_poly_uniform (sample.c:39-57), for batches that fill a warp exactly (16), leave
    idle pairs in the last warp (1, 9, 37) and span several warps; the counter chain runs on across the primes."""
    torch = torch_cuda
    ctx = ctxs(n, np_, False)
    ctx.set_option("uniform_pair", 1)
    seeds = oracle_mod.make_seeds(batch, b"uniform-pair-%d" % n)
    d_seeds = dev(torch, seeds)
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    d_out = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
    try:
        for p in range(np_):
            ctx.sample_uniform_device(d_seeds, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
        torch.cuda.synchronize()
    finally:
        ctx.set_option("uniform_pair", -1)
    out = host(d_out, np.uint32).reshape(batch, np_, n)
    ctr = host(d_ctr, np.uint32)
    for b in range(batch):
        c = 0
        for p, q in enumerate(ctx.primes):
            exp, c = orc.sample_uniform(n, q, seeds[b], c)
            assert np.array_equal(out[b, p], exp), (n, batch, b, p)
        assert ctr[b] == c


@pytest.mark.parametrize("lanes", [4, 8, 32, "stream2", "stream4", "stream"])
@pytest.mark.parametrize("n,np_", [(1024, 1), (2048, 1), (4096, 3), (16384, 2)])
def test_sampler_uniform_fixup_lanes(n, np_, lanes, seb, torch_cuda, oracle_mod, orc, ctxs):
    """The fix-up with 4 / 8 / 32 lanes per ciphertext (k_uniform_fix_sub / k_uniform_fix, forced with the
    "uniform_fix_lanes" option) and as a stream over 8 ciphertexts per warp (k_uniform_fix_stream, "uniform_fix_stream"):
    same polynomials and counters as sample_poly_uniform (sample.c:39-57) whether a ciphertext's rejected words are
    served in one wave (n = 1024: ~1.5 of them), in twenty (n = 4096 with 4 lanes) or by lanes of waves it shares with
    its neighbours, for a batch that leaves the last warp partly empty."""
    torch = torch_cuda
    if n == 16384 and lanes in (4, 8):
        pytest.skip("hundreds of dependent small waves: covered at n = 4096")
    ctx = ctxs(n, np_, False)
    batch = 45
    seeds = oracle_mod.make_seeds(batch, b"uniform-lanes-%d" % n)
    d_seeds = dev(torch, seeds)
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    d_out = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
    stream = {"stream": 8, "stream2": 2, "stream4": 4}.get(lanes, 0)  # ciphertexts per warp of the streamed form
    ctx.set_option("uniform_fix_lanes", 32 if stream else lanes)
    ctx.set_option("uniform_fix_stream", stream)
    ctx.set_option("uniform_fix_wide", 0)
    try:
        for p in range(np_):
            ctx.sample_uniform_device(d_seeds, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
        torch.cuda.synchronize()
    finally:
        ctx.set_option("uniform_fix_lanes", -1)
        ctx.set_option("uniform_fix_stream", -1)
        ctx.set_option("uniform_fix_wide", -1)
    out = host(d_out, np.uint32).reshape(batch, np_, n)
    ctr = host(d_ctr, np.uint32)
    for b in range(batch):
        c = 0
        for p, q in enumerate(ctx.primes):
            exp, c = orc.sample_uniform(n, q, seeds[b], c)
            assert np.array_equal(out[b, p], exp), (n, lanes, b, p)
        assert ctr[b] == c


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_sampler_uniform(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """sample.c:39-57: bulk draw, ordered redraws, counter running on across primes (ckks_sym.c:219)."""
    torch = torch_cuda
    ctx = ctxs(n, np_, False)
    batch = 9
    seeds = oracle_mod.make_seeds(batch, b"uniform-%d" % n)
    d_seeds = dev(torch, seeds)
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    d_out = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
    for p in range(np_):
        ctx.sample_uniform_device(d_seeds, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
    torch.cuda.synchronize()
    out = host(d_out, np.uint32).reshape(batch, np_, n)
    ctr = host(d_ctr, np.uint32)
    for b in range(batch):
        c = 0
        for p, q in enumerate(ctx.primes):
            exp, c = orc.sample_uniform(n, q, seeds[b], c)
            assert np.array_equal(out[b, p], exp), (n, b, p)
        assert ctr[b] == c


@pytest.mark.parametrize("n", [1024, 4096, 16384])
def test_sampler_uniform_row_alignment(n, seb, torch_cuda, oracle_mod, orc, ctxs):
    """The bulk squeeze writes each polynomial with 64-bit stores into rows that are only 8-byte aligned in general (rows at
    an odd multiple of 8 bytes, every other row with a stride of 4k + 2 words), and nothing outside them; a batch above
    the warp-cooperative kernel's range so that the thread-per-sponge kernel runs.  n / 2 lanes are 30 full blocks + 2
    lanes at n = 1024, 120 + 8 at n = 4096, 481 + 15 at n = 16384."""
    torch = torch_cuda
    ctx = ctxs(n, 1, False)
    batch = 1500
    seeds = oracle_mod.make_seeds(batch, b"uniform-align-%d" % n)
    d_seeds = dev(torch, seeds)
    ctx.set_option("uniform_coop", 0)
    ctx.set_option("uniform_pair", 0)
    try:
        for offset, stride in ((0, n), (2, n + 2), (0, n + 2), (2, n + 4)):
            d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
            d_buf = torch.full((batch * stride + 8,), -1, dtype=torch.int32, device="cuda")
            ctx.sample_uniform_device(d_seeds, d_ctr, 0, batch, d_buf.data_ptr() + 4 * offset, stride)
            torch.cuda.synchronize()
            buf = host(d_buf, np.uint32)
            rows = buf[offset:offset + batch * stride].reshape(batch, stride)
            for b in list(range(0, batch, 97)) + [1, batch - 1]:
                exp, c = orc.sample_uniform(n, ctx.primes[0], seeds[b], 0)
                assert np.array_equal(rows[b, :n], exp), (n, offset, stride, b)
                assert int(host(d_ctr, np.uint32)[b]) == c
            # nothing outside the rows was touched
            assert np.all(rows[:, n:] == 0xFFFFFFFF) and np.all(buf[:offset] == 0xFFFFFFFF)
            assert np.all(buf[offset + batch * stride:] == 0xFFFFFFFF)
    finally:
        ctx.set_option("uniform_coop", -1)
        ctx.set_option("uniform_pair", -1)


@pytest.mark.parametrize("n,np_,batch", [(1024, 1, 20000), (4096, 2, 19800)])
def test_sampler_uniform_mixed_squeeze(n, np_, batch, seb, torch_cuda, oracle_mod, orc, ctxs):
    """Above one full layer of thread-kernel warps (one per SM sub-partition: 18944 sponges on 148 SMs) the sponges beyond
    the full layers are squeezed by the two-lane kernel on a second stream ("uniform_mix"): same polynomials, counters and
    reject handling as the one-kernel squeeze, and as the oracle on rows from both parts."""
    torch = torch_cuda
    ctx = ctxs(n, np_, False)
    seeds = oracle_mod.make_seeds(batch, b"uniform-mix-%d" % n)
    d_seeds = dev(torch, seeds)
    res = []
    try:
        for mix in (0, 1):
            ctx.set_option("uniform_mix", mix)
            d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
            d_out = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
            for p in range(np_):
                ctx.sample_uniform_device(d_seeds, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
            torch.cuda.synchronize()
            res.append((host(d_out, np.uint32).reshape(batch, np_, n), host(d_ctr, np.uint32)))
    finally:
        ctx.set_option("uniform_mix", -1)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    for b in (0, 18943, 18944, 18945, batch - 1):
        c = 0
        for p, q in enumerate(ctx.primes):
            exp, c = orc.sample_uniform(n, q, seeds[b], c)
            assert np.array_equal(res[1][0][b, p], exp), (n, b, p)
        assert res[1][1][b] == c


@pytest.mark.parametrize("wide", ["0", "1", "stream"])
@pytest.mark.parametrize("cap", ["0", "3", "70"])
def test_sampler_uniform_list_overflow(cap, wide, seb, torch_cuda, oracle_mod, orc, monkeypatch):
    """The uniform sampler's reject lists have a fixed capacity; ciphertexts that overflow it take the
    scanning fix-up.  Forced here with tiny capacities (none / most / some ciphertexts overflow at
    n = 4096, where ~76 of 4096 words are rejected per prime) and through the full symmetric path, for the
    warp-per-ciphertext, the CTA-per-ciphertext and the streamed fix-up (which serves overflowed ciphertexts on the
    spot, between the ones it streams)."""
    torch = torch_cuda
    monkeypatch.setenv("SEB_UNIFORM_LIST_CAP", cap)
    monkeypatch.setenv("SEB_UNIFORM_FIX_WIDE", "1" if wide == "1" else "0")
    n, np_, batch = 4096, 3, 7 if wide != "stream" else 21
    ctx = seb.Context(n, np_, False, device=0)
    if wide == "stream":
        ctx.set_option("uniform_fix_stream", 1)
    try:
        seeds = oracle_mod.make_seeds(batch, b"uniform-cap")
        d_seeds = dev(torch, seeds)
        d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
        d_out = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
        for p in range(np_):
            ctx.sample_uniform_device(d_seeds, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
        torch.cuda.synchronize()
        out = host(d_out, np.uint32).reshape(batch, np_, n)
        ctr = host(d_ctr, np.uint32)
        for b in range(batch):
            c = 0
            for p, q in enumerate(ctx.primes):
                exp, c = orc.sample_uniform(n, q, seeds[b], c)
                assert np.array_equal(out[b, p], exp), (cap, b, p)
            assert ctr[b] == c
        sk = oracle_mod.make_sk(n)
        ctx.set_secret_key(sk)
        vals = oracle_mod.make_values(batch, n // 2, seed=17)
        eseeds = oracle_mod.make_seeds(batch, b"uniform-cap-e")
        d_ct = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
        ctx.encrypt_sym_device(dev(torch, vals), n // 2, d_seeds, dev(torch, eseeds), batch, d_ct, False)
        assert ctx.encode_failures() == 0
        got = host(d_ct, np.uint32).reshape(batch, np_, 2, n)
        for b in range(batch):
            ok, exp = orc.encrypt_sym(n, np_, vals[b], seeds[b], eseeds[b], sk)
            assert ok and np.array_equal(got[b], exp), (cap, b)
    finally:
        ctx.close()


@pytest.mark.parametrize("n,np_", [(1024, 1), (2048, 1), (4096, 3), (8192, 6), (16384, 13)])
def test_ntt(n, np_, seb, torch_cuda, orc, ctxs):
    """ntt.c:168-189 for every tabulated (n, q): random residues plus the delta and all-(q-1) edge
    polynomials (device/test/ntt_tests.c:130-317 cases)."""
    torch = torch_cuda
    ctx = ctxs(n, np_, True)
    rng = np.random.default_rng(n)
    batch = 4
    polys = np.zeros((batch, np_, n), np.uint32)
    for p, q in enumerate(ctx.primes):
        polys[0, p] = rng.integers(0, q, n)
        polys[1, p, 0] = 1
        polys[2, p, :] = q - 1
        polys[3, p] = rng.integers(0, q, n)
    d = dev(torch, polys)
    ctx.ntt_device(d, batch)
    torch.cuda.synchronize()
    got = host(d, np.uint32).reshape(batch, np_, n)
    for p, q in enumerate(ctx.primes):
        for b in range(batch):
            assert np.array_equal(got[b, p], orc.ntt(n, q, polys[b, p])), (n, q, b)
        assert np.all(got[1, p] == 1)  # NTT(delta_0) = all ones


@pytest.mark.parametrize("n", [1024, 4096])
def test_ntt_custom_prime_chain(n, seb, torch_cuda, orc):
    """Config E asks for 1..8 primes at n = 1024 and 4096, beyond the reference's parameter sets
    (SURVEY 0.9): a caller-supplied chain of eight 30-bit primes with the library's minimal 2n-th roots
    (the custom-prime path the reference leaves broken, 0.8) against the oracle's ntt_inpl
    restatement run with the same roots."""
    torch = torch_cuda
    primes = orc.primes(16384, 8)
    ctx = seb.Context(n, 8, True, device=0, primes=primes)
    try:
        assert ctx.primes == primes
        batch = 3
        rng = np.random.default_rng(n)
        x = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint32) for q in primes]) for _ in range(batch)])
        d = dev(torch, x)
        ctx.ntt_device(d, batch)
        torch.cuda.synchronize()
        got = host(d, np.uint32).reshape(batch, 8, n)
        for b in range(batch):
            for p, q in enumerate(primes):
                psi = seb.api.minimal_psi(n, q)
                assert np.array_equal(got[b, p], orc.ntt(n, q, x[b, p], psi=psi)), (n, b, p)
        ctx.intt_device(d, batch)
        torch.cuda.synchronize()
        assert np.array_equal(host(d, np.uint32).reshape(batch, 8, n), x)
    finally:
        ctx.close()


@pytest.mark.parametrize("n,np_", [(1024, 1), (2048, 1), (4096, 3), (8192, 6), (16384, 13)])
def test_intt(n, np_, seb, torch_cuda, orc, ctxs):
    """Verifier INTT (inverse of ntt_inpl; reference: intt.c:226-501): equals the oracle's inverse and undoes
    the forward kernel on every tabulated (n, q)."""
    torch = torch_cuda
    ctx = ctxs(n, np_, True)
    rng = np.random.default_rng(n + 5)
    batch = 3
    x = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint32) for q in ctx.primes]) for _ in range(batch)])
    x[0, 0, :3] = (0, 1, ctx.primes[0] - 1)
    d = dev(torch, x)
    ctx.intt_device(d, batch)
    torch.cuda.synchronize()
    got = host(d, np.uint32).reshape(batch, np_, n)
    for b in range(batch):
        for p, q in enumerate(ctx.primes):
            assert np.array_equal(got[b, p], orc.intt(n, q, x[b, p])), (n, b, p)
    ctx.ntt_device(d, batch)  # ntt(intt(x)) = x
    torch.cuda.synchronize()
    assert np.array_equal(host(d, np.uint32).reshape(batch, np_, n), x)


@pytest.mark.parametrize("n,np_,asym", [(1024, 1, False), (4096, 3, True), (4096, 3, False), (8192, 4, True),
                                        (16384, 6, False)])
def test_decrypt_decode(n, np_, asym, seb, torch_cuda, oracle_mod, orc, ctxs):
    """Verifier decrypt + decode (device/test/ckks_tests_common.c:59-171) against the oracle's, every prime,
    ragged vlen; and the reference's acceptance criterion: decoded values within 0.1 of the message."""
    torch = torch_cuda
    ctx = ctxs(n, np_, asym)
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    ctx.set_secret_key(sk)
    batch, vlen = 4, n // 2
    vals = oracle_mod.make_values(batch, vlen, seed=31 * n)
    seeds = oracle_mod.make_seeds(batch, b"dd-%d" % n)
    sseeds = oracle_mod.make_seeds(batch, b"dd-share-%d" % n)
    d_ct = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
    if asym:
        ctx.set_public_key(pk0, pk1)
        ctx.encrypt_asym_device(dev(torch, vals), vlen, dev(torch, seeds), batch, d_ct)
    else:
        ctx.encrypt_sym_device(dev(torch, vals), vlen, dev(torch, sseeds), dev(torch, seeds), batch, d_ct, False)
    assert ctx.encode_failures() == 0
    ct = host(d_ct, np.uint32).reshape(batch, np_, 2, n)
    for p in range(np_):
        for vl in (vlen, 37):
            d_dec = torch.full((batch, vl), 1e9, dtype=torch.float32, device="cuda")
            ctx.decrypt_decode_device(d_ct, batch, p, vl, d_dec)
            torch.cuda.synchronize()
            dec = d_dec.cpu().numpy()
            for b in range(batch):
                exp = orc.decrypt_decode(n, np_, ct[b], sk, vl, prime_idx=p)
                assert np.allclose(dec[b], exp, rtol=0, atol=1e-4), (n, p, b, np.abs(dec[b] - exp).max())
            assert np.abs(dec - vals[:, :vl]).max() < 0.1


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_gen_public_key(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """gen_pk (ckks_asym.c:159-171) on the GPU equals the oracle's (which equals the reference's:
    tests/test_oracle.py::test_gen_pk_matches_reference), on asymmetric and symmetric contexts."""
    sk = oracle_mod.make_sk(n)
    exp0, exp1 = orc.gen_pk(n, np_, sk)
    for asym in (True, False):
        pk0, pk1 = ctxs(n, np_, asym).gen_public_key(sk)
        assert np.array_equal(pk0, exp0) and np.array_equal(pk1, exp1), (n, asym)
    e0, e1 = orc.gen_pk(n, np_, sk, ep_seed=bytes(range(64)), seed_base=bytes(range(100, 164)))
    g0, g1 = ctxs(n, np_, True).gen_public_key(sk, ep_seed=bytes(range(64)), a_seed_base=bytes(range(100, 164)))
    assert np.array_equal(g0, e0) and np.array_equal(g1, e1)


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_encrypt_asym(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """seal_embedded.c:98-215 asymmetric branch == ckks_asym.c:173-286, full ciphertext bit-exact."""
    torch = torch_cuda
    ctx = ctxs(n, np_, True)
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    ctx.set_public_key(pk0, pk1)
    batch = 5
    vlen = n // 2
    vals = oracle_mod.make_values(batch, vlen, seed=n)
    seeds = oracle_mod.make_seeds(batch, b"asym-%d" % n)
    d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
    ctx.encrypt_asym_device(dev(torch, vals), vlen, dev(torch, seeds), batch, d_out)
    assert ctx.encode_failures() == 0
    got = host(d_out, np.uint32).reshape(batch, np_, 2, n)
    for b in range(batch):
        ok, exp = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
        assert ok and np.array_equal(got[b], exp), (n, b)
    # decrypt + decode round trip (device/test/ckks_tests_asym.c: within 0.1)
    dec = orc.decrypt_decode(n, np_, got[0], sk, vlen)
    assert np.abs(dec - vals[0]).max() < 0.1


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_encrypt_sym(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """ckks_sym.c:181-301: c1 = a, c0 = -a*s + m + e; and the reference's se_encrypt byte stream
    where c1 holds ntt(m+e) when ref_quirk is on."""
    torch = torch_cuda
    ctx = ctxs(n, np_, False)
    sk = oracle_mod.make_sk(n)
    ctx.set_secret_key(sk)
    batch = 4
    vlen = n // 2
    vals = oracle_mod.make_values(batch, vlen, seed=n + 1)
    seeds = oracle_mod.make_seeds(batch, b"sym-%d" % n)
    sseeds = oracle_mod.make_seeds(batch, b"share-%d" % n)
    d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
    for quirk in (False, True):
        ctx.encrypt_sym_device(dev(torch, vals), vlen, dev(torch, sseeds), dev(torch, seeds), batch, d_out, quirk)
        assert ctx.encode_failures() == 0
        got = host(d_out, np.uint32).reshape(batch, np_, 2, n)
        for b in range(batch):
            ok, exp = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk, ref_quirk=quirk)
            assert ok and np.array_equal(got[b], exp), (n, b, quirk)
        if not quirk:
            dec = orc.decrypt_decode(n, np_, got[0], sk, vlen)
            assert np.abs(dec - vals[0]).max() < 0.1


@pytest.mark.parametrize("mode", ["serial", "speculative", "speculative-narrow"])
@pytest.mark.parametrize("n,np_", [(4096, 3), (8192, 4), (16384, 6)])
def test_encrypt_sym_lone_call_paths(n, np_, mode, seb, torch_cuda, oracle_mod, orc, monkeypatch):
    """Lone symmetric calls run every prime's uniform squeeze at once on speculated PRNG counters
    (seb_launch_uniform_chain_spec).  Ciphertexts must equal the oracle's whichever way the chain is walked:
    prime after prime (SEB_UNIFORM_SPEC=0), speculatively with the 5-sigma windows (no miss expected), and
    speculatively with windows narrowed to +-2 counters so that most true counters miss and are re-squeezed on
    the spot (the fallback)."""
    torch = torch_cuda
    monkeypatch.setenv("SEB_UNIFORM_SPEC", "0" if mode == "serial" else "1")
    if mode == "speculative-narrow":
        monkeypatch.setenv("SEB_UNIFORM_SPEC_SIGMAS", "0")
    ctx = seb.Context(n, np_, False, device=0)
    try:
        sk = oracle_mod.make_sk(n)
        ctx.set_secret_key(sk)
        vlen = n // 2
        for batch in (1, 3, 6):
            vals = oracle_mod.make_values(batch, vlen, seed=n + batch)
            seeds = oracle_mod.make_seeds(batch, b"lone-%d-%d" % (n, batch))
            sseeds = oracle_mod.make_seeds(batch, b"lone-share-%d-%d" % (n, batch))
            d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
            ctx.encrypt_sym_device(dev(torch, vals), vlen, dev(torch, sseeds), dev(torch, seeds), batch, d_out, False)
            assert ctx.encode_failures() == 0
            got = host(d_out, np.uint32).reshape(batch, np_, 2, n)
            for b in range(batch):
                ok, exp = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk)
                assert ok and np.array_equal(got[b], exp), (n, mode, batch, b)
        misses = ctx.uniform_spec_misses()
        if mode == "speculative":
            assert misses == 0
        elif mode == "speculative-narrow":
            assert misses > 0  # +-2 counters around the mean against a standard deviation of 9..40
        else:
            assert misses == 0
    finally:
        ctx.close()


@pytest.mark.parametrize("n,np_", CONFIGS)
def test_encrypt_sym_seed_compressed(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """SURVEY 8f-2 (the reference's unfinished SE_ENABLE_SYM_SEED_CT, seal_embedded.c:184-194): the
    seed-compressed call emits c0 only, bit-identical to the full ciphertext's c0; expanding the shareable
    seeds on the receiving side rebuilds c1 = a (sample.c:39-57) bit-identically; the expanded ciphertext
    decrypts.  Device and host flavours, odd batch so the host path's last chunk is ragged."""
    torch = torch_cuda
    ctx = ctxs(n, np_, False)
    sk = oracle_mod.make_sk(n)
    ctx.set_secret_key(sk)
    batch = 5
    vlen = n // 2
    vals = oracle_mod.make_values(batch, vlen, seed=n + 7)
    seeds = oracle_mod.make_seeds(batch, b"seedct-%d" % n)
    sseeds = oracle_mod.make_seeds(batch, b"seedct-share-%d" % n)
    exp = np.stack([orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk)[1] for b in range(batch)])
    d_c0 = torch.zeros(batch * np_ * n, dtype=torch.int32, device="cuda")
    d_ss = dev(torch, sseeds)
    ctx.encrypt_sym_seedct_device(dev(torch, vals), vlen, d_ss, dev(torch, seeds), batch, d_c0)
    assert ctx.encode_failures() == 0
    c0 = host(d_c0, np.uint32).reshape(batch, np_, n)
    assert np.array_equal(c0, exp[:, :, 0, :])
    d_full = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
    ctx.expand_seedct_device(d_ss, d_c0, batch, d_full)
    torch.cuda.synchronize()
    full = host(d_full, np.uint32).reshape(batch, np_, 2, n)
    assert np.array_equal(full, exp)
    dec = orc.decrypt_decode(n, np_, full[batch - 1], sk, vlen)
    assert np.abs(dec - vals[batch - 1]).max() < 0.1
    # c0 == NULL: only the c1 slots are written
    d_full.fill_(-1)
    ctx.expand_seedct_device(d_ss, None, batch, d_full)
    torch.cuda.synchronize()
    part = host(d_full, np.uint32).reshape(batch, np_, 2, n)
    assert np.array_equal(part[:, :, 1, :], exp[:, :, 1, :]) and (part[:, :, 0, :] == 0xFFFFFFFF).all()
    # host flavour (half the device-to-host bytes of encrypt_sym_host), then the full-size call again:
    # the staging buffers are shared between the two output sizes
    c0h = ctx.encrypt_sym_seedct_host(vals, sseeds, seeds)
    assert np.array_equal(c0h, exp[:, :, 0, :])
    assert np.array_equal(ctx.encrypt_sym_host(vals, sseeds, seeds), exp)
    # an asymmetric context refuses
    actx = ctxs(n, np_, True)
    with pytest.raises(seb.api.SebError):
        actx.encrypt_sym_seedct_device(dev(torch, vals), vlen, d_ss, dev(torch, seeds), batch, d_c0)
    with pytest.raises(seb.api.SebError):
        actx.expand_seedct_device(d_ss, d_c0, batch, d_full)


def test_encrypt_asym_27bit_primes(seb, torch_cuda, oracle_mod, orc):
    """The reference's SE_DEFAULT_4K_27BIT parameter set (parameters.c:204-209: n = 4096 with the three
    27-bit primes, roots ntt.c:213-225) as a run-time option: explicit primes, tabulated roots.  The
    expected ciphertext is composed from the oracle's stage functions (ckks_asym.c:205-286)."""
    torch = torch_cuda
    n, primes = 4096, [134012929, 134111233, 134176769]
    np_ = len(primes)
    ctx = seb.Context(n, np_, True, device=0, primes=primes, scale=2.0 ** 20)
    try:
        assert ctx.primes == primes
        rng = np.random.default_rng(27)
        pk0 = np.stack([rng.integers(0, q, n, dtype=np.uint32) for q in primes])
        pk1 = np.stack([rng.integers(0, q, n, dtype=np.uint32) for q in primes])
        ctx.set_public_key(pk0, pk1)
        batch, vlen = 3, n // 2
        vals = oracle_mod.make_values(batch, vlen, seed=27)
        vals[2] *= np.float32(1.0e6)  # coefficients far above these small primes: 64-bit reduction path
        seeds = oracle_mod.make_seeds(batch, b"27bit")
        d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
        ctx.encrypt_asym_device(dev(torch, vals), vlen, dev(torch, seeds), batch, d_out)
        assert ctx.encode_failures() == 0
        got = host(d_out, np.uint32).reshape(batch, np_, 2, n)
        for b in range(batch):
            ok, pt = orc.encode(n, vals[b], scale=2.0 ** 20)
            assert ok
            u, ctr = orc.sample_ternary_small(n, seeds[b], 0)
            e0, ctr = orc.sample_cbd(n, seeds[b], ctr)
            e1, ctr = orc.sample_cbd(n, seeds[b], ctr)
            pte = pt + e0.astype(np.int64)
            for p, q in enumerate(primes):
                uh = orc.ntt(n, q, orc.expand_ternary(n, q, u)).astype(np.uint64)
                e1h = orc.ntt(n, q, orc.reduce_small(n, q, e1)).astype(np.uint64)
                mh = orc.ntt(n, q, orc.reduce_pte(n, q, pte)).astype(np.uint64)
                c0 = (pk0[p].astype(np.uint64) * uh + mh) % np.uint64(q)
                c1 = (pk1[p].astype(np.uint64) * uh + e1h) % np.uint64(q)
                assert np.array_equal(got[b, p, 0], c0.astype(np.uint32)), (b, p)
                assert np.array_equal(got[b, p, 1], c1.astype(np.uint32)), (b, p)
    finally:
        ctx.close()


@pytest.mark.parametrize("n,np_,asym", [(4096, 3, True), (1024, 1, False), (8192, 4, True)])
def test_encrypt_large_magnitudes(n, np_, asym, seb, torch_cuda, oracle_mod, orc, ctxs):
    """reduce_set_pte (ckks_common.c:224-245) over the whole int64 range: messages from tiny to ~1e11 give
    plaintext coefficients from below 2^32 (32-bit reduction path) up to ~2^62 (64-bit path), mixed
    inside one warp, positive and negative."""
    torch = torch_cuda
    ctx = ctxs(n, np_, asym)
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    batch, vlen = 6, n // 2
    rng = np.random.default_rng(n)
    vals = oracle_mod.make_values(batch, vlen, seed=5 * n)
    vals[1] *= np.float32(1.0e4)
    vals[2] *= np.float32(1.0e9)
    vals[3] = (rng.standard_normal(vlen) * 10.0 ** rng.uniform(-3, 10, vlen)).astype(np.float32)
    vals[4, ::7] *= np.float32(3.0e7)
    vals[5] = 0
    vals[5, 3] = np.float32(2.0e10)
    seeds = oracle_mod.make_seeds(batch, b"big-%d" % n)
    sseeds = oracle_mod.make_seeds(batch, b"big-share-%d" % n)
    d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
    if asym:
        ctx.set_public_key(pk0, pk1)
        ctx.encrypt_asym_device(dev(torch, vals), vlen, dev(torch, seeds), batch, d_out)
    else:
        ctx.set_secret_key(sk)
        ctx.encrypt_sym_device(dev(torch, vals), vlen, dev(torch, sseeds), dev(torch, seeds), batch, d_out, False)
    assert ctx.encode_failures() == 0
    got = host(d_out, np.uint32).reshape(batch, np_, 2, n)
    big = 0
    for b in range(batch):
        ok_pt, pt = orc.encode(n, vals[b])
        big += int((np.abs(pt) >= (1 << 32)).sum())
        if asym:
            ok, exp = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
        else:
            ok, exp = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk)
        assert ok and ok_pt and np.array_equal(got[b], exp), (n, b)
    assert big > 1000  # the 64-bit path really ran


@pytest.mark.parametrize("n,np_", [(4096, 3), (2048, 1)])
def test_encrypt_reduction_path_boundary(n, np_, seb, torch_cuda, oracle_mod, orc, ctxs):
    """The encrypt kernels pick a 32-bit reduction of m + e0 when max |m| < 2q - 21 and the 64-bit one
    otherwise: constant messages put |m[0]| right at that switch (both signs, both sides) for the
    30-bit and the 27-bit primes."""
    torch = torch_cuda
    ctx = ctxs(n, np_, True)
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    ctx.set_public_key(pk0, pk1)
    q0 = ctx.primes[0]
    centre = np.float32((2 * q0 - 21) / ctx.scale)
    cs = [centre]
    for _ in range(3):
        cs = [np.nextafter(cs[0], np.float32(0)), *cs, np.nextafter(cs[-1], np.float32(1e9))]
    cs = cs + [-c for c in cs] + [np.float32(q0 / ctx.scale), np.float32(-4.0 * q0 / ctx.scale)]
    batch, vlen = len(cs), n // 2
    vals = np.stack([np.full(vlen, c, np.float32) for c in cs])
    seeds = oracle_mod.make_seeds(batch, b"edge-%d" % n)
    d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
    ctx.encrypt_asym_device(dev(torch, vals), vlen, dev(torch, seeds), batch, d_out)
    assert ctx.encode_failures() == 0
    got = host(d_out, np.uint32).reshape(batch, np_, 2, n)
    below = above = 0
    for b in range(batch):
        ok_pt, pt = orc.encode(n, vals[b])
        mx = int(np.abs(pt).max())
        below += mx < 2 * q0 - 21
        above += mx >= 2 * q0 - 21
        ok, exp = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
        assert ok and ok_pt and np.array_equal(got[b], exp), (n, b, mx)
    assert below >= 3 and above >= 3


def test_contexts_of_several_degrees_coexist(seb, torch_cuda, ctxs):
    """Kernel attributes are per kernel, not per context: creating a small-degree context after a large one
    must not shrink what the large one may launch (regression: the verifier's dynamic shared memory limit)."""
    torch = torch_cuda
    big = ctxs(16384, 6, False)
    small = seb.Context(1024, 1, True, device=0)
    try:
        rng = np.random.default_rng(5)
        x = np.stack([rng.integers(0, q, 16384, dtype=np.uint32) for q in big.primes])[None]
        d = dev(torch, x)
        big.ntt_device(d, 1)
        big.intt_device(d, 1)
        torch.cuda.synchronize()
        assert np.array_equal(host(d, np.uint32).reshape(1, 6, 16384), x)
        y = rng.integers(0, small.primes[0], (2, 1, 1024), dtype=np.uint32)
        d2 = dev(torch, y)
        small.ntt_device(d2, 2)
        small.intt_device(d2, 2)
        torch.cuda.synchronize()
        assert np.array_equal(host(d2, np.uint32).reshape(2, 1, 1024), y)
    finally:
        small.close()


def test_context_lifecycle_does_not_leak(seb, torch_cuda, oracle_mod, orc):
    """seb_create / work / seb_destroy twenty times over (both encryption types, device and host-pointer
    calls, the seed-compressed scratch, the verifier): free device memory returns to where it started."""
    torch = torch_cuda
    n, np_ = 2048, 1
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    vals = oracle_mod.make_values(64, n // 2, seed=5)
    seeds = oracle_mod.make_seeds(64, b"leak")
    sseeds = oracle_mod.make_seeds(64, b"leak-share")

    def cycle():
        for asym in (True, False):
            ctx = seb.Context(n, np_, asym, device=0)
            ctx.set_secret_key(sk)
            if asym:
                ctx.set_public_key(pk0, pk1)
                ctx.encrypt_asym_host(vals, seeds)
            else:
                ctx.encrypt_sym_host(vals, sseeds, seeds)
                ctx.encrypt_sym_seedct_host(vals, sseeds, seeds)
            ctx.close()

    cycle()  # first use pays for lazily created CUDA state (module load, constant banks)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(20):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20), f"{(free0 - free1) >> 20} MiB of device memory not returned"


def test_edge_cases_and_errors(seb, torch_cuda, oracle_mod, orc):
    """Empty batch, empty message (vlen = 0 encrypts the zero message), and the error behaviour of the
    seb_* layer: missing key material, vlen beyond n/2, key coefficients outside [0, q)."""
    torch = torch_cuda
    n, np_ = 1024, 1
    ctx = seb.Context(n, np_, True, device=0)
    try:
        d_vals = torch.zeros((4, n // 2), dtype=torch.float32, device="cuda")
        d_seeds = dev(torch, oracle_mod.make_seeds(4, b"edge"))
        d_out = torch.full((4, np_, 2, n), -1, dtype=torch.int32, device="cuda")
        with pytest.raises(seb.SebError, match="no public key"):
            ctx.encrypt_asym_device(d_vals, n // 2, d_seeds, 4, d_out)
        sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
        bad = pk0.copy()
        bad[0, 5] = ctx.primes[0]
        with pytest.raises(seb.SebError, match="modulus"):
            ctx.set_public_key(bad, pk1)
        ctx.set_public_key(pk0, pk1)
        with pytest.raises(seb.SebError, match="exceeds"):
            ctx.encrypt_asym_device(d_vals, n // 2 + 1, d_seeds, 4, d_out)
        with pytest.raises(seb.SebError, match="no secret key"):
            ctx.decrypt_decode_device(d_out, 4, 0, n // 2, d_vals)
        ctx.encrypt_asym_device(d_vals, n // 2, d_seeds, 0, d_out)  # empty batch: nothing is written
        torch.cuda.synchronize()
        assert ctx.encode_failures() == 0 and int((d_out != -1).sum()) == 0
        ctx.encrypt_asym_device(d_vals, 0, d_seeds, 4, d_out)  # vlen = 0: the zero message
        assert ctx.encode_failures() == 0
        got = host(d_out, np.uint32).reshape(4, np_, 2, n)
        seeds = host(d_seeds, np.uint8).reshape(4, 64)
        for b in range(4):
            ok, exp = orc.encrypt_asym(n, np_, np.zeros(0, np.float32), seeds[b], pk0, pk1)
            assert ok and np.array_equal(got[b], exp)
        assert ctx.encrypt_asym_host(np.zeros((0, 7), np.float32), np.zeros((0, 64), np.uint8)).shape == (0, np_, 2, n)
    finally:
        ctx.close()
    with pytest.raises(seb.SebError):
        seb.Context(4096, 4, True, device=0)  # parameters.c:204-213: n = 4096 takes at most 3 primes
    with pytest.raises(seb.SebError):
        seb.Context(4096, 1, True, device=0, primes=[1053818881], psis=[12345])  # not a primitive 2n-th root


def test_host_api_chunked(seb, torch_cuda, oracle_mod, orc, ctxs, monkeypatch):
    """Host-pointer batch API: pageable and pinned buffers, several chunks in flight (the "host_chunk" option forces
    3 balanced chunks of 500 items; the default policy would take 1500 items in one), ragged vlen, and
    the symmetric flavours (full and seed-compressed) through the same pipeline."""
    torch = torch_cuda
    n, np_ = 4096, 3
    ctx = ctxs(n, np_, True)
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    ctx.set_public_key(pk0, pk1)
    batch = 1500
    vlen = 100
    vals = oracle_mod.make_values(batch, vlen, seed=99)
    seeds = oracle_mod.make_seeds(batch, b"host")
    one = ctx.encrypt_asym_host(vals, seeds)  # default policy: a single chunk
    ctx.set_option("host_chunk", 600)
    out = ctx.encrypt_asym_host(vals, seeds)  # pageable buffers
    assert np.array_equal(out, one)
    pv = torch.from_numpy(vals).pin_memory()
    ps = torch.from_numpy(seeds).pin_memory()
    po = torch.empty((batch, np_, 2, n), dtype=torch.int32).pin_memory()
    ctx.lib.seb_encrypt_asym_host(ctx.h, pv.data_ptr(), vlen, ps.data_ptr(), batch, po.data_ptr())
    assert np.array_equal(out, po.numpy().view(np.uint32))
    ctx.set_option("host_chunk", 0)
    for b in (0, 1, 499, 500, 501, 999, 1000, 1499):
        ok, exp = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
        assert ok and np.array_equal(out[b], exp), b
    # symmetric: chunks of 40 out of 100 items -> 3 chunks of 34/34/32
    sctx = ctxs(n, np_, False)
    sctx.set_secret_key(sk)
    sctx.set_option("host_chunk", 40)
    sb = 100
    sseeds = oracle_mod.make_seeds(sb, b"host-share")
    full = sctx.encrypt_sym_host(vals[:sb], sseeds, seeds[:sb])
    c0 = sctx.encrypt_sym_seedct_host(vals[:sb], sseeds, seeds[:sb])
    sctx.set_option("host_chunk", 0)
    assert np.array_equal(c0, full[:, :, 0, :])
    for b in (0, 33, 34, 67, 68, 99):
        ok, exp = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk)
        assert ok and np.array_equal(full[b], exp), b


def _send_collector(chunks):
    def send(data: bytes) -> int:
        chunks.append(data)
        return len(data)

    return send


@pytest.mark.parametrize("asym", [False, True])
def test_se_api_dropin(asym, seb, torch_cuda, oracle_mod, orc, tmp_path, monkeypatch, capfd):
    """The reference's own API surface: se_setup reads adapter_output_data/, se_encrypt_seeded calls
    send(c0), send(c1) per prime with n*4 bytes each (seal_embedded.c:180-204); compared with the
    oracle and, when its .so is present, with the compiled reference run on the same files."""
    n, np_ = 4096, 3
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    primes = orc.primes(n, np_)
    oracle_mod.write_key_files(str(tmp_path), n, primes, sk, pk0, pk1)
    monkeypatch.chdir(tmp_path)
    se = seb.SealEmbedded()
    se.se_setup(n, np_, 12345.0, seb.api.SE_ASYM_ENCR if asym else seb.api.SE_SYM_ENCR)
    try:
        p = se.parms
        assert p.coeff_count == n and p.nprimes == np_ and p.logn == 12
        assert p.scale == 2.0 ** 25  # user scale overridden like set_parms_ckks (parameters.c:214)
        assert [p.moduli[i].value for i in range(np_)] == primes
        assert [(p.moduli[i].const_ratio[0], p.moduli[i].const_ratio[1]) for i in range(np_)] == \
            [orc.const_ratio(q) for q in primes]
        vals = oracle_mod.make_values(1, n // 2, seed=3)[0]
        seed = oracle_mod.make_seeds(1, b"api")[0]
        sseed = oracle_mod.make_seeds(1, b"api-share")[0]
        se.set_reference_quirk(True)
        chunks = []
        assert se.se_encrypt_seeded(sseed, seed, _send_collector(chunks), vals)
        assert len(chunks) == 2 * np_ and all(len(c) == 4 * n for c in chunks)
        got = np.frombuffer(b"".join(chunks), np.uint32).reshape(np_, 2, n)
        if asym:
            ok, exp = orc.encrypt_asym(n, np_, vals, seed, pk0, pk1)
        else:
            ok, exp = orc.encrypt_sym(n, np_, vals, sseed, seed, sk, ref_quirk=True)
        assert ok and np.array_equal(got, exp)
        if oracle_mod.have_reference():
            ref = oracle_mod.ReferenceLib()
            ref.setup(n, np_, asym, sk=sk, pk0=pk0, pk1=pk1, primes=primes)
            try:
                okr, ctr = ref.encrypt_seeded(sseed, seed, vals)
            finally:
                ref.close()
                os.chdir(tmp_path)
            assert okr and np.array_equal(got, ctr)
        # print = true: the reference's print_poly text (seal_embedded.c:161-165, util_print.h:478-489).
        # Default build flavour: 8 values then "... }"; full flavour: what the adapter's
        # ct_string_file_load / poly_string_file_load parse (adapter/fileops.h:221-301).
        capfd.readouterr()
        assert se.se_encrypt_seeded(sseed, seed, None, vals, print_=True)
        import ctypes
        ctypes.CDLL(None).fflush(None)
        lines = [ln for ln in capfd.readouterr().out.splitlines() if ln.startswith("c")]
        assert len(lines) == 2 * np_
        for p_ in range(np_):
            for k, nm in enumerate(("c0: ", "c1: ")):
                first8 = ", ".join(str(int(x)) for x in exp[p_, k, :8])
                assert lines[2 * p_ + k] == f"{nm} : {{ {first8}, ... }}", lines[2 * p_ + k]
        se.set_print_full(True)
        assert se.se_encrypt_seeded(sseed, seed, None, vals, print_=True)
        ctypes.CDLL(None).fflush(None)
        text = capfd.readouterr().out
        se.set_print_full(False)
        parsed, pos = [], 0
        while True:  # poly_string_file_load: find '{', read whitespace-separated tokens up to '}', strip commas
            i = text.find("{", pos)
            if i < 0:
                break
            j = text.find("}", i)
            parsed.append([int(tok.replace(",", "")) for tok in text[i + 1:j].split()])
            pos = j + 1
        assert len(parsed) == 2 * np_ and all(len(v) == n for v in parsed)
        assert np.array_equal(np.array(parsed, np.uint32).reshape(np_, 2, n), exp)
        # short input keeps earlier slots only (seal_embedded.c:108-111); NULL seeds draw randomness
        chunks2 = []
        assert se.se_encrypt(_send_collector(chunks2), vals[:16])
        assert len(chunks2) == 2 * np_
        assert b"".join(chunks2) != b"".join(chunks)
        # batch extension
        se.set_reference_quirk(False)
        bvals = oracle_mod.make_values(3, n // 2, seed=4)
        bseeds = oracle_mod.make_seeds(3, b"batch")
        bshare = oracle_mod.make_seeds(3, b"batch-share")
        ok, out = se.se_encrypt_batch_seeded(None if asym else bshare, bseeds, bvals)
        assert ok
        for b in range(3):
            if asym:
                _, exp = orc.encrypt_asym(n, np_, bvals[b], bseeds[b], pk0, pk1)
            else:
                _, exp = orc.encrypt_sym(n, np_, bvals[b], bshare[b], bseeds[b], sk)
            assert np.array_equal(out[b], exp)
        if not asym:
            # seed-compressed protocol: per prime send(seed, 64) then send(c0, 4n); batch form = c0 only
            se.set_sym_seed_ct(True)
            try:
                chunks3 = []
                assert se.se_encrypt_seeded(sseed, seed, _send_collector(chunks3), vals)
                assert [len(c) for c in chunks3] == [64, 4 * n] * np_
                _, full = orc.encrypt_sym(n, np_, vals, sseed, seed, sk)
                for p_ in range(np_):
                    assert chunks3[2 * p_] == sseed.tobytes()
                    assert np.array_equal(np.frombuffer(chunks3[2 * p_ + 1], np.uint32), full[p_, 0])
                # NULL shareable seed: the library draws one and sends it; the receiver can rebuild a from it
                chunks4 = []
                assert se.se_encrypt(_send_collector(chunks4), vals)
                drawn = np.frombuffer(chunks4[0], np.uint8)
                assert all(chunks4[2 * p_] == chunks4[0] for p_ in range(np_)) and chunks4[0] != sseed.tobytes()
                a, ctr = [], 0  # the receiver's expansion: one PRNG, counter running on across primes
                for q in primes:
                    ap, ctr = orc.sample_uniform(n, q, drawn, ctr)
                    a.append(ap)
                ct = np.stack([np.stack([np.frombuffer(chunks4[2 * p_ + 1], np.uint32), a[p_]]) for p_ in range(np_)])
                dec = orc.decrypt_decode(n, np_, ct, sk, n // 2)
                assert np.abs(dec - vals).max() < 0.1
                ok, c0s = se.se_encrypt_batch_seedct(bshare, bseeds, bvals)
                assert ok
                for b in range(3):
                    _, exp = orc.encrypt_sym(n, np_, bvals[b], bshare[b], bseeds[b], sk)
                    assert np.array_equal(c0s[b], exp[:, 0, :])
            finally:
                se.set_sym_seed_ct(False)
    finally:
        se.se_cleanup()


@pytest.mark.parametrize("mode,n,np_", [("asym", 4096, 3), ("sym", 1024, 1), ("sym", 8192, 4)])
def test_c_application_dropin(mode, n, np_, seb, torch_cuda, oracle_mod, orc, tmp_path):
    """The drop-in claim end to end, from C: examples/se_encrypt_demo.c (plain C11 against the reference's
    API names, key files in ./adapter_output_data like fileops.c:140-204) is compiled with gcc, linked against
    the library and run as its own process; the bytes its send callback received and the batch extension's
    output equal the oracle's ciphertexts."""
    import subprocess
    from test_abi import _build_demo

    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    primes = orc.primes(n, np_)
    oracle_mod.write_key_files(str(tmp_path), n, primes, sk, pk0, pk1)
    exe = _build_demo(tmp_path)
    out_file = tmp_path / "ct.bin"
    r = subprocess.run([exe, mode, str(n), str(np_), str(out_file)], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    got = np.fromfile(out_file, dtype=np.uint32)
    assert got.size == 4 * np_ * 2 * n
    got = got.reshape(4, np_, 2, n)
    vlen = n // 2
    i = np.arange(vlen)
    msgs = [(((i + b) % 17) - 8).astype(np.float32) + (i.astype(np.float32) / np.float32(1024.0)) for b in range(3)]
    seed = np.full(64, 0x5A, np.uint8)
    share = np.full(64, 0xA5, np.uint8)

    def expect(v, sd, ss):
        if mode == "asym":
            ok, ct = orc.encrypt_asym(n, np_, v, sd, pk0, pk1)
        else:
            ok, ct = orc.encrypt_sym(n, np_, v, ss, sd, sk)
        assert ok
        return ct

    assert np.array_equal(got[0], expect(msgs[0], seed, share))
    for b in range(3):
        sd, ss = seed.copy(), share.copy()
        sd[0], ss[0] = b, b + 100
        assert np.array_equal(got[1 + b], expect(msgs[b], sd, ss)), b


def test_golden_fixtures(seb, torch_cuda, oracle_mod, ctxs):
    """Committed vectors generated from the compiled reference (tests/golden/make_golden.py)."""
    torch = torch_cuda
    g = np.load(os.path.join(GOLDEN, "encrypt_golden.npz"))
    for key in [k[:-len("_digest")] for k in g.files if k.endswith("_digest")]:
        n, np_, asym = (int(x) for x in g[key + "_cfg"])
        vals, seeds, sseeds = g[key + "_values"], g[key + "_seeds"], g[key + "_sseeds"]
        sk = oracle_mod.make_sk(n)
        ctx = ctxs(n, np_, bool(asym))
        batch = vals.shape[0]
        d_out = torch.zeros(batch * np_ * 2 * n, dtype=torch.int32, device="cuda")
        if asym:
            pk0, pk1 = oracle_mod.Oracle().gen_pk(n, np_, sk)
            assert hashlib.sha256(pk0.tobytes() + pk1.tobytes()).digest() == g[key + "_pkdigest"].tobytes()
            ctx.set_public_key(pk0, pk1)
            ctx.encrypt_asym_device(dev(torch, vals), vals.shape[1], dev(torch, seeds), batch, d_out)
        else:
            ctx.set_secret_key(sk)
            ctx.encrypt_sym_device(dev(torch, vals), vals.shape[1], dev(torch, sseeds), dev(torch, seeds), batch,
                                   d_out, True)
        assert ctx.encode_failures() == 0
        got = host(d_out, np.uint32).reshape(batch, -1)
        for b in range(batch):
            assert hashlib.sha256(got[b].tobytes()).digest() == g[key + "_digest"][b].tobytes(), (key, b)
        if key + "_ct0" in g.files:
            assert np.array_equal(got[0], g[key + "_ct0"].reshape(-1))


@pytest.mark.parametrize("n,np_,asym,batch", [(4096, 3, True, 65536), (8192, 4, True, 32768), (16384, 6, False, 16384),
                                              (1024, 1, False, 65536)],
                         ids=["B-4096x3-asym-65536", "C-8192x4-asym-32768", "D-16384x6-sym-16384", "A-1024x1-sym-65536"])
def test_full_size_properties(n, np_, asym, batch, seb, torch_cuda, oracle_mod, orc, ctxs):
    """BASELINE.json's configurations at full per-GPU size (B whole; C and D as their 1/8 shards; A batched):
    size-independent properties — (1) items are independent of batch position and grid shape (re-encrypting
    slices reproduces the same per-item checksums), (2) the first 32, the middle and the last 32 items equal
    the oracle bit for bit, (3) every residue is < q, (4) EVERY item decrypts and decodes to its message within the
    reference's 0.1 on the GPU verifier."""
    torch = torch_cuda
    ctx = ctxs(n, np_, asym)
    sk, pk0, pk1 = keys_for(oracle_mod, orc, n, np_)
    if asym:
        ctx.set_public_key(pk0, pk1)
    ctx.set_secret_key(sk)
    vlen = n // 2
    gen = torch.Generator(device="cuda").manual_seed(1)
    d_vals = torch.rand((batch, vlen), generator=gen, device="cuda", dtype=torch.float32) * 32 - 16
    d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")

    def encrypt(lo, hi, out):
        v, sd, ss = d_vals[lo:hi].contiguous(), d_seeds[lo:hi].contiguous(), d_ss[lo:hi].contiguous()
        if asym:
            ctx.encrypt_asym_device(v, vlen, sd, hi - lo, out)
        else:
            ctx.encrypt_sym_device(v, vlen, ss, sd, hi - lo, out, False)
        assert ctx.encode_failures() == 0

    encrypt(0, batch, d_out)
    sums = d_out.view(batch, -1).to(torch.int64).sum(dim=1)
    xors = d_out.view(batch, -1)[:, ::97].to(torch.int64).sum(dim=1)
    for p, q in enumerate(ctx.primes):
        assert int(d_out[:, p].max()) < q and int(d_out[:, p].min()) >= 0
    # (1) slices with other batch sizes / offsets
    for lo, hi in ((0, 1), (5, 777), (batch - 4099, batch)):
        d_o2 = torch.empty((hi - lo, np_, 2, n), dtype=torch.int32, device="cuda")
        encrypt(lo, hi, d_o2)
        assert torch.equal(d_o2.view(hi - lo, -1).to(torch.int64).sum(dim=1), sums[lo:hi])
        assert torch.equal(d_o2.view(hi - lo, -1)[:, ::97].to(torch.int64).sum(dim=1), xors[lo:hi])
        del d_o2
    # (4) every item of the batch round-trips on the GPU verifier (device/test/ckks_tests_common.c:228),
    # under each prime
    d_dec = torch.empty((batch, vlen), dtype=torch.float32, device="cuda")
    for p in range(np_):
        ctx.decrypt_decode_device(d_out, batch, p, vlen, d_dec)
        assert float((d_dec - d_vals).abs().max()) < 0.1, p
    # (2) the oracle on the ends and the middle
    vals = d_vals.cpu().numpy()
    seeds = d_seeds.cpu().numpy()
    sseeds = d_ss.cpu().numpy()
    # the first and last items of the batch byte for byte (32 + 32 here), plus the middle
    picks = list(range(32)) + [batch // 2] + list(range(batch - 32, batch))
    for b in picks:
        got = host(d_out[b], np.uint32).reshape(np_, 2, n)
        if asym:
            ok, exp = orc.encrypt_asym(n, np_, vals[b], seeds[b], pk0, pk1)
        else:
            ok, exp = orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk)
        assert ok and np.array_equal(got, exp), b
        if b in (0, batch // 2, batch - 1):
            dec = orc.decrypt_decode(n, np_, got, sk, vlen)
            assert np.abs(dec - vals[b]).max() < 0.1
    del d_out, d_dec
    torch.cuda.empty_cache()
