"""ctypes mirror of include/seal_embedded_b200.h.

Two classes:
  * ``Context``      — the batch / device-pointer extension (``seb_*``), one per GPU per process.
  * ``SealEmbedded`` — the reference's own API (``se_setup`` / ``se_encrypt_seeded`` / ``se_cleanup``,
    reference: device/lib/seal_embedded.h:91-130), including the send-callback protocol.

Device buffers are passed as integer addresses or as any object with ``data_ptr()`` (torch tensors);
PyTorch is only the allocator/stream provider here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = _build.LIB

SE_SYM_ENCR, SE_ASYM_ENCR = 0, 1
SEED_BYTES = 64


class SebError(RuntimeError):
    pass


def build_library(force: bool = False) -> str:
    return _build.build(force=force)


_lib = None


class _Modulus(C.Structure):
    _fields_ = [("value", C.c_uint32), ("const_ratio", C.c_uint32 * 2)]


class _Parms(C.Structure):
    _fields_ = [
        ("coeff_count", C.c_size_t),
        ("logn", C.c_size_t),
        ("moduli", C.POINTER(_Modulus)),
        ("curr_modulus", C.POINTER(_Modulus)),
        ("curr_modulus_idx", C.c_size_t),
        ("nprimes", C.c_size_t),
        ("scale", C.c_double),
        ("is_asymmetric", C.c_bool),
        ("pk_from_file", C.c_bool),
        ("sample_s", C.c_bool),
        ("small_s", C.c_bool),
        ("small_u", C.c_bool),
    ]


class _SePtrs(C.Structure):
    _fields_ = [
        ("conj_vals", C.c_void_p),
        ("ifft_roots", C.c_void_p),
        ("values", C.POINTER(C.c_float)),
        ("ternary", C.POINTER(C.c_uint32)),
        ("conj_vals_int_ptr", C.POINTER(C.c_int64)),
        ("c0_ptr", C.POINTER(C.c_uint32)),
        ("c1_ptr", C.POINTER(C.c_uint32)),
        ("index_map_ptr", C.POINTER(C.c_uint16)),
        ("ntt_roots_ptr", C.POINTER(C.c_uint32)),
        ("ntt_pte_ptr", C.POINTER(C.c_uint32)),
        ("e1_ptr", C.POINTER(C.c_int8)),
    ]


class _SeParms(C.Structure):
    _fields_ = [("parms", C.POINTER(_Parms)), ("se_ptrs", C.POINTER(_SePtrs))]


SEND_FNCT = C.CFUNCTYPE(C.c_size_t, C.c_void_p, C.c_size_t)

# every symbol include/seal_embedded_b200.h declares
EXPORTED_SYMBOLS = [
    "se_setup_custom", "se_setup", "se_setup_default", "se_encrypt_seeded", "se_encrypt", "se_cleanup",
    "se_encrypt_batch_seeded", "se_b200_set_reference_quirk", "se_b200_set_print_full", "se_b200_context",
    "seb_last_error", "seb_create", "seb_destroy", "seb_set_stream", "seb_set_public_key", "seb_set_secret_key",
    "seb_reserve", "seb_degree", "seb_nprimes", "seb_scale", "seb_prime", "seb_launch_count",
    "seb_encrypt_asym_device", "seb_encrypt_sym_device", "seb_encode_failures", "seb_encrypt_asym_host",
    "seb_encrypt_sym_host", "seb_encode_device", "seb_sample_asym_device", "seb_sample_cbd_device",
    "seb_sample_uniform_device", "seb_ntt_device", "seb_prng_blocks_device", "seb_profile_begin", "seb_profile_end",
    "seb_intt_device", "seb_decrypt_decode_device", "seb_gen_public_key",
    "seb_encrypt_sym_seedct_device", "seb_encrypt_sym_seedct_host", "seb_expand_seedct_device",
    "se_b200_set_sym_seed_ct", "se_encrypt_batch_seedct", "seb_minimal_psi", "seb_uniform_spec_misses",
    "seb_packed30_words", "seb_encrypt_asym_host_packed30", "seb_encrypt_sym_host_packed30", "seb_unpack30",
    "seb_unpack30_device", "seb_measure_ceilings", "seb_set_option", "seb_gen_secret_key", "seb_digest_device", "seb_ct_to_seal_layout", "seb_ct_from_seal_layout",
]


def load_library(path: str | None = None) -> C.CDLL:
    """Load the CUDA library.  There is no fallback: a missing or unloadable .so is an error."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # SEB_LIBRARY_PATH: load another build of the same library (A/B measurements of compile-time variants)
    p = path or os.environ.get("SEB_LIBRARY_PATH") or LIB_PATH
    if not os.path.exists(p):
        raise SebError(f"{p} is missing: build it with `python seal-embedded_b200/build.py` "
                       "(the CUDA extension is the only implementation; there is no CPU path)")
    L = C.CDLL(p)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    L.seb_last_error.restype = C.c_char_p
    L.seb_create.argtypes = [sz, sz, vp, vp, C.c_double, i32, i32]
    L.seb_create.restype = vp
    L.seb_minimal_psi.argtypes = [sz, u32]
    L.seb_minimal_psi.restype = u32
    L.seb_destroy.argtypes = [vp]
    L.seb_destroy.restype = None
    L.seb_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.seb_measure_ceilings.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.seb_packed30_words.argtypes = [vp]
    L.seb_packed30_words.restype = sz
    L.seb_encrypt_asym_host_packed30.argtypes = [vp, vp, sz, vp, sz, vp]
    L.seb_encrypt_sym_host_packed30.argtypes = [vp, vp, sz, vp, vp, sz, vp, i32]
    L.seb_unpack30.argtypes = [vp, sz, vp]
    L.seb_unpack30_device.argtypes = [vp, vp, sz, vp]
    L.seb_gen_secret_key.argtypes = [vp, vp, vp]
    L.seb_digest_device.argtypes = [vp, vp, sz, sz, vp]
    L.seb_ct_to_seal_layout.argtypes = [vp, sz, sz, sz, vp]
    L.seb_ct_from_seal_layout.argtypes = [vp, sz, sz, sz, vp]
    L.seb_set_stream.argtypes = [vp, vp]
    L.seb_set_public_key.argtypes = [vp, vp, vp]
    L.seb_set_secret_key.argtypes = [vp, vp]
    L.seb_reserve.argtypes = [vp, sz]
    L.seb_degree.argtypes = [vp]
    L.seb_degree.restype = sz
    L.seb_nprimes.argtypes = [vp]
    L.seb_nprimes.restype = sz
    L.seb_scale.argtypes = [vp]
    L.seb_scale.restype = C.c_double
    L.seb_prime.argtypes = [vp, sz]
    L.seb_prime.restype = u32
    L.seb_launch_count.argtypes = [vp]
    L.seb_launch_count.restype = C.c_uint64
    L.seb_encrypt_asym_device.argtypes = [vp, vp, sz, vp, sz, vp]
    L.seb_encrypt_sym_device.argtypes = [vp, vp, sz, vp, vp, sz, vp, i32]
    L.seb_encode_failures.argtypes = [vp]
    L.seb_uniform_spec_misses.argtypes = [vp]
    L.seb_uniform_spec_misses.restype = C.c_long
    L.seb_encrypt_asym_host.argtypes = [vp, vp, sz, vp, sz, vp]
    L.seb_encrypt_sym_host.argtypes = [vp, vp, sz, vp, vp, sz, vp, i32]
    L.seb_encrypt_sym_seedct_device.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.seb_encrypt_sym_seedct_host.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.seb_expand_seedct_device.argtypes = [vp, vp, vp, sz, vp]
    L.seb_encode_device.argtypes = [vp, vp, sz, sz, vp]
    L.seb_sample_asym_device.argtypes = [vp, vp, sz, vp, vp, vp]
    L.seb_sample_cbd_device.argtypes = [vp, vp, vp, sz, sz, vp]
    L.seb_sample_uniform_device.argtypes = [vp, vp, vp, sz, sz, vp, sz]
    L.seb_ntt_device.argtypes = [vp, vp, sz]
    L.seb_prng_blocks_device.argtypes = [vp, vp, vp, sz, vp]
    L.seb_gen_public_key.argtypes = [vp, vp, vp, vp, vp, vp]
    L.seb_intt_device.argtypes = [vp, vp, sz]
    L.seb_decrypt_decode_device.argtypes = [vp, vp, sz, sz, sz, vp]
    L.seb_profile_begin.argtypes = [vp, i32]
    L.seb_profile_end.argtypes = [vp, vp]
    L.se_setup_custom.argtypes = [sz, sz, vp, vp, C.c_double, i32]
    L.se_setup_custom.restype = C.POINTER(_SeParms)
    L.se_setup.argtypes = [sz, sz, C.c_double, i32]
    L.se_setup.restype = C.POINTER(_SeParms)
    L.se_setup_default.argtypes = [i32]
    L.se_setup_default.restype = C.POINTER(_SeParms)
    L.se_encrypt_seeded.argtypes = [vp, vp, SEND_FNCT, vp, sz, C.c_bool, C.POINTER(_SeParms)]
    L.se_encrypt_seeded.restype = C.c_bool
    L.se_encrypt.argtypes = [SEND_FNCT, vp, sz, C.c_bool, C.POINTER(_SeParms)]
    L.se_encrypt.restype = C.c_bool
    L.se_cleanup.argtypes = [C.POINTER(_SeParms)]
    L.se_cleanup.restype = None
    L.se_encrypt_batch_seeded.argtypes = [vp, vp, vp, sz, sz, vp, C.POINTER(_SeParms)]
    L.se_encrypt_batch_seeded.restype = C.c_bool
    L.se_b200_set_reference_quirk.argtypes = [i32]
    L.se_b200_set_reference_quirk.restype = None
    L.se_b200_set_print_full.argtypes = [i32]
    L.se_b200_set_print_full.restype = None
    L.se_b200_set_sym_seed_ct.argtypes = [i32]
    L.se_b200_set_sym_seed_ct.restype = None
    L.se_encrypt_batch_seedct.argtypes = [vp, vp, vp, sz, sz, vp, C.POINTER(_SeParms)]
    L.se_encrypt_batch_seedct.restype = C.c_bool
    L.se_b200_context.argtypes = [C.POINTER(_SeParms)]
    L.se_b200_context.restype = vp
    if path is None:
        _lib = L
    return L


def ct_to_seal_layout(ct: np.ndarray) -> np.ndarray:
    """[batch][nprimes][2][n] u32 (device-library stream) -> [batch][2][nprimes][n] u64 (seal::Ciphertext data)."""
    ct = np.ascontiguousarray(ct, dtype=np.uint32)
    b, np_, two, n = ct.shape
    assert two == 2
    out = np.empty((b, 2, np_, n), np.uint64)
    rc = load_library().seb_ct_to_seal_layout(_addr(ct), b, np_, n, _addr(out))
    if rc:
        raise SebError(f"seb_ct_to_seal_layout: {rc}")
    return out


def ct_from_seal_layout(seal: np.ndarray) -> np.ndarray:
    seal = np.ascontiguousarray(seal, dtype=np.uint64)
    b, two, np_, n = seal.shape
    assert two == 2
    out = np.empty((b, np_, 2, n), np.uint32)
    rc = load_library().seb_ct_from_seal_layout(_addr(seal), b, np_, n, _addr(out))
    if rc:
        raise SebError(f"seb_ct_from_seal_layout: {rc} (a coefficient does not fit 32 bits)")
    return out


def minimal_psi(n: int, q: int) -> int:
    """Smallest primitive 2n-th root of unity mod q (host arithmetic inside the library; 0 if none)."""
    return int(load_library().seb_minimal_psi(n, q))


def _addr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    raise TypeError(type(x))


class Context:
    """One GPU context (``seb_ctx``): resident tables + keys + per-batch scratch."""

    def __init__(self, n: int, nprimes: int, asym: bool, device: int = -1, primes=None, psis=None, scale: float = 0.0,
                 handle: int | None = None):
        self.lib = load_library()
        self._owned = handle is None
        if handle is None:
            pa = np.ascontiguousarray(primes, dtype=np.uint32) if primes is not None else None
            ps = np.ascontiguousarray(psis, dtype=np.uint32) if psis is not None else None
            handle = self.lib.seb_create(n, nprimes, _addr(pa), _addr(ps), scale, int(asym), device)
            if not handle:
                raise SebError(self.lib.seb_last_error().decode())
        self.h = handle
        self.n = int(self.lib.seb_degree(self.h))
        self.nprimes = int(self.lib.seb_nprimes(self.h))
        self.scale = float(self.lib.seb_scale(self.h))
        self.primes = [int(self.lib.seb_prime(self.h, i)) for i in range(self.nprimes)]
        self.asym = asym

    def close(self) -> None:
        if self.h and self._owned:
            self.lib.seb_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise SebError(f"[{rc}] {self.lib.seb_last_error().decode()}")
        return rc

    # -- configuration
    def set_stream(self, cuda_stream: int | None) -> None:
        self._check(self.lib.seb_set_stream(self.h, cuda_stream or 0))

    def set_public_key(self, pk0: np.ndarray, pk1: np.ndarray) -> None:
        pk0 = np.ascontiguousarray(pk0, dtype=np.uint32)
        pk1 = np.ascontiguousarray(pk1, dtype=np.uint32)
        assert pk0.shape == pk1.shape == (self.nprimes, self.n)
        self._check(self.lib.seb_set_public_key(self.h, _addr(pk0), _addr(pk1)))

    def set_secret_key(self, sk_packed: np.ndarray) -> None:
        sk = np.ascontiguousarray(sk_packed, dtype=np.uint8)
        assert sk.size == self.n // 4
        self._check(self.lib.seb_set_secret_key(self.h, _addr(sk)))

    def gen_public_key(self, sk_packed: np.ndarray, ep_seed=bytes([7]) * 64, a_seed_base=bytes([9]) * 64):
        """gen_pk on the GPU (ckks_asym.c:159-171); returns (pk0, pk1) [nprimes][n] and installs them on an
        asymmetric context.  Default seeds = the recipe of the test key material."""
        sk = np.ascontiguousarray(sk_packed, dtype=np.uint8)
        assert sk.size == self.n // 4
        es = np.frombuffer(bytes(ep_seed), np.uint8).copy()
        as_ = np.frombuffer(bytes(a_seed_base), np.uint8).copy()
        assert es.size == SEED_BYTES and as_.size == SEED_BYTES
        pk0 = np.empty((self.nprimes, self.n), np.uint32)
        pk1 = np.empty((self.nprimes, self.n), np.uint32)
        self._check(self.lib.seb_gen_public_key(self.h, _addr(sk), _addr(es), _addr(as_), _addr(pk0), _addr(pk1)))
        return pk0, pk1

    def gen_secret_key(self, seed) -> np.ndarray:
        """sample_s (ckks_sym.c:162-173) on the GPU: returns the packed n/4-byte key and installs it."""
        sd = np.frombuffer(bytes(seed), np.uint8).copy()
        assert sd.size == SEED_BYTES
        sk = np.empty(self.n // 4, np.uint8)
        self._check(self.lib.seb_gen_secret_key(self.h, _addr(sd), _addr(sk)))
        return sk

    def set_option(self, name: str, value: int) -> None:
        """A/B switches ("uniform_coop", "uniform_fix_wide", "uniform_spec", "uniform_pair", "host_chunk"); < 0 = auto."""
        self._check(self.lib.seb_set_option(self.h, name.encode(), int(value)))

    def measure_ceilings(self) -> tuple[float, float]:
        """(Keccak-f[1600]/s, lazy butterflies/s) register-only on this device, measured now."""
        k, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.seb_measure_ceilings(self.h, C.byref(k), C.byref(b)))
        return k.value, b.value

    def digest_device(self, d_words, words_per_item: int, items: int, d_digests) -> None:
        self._check(self.lib.seb_digest_device(self.h, _addr(d_words), words_per_item, items, _addr(d_digests)))

    def reserve(self, batch: int) -> None:
        self._check(self.lib.seb_reserve(self.h, batch))

    @property
    def launch_count(self) -> int:
        return int(self.lib.seb_launch_count(self.h))

    # -- full path
    def encrypt_asym_device(self, d_values, vlen: int, d_seeds, batch: int, d_out) -> None:
        self._check(self.lib.seb_encrypt_asym_device(self.h, _addr(d_values), vlen, _addr(d_seeds), batch,
                                                     _addr(d_out)))

    def encrypt_sym_device(self, d_values, vlen: int, d_share_seeds, d_seeds, batch: int, d_out,
                           ref_quirk: bool = False) -> None:
        self._check(self.lib.seb_encrypt_sym_device(self.h, _addr(d_values), vlen, _addr(d_share_seeds),
                                                    _addr(d_seeds), batch, _addr(d_out), int(ref_quirk)))

    def uniform_spec_misses(self) -> int:
        return self._check(int(self.lib.seb_uniform_spec_misses(self.h)))

    def encode_failures(self) -> int:
        return self._check(self.lib.seb_encode_failures(self.h))

    def encrypt_asym_host(self, values: np.ndarray, seeds: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        batch, vlen = values.shape
        if out is None:
            out = np.empty((batch, self.nprimes, 2, self.n), np.uint32)
        self._check(self.lib.seb_encrypt_asym_host(self.h, _addr(values), vlen, _addr(seeds), batch, _addr(out)))
        return out

    def encrypt_sym_host(self, values: np.ndarray, share_seeds: np.ndarray, seeds: np.ndarray,
                         out: np.ndarray | None = None, ref_quirk: bool = False) -> np.ndarray:
        batch, vlen = values.shape
        if out is None:
            out = np.empty((batch, self.nprimes, 2, self.n), np.uint32)
        self._check(self.lib.seb_encrypt_sym_host(self.h, _addr(values), vlen, _addr(share_seeds), _addr(seeds),
                                                  batch, _addr(out), int(ref_quirk)))
        return out

    # -- optional packed wire form: 30 bits per residue (15 words per 16 residues)
    def packed30_words(self) -> int:
        return int(self.lib.seb_packed30_words(self.h))

    def encrypt_asym_host_packed30_raw(self, values_ptr: int, vlen: int, seeds_ptr: int, batch: int, out_ptr: int) -> None:
        self._check(self.lib.seb_encrypt_asym_host_packed30(self.h, values_ptr, vlen, seeds_ptr, batch, out_ptr))

    def encrypt_asym_host_packed30(self, values: np.ndarray, seeds: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        batch, vlen = values.shape
        if out is None:
            out = np.empty((batch, self.packed30_words()), np.uint32)
        self.encrypt_asym_host_packed30_raw(_addr(values), vlen, _addr(seeds), batch, _addr(out))
        return out

    def encrypt_sym_host_packed30(self, values: np.ndarray, share_seeds: np.ndarray, seeds: np.ndarray,
                                  ref_quirk: bool = False) -> np.ndarray:
        batch, vlen = values.shape
        out = np.empty((batch, self.packed30_words()), np.uint32)
        self._check(self.lib.seb_encrypt_sym_host_packed30(self.h, _addr(values), vlen, _addr(share_seeds), _addr(seeds),
                                                           batch, _addr(out), int(ref_quirk)))
        return out

    def unpack30_host(self, packed: np.ndarray) -> np.ndarray:
        """[batch][packed30_words] -> [batch][nprimes][2][n] on the host (seb_unpack30, no GPU)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint32)
        batch = packed.shape[0]
        out = np.empty((batch, self.nprimes, 2, self.n), np.uint32)
        rc = self.lib.seb_unpack30(_addr(packed), batch * 2 * self.nprimes * self.n, _addr(out))
        if rc:
            raise SebError(f"seb_unpack30: {rc}")
        return out

    def unpack30_device(self, d_packed, words: int, d_out) -> None:
        self._check(self.lib.seb_unpack30_device(self.h, _addr(d_packed), words, _addr(d_out)))

    # -- seed-compressed symmetric ciphertexts: c0 only; c1 = a is rebuilt from the shareable seed
    def encrypt_sym_seedct_device(self, d_values, vlen: int, d_share_seeds, d_seeds, batch: int, d_c0_out) -> None:
        self._check(self.lib.seb_encrypt_sym_seedct_device(self.h, _addr(d_values), vlen, _addr(d_share_seeds),
                                                           _addr(d_seeds), batch, _addr(d_c0_out)))

    def encrypt_sym_seedct_host(self, values: np.ndarray, share_seeds: np.ndarray, seeds: np.ndarray,
                                out: np.ndarray | None = None) -> np.ndarray:
        batch, vlen = values.shape
        if out is None:
            out = np.empty((batch, self.nprimes, self.n), np.uint32)
        self._check(self.lib.seb_encrypt_sym_seedct_host(self.h, _addr(values), vlen, _addr(share_seeds), _addr(seeds),
                                                         batch, _addr(out)))
        return out

    def expand_seedct_device(self, d_share_seeds, d_c0, batch: int, d_out) -> None:
        self._check(self.lib.seb_expand_seedct_device(self.h, _addr(d_share_seeds), _addr(d_c0), batch, _addr(d_out)))

    PROFILE_SEGMENTS = {True: ("encode", "sample_ternary", "sample_cbd", "encrypt"),
                        False: ("encode", "sample_cbd", "sample_uniform", "encrypt")}

    def profile_begin(self, max_steps: int) -> None:
        self._prof_max = max_steps
        self._check(self.lib.seb_profile_begin(self.h, max_steps))

    def profile_end(self) -> np.ndarray:
        """ms[step][4] per-kernel durations (CUDA events on the launching stream)."""
        ms = np.zeros((max(self._prof_max, 1), 4), np.float32)
        steps = self._check(self.lib.seb_profile_end(self.h, _addr(ms)))
        return ms[:steps]

    # -- stage level
    def encode_device(self, d_values, vlen: int, batch: int, d_pt) -> None:
        self._check(self.lib.seb_encode_device(self.h, _addr(d_values), vlen, batch, _addr(d_pt)))

    def sample_asym_device(self, d_seeds, batch: int, d_u, d_e, d_ctr) -> None:
        self._check(self.lib.seb_sample_asym_device(self.h, _addr(d_seeds), batch, _addr(d_u), _addr(d_e),
                                                    _addr(d_ctr)))

    def sample_cbd_device(self, d_seeds, d_ctr, npoly: int, batch: int, d_e) -> None:
        self._check(self.lib.seb_sample_cbd_device(self.h, _addr(d_seeds), _addr(d_ctr), npoly, batch, _addr(d_e)))

    def sample_uniform_device(self, d_seeds, d_ctr, prime_idx: int, batch: int, d_out, ct_stride: int) -> None:
        self._check(self.lib.seb_sample_uniform_device(self.h, _addr(d_seeds), _addr(d_ctr), prime_idx, batch,
                                                       _addr(d_out), ct_stride))

    def ntt_device(self, d_polys, batch: int) -> None:
        self._check(self.lib.seb_ntt_device(self.h, _addr(d_polys), batch))

    def intt_device(self, d_polys, batch: int) -> None:
        self._check(self.lib.seb_intt_device(self.h, _addr(d_polys), batch))

    def decrypt_decode_device(self, d_ct, batch: int, prime_idx: int, vlen: int, d_values_out) -> None:
        self._check(self.lib.seb_decrypt_decode_device(self.h, _addr(d_ct), batch, prime_idx, vlen, _addr(d_values_out)))

    def prng_blocks_device(self, d_seeds, d_counters, count: int, d_out) -> None:
        self._check(self.lib.seb_prng_blocks_device(self.h, _addr(d_seeds), _addr(d_counters), count, _addr(d_out)))


class SealEmbedded:
    """The reference's API, name for name (device/lib/seal_embedded.h:91-130).

    Key files are read from ``./adapter_output_data`` exactly like the reference
    (``sk_<n>.dat`` at setup for symmetric, ``pk{0,1}_ntt_<n>_<q>.dat`` for asymmetric).
    One instance per process (static state, as in the reference).
    """

    def __init__(self):
        self.lib = load_library()
        self.se_parms = None

    def se_setup(self, degree: int, nprimes: int, scale: float, encrypt_type: int):
        self.se_parms = self.lib.se_setup(degree, nprimes, scale, encrypt_type)
        return self.se_parms

    def se_setup_default(self, encrypt_type: int):
        self.se_parms = self.lib.se_setup_default(encrypt_type)
        return self.se_parms

    def se_setup_custom(self, degree, nprimes, modulus_vals, ratios, scale, encrypt_type):
        mv = np.ascontiguousarray(modulus_vals, dtype=np.uint32) if modulus_vals is not None else None
        rt = np.ascontiguousarray(ratios, dtype=np.uint32) if ratios is not None else None
        self.se_parms = self.lib.se_setup_custom(degree, nprimes, _addr(mv), _addr(rt), scale, encrypt_type)
        return self.se_parms

    @property
    def parms(self):
        return self.se_parms.contents.parms.contents

    def context(self) -> Context:
        p = self.parms
        return Context(p.coeff_count, p.nprimes, bool(p.is_asymmetric), handle=self.lib.se_b200_context(self.se_parms))

    def se_encrypt_seeded(self, shareable_seed, seed, send, v: np.ndarray, print_: bool = False) -> bool:
        """``send(data: bytes) -> int`` is called 2*nprimes times: c0 then c1 per prime."""
        v = np.ascontiguousarray(v, dtype=np.float32)
        ss = np.frombuffer(bytes(shareable_seed), np.uint8).copy() if shareable_seed is not None else None
        sd = np.frombuffer(bytes(seed), np.uint8).copy() if seed is not None else None

        def _cb(ptr, nbytes):
            return int(send(C.string_at(ptr, nbytes)))

        cb = SEND_FNCT(_cb) if send is not None else C.cast(None, SEND_FNCT)
        return bool(self.lib.se_encrypt_seeded(_addr(ss), _addr(sd), cb, _addr(v), v.size * 4, print_, self.se_parms))

    def se_encrypt(self, send, v: np.ndarray, print_: bool = False) -> bool:
        return self.se_encrypt_seeded(None, None, send, v, print_)

    def se_encrypt_batch_seeded(self, shareable_seeds, seeds, v: np.ndarray, out: np.ndarray | None = None):
        v = np.ascontiguousarray(v, dtype=np.float32)
        batch, vlen = v.shape
        p = self.parms
        if out is None:
            out = np.empty((batch, p.nprimes, 2, p.coeff_count), np.uint32)
        ok = self.lib.se_encrypt_batch_seeded(_addr(shareable_seeds), _addr(seeds), _addr(v), vlen, batch, _addr(out),
                                              self.se_parms)
        return bool(ok), out

    def set_reference_quirk(self, on: bool) -> None:
        self.lib.se_b200_set_reference_quirk(int(on))

    def set_print_full(self, on: bool) -> None:
        self.lib.se_b200_set_print_full(int(on))

    def set_sym_seed_ct(self, on: bool) -> None:
        self.lib.se_b200_set_sym_seed_ct(int(on))

    def se_encrypt_batch_seedct(self, shareable_seeds, seeds, v: np.ndarray, out: np.ndarray | None = None):
        v = np.ascontiguousarray(v, dtype=np.float32)
        batch, vlen = v.shape
        p = self.parms
        if out is None:
            out = np.empty((batch, p.nprimes, p.coeff_count), np.uint32)
        ok = self.lib.se_encrypt_batch_seedct(_addr(shareable_seeds), _addr(seeds), _addr(v), vlen, batch, _addr(out),
                                              self.se_parms)
        return bool(ok), out

    def se_cleanup(self) -> None:
        if self.se_parms is not None:
            self.lib.se_cleanup(self.se_parms)
            self.se_parms = None
