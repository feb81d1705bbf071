"""seal-embedded_b200 — B200-native CKKS encode+encrypt behind SEAL-Embedded's se_setup/se_encrypt C ABI.

The product is the shared library ``libseal_embedded_b200.so`` (hand-written sm_100a CUDA kernels +
a C host layer, see ``csrc/`` and ``host/``; C ABI in ``include/seal_embedded_b200.h``).  This Python
package is a thin ctypes mirror of that ABI for tests, the benchmark and multi-GPU launch plumbing;
it contains no arithmetic and no CPU fallback: if the library is missing or fails to load, importing
the bindings raises.

The directory name carries a hyphen, so import it with
``importlib.import_module("seal-embedded_b200")``.
"""
from .api import (  # noqa: F401
    LIB_PATH,
    Context,
    SealEmbedded,
    SebError,
    build_library,
    ct_from_seal_layout,
    ct_to_seal_layout,
    load_library,
    minimal_psi,
)
from .shard import Shard, all_gather_ciphertexts, bind_to_gpu_numa, owner_of, shard_range  # noqa: F401
