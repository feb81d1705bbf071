/*
 * seal_embedded.h — the reference's header NAME (device/lib/seal_embedded.h), so that an application
 * written for SEAL-Embedded's device library compiles unchanged with -I<this directory>:
 * the declarations it needs (se_setup*, se_encrypt*, se_cleanup, Parms, Modulus, SE_PTRS, SE_PARMS,
 * EncryptType, SEND_FNCT_PTR, ZZ, flpt) are in seal_embedded_b200.h.
 * The reverse also holds and is tested (tests/test_abi.py, tests/test_gpu_round2.py): the same
 * application compiled against the REFERENCE's own seal_embedded.h links and runs against
 * libseal_embedded_b200.so, because the struct layouts and prototypes are identical.
 */
#ifndef SEAL_EMBEDDED_COMPAT_H
#define SEAL_EMBEDDED_COMPAT_H
#include "seal_embedded_b200.h"
#endif
