/* Prints sizeof / offsetof of every caller-visible struct of the se_* API, one line per fact.  Compiled once
 * against the reference's own headers (-I/root/reference/device/lib) and once against include/; the two outputs
 * must be identical (tests/test_abi.py), and the reference's output is committed as tests/golden/ref_struct_layout.txt
 * for machines where /root/reference is not mounted. */
#include <stddef.h>
#include <stdio.h>

#include "seal_embedded.h"

#define SZ(T) printf("sizeof(" #T ") = %zu\n", sizeof(T))
#define OFF(T, f) printf("offsetof(" #T ", " #f ") = %zu size %zu\n", offsetof(T, f), sizeof(((T *)0)->f))

int main(void)
{
    SZ(ZZ);
    SZ(flpt);
    SZ(Modulus);
    OFF(Modulus, value);
    OFF(Modulus, const_ratio);
    SZ(Parms);
    OFF(Parms, coeff_count);
    OFF(Parms, logn);
    OFF(Parms, moduli);
    OFF(Parms, curr_modulus);
    OFF(Parms, curr_modulus_idx);
    OFF(Parms, nprimes);
    OFF(Parms, scale);
    OFF(Parms, is_asymmetric);
    OFF(Parms, pk_from_file);
    OFF(Parms, sample_s);
    OFF(Parms, small_s);
    OFF(Parms, small_u);
    SZ(SE_PTRS);
    OFF(SE_PTRS, conj_vals);
    OFF(SE_PTRS, ifft_roots);
    OFF(SE_PTRS, values);
    OFF(SE_PTRS, ternary);
    OFF(SE_PTRS, conj_vals_int_ptr);
    OFF(SE_PTRS, c0_ptr);
    OFF(SE_PTRS, c1_ptr);
    OFF(SE_PTRS, index_map_ptr);
    OFF(SE_PTRS, ntt_roots_ptr);
    OFF(SE_PTRS, ntt_pte_ptr);
    OFF(SE_PTRS, e1_ptr);
    SZ(SE_PARMS);
    OFF(SE_PARMS, parms);
    OFF(SE_PARMS, se_ptrs);
    SZ(EncryptType);
    printf("SE_SYM_ENCR = %d\nSE_ASYM_ENCR = %d\n", (int)SE_SYM_ENCR, (int)SE_ASYM_ENCR);
    SZ(SEND_FNCT_PTR);
    printf("SE_PRNG_SEED_BYTE_COUNT = %d\n", (int)SE_PRNG_SEED_BYTE_COUNT);
    printf("SE_SUCCESS = %d SE_ERR_NO_MEMORY = %d SE_ERR_INVALD_ARGUMENT = %d SE_ERR_UNKNOWN = %d SE_ERR_MINIMUM = %d\n",
           SE_SUCCESS, SE_ERR_NO_MEMORY, SE_ERR_INVALD_ARGUMENT, SE_ERR_UNKNOWN, SE_ERR_MINIMUM);
    return 0;
}
