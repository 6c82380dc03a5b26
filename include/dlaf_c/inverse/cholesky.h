/* dlaf_c/inverse/cholesky.h — inverse of a Hermitian positive definite matrix from its Cholesky factor (POTRI), same
 * names, argument meaning and memory contract as the reference's include/dlaf_c/inverse/cholesky.h:32-48 and :76-89:
 *
 *   a     caller-owned HOST pointer to this rank's local part of the block-cyclic matrix, column-major, leading
 *         dimension desc.ld; on entry the Cholesky factor in the `uplo` triangle (the output of
 *         dlaf_cholesky_factorization_* / dlaf_p?potrf), on exit the `uplo` triangle of inv(A); the other triangle is
 *         left untouched (src/c_api/inverse/cholesky.h:40-59).
 *   uplo  'L' or 'U'.
 *   return / *info: 0.
 * Synchronous: returns when the result is in `a`. Collective over the grid of the context. */
#pragma once

#include <dlaf_c/desc.h>
#include <dlaf_c/utils.h>

DLAF_EXTERN_C int dlaf_inverse_from_cholesky_factor_s(const int dlaf_context, const char uplo, float* a,
                                                      const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;
DLAF_EXTERN_C int dlaf_inverse_from_cholesky_factor_d(const int dlaf_context, const char uplo, double* a,
                                                      const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;
DLAF_EXTERN_C int dlaf_inverse_from_cholesky_factor_c(const int dlaf_context, const char uplo,
                                                      dlaf_complex_c* a,
                                                      const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;
DLAF_EXTERN_C int dlaf_inverse_from_cholesky_factor_z(const int dlaf_context, const char uplo,
                                                      dlaf_complex_z* a,
                                                      const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;

/* ScaLAPACK-like: desca = {1, ctxt, m, n, mb, nb, rsrc, csrc, lld}; ia == ja == 1 required
 * (src/c_api/inverse/cholesky.h:63-75). The context is the value returned by dlaf_create_grid. Always available
 * (the reference: only with DLAF_WITH_SCALAPACK). */
DLAF_EXTERN_C void dlaf_pspotri(const char uplo, const int n, float* a, const int ia, const int ja,
                                const int desca[9], int* info) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_pdpotri(const char uplo, const int n, double* a, const int ia, const int ja,
                                const int desca[9], int* info) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_pcpotri(const char uplo, const int n, dlaf_complex_c* a, const int ia,
                                const int ja, const int desca[9], int* info) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_pzpotri(const char uplo, const int n, dlaf_complex_z* a, const int ia,
                                const int ja, const int desca[9], int* info) DLAF_NOEXCEPT;
