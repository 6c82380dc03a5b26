/* dlaf_c/factorization/cholesky.h — the POTRF entry points, same names, argument meaning and memory
 * contract as the reference's include/dlaf_c/factorization/cholesky.h:32-47 and :74-87:
 *
 *   a     caller-owned HOST pointer to this rank's local part of the block-cyclic matrix, column-major,
 *         leading dimension desc.ld; overwritten in place with the factor in the `uplo` triangle; the
 *         other triangle is left untouched (src/c_api/factorization/cholesky.h:32-58).
 *   uplo  'L' or 'U'.
 *   return / *info: 0 on success. Improvement over the reference (which always returns 0 and
 *         std::terminate()s / __trap()s on a non-SPD input): a LAPACK-style k > 0 when the leading minor
 *         of order k is not positive definite.
 * Synchronous: returns when the result is in `a`. Not re-entrant (like the reference's global grid map). */
#pragma once

#include <dlaf_c/desc.h>
#include <dlaf_c/utils.h>

DLAF_EXTERN_C int dlaf_cholesky_factorization_s(const int dlaf_context, const char uplo, float* a,
                                                const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;
DLAF_EXTERN_C int dlaf_cholesky_factorization_d(const int dlaf_context, const char uplo, double* a,
                                                const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;
DLAF_EXTERN_C int dlaf_cholesky_factorization_c(const int dlaf_context, const char uplo,
                                                dlaf_complex_c* a,
                                                const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;
DLAF_EXTERN_C int dlaf_cholesky_factorization_z(const int dlaf_context, const char uplo,
                                                dlaf_complex_z* a,
                                                const struct DLAF_descriptor dlaf_desca) DLAF_NOEXCEPT;

/* ScaLAPACK-like: desca = {1, ctxt, m, n, mb, nb, rsrc, csrc, lld}; ia == ja == 1 required
 * (src/c_api/factorization/cholesky.h:62-73). The context is the value returned by dlaf_create_grid. */
DLAF_EXTERN_C void dlaf_pspotrf(const char uplo, const int n, float* a, const int ia, const int ja,
                                const int desca[9], int* info) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_pdpotrf(const char uplo, const int n, double* a, const int ia, const int ja,
                                const int desca[9], int* info) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_pcpotrf(const char uplo, const int n, dlaf_complex_c* a, const int ia,
                                const int ja, const int desca[9], int* info) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_pzpotrf(const char uplo, const int n, dlaf_complex_z* a, const int ia,
                                const int ja, const int desca[9], int* info) DLAF_NOEXCEPT;
