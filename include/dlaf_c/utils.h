/* dlaf_c/utils.h — B200-native drop-in for DLA-Future's C API (POTRF path only).
 * Replaces: include/dlaf_c/utils.h:23-52 of the reference (complex typedefs, make_dlaf_descriptor). */
#pragma once

#ifdef __cplusplus
#define DLAF_EXTERN_C extern "C"
#define DLAF_NOEXCEPT noexcept
#include <complex>
typedef std::complex<float> dlaf_complex_c;
typedef std::complex<double> dlaf_complex_z;
#else
#define DLAF_EXTERN_C
#define DLAF_NOEXCEPT
#include <complex.h>
typedef float complex dlaf_complex_c;
typedef double complex dlaf_complex_z;
#endif

#include <dlaf_c/desc.h>

/* ScaLAPACK descriptor {dtype=1, ctxt, m, n, mb, nb, rsrc, csrc, lld} + (i, j) 1-based -> DLAF_descriptor
 * (reference: src/c_api/utils.cpp:26-34). */
DLAF_EXTERN_C struct DLAF_descriptor make_dlaf_descriptor(const int m, const int n, const int i,
                                                          const int j, const int desc[9]) DLAF_NOEXCEPT;
