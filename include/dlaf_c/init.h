/* dlaf_c/init.h — replaces include/dlaf_c/init.h:27-35 of the reference.
 * The pika arguments are accepted and ignored (there is no pika runtime: the scheduler is a thin
 * stream/event issuer inside the library). Recognised --dlaf: options / DLAF_ environment variables:
 *   --dlaf:print-config                       (src/init.cpp:377-383)
 *   --dlaf:device=<ordinal> / DLAF_B200_DEVICE   CUDA device of this rank (reference hard-codes 0,
 *                                                src/init.cpp:126; default here: LOCAL_RANK or 0)
 * Idempotent like the reference (src/c_api/init.cpp:19-50). */
#pragma once

#include <dlaf_c/utils.h>

DLAF_EXTERN_C void dlaf_initialize(int argc_pika, const char** argv_pika, int argc_dlaf,
                                   const char** argv_dlaf) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_finalize(void) DLAF_NOEXCEPT;
