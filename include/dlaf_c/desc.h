/* dlaf_c/desc.h — matrix descriptor of the C API (same fields and meaning as the reference's
 * include/dlaf_c/desc.h:16-26). */
#pragma once

struct DLAF_descriptor {
  int m;    /* rows of the global matrix */
  int n;    /* columns of the global matrix */
  int mb;   /* row blocking factor */
  int nb;   /* column blocking factor */
  int isrc; /* process row owning the first row of the global matrix */
  int jsrc; /* process column owning the first column of the global matrix */
  int i;    /* first row of the sub-matrix, must be 0 */
  int j;    /* first column of the sub-matrix, must be 0 */
  int ld;   /* leading dimension of the local matrix */
};
