/* A real C translation unit (like the reference's test_cholesky_c_api_wrapper.c): proves that the public
 * headers are valid C and that the library links with C linkage. No compute here (runs without a GPU). */
#include <dlaf_c/b200_ext.h>
#include <dlaf_c/desc.h>
#include <dlaf_c/factorization/cholesky.h>
#include <dlaf_c/grid.h>
#include <dlaf_c/init.h>
#include <dlaf_c/utils.h>
#include <stdio.h>

int main(void) {
  int desca[9] = {1, 0, 300, 300, 64, 64, 0, 0, 300};
  struct DLAF_descriptor d;
  int ctx, out[4];
  dlaf_initialize(0, NULL, 0, NULL);
  ctx = dlaf_create_grid(NULL, 1, 1, 'R');
  desca[1] = ctx;
  d = make_dlaf_descriptor(300, 300, 1, 1, desca);
  dlaf_b200_grid_info(ctx, out);
  printf("ctx_ok %d desc %d %d %d %d ld %d grid %dx%d rank %d,%d local %dx%d\n", ctx > 0, d.m, d.n, d.mb, d.nb, d.ld,
         out[0], out[1], out[2], out[3], dlaf_b200_local_rows(ctx, d), dlaf_b200_local_cols(ctx, d));
  {
    /* taking the addresses is enough to require the symbols at link time */
    void (*volatile fp[5])(void) = {(void (*)(void)) dlaf_pdpotrf, (void (*)(void)) dlaf_cholesky_factorization_z,
                                    (void (*)(void)) dlaf_pspotrf, (void (*)(void)) dlaf_pcpotrf,
                                    (void (*)(void)) dlaf_pzpotrf};
    int k, n = 0;
    for (k = 0; k < 5; ++k)
      n += fp[k] != 0;
    printf("symbols %d\n", n == 5);
  }
  dlaf_free_grid(ctx);
  dlaf_finalize();
  return 0;
}
