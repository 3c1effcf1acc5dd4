/* c_abi_smoke.c -- the drop-in boundary is plain C: this file is compiled with `gcc -std=c99 -Wall -Werror`,
 * linked against libb200node.so and run by tests/test_abi.py (no GPU needed: only the entry points that do not touch
 * a device are called; on a machine without a GPU b200_ctx_create must fail loudly, never fall back). */
#include <stdio.h>
#include <string.h>

#include "b200node.h"

int main(void) {
  b200_opts o;
  b200_ctx* ctx = NULL;
  int first[8], last[8], n, rc;

  if (b200_abi_version() != B200_ABI_VERSION) { printf("abi version mismatch\n"); return 1; }
  b200_opts_default(&o);
  if (!(o.abstol > 0.f) || !(o.reltol > 0.f) || o.maxiters < 1) { printf("bad defaults\n"); return 2; }
  /* group_ranges(10, 4): Julia 1:4, 4:7, 7:10 -> 0-based inclusive (0,3) (3,6) (6,9); pure host arithmetic */
  n = b200_group_ranges(10, 4, first, last, 8);
  if (n != 3 || first[0] != 0 || last[0] != 3 || first[2] != 6 || last[2] != 9) { printf("group_ranges: %d\n", n); return 3; }
  if (b200_group_ranges(10, 1, first, last, 8) >= 0) { printf("group_size 1 must be rejected\n"); return 4; }
  rc = b200_ctx_create(0, NULL, &ctx);
  if (rc == 0) {
    printf("ctx ok (GPU present)\n");
    b200_ctx_destroy(ctx);
  } else if (rc == B200_ERR_CUDA && ctx == NULL) {
    printf("no GPU: ctx_create failed loudly (%d)\n", rc);
  } else {
    printf("unexpected ctx_create result %d\n", rc);
    return 5;
  }
  printf("c abi ok, sizeof(b200_opts)=%u sizeof(b200_stats)=%u\n", (unsigned)sizeof(b200_opts), (unsigned)sizeof(b200_stats));
  return 0;
}
