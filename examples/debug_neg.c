/* minimal C-ABI client: variable -> neg -> evaluate_gradients_once */
#include <stdio.h>
#include "gadcu.h"
int main(void) {
  gadcu_ctx* ctx; gadcu_algebra* g; gadcu_array *x, *seed; gadcu_value a, b, ga; gadcu_store* st; int found;
  float hx[12], hs[12], out[12];
  for (int i = 0; i < 12; i++) { hx[i] = i - 5.f; hs[i] = 1.f + i; }
  gadcu_dim4 d = {{4, 3, 1, 1}};
  { int rc0 = gadcu_ctx_create(0, &ctx); printf("ctx %d\n", rc0); if (rc0) return 1; }
  printf("alg %d\n", gadcu_algebra_create(ctx, GADCU_GRAPH1, &g));
  printf("arr %d\n", gadcu_array_from_host(ctx, GADCU_F32, &d, hx, &x));
  printf("arr %d\n", gadcu_array_from_host(ctx, GADCU_F32, &d, hs, &seed));
  printf("var %d\n", gadcu_variable(g, x, &a));
  printf("neg %d id=%u\n", gadcu_neg(g, &a, &b), b.index);
  int rc = gadcu_evaluate_gradients(g, b.arena, b.index, seed, 1, &st);
  printf("sweep %d %s\n", rc, rc ? gadcu_last_error() : "");
  if (!rc) {
    gadcu_store_get(st, a.arena, a.index, &ga, &found);
    gadcu_array_to_host(ga.data, out);
    for (int i = 0; i < 12; i++) printf("%g ", out[i]);
    printf("\n");
  }
  return 0;
}
