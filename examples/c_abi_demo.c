/* The drop-in boundary from C: c = a + a*a on a Graph1 tape, gradient with seed 2 (gad/tests/graph.rs:8-26).
 *   cc examples/c_abi_demo.c -Iinclude -Lgad_b200/lib -lgadcu -Wl,-rpath,$PWD/gad_b200/lib -o /tmp/c_abi_demo && /tmp/c_abi_demo
 * Needs a B200 (gadcu_ctx_create fails with GADCU_ERR_CUDA elsewhere: there is no CPU fallback). */
#include <stdio.h>

#include "gadcu.h"

#define OK(call)                                                                  \
  do {                                                                            \
    gadcu_status st__ = (call);                                                   \
    if (st__ != GADCU_OK) {                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #call, st__, gadcu_last_error());         \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

int main(void) {
  gadcu_ctx* ctx;
  OK(gadcu_ctx_create(0, &ctx));
  gadcu_algebra* g;
  OK(gadcu_algebra_create(ctx, GADCU_GRAPH1, &g));
  const float host_a[3] = {1.0f, 2.0f, 3.0f}, host_seed[3] = {2.0f, 2.0f, 2.0f};
  const gadcu_dim4 dims = {{3, 1, 1, 1}};
  gadcu_array *a_data, *seed;
  OK(gadcu_array_from_host(ctx, GADCU_F32, &dims, host_a, &a_data));
  OK(gadcu_array_from_host(ctx, GADCU_F32, &dims, host_seed, &seed));
  gadcu_value a, aa, c;
  OK(gadcu_variable(g, a_data, &a));
  OK(gadcu_mul(g, &a, &a, &aa));
  OK(gadcu_add(g, &a, &aa, &c));
  gadcu_store* store;
  OK(gadcu_evaluate_gradients(g, c.arena, c.index, seed, /*once=*/1, &store));
  gadcu_value grad;
  int found = 0;
  OK(gadcu_store_get(store, a.arena, a.index, &grad, &found));
  float out[3];
  OK(gadcu_array_to_host(grad.data, out));
  printf("d(a + a*a)/da * 2 at a = (1, 2, 3): %g %g %g   (expected 6 10 14)\n", out[0], out[1], out[2]);
  /* a dimension error is a value, decided on the host before any launch */
  const gadcu_dim4 other = {{2, 1, 1, 1}};
  gadcu_array* b_data;
  OK(gadcu_array_from_host(ctx, GADCU_F32, &other, host_a, &b_data));
  gadcu_value b, bad;
  OK(gadcu_constant(g, b_data, &b));
  gadcu_status st = gadcu_add(g, &a, &b, &bad);
  printf("add of [3] and [2]: status %d (GADCU_ERR_DIMENSIONS = %d): %s\n", st, GADCU_ERR_DIMENSIONS, gadcu_last_error());
  gadcu_value_release(&grad);
  gadcu_value_release(&a);
  gadcu_value_release(&aa);
  gadcu_value_release(&c);
  gadcu_value_release(&b);
  gadcu_store_destroy(store);
  gadcu_array_release(a_data);
  gadcu_array_release(b_data);
  gadcu_array_release(seed);
  gadcu_algebra_destroy(g);
  gadcu_ctx_destroy(ctx);
  return 0;
}
