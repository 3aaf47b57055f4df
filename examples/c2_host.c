/* C2 through the drop-in boundary from a COMPILED host, the way gad itself would drive it: every step records a fresh
 * scalar tape (gad/src/net_ext.rs:52 builds a new Graph1 per example) — Rosenbrock-64 as SURVEY App. D writes it, 64
 * variables + 63 x 8 nodes + add_all = 569 nodes, ~1 100 calls into libgadcu — over `lanes` lanes at once, sweeps it
 * (Graph1::evaluate_gradients_once, graph.rs:279-290) and reads the 64 gradient columns. What is timed is that whole
 * call sequence, recording included. bench.py runs this program for the API-level batched-tape figure: a Python host pays
 * several microseconds of ctypes marshalling per call, which is an artefact of the test host, not of the boundary.
 *
 *   cc -O2 examples/c2_host.c -Iinclude -Lgad_b200/lib -lgadcu -Wl,-rpath,$PWD/gad_b200/lib -o examples/c2_host
 *   examples/c2_host [device [lanes [reps [seed0]]]]      -> one JSON line
 * Needs a B200 (no CPU fallback: gadcu_ctx_create fails elsewhere).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gadcu.h"

#define N 64
#define OK(call)                                                                  \
  do {                                                                            \
    gadcu_status st__ = (call);                                                   \
    if (st__ != GADCU_OK) {                                                       \
      fprintf(stderr, "%s -> %d: %s\n", #call, st__, gadcu_last_error());         \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

static double now_ms(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

/* one step: record, sweep, collect the gradient columns (grads[i] owned by the caller) */
static void step(gadcu_ctx* ctx, uint64_t lanes, gadcu_array* const* cols, gadcu_array* seed, gadcu_array** grads) {
  gadcu_algebra* g;
  OK(gadcu_batched_create(ctx, GADCU_F64, lanes, &g));
  gadcu_value x[N], terms[N - 1];
  for (int i = 0; i < N; i++) OK(gadcu_variable(g, cols[i], &x[i]));
  for (int i = 0; i + 1 < N; i++) { /* 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2, op for op as in SURVEY App. D */
    gadcu_value sq, d, d2, t1, nx, e, e2;
    OK(gadcu_mul(g, &x[i], &x[i], &sq));
    OK(gadcu_sub(g, &x[i + 1], &sq, &d));
    OK(gadcu_mul(g, &d, &d, &d2));
    OK(gadcu_mulc(g, &d2, 100.0, 1, &t1));
    OK(gadcu_neg(g, &x[i], &nx));
    OK(gadcu_addc(g, &nx, 1.0, 1, &e));
    OK(gadcu_mul(g, &e, &e, &e2));
    OK(gadcu_add(g, &t1, &e2, &terms[i]));
    gadcu_value_release(&sq); gadcu_value_release(&d); gadcu_value_release(&d2); gadcu_value_release(&t1);
    gadcu_value_release(&nx); gadcu_value_release(&e); gadcu_value_release(&e2);
  }
  const gadcu_value* ptrs[N - 1];
  for (int i = 0; i + 1 < N; i++) ptrs[i] = &terms[i];
  gadcu_value f;
  OK(gadcu_add_all(g, ptrs, N - 1, &f));
  gadcu_store* store;
  OK(gadcu_evaluate_gradients(g, f.arena, f.index, seed, /*once=*/1, &store));
  for (int i = 0; i < N; i++) {
    gadcu_value gv;
    int found = 0;
    OK(gadcu_store_get(store, x[i].arena, x[i].index, &gv, &found));
    if (!found) { fprintf(stderr, "no gradient for x%d\n", i); exit(1); }
    grads[i] = gv.data;
  }
  OK(gadcu_store_destroy(store));
  for (int i = 0; i + 1 < N; i++) gadcu_value_release(&terms[i]);
  for (int i = 0; i < N; i++) gadcu_value_release(&x[i]);
  gadcu_value_release(&f);
  OK(gadcu_algebra_destroy(g));
}

int main(int argc, char** argv) {
  const int device = argc > 1 ? atoi(argv[1]) : 0;
  const uint64_t lanes = argc > 2 ? strtoull(argv[2], NULL, 10) : (1ull << 20);
  const int reps = argc > 3 ? atoi(argv[3]) : 20;
  const uint64_t seed0 = argc > 4 ? strtoull(argv[4], NULL, 10) : 200;
  gadcu_ctx* ctx;
  OK(gadcu_ctx_create(device, &ctx));
  const gadcu_dim4 dims = {{lanes, 1, 1, 1}};
  gadcu_array* cols[N];
  for (int i = 0; i < N; i++) OK(gadcu_array_random(ctx, GADCU_F64, &dims, seed0 + (uint64_t)i, 0, -1.2, 1.2, &cols[i]));
  gadcu_array* seed;
  OK(gadcu_scalar_from_host(ctx, GADCU_F64, 1.0, &seed));
  gadcu_array* grads[N];
  double first_ms = 0.0;
  for (int w = 0; w < 3; w++) { /* the first call compiles the lane program (NVRTC) and fills the allocator's free lists */
    const double t0 = now_ms();
    step(ctx, lanes, cols, seed, grads);
    OK(gadcu_ctx_synchronize(ctx));
    if (w == 0) first_ms = now_ms() - t0;
    if (w < 2) for (int i = 0; i < N; i++) gadcu_array_release(grads[i]);
  }
  /* order-independent checksum of the first 4096 lanes of every gradient column: the sum of their bit patterns mod 2^64
   * (bench.py compares it with the same sum over the tape recorded from Python: equal sums = the same bits) */
  const uint64_t sample = lanes < 4096 ? lanes : 4096;
  double* host = (double*)malloc(lanes * sizeof(double));
  uint64_t checksum = 0;
  for (int i = 0; i < N; i++) {
    OK(gadcu_array_to_host(grads[i], host));
    for (uint64_t l = 0; l < sample; l++) {
      uint64_t bits;
      memcpy(&bits, &host[l], 8);
      checksum += bits;
    }
    gadcu_array_release(grads[i]);
  }
  free(host);
  uint64_t launches0 = 0, launches1 = 0;
  OK(gadcu_ctx_memory_stats(ctx, NULL, NULL, &launches0));
  OK(gadcu_ctx_synchronize(ctx));
  const double t0 = now_ms();
  double issue_ms = 0.0;
  for (int r = 0; r < reps; r++) {
    const double a = now_ms();
    step(ctx, lanes, cols, seed, grads);
    issue_ms += now_ms() - a;
    for (int i = 0; i < N; i++) gadcu_array_release(grads[i]);
  }
  OK(gadcu_ctx_synchronize(ctx));
  const double ms = (now_ms() - t0) / reps;
  OK(gadcu_ctx_memory_stats(ctx, NULL, NULL, &launches1));
  printf("{\"host\": \"C through include/gadcu.h\", \"lanes\": %llu, \"reps\": %d, \"ms_per_call\": %.6f, \"host_issue_ms_per_call\": %.6f, "
         "\"first_call_ms\": %.3f, \"grads_per_s\": %.6e, \"launches_per_call\": %.2f, \"bits_checksum_first_4096_lanes\": \"%llu\"}\n",
         (unsigned long long)lanes, reps, ms, issue_ms / reps, first_ms, (double)lanes / (ms * 1e-3), (double)(launches1 - launches0) / reps, (unsigned long long)checksum);
  for (int i = 0; i < N; i++) gadcu_array_release(cols[i]);
  gadcu_array_release(seed);
  OK(gadcu_ctx_destroy(ctx));
  return 0;
}
