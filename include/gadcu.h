/* gadcu.h — C ABI of the B200-native array backend for gad (facebookresearch/gad).
 *
 * This is the drop-in boundary: the entry points a Rust binding crate would declare in an
 * `extern "C"` block to implement gad's algebra traits for a device-resident CUDA array type
 * (see INTEGRATION.md for the binding). Every function cites the reference interface it serves
 * (paths relative to the reference checkout, `gad/src/...`).
 *
 * Conventions
 *  - All functions return a gadcu_status. 0 = OK; 1..7 are gad's own Error variants
 *    (gad/src/error.rs:9-40) with the same meaning; >= 8 are backend errors. Nothing aborts.
 *    Dimension checks run on the host BEFORE any kernel is launched (as `Eval` calls `Check`
 *    first: gad/src/arith.rs:66, matrix.rs:52, array.rs:71,77,88,107,120), so a failing call
 *    launches nothing. gadcu_last_error() returns a thread-local description (variant + dims).
 *  - Arrays are immutable, refcounted, column-major, Dim4-shaped (dim 0 fastest; af::Dim4
 *    convention, gad/src/core.rs:120-136) device buffers of f32 or f64. Every op returns a new
 *    handle (+1 reference owned by the caller). Clone = gadcu_array_clone, O(1) (a new handle on the same buffer).
 *  - Scalars produced by array ops (`dot`, `as_scalar`, gad/src/array.rs:106-126) stay on the
 *    device as 1-element "scalar" arrays; gadcu_scalar_to_host reads (and synchronises) on demand.
 *  - A gadcu_value is gad's `Value<D>` (gad/src/graph.rs:35-43): data + optional node id
 *    (index == 0 means constant / `id: None`). Ids are (arena, 1-based index) ordered
 *    lexicographically (gad/src/store.rs:12-16, 155-178).
 *  - An algebra handle is one of gad's default algebras (gad/src/lib.rs:519-532): Eval, Graph1,
 *    GraphN — the SAME op entry points serve all three, mirroring how one generic gad program is
 *    instantiated with different algebras (gad/src/arrayfire.rs:39-61). `Check` works on dims only
 *    and needs no device: the gadcu_check_* functions.
 *  - Thread safety: handles may be used from several threads; evaluate_gradients (non-once) does
 *    not modify the tape (gad/src/graph.rs:263-274). All work of a ctx is issued on its stream.
 */
#ifndef GADCU_H_
#define GADCU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GADCU_API __attribute__((visibility("default")))

typedef int gadcu_status;
enum {
  GADCU_OK = 0,
  /* gad/src/error.rs:9-40 */
  GADCU_ERR_DIMENSIONS = 1,
  GADCU_ERR_REDUCED_DIMENSIONS = 2,
  GADCU_ERR_LENGTHS = 3,
  GADCU_ERR_EMPTY = 4,
  GADCU_ERR_MISSING_ID = 5,
  GADCU_ERR_MISSING_GRADIENT = 6,
  GADCU_ERR_MISSING_NODE = 7,
  /* backend */
  GADCU_ERR_CUDA = 8,
  GADCU_ERR_NCCL = 9,
  GADCU_ERR_UNSUPPORTED = 10,
  GADCU_ERR_INVALID = 11,
  GADCU_ERR_TYPE = 12 /* "indices should have a unique type" (store.rs:75-78) / dtype mismatch */
};

typedef enum { GADCU_F32 = 0, GADCU_F64 = 1 } gadcu_dtype;
typedef enum { GADCU_EVAL = 1, GADCU_GRAPH1 = 2, GADCU_GRAPHN = 3, GADCU_BATCHED1 = 4 } gadcu_algebra_kind;

typedef struct { uint64_t d[4]; } gadcu_dim4;
typedef struct { uint8_t transposed, conjugated; } gadcu_matprop; /* gad/src/matrix.rs:13-17 */

typedef struct gadcu_ctx gadcu_ctx;
typedef struct gadcu_array gadcu_array;
typedef struct gadcu_algebra gadcu_algebra;
typedef struct gadcu_store gadcu_store;
typedef struct gadcu_comm gadcu_comm;
typedef struct gadcu_capture gadcu_capture;

typedef struct {
  gadcu_array* data; /* owned reference (release with gadcu_value_release); NULL for lazy batched values */
  uint32_t arena;
  uint32_t index;    /* 0 = constant */
} gadcu_value;

GADCU_API const char* gadcu_last_error(void);
GADCU_API const char* gadcu_version(void);
/* Host-only introspection (no device needed): the work list the tensor-core GEMM would give every 2-CTA cluster for an
 * f32 product [M,K]x[K,N] on a GPU with `sm_count` SMs, from the kernel's own scheduling code. rows6: one row per work
 * item = {cluster, tile row, tile column, first k-block, end k-block, part (0 whole unit, 1 tail, 2 head of a split
 * tile)}; info8 = {clusters, k splits, last-wave tiles, tail k-blocks, k-blocks per tile, band, band along N, tiles}.
 * push_splits: 0, or the k splits of the tile-pushing data-parallel mode. Returns the number of work items (may exceed
 * cap). Used by tests/test_gemm_schedule.py to check coverage and pairing for many shapes without a GPU. */
GADCU_API size_t gadcu_debug_gemm_schedule(int sm_count, uint64_t M, uint64_t N, uint64_t K, int push_splits, uint32_t* rows6,
                                           size_t cap, uint32_t* info8);

/* ---- context: device, stream, caching allocator, error slot -------------------------------- */
GADCU_API gadcu_status gadcu_ctx_create(int device, gadcu_ctx** out);
GADCU_API gadcu_status gadcu_ctx_destroy(gadcu_ctx* ctx);
GADCU_API gadcu_status gadcu_ctx_synchronize(gadcu_ctx* ctx);
GADCU_API void* gadcu_ctx_stream(gadcu_ctx* ctx); /* cudaStream_t */
/* fuse != 0 (default): the Graph1 sweep runs each VJP + gradient accumulation as one kernel;
 * 0: one kernel per primitive, exactly the reference's op sequence (bit-identical results). */
GADCU_API gadcu_status gadcu_ctx_set_fusion(gadcu_ctx* ctx, int fuse);
/* The reference's matmul VJP (gad/src/matrix.rs:164-176) is the correct adjoint only for operands that
 * are not themselves transposed; for a transposed operand it yields a Dimensions error (or the
 * transposed adjoint when square). strict != 0 reproduces that as is; the default (0) computes the
 * correct adjoint in those cases, which GraphN needs to differentiate through recorded backward
 * GEMMs (Hessian-vector products). Identical wherever the reference formula is valid. */
GADCU_API gadcu_status gadcu_ctx_set_strict_reference(gadcu_ctx* ctx, int strict);
/* Deferred elementwise regions (on by default when fusion is on and NVRTC loads): elementwise ops on arrays of at
 * least `min_elements` elements are recorded and run as ONE specialised kernel per dependency chain when their
 * bytes are needed (a reduction, a GEMM, a host read, the end of a reverse sweep). Values are bit-identical to
 * the eager kernels'. enable = 0 restores one kernel per op / per VJP contribution. */
GADCU_API gadcu_status gadcu_ctx_set_lazy(gadcu_ctx* ctx, int enable, uint64_t min_elements);
/* Fused GEMM epilogues (on by default, with deferred regions): a tensor-core matmul is launched when its result is needed,
 * and the elementwise ops a layer puts right behind it — add(., tile_as(bias)) -> tanh (core.rs:179-182, analytic.rs:96-98), the
 * tanh VJP on an incoming dA (analytic.rs:385-395), sub(target, add(., tile_as(bias))) (net_ext.rs:110-118) — run on the
 * accumulator in the GEMM's epilogue, which also writes the TF32 operand copies the next product reads. Same op order, one
 * rounding each: values are bit-identical to enable = 0 (every product launched where it is recorded, tails as separate kernels).
 * enable = 1 (default): where it pays — products whose 256 x 256 output tiles keep at least half of the chip's SM pairs busy;
 * 2: always; 0: never. */
GADCU_API gadcu_status gadcu_ctx_set_gemm_epilogue(gadcu_ctx* ctx, int enable);
GADCU_API gadcu_status gadcu_ctx_memory_stats(gadcu_ctx* ctx, uint64_t* bytes_in_use, uint64_t* bytes_reserved,
                                              uint64_t* kernel_launches);

/* Per-category kernel timing with CUDA events on the context's stream, for roofline reporting:
 * categories 0 = matmul, 1 = elementwise (incl. fused VJPs), 2 = reductions, 3 = tile/transpose. */
GADCU_API gadcu_status gadcu_ctx_profile_begin(gadcu_ctx* ctx);
GADCU_API gadcu_status gadcu_ctx_profile_end(gadcu_ctx* ctx, double* ms_by_category4, uint64_t* count_by_category4);

/* ---- arrays: af::Array::new / host / dims / clone / drop -------------------------------------
 * gad/src/arrayfire.rs:131-161 (host, Array::new), core.rs:120-136 (dims). */
GADCU_API gadcu_status gadcu_array_from_host(gadcu_ctx* ctx, gadcu_dtype dt, const gadcu_dim4* dims, const void* host,
                                             gadcu_array** out);
GADCU_API gadcu_status gadcu_array_to_host(const gadcu_array* a, void* host); /* synchronises */
GADCU_API gadcu_status gadcu_array_alloc(gadcu_ctx* ctx, gadcu_dtype dt, const gadcu_dim4* dims, gadcu_array** out);
/* async H2D into an existing, uniquely owned array (static input buffers of a captured step) */
GADCU_API gadcu_status gadcu_array_upload(gadcu_array* a, const void* pinned_host, void* stream_or_null);
/* Prefetch: the copy runs on the context's copy stream once everything already issued on the compute stream has
 * finished (the target may still be read by it), overlapping the launches issued afterwards. Work that reads `a`
 * must be issued after gadcu_ctx_wait_uploads(), which makes the compute stream wait for the copies (no host sync). */
GADCU_API gadcu_status gadcu_array_upload_async(gadcu_array* a, const void* pinned_host);
GADCU_API gadcu_status gadcu_ctx_wait_uploads(gadcu_ctx* ctx);
/* The read-back mirror of the prefetch: the device -> host copy of `a` into caller-owned (pinned) memory is queued on the
 * compute stream behind the work that produces `a`, and the call returns at once with a ticket; the host looks at the bytes
 * after gadcu_ctx_wait_download(ticket) — typically one step later, so that reading a step's loss (what a training loop logs)
 * does not keep the host from issuing the next step. Tickets complete in order; waiting for a retired one returns at once. */
GADCU_API gadcu_status gadcu_array_download_async(const gadcu_array* a, void* pinned_host, uint64_t* ticket);
GADCU_API gadcu_status gadcu_ctx_wait_download(gadcu_ctx* ctx, uint64_t ticket);
/* af::Array::clone: a NEW handle on the same device buffer, O(1) (what gad's pervasive `.clone()` costs: graph.rs:152-155,
 * net.rs:100-130 get_weights). Handles made this way are independent values: gadcu_array_add_assign on one of them
 * leaves the others' contents alone (copy-on-write). gadcu_array_retain, by contrast, adds a reference to the SAME handle. */
GADCU_API gadcu_status gadcu_array_clone(const gadcu_array* a, gadcu_array** out);
GADCU_API gadcu_status gadcu_array_retain(gadcu_array* a);
GADCU_API gadcu_status gadcu_array_release(gadcu_array* a);
GADCU_API gadcu_status gadcu_array_dims(const gadcu_array* a, gadcu_dim4* out);
GADCU_API gadcu_dtype gadcu_array_dtype(const gadcu_array* a);
GADCU_API int gadcu_array_is_scalar(const gadcu_array* a);
GADCU_API void* gadcu_array_device_ptr(const gadcu_array* a); /* NULL for lazy broadcasts / unforced lane values */
GADCU_API gadcu_status gadcu_scalar_from_host(gadcu_ctx* ctx, gadcu_dtype dt, double v, gadcu_array** out);
GADCU_API gadcu_status gadcu_scalar_to_host(const gadcu_array* a, double* out); /* synchronises */
GADCU_API gadcu_status gadcu_value_release(gadcu_value* v);
/* seeded synthetic data generated on the device (bench / tests): U[lo,hi) or N(mean, std) */
GADCU_API gadcu_status gadcu_array_random(gadcu_ctx* ctx, gadcu_dtype dt, const gadcu_dim4* dims, uint64_t seed,
                                          int normal, double a, double b, gadcu_array** out);
/* WeightOps (gad/src/net.rs:161-180): `self += other`, and `self * lambda`. `+=` follows af::Array's copy-on-write:
 * the sum is written in place when `self` is the only holder of its buffer; otherwise it goes to a fresh buffer and
 * `self` is rebound to it, so clones (gadcu_array_clone), weight snapshots and values captured by a tape keep the old
 * contents. */
GADCU_API gadcu_status gadcu_array_add_assign(gadcu_array* self, const gadcu_array* other);
GADCU_API gadcu_status gadcu_array_scale(const gadcu_array* a, double lambda, gadcu_array** out);

/* ---- Check: the dimension algebra (gad/src/lib.rs:519-520); pure host functions ---------------
 * Scalars have dims `()` and always check OK (core.rs:75-93 etc.), so only the Dim4 rules appear. */
GADCU_API gadcu_status gadcu_check_equal(const gadcu_dim4* const* dims, size_t n, gadcu_dim4* out); /* add sub mul div pow min max add_all: error.rs:139-153 */
GADCU_API gadcu_status gadcu_check_select_argmax(const gadcu_dim4* v0, const gadcu_dim4* v1, const gadcu_dim4* r0_or_null,
                                                 const gadcu_dim4* r1_or_null, gadcu_dim4* out); /* compare.rs:128-144 */
GADCU_API gadcu_status gadcu_check_flat(const gadcu_dim4* v, gadcu_dim4* out);                          /* array.rs:133-136 */
GADCU_API gadcu_status gadcu_check_moddims(const gadcu_dim4* v, const gadcu_dim4* d, gadcu_dim4* out);  /* array.rs:138-145 */
GADCU_API gadcu_status gadcu_check_tile_as(const gadcu_dim4* v, const gadcu_dim4* r, gadcu_dim4* out);  /* array.rs:147-157 */
GADCU_API gadcu_status gadcu_check_sum_as(const gadcu_dim4* v, const gadcu_dim4* r, gadcu_dim4* out);   /* array.rs:159-170 */
GADCU_API gadcu_status gadcu_check_as_scalar(const gadcu_dim4* v);                                     /* array.rs:177-181 */
GADCU_API gadcu_status gadcu_check_dot(const gadcu_dim4* a, const gadcu_dim4* b);                       /* array.rs:188-192 */
GADCU_API gadcu_status gadcu_check_reduced(const gadcu_dim4* v, const gadcu_dim4* r, int keeps_shape, gadcu_dim4* out); /* max_as / argmax_as / softmax_as: array_compare.rs:84-101, error.rs:175-189 */
GADCU_API gadcu_status gadcu_check_transpose(const gadcu_dim4* v, gadcu_dim4* out);                     /* matrix.rs:118-125 */
GADCU_API gadcu_status gadcu_check_matmul(const gadcu_dim4* a, const gadcu_dim4* b, gadcu_matprop p1, gadcu_matprop p2,
                                          gadcu_dim4* out);                                            /* matrix.rs:88-116 */

/* ---- algebras -------------------------------------------------------------------------------- */
GADCU_API gadcu_status gadcu_algebra_create(gadcu_ctx* ctx, gadcu_algebra_kind kind, gadcu_algebra** out); /* Eval::default / Graph::new: lib.rs:523-532, graph.rs:77-84 */
GADCU_API gadcu_status gadcu_algebra_clone(const gadcu_algebra* alg, gadcu_algebra** out);                 /* graph.rs:353-360 (arena id preserved) */
GADCU_API gadcu_status gadcu_algebra_destroy(gadcu_algebra* alg);
GADCU_API gadcu_algebra_kind gadcu_algebra_get_kind(const gadcu_algebra* alg);
GADCU_API size_t gadcu_algebra_num_nodes(const gadcu_algebra* alg);

/* CoreAlgebra — core.rs:12-38 (trait), 162-183 (Eval for arrays), 203-262 (Graph) */
GADCU_API gadcu_status gadcu_variable(gadcu_algebra* g, gadcu_array* data, gadcu_value* out);
GADCU_API gadcu_status gadcu_constant(gadcu_algebra* g, gadcu_array* data, gadcu_value* out);
GADCU_API gadcu_status gadcu_add(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_add_all(gadcu_algebra* g, const gadcu_value* const* vs, size_t n, gadcu_value* out);
/* ArithAlgebra — arith.rs:14-32, 40-75, 153-234 */
GADCU_API gadcu_status gadcu_sub(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_mul(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_zeros(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_ones(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_neg(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
/* ConstArithAlgebra<_, T> and <_, i16> — const_arith.rs:14-26, 33-56, 120-187.
 * c_is_int != 0 selects the i16 instantiation (the constant is converted with T::from). */
GADCU_API gadcu_status gadcu_setc(gadcu_algebra* g, const gadcu_value* v, double c, int c_is_int, gadcu_value* out);
GADCU_API gadcu_status gadcu_addc(gadcu_algebra* g, const gadcu_value* v, double c, int c_is_int, gadcu_value* out);
GADCU_API gadcu_status gadcu_mulc(gadcu_algebra* g, const gadcu_value* v, double c, int c_is_int, gadcu_value* out);
GADCU_API gadcu_status gadcu_powc(gadcu_algebra* g, const gadcu_value* v, double c, int c_is_int, gadcu_value* out);
/* AnalyticAlgebra — analytic.rs:16-56, 64-128, 287-484 */
GADCU_API gadcu_status gadcu_exp(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_log(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_log1p(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_sin(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_cos(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_tanh(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_sigmoid(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_reciprocal(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_sqrt(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_div(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_pow(gadcu_algebra* g, const gadcu_value* v, const gadcu_value* p, gadcu_value* out);
/* CompareAlgebra — compare.rs:15-66, 74-115, 190-251 */
GADCU_API gadcu_status gadcu_select_argmax(gadcu_algebra* g, const gadcu_value* v0, const gadcu_value* v1,
                                           const gadcu_value* r0_or_null, const gadcu_value* r1_or_null, gadcu_value* out);
GADCU_API gadcu_status gadcu_min(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_max(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_abs(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_sign(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_relu(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
/* ArrayAlgebra — array.rs:13-45, 57-127, 196-357 (Scalar values are 1-element scalar arrays) */
GADCU_API gadcu_status gadcu_flat(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_moddims(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* dims, gadcu_value* out);
GADCU_API gadcu_status gadcu_tile_as(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* dims, gadcu_value* out);
GADCU_API gadcu_status gadcu_sum_as(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* dims, gadcu_value* out);
GADCU_API gadcu_status gadcu_constant_as(gadcu_algebra* g, const gadcu_value* scalar, const gadcu_dim4* dims, gadcu_value* out);
GADCU_API gadcu_status gadcu_as_scalar(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_scale(gadcu_algebra* g, const gadcu_value* lambda, const gadcu_value* v, gadcu_value* out);
GADCU_API gadcu_status gadcu_dot(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_norm2(gadcu_algebra* g, const gadcu_value* v, gadcu_value* out);
/* MatrixAlgebra — matrix.rs:20-32, 40-61, 141-199 */
GADCU_API gadcu_status gadcu_matmul(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_matprop p1,
                                    gadcu_matprop p2, gadcu_value* out);
GADCU_API gadcu_status gadcu_matmul_nn(gadcu_algebra* g, const gadcu_value* a, const gadcu_value* b, gadcu_value* out);
GADCU_API gadcu_status gadcu_transpose(gadcu_algebra* g, const gadcu_value* v, int conjugate, gadcu_value* out);
/* ArrayCompareAlgebra — array_compare.rs:16-22, 30-82, 104-174 */
GADCU_API gadcu_status gadcu_max_as(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* dims, gadcu_value* out);
GADCU_API gadcu_status gadcu_argmax_as(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* dims, gadcu_value* out);
GADCU_API gadcu_status gadcu_softmax_as(gadcu_algebra* g, const gadcu_value* v, const gadcu_dim4* dims, gadcu_value* out);

/* ---- user-defined operators: Graph::make_node + GradientStore::add_gradient ---------------------
 * gad/src/graph.rs:106-165, store.rs:44-55; used like gad/tests/extension.rs:61-96. The callback
 * receives the gradient algebra (Eval for Graph1, the graph itself for GraphN), the store, and the
 * node's own gradient (already dimension-checked against the forward dims). */
typedef gadcu_status (*gadcu_vjp_fn)(void* user, gadcu_algebra* grad_algebra, gadcu_store* store, const gadcu_value* gradient);
GADCU_API gadcu_status gadcu_make_node(gadcu_algebra* g, gadcu_array* data, const gadcu_value* const* inputs, size_t n,
                                       gadcu_vjp_fn vjp, void* user, void (*drop_user)(void*), gadcu_value* out);
GADCU_API gadcu_status gadcu_store_add_gradient(gadcu_store* store, gadcu_algebra* grad_algebra, uint32_t arena,
                                                uint32_t index, const gadcu_value* value);

/* ---- reverse sweep ------------------------------------------------------------------------------
 * Graph1::evaluate_gradients / evaluate_gradients_once (graph.rs:263-290): gradients are data.
 * GraphN::compute_gradients (graph.rs:307-318): gradients are values recorded on the tape.      */
GADCU_API gadcu_status gadcu_evaluate_gradients(gadcu_algebra* g, uint32_t arena, uint32_t index, gadcu_array* seed,
                                                int once, gadcu_store** out);
GADCU_API gadcu_status gadcu_compute_gradients(gadcu_algebra* g, uint32_t arena, uint32_t index, const gadcu_value* seed,
                                               gadcu_store** out);
/* GradientReader::read (store.rs:27-29): *found = 0 when the id has no gradient (absent != zero). */
GADCU_API gadcu_status gadcu_store_get(const gadcu_store* s, uint32_t arena, uint32_t index, gadcu_value* out, int* found);
GADCU_API size_t gadcu_store_ids(const gadcu_store* s, uint32_t* out_indices, size_t cap);
GADCU_API gadcu_status gadcu_store_destroy(gadcu_store* s);

/* ---- automatic CUDA-graph replay keyed by tape structure ------------------------------------------
 * gad rebuilds its tape every step (net_ext.rs:52), so a captured graph has no tape to live on; but the kernels a step
 * issues are the same sequence every time. With auto replay on, what is issued on the context's stream is recorded
 * instead of launched, and runs as ONE graph launch at the next point where the host needs results or order: a read, a
 * synchronise, an upload, a collective, the end of a reverse sweep (gadcu_evaluate_gradients / gadcu_compute_gradients),
 * gadcu_ctx_flush. The recorded sequence is laid over an executable graph of the same shape kept from an earlier step
 * (same kernels in the same order; pointers and scalars are updated), or instantiated when there is none: the structural
 * key is the sequence itself and nothing is assumed about the program. Results are identical to eager issue. Off by
 * default: while it is on, work issued through this API reaches gadcu_ctx_stream() only at those points (taking the
 * stream with gadcu_ctx_stream flushes). Profiled stretches (gadcu_ctx_profile_begin) and data-parallel sweeps are
 * issued eagerly. */
GADCU_API gadcu_status gadcu_ctx_set_auto_replay(gadcu_ctx* ctx, int enable);
GADCU_API gadcu_status gadcu_ctx_flush(gadcu_ctx* ctx);
/* segments run so far; how many reused an earlier executable graph; how many had to instantiate one */
GADCU_API gadcu_status gadcu_ctx_replay_stats(gadcu_ctx* ctx, uint64_t* segments, uint64_t* reused, uint64_t* instantiated);

/* ---- CUDA-graph capture of a whole step (forward + sweep + update) -------------------------------
 * Between begin and end every launch of the ctx is recorded instead of executed; buffers allocated
 * meanwhile stay reserved for the capture, so handles created inside remain valid and are refilled
 * by every gadcu_capture_launch. */
GADCU_API gadcu_status gadcu_capture_begin(gadcu_ctx* ctx);
GADCU_API gadcu_status gadcu_capture_end(gadcu_ctx* ctx, gadcu_capture** out);
GADCU_API gadcu_status gadcu_capture_launch(gadcu_capture* c);
GADCU_API gadcu_status gadcu_capture_destroy(gadcu_capture* c);

/* ---- batched scalar tapes (GADCU_BATCHED1): one gad scalar tape per lane, SoA ---------------------
 * Variables/constants are [lanes] columns; ops only record; the sweep compiles forward + reverse
 * into one per-lane program and runs it with one lane per thread. */
GADCU_API gadcu_status gadcu_batched_create(gadcu_ctx* ctx, gadcu_dtype dt, uint64_t lanes, gadcu_algebra** out);
GADCU_API gadcu_status gadcu_batched_force(gadcu_algebra* g, const gadcu_value* v, gadcu_array** out); /* forward value column */
/* The same tapes at new points: `columns[i]` replaces the data of the i-th variable (creation order); the compiled
 * lane program runs again — one kernel launch, nothing is re-recorded — and gradients_out[i] receives the gradient
 * column of the i-th variable (+1 reference; NULL where the sweep produced none). `s` must come from a batched sweep. */
GADCU_API gadcu_status gadcu_batched_rerun(gadcu_store* s, gadcu_array* const* columns, size_t n, gadcu_array** gradients_out);
GADCU_API gadcu_status gadcu_batched_program_stats(const gadcu_store* s, uint64_t* instructions, uint64_t* slots);
/* A batched sweep over a tape structure seen before (same recorded instructions, nodes, root and seed kind) is not walked
 * again: the instructions it appended and the store it filled are replayed from a process-wide table (GADCU_SWEEP_MEMO=0
 * turns that off). Counters since process start: sweeps served from the table, sweeps walked, structures held. */
GADCU_API gadcu_status gadcu_batched_sweep_memo_stats(uint64_t* replayed, uint64_t* walked, uint64_t* structures);

/* ---- data parallelism over the batch dimension (new: the reference is single-process) ------------
 * apply_gradient_step sums per-example gradients (net_ext.rs:62-65), so sharding the batch and
 * all-reducing the weight gradients reproduces it up to summation order. */
GADCU_API gadcu_status gadcu_comm_unique_id(void* out_128_bytes);
GADCU_API gadcu_status gadcu_comm_init(gadcu_ctx* ctx, int nranks, int rank, const void* id_128_bytes, gadcu_comm** out);
GADCU_API gadcu_status gadcu_comm_allreduce_sum(gadcu_comm* comm, gadcu_array* const* arrays, size_t n); /* in place */
/* Overlapped form: the collectives run on the communicator's own stream after everything already issued on the
 * compute stream; gadcu_comm_join makes the compute stream wait for them (no host synchronisation). */
GADCU_API gadcu_status gadcu_comm_allreduce_sum_async(gadcu_comm* comm, gadcu_array* const* arrays, size_t n);
GADCU_API gadcu_status gadcu_comm_join(gadcu_comm* comm);
/* evaluate_gradients[_once] of one batch shard (net_ext.rs:47-73 sums per-example gradients): every variable's gradient
 * is all-reduced as soon as the sweep has completed it, overlapping the rest of the sweep; joined before returning.
 * comm == NULL or a 1-rank communicator: identical to gadcu_evaluate_gradients. */
GADCU_API gadcu_status gadcu_evaluate_gradients_dp(gadcu_algebra* g, uint32_t arena, uint32_t index, gadcu_array* seed, int once,
                                                   gadcu_comm* comm, gadcu_store** out);
/* Same, and `also` (e.g. the loss of the step) is summed over the ranks in the same small exchange as the bias gradients. */
GADCU_API gadcu_status gadcu_evaluate_gradients_dp_with(gadcu_algebra* g, uint32_t arena, uint32_t index, gadcu_array* seed, int once,
                                                        gadcu_comm* comm, gadcu_array* const* also, size_t n_also, gadcu_store** out);
/* GraphN::compute_gradients (graph.rs:307-318) of one batch shard: the backward pass is recorded on this rank's tape as
 * usual; each variable's gradient is summed over the ranks as soon as the sweep has completed it (on the communicator's
 * stream, overlapping the rest of the sweep) and the store returns that sum as a constant. Joined before returning. */
GADCU_API gadcu_status gadcu_compute_gradients_dp(gadcu_algebra* g, uint32_t arena, uint32_t index, const gadcu_value* seed,
                                                  gadcu_comm* comm, gadcu_store** out);
GADCU_API gadcu_status gadcu_comm_destroy(gadcu_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* GADCU_H_ */
