//! `extern "C"` declarations of include/gadcu.h — one entry point per `Eval for af::Array<T>` method of
//! the reference (gad/src/{core,arith,const_arith,analytic,compare,array,array_compare,matrix}.rs),
//! plus the native tape / store / sweep entry points (gad/src/graph.rs, store.rs).
//! SOURCE ONLY (no Rust toolchain in the build image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub type gadcu_status = c_int;
pub const GADCU_OK: gadcu_status = 0;
// gad/src/error.rs:9-40
pub const GADCU_ERR_DIMENSIONS: gadcu_status = 1;
pub const GADCU_ERR_REDUCED_DIMENSIONS: gadcu_status = 2;
pub const GADCU_ERR_LENGTHS: gadcu_status = 3;
pub const GADCU_ERR_EMPTY: gadcu_status = 4;
pub const GADCU_ERR_MISSING_ID: gadcu_status = 5;
pub const GADCU_ERR_MISSING_GRADIENT: gadcu_status = 6;
pub const GADCU_ERR_MISSING_NODE: gadcu_status = 7;
// new with this backend
pub const GADCU_ERR_CUDA: gadcu_status = 8;
pub const GADCU_ERR_NCCL: gadcu_status = 9;
pub const GADCU_ERR_UNSUPPORTED: gadcu_status = 10;
pub const GADCU_ERR_INVALID: gadcu_status = 11;
pub const GADCU_ERR_TYPE: gadcu_status = 12;

#[repr(C)] pub struct gadcu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct gadcu_array { _p: [u8; 0] }
#[repr(C)] pub struct gadcu_algebra { _p: [u8; 0] }
#[repr(C)] pub struct gadcu_store { _p: [u8; 0] }
#[repr(C)] pub struct gadcu_comm { _p: [u8; 0] }
#[repr(C)] pub struct gadcu_capture { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct gadcu_dim4 { pub d: [u64; 4] }
#[repr(C)] #[derive(Clone, Copy)]
pub struct gadcu_matprop { pub transposed: u8, pub conjugated: u8 }
/// gad's `Value<D>` (graph.rs:35-43): data + optional id (`index == 0` is `id: None`).
#[repr(C)] #[derive(Clone, Copy)]
pub struct gadcu_value { pub data: *mut gadcu_array, pub arena: u32, pub index: u32 }

pub const GADCU_F32: c_int = 0;
pub const GADCU_F64: c_int = 1;
pub const GADCU_EVAL: c_int = 1;
pub const GADCU_GRAPH1: c_int = 2;
pub const GADCU_GRAPHN: c_int = 3;
pub const GADCU_BATCHED1: c_int = 4;

type A = *mut gadcu_algebra;
type V = *const gadcu_value;
type O = *mut gadcu_value;
type D = *const gadcu_dim4;
/// VJP of a user-defined node (gad/src/graph.rs:106-165; tests/extension.rs:61-96): gets the gradient algebra
/// (Eval for Graph1, the graph itself for GraphN), the store and the node's own gradient.
pub type gadcu_vjp_fn = unsafe extern "C" fn(user: *mut c_void, grad_algebra: A, store: *mut gadcu_store, gradient: V) -> gadcu_status;

extern "C" {
    pub fn gadcu_last_error() -> *const c_char;
    pub fn gadcu_ctx_create(device: c_int, out: *mut *mut gadcu_ctx) -> gadcu_status;
    pub fn gadcu_ctx_destroy(ctx: *mut gadcu_ctx) -> gadcu_status;
    pub fn gadcu_ctx_synchronize(ctx: *mut gadcu_ctx) -> gadcu_status;
    // arrays: af::Array::new / host / dims / clone / drop
    pub fn gadcu_array_from_host(ctx: *mut gadcu_ctx, dt: c_int, dims: D, host: *const c_void, out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_array_to_host(a: *const gadcu_array, host: *mut c_void) -> gadcu_status;
    pub fn gadcu_array_clone(a: *const gadcu_array, out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_array_retain(a: *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_array_release(a: *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_array_dims(a: *const gadcu_array, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_scalar_from_host(ctx: *mut gadcu_ctx, dt: c_int, v: f64, out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_scalar_to_host(a: *const gadcu_array, out: *mut f64) -> gadcu_status;
    pub fn gadcu_value_release(v: O) -> gadcu_status;
    // WeightOps (net.rs:161-180)
    pub fn gadcu_array_add_assign(this: *mut gadcu_array, other: *const gadcu_array) -> gadcu_status;
    pub fn gadcu_array_scale(a: *const gadcu_array, lambda: f64, out: *mut *mut gadcu_array) -> gadcu_status;
    // Check (dims only; the per-family `impl ... for Check` blocks)
    pub fn gadcu_check_equal(dims: *const D, n: usize, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_moddims(v: D, d: D, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_tile_as(v: D, r: D, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_sum_as(v: D, r: D, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_as_scalar(v: D) -> gadcu_status;
    pub fn gadcu_check_dot(a: D, b: D) -> gadcu_status;
    pub fn gadcu_check_reduced(v: D, r: D, keeps_shape: c_int, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_transpose(v: D, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_matmul(a: D, b: D, p1: gadcu_matprop, p2: gadcu_matprop, out: *mut gadcu_dim4) -> gadcu_status;
    // algebras: Eval::default / Graph1::new / GraphN::new (lib.rs:523-532, graph.rs:77-92), Clone (graph.rs:353-360)
    pub fn gadcu_algebra_create(ctx: *mut gadcu_ctx, kind: c_int, out: *mut A) -> gadcu_status;
    pub fn gadcu_algebra_clone(alg: *const gadcu_algebra, out: *mut A) -> gadcu_status;
    pub fn gadcu_algebra_destroy(alg: A) -> gadcu_status;
    // CoreAlgebra (core.rs:12-38)
    pub fn gadcu_variable(g: A, data: *mut gadcu_array, out: O) -> gadcu_status;
    pub fn gadcu_constant(g: A, data: *mut gadcu_array, out: O) -> gadcu_status;
    pub fn gadcu_add(g: A, a: V, b: V, out: O) -> gadcu_status;
    pub fn gadcu_add_all(g: A, vs: *const V, n: usize, out: O) -> gadcu_status;
    // ArithAlgebra (arith.rs:14-32)
    pub fn gadcu_sub(g: A, a: V, b: V, out: O) -> gadcu_status;
    pub fn gadcu_mul(g: A, a: V, b: V, out: O) -> gadcu_status;
    pub fn gadcu_zeros(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_ones(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_neg(g: A, v: V, out: O) -> gadcu_status;
    // ConstArithAlgebra (const_arith.rs:14-26); c_is_int = the constant has type i16
    pub fn gadcu_setc(g: A, v: V, c: f64, c_is_int: c_int, out: O) -> gadcu_status;
    pub fn gadcu_addc(g: A, v: V, c: f64, c_is_int: c_int, out: O) -> gadcu_status;
    pub fn gadcu_mulc(g: A, v: V, c: f64, c_is_int: c_int, out: O) -> gadcu_status;
    pub fn gadcu_powc(g: A, v: V, c: f64, c_is_int: c_int, out: O) -> gadcu_status;
    // AnalyticAlgebra (analytic.rs:16-56)
    pub fn gadcu_exp(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_log(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_log1p(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_sin(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_cos(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_tanh(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_sigmoid(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_reciprocal(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_sqrt(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_div(g: A, a: V, b: V, out: O) -> gadcu_status;
    pub fn gadcu_pow(g: A, v: V, p: V, out: O) -> gadcu_status;
    // CompareAlgebra (compare.rs:15-66)
    pub fn gadcu_select_argmax(g: A, v0: V, v1: V, r0_or_null: V, r1_or_null: V, out: O) -> gadcu_status;
    // ArrayAlgebra (array.rs:13-45)
    pub fn gadcu_flat(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_moddims(g: A, v: V, dims: D, out: O) -> gadcu_status;
    pub fn gadcu_tile_as(g: A, v: V, dims: D, out: O) -> gadcu_status;
    pub fn gadcu_sum_as(g: A, v: V, dims: D, out: O) -> gadcu_status;
    pub fn gadcu_constant_as(g: A, scalar: V, dims: D, out: O) -> gadcu_status;
    pub fn gadcu_as_scalar(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_scale(g: A, lambda: V, v: V, out: O) -> gadcu_status;
    pub fn gadcu_dot(g: A, a: V, b: V, out: O) -> gadcu_status;
    // MatrixAlgebra (matrix.rs:20-32)
    pub fn gadcu_matmul(g: A, a: V, b: V, p1: gadcu_matprop, p2: gadcu_matprop, out: O) -> gadcu_status;
    pub fn gadcu_transpose(g: A, v: V, conjugate: c_int, out: O) -> gadcu_status;
    // ArrayCompareAlgebra (array_compare.rs:16-22)
    pub fn gadcu_max_as(g: A, v: V, dims: D, out: O) -> gadcu_status;
    pub fn gadcu_argmax_as(g: A, v: V, dims: D, out: O) -> gadcu_status;
    pub fn gadcu_softmax_as(g: A, v: V, dims: D, out: O) -> gadcu_status;
    // reverse sweep (graph.rs:263-318) and GradientReader::read (store.rs:27-29)
    pub fn gadcu_evaluate_gradients(g: A, arena: u32, index: u32, seed: *mut gadcu_array, once: c_int, out: *mut *mut gadcu_store) -> gadcu_status;
    pub fn gadcu_compute_gradients(g: A, arena: u32, index: u32, seed: V, out: *mut *mut gadcu_store) -> gadcu_status;
    pub fn gadcu_store_get(s: *const gadcu_store, arena: u32, index: u32, out: O, found: *mut c_int) -> gadcu_status;
    pub fn gadcu_store_destroy(s: *mut gadcu_store) -> gadcu_status;
    // data parallelism (new)
    pub fn gadcu_comm_unique_id(out_128_bytes: *mut c_void) -> gadcu_status;
    pub fn gadcu_comm_init(ctx: *mut gadcu_ctx, nranks: c_int, rank: c_int, id: *const c_void, out: *mut *mut gadcu_comm) -> gadcu_status;
    pub fn gadcu_comm_allreduce_sum(comm: *mut gadcu_comm, arrays: *const *mut gadcu_array, n: usize) -> gadcu_status;
    pub fn gadcu_comm_destroy(comm: *mut gadcu_comm) -> gadcu_status;
}

// ---- the rest of include/gadcu.h: helpers of the derived trait methods, user nodes, and the entry points that have no
// counterpart in the reference (tests/test_check_and_abi.py keeps this file and the header in step)
extern "C" {
    pub fn gadcu_version() -> *const c_char;
    pub fn gadcu_debug_gemm_schedule(sm_count: c_int, m: u64, n: u64, k: u64, push_splits: c_int, rows6: *mut u32, cap: usize, info8: *mut u32) -> usize;
    pub fn gadcu_ctx_stream(ctx: *mut gadcu_ctx) -> *mut c_void;
    pub fn gadcu_ctx_set_fusion(ctx: *mut gadcu_ctx, fuse: c_int) -> gadcu_status;
    pub fn gadcu_ctx_set_lazy(ctx: *mut gadcu_ctx, enable: c_int, min_elements: u64) -> gadcu_status;
    pub fn gadcu_ctx_set_strict_reference(ctx: *mut gadcu_ctx, strict: c_int) -> gadcu_status;
    pub fn gadcu_ctx_set_gemm_epilogue(ctx: *mut gadcu_ctx, enable: c_int) -> gadcu_status;
    pub fn gadcu_ctx_set_auto_replay(ctx: *mut gadcu_ctx, enable: c_int) -> gadcu_status;
    pub fn gadcu_ctx_flush(ctx: *mut gadcu_ctx) -> gadcu_status;
    pub fn gadcu_ctx_replay_stats(ctx: *mut gadcu_ctx, segments: *mut u64, reused: *mut u64, instantiated: *mut u64) -> gadcu_status;
    pub fn gadcu_ctx_memory_stats(ctx: *mut gadcu_ctx, bytes_in_use: *mut u64, bytes_reserved: *mut u64, kernel_launches: *mut u64) -> gadcu_status;
    pub fn gadcu_ctx_profile_begin(ctx: *mut gadcu_ctx) -> gadcu_status;
    pub fn gadcu_ctx_profile_end(ctx: *mut gadcu_ctx, ms_by_category4: *mut f64, count_by_category4: *mut u64) -> gadcu_status;
    pub fn gadcu_ctx_wait_uploads(ctx: *mut gadcu_ctx) -> gadcu_status;
    pub fn gadcu_ctx_wait_download(ctx: *mut gadcu_ctx, ticket: u64) -> gadcu_status;
    // arrays
    pub fn gadcu_array_alloc(ctx: *mut gadcu_ctx, dt: c_int, dims: D, out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_array_random(ctx: *mut gadcu_ctx, dt: c_int, dims: D, seed: u64, normal: c_int, a: f64, b: f64, out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_array_upload(a: *mut gadcu_array, pinned_host: *const c_void, stream_or_null: *mut c_void) -> gadcu_status;
    pub fn gadcu_array_upload_async(a: *mut gadcu_array, pinned_host: *const c_void) -> gadcu_status;
    pub fn gadcu_array_download_async(a: *const gadcu_array, pinned_host: *mut c_void, ticket: *mut u64) -> gadcu_status;
    pub fn gadcu_array_dtype(a: *const gadcu_array) -> c_int;
    pub fn gadcu_array_is_scalar(a: *const gadcu_array) -> c_int;
    pub fn gadcu_array_device_ptr(a: *const gadcu_array) -> *mut c_void;
    // Check
    pub fn gadcu_check_select_argmax(v0: D, v1: D, r0_or_null: D, r1_or_null: D, out: *mut gadcu_dim4) -> gadcu_status;
    pub fn gadcu_check_flat(v: D, out: *mut gadcu_dim4) -> gadcu_status;
    // algebras
    pub fn gadcu_algebra_get_kind(alg: *const gadcu_algebra) -> c_int;
    pub fn gadcu_algebra_num_nodes(alg: *const gadcu_algebra) -> usize;
    // derived CompareAlgebra / ArrayAlgebra / MatrixAlgebra methods (compare.rs:24-66, array.rs:42-44, matrix.rs:28-31)
    pub fn gadcu_min(g: A, a: V, b: V, out: O) -> gadcu_status;
    pub fn gadcu_max(g: A, a: V, b: V, out: O) -> gadcu_status;
    pub fn gadcu_abs(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_sign(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_relu(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_norm2(g: A, v: V, out: O) -> gadcu_status;
    pub fn gadcu_matmul_nn(g: A, a: V, b: V, out: O) -> gadcu_status;
    // user-defined nodes (graph.rs:106-165, store.rs:44-55)
    pub fn gadcu_make_node(g: A, data: *mut gadcu_array, inputs: *const V, n: usize, vjp: gadcu_vjp_fn, user: *mut c_void,
                           drop_user: Option<unsafe extern "C" fn(*mut c_void)>, out: O) -> gadcu_status;
    pub fn gadcu_store_add_gradient(store: *mut gadcu_store, grad_algebra: A, arena: u32, index: u32, value: V) -> gadcu_status;
    pub fn gadcu_store_ids(s: *const gadcu_store, out_indices: *mut u32, cap: usize) -> usize;
    // CUDA-graph capture of everything issued between begin and end
    pub fn gadcu_capture_begin(ctx: *mut gadcu_ctx) -> gadcu_status;
    pub fn gadcu_capture_end(ctx: *mut gadcu_ctx, out: *mut *mut gadcu_capture) -> gadcu_status;
    pub fn gadcu_capture_launch(c: *mut gadcu_capture) -> gadcu_status;
    pub fn gadcu_capture_destroy(c: *mut gadcu_capture) -> gadcu_status;
    // batched scalar tapes: one Graph1 tape per lane
    pub fn gadcu_batched_create(ctx: *mut gadcu_ctx, dt: c_int, lanes: u64, out: *mut A) -> gadcu_status;
    pub fn gadcu_batched_force(g: A, v: V, out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_batched_rerun(s: *mut gadcu_store, columns: *const *mut gadcu_array, n: usize, gradients_out: *mut *mut gadcu_array) -> gadcu_status;
    pub fn gadcu_batched_program_stats(s: *const gadcu_store, instructions: *mut u64, slots: *mut u64) -> gadcu_status;
    pub fn gadcu_batched_sweep_memo_stats(replayed: *mut u64, walked: *mut u64, structures: *mut u64) -> gadcu_status;
    // data parallelism
    pub fn gadcu_comm_allreduce_sum_async(comm: *mut gadcu_comm, arrays: *const *mut gadcu_array, n: usize) -> gadcu_status;
    pub fn gadcu_comm_join(comm: *mut gadcu_comm) -> gadcu_status;
    pub fn gadcu_evaluate_gradients_dp(g: A, arena: u32, index: u32, seed: *mut gadcu_array, once: c_int, comm: *mut gadcu_comm,
                                       out: *mut *mut gadcu_store) -> gadcu_status;
    pub fn gadcu_evaluate_gradients_dp_with(g: A, arena: u32, index: u32, seed: *mut gadcu_array, once: c_int, comm: *mut gadcu_comm,
                                            also: *const *mut gadcu_array, n_also: usize, out: *mut *mut gadcu_store) -> gadcu_status;
    pub fn gadcu_compute_gradients_dp(g: A, arena: u32, index: u32, seed: V, comm: *mut gadcu_comm, out: *mut *mut gadcu_store) -> gadcu_status;
}
