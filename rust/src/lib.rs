//! gad-cuda: `CuArray<T>` + `impl XAlgebra<CuArray<T>> for gad::{Eval, Check}`.
//!
//! With these impls in scope, gad's `Graph1`/`GraphN` exist for `CuArray<T>` through gad's own
//! `impl_graph!`-style generic impls (`Graph<C>: XAlgebra<Value<D>>` whenever `Eval: XAlgebra<D>`,
//! e.g. gad/src/arith.rs:153-234) — exactly how gad/tests/symbolic.rs:40-93 derives `SymGraph1` from
//! `SymEval`. A gad program generic over `A: CuAlgebra<T>` (the analogue of `AfAlgebra<T>`,
//! gad/src/arrayfire.rs:9-37) switches backends by changing the array type parameter only.
//!
//! `CuGraph1` (bottom) additionally exposes the NATIVE device tape (fused VJP kernels, device store)
//! behind the same traits for callers who want the sweep itself to run inside libgadcu.
//!
//! SOURCE ONLY: never compiled (no Rust toolchain in the build image). Orphan-rule note: these impls
//! must live either in the gad crate behind a `cuda` feature (as `arrayfire` does, gad/src/lib.rs:515)
//! or on local newtypes of `Eval`/`Check`; INTEGRATION.md shows the in-crate form.
pub mod ffi;

use ffi::*;
use gad::prelude::*;
use gad::error::{Error, Result};
use std::marker::PhantomData;

/// f32 / f64 — the element types `AfAlgebra<T>` is instantiated with (gad/src/arrayfire.rs:64-87).
pub trait CuFloat: Copy + Into<f64> + 'static + Send + Sync { const DTYPE: i32; fn from_f64(v: f64) -> Self; }
impl CuFloat for f32 { const DTYPE: i32 = GADCU_F32; fn from_f64(v: f64) -> f32 { v as f32 } }
impl CuFloat for f64 { const DTYPE: i32 = GADCU_F64; fn from_f64(v: f64) -> f64 { v } }

/// Column-major Dim4 (af::Dim4 convention). `Default` is all ones = "reduce to a scalar" (array.rs:283).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct Dims(pub [u64; 4]);
impl Default for Dims { fn default() -> Self { Dims([1, 1, 1, 1]) } }
impl Dims { fn c(&self) -> gadcu_dim4 { gadcu_dim4 { d: self.0 } } }

/// Process-wide context + Eval algebra handle (one CUDA stream; `Eval` itself is a zero-sized type in gad).
pub struct Backend { pub ctx: *mut gadcu_ctx, pub eval: *mut gadcu_algebra }
unsafe impl Send for Backend {}
unsafe impl Sync for Backend {}
pub fn backend() -> &'static Backend {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut B: Option<Backend> = None;
    unsafe {
        INIT.call_once(|| {
            let mut ctx = std::ptr::null_mut();
            assert_eq!(gadcu_ctx_create(0, &mut ctx), GADCU_OK);
            let mut eval = std::ptr::null_mut();
            assert_eq!(gadcu_algebra_create(ctx, GADCU_EVAL, &mut eval), GADCU_OK);
            B = Some(Backend { ctx, eval });
        });
        B.as_ref().unwrap()
    }
}

/// Device-resident, immutable, refcounted array: `Clone` is O(1) (gad clones values into every VJP closure).
pub struct CuArray<T: CuFloat> { h: *mut gadcu_array, _t: PhantomData<T> }
unsafe impl<T: CuFloat> Send for CuArray<T> {}
unsafe impl<T: CuFloat> Sync for CuArray<T> {}
impl<T: CuFloat> Clone for CuArray<T> {
    fn clone(&self) -> Self { unsafe { gadcu_array_retain(self.h); } CuArray { h: self.h, _t: PhantomData } }
}
impl<T: CuFloat> Drop for CuArray<T> { fn drop(&mut self) { unsafe { gadcu_array_release(self.h); } } }
impl<T: CuFloat> CuArray<T> {
    pub fn new(values: &[T], dims: Dims) -> Self {
        let mut h = std::ptr::null_mut();
        let st = unsafe { gadcu_array_from_host(backend().ctx, T::DTYPE, &dims.c(), values.as_ptr() as *const _, &mut h) };
        assert_eq!(st, GADCU_OK);
        CuArray { h, _t: PhantomData }
    }
    pub fn host(&self, out: &mut [T]) { unsafe { gadcu_array_to_host(self.h, out.as_mut_ptr() as *mut _); } }
    fn from_value(v: gadcu_value) -> Self { CuArray { h: v.data, _t: PhantomData } }
    fn as_value(&self) -> gadcu_value { gadcu_value { data: self.h, arena: 0, index: 0 } }
}
impl<T: CuFloat> HasDims for CuArray<T> {
    type Dims = Dims;
    fn dims(&self) -> Dims { let mut d = gadcu_dim4 { d: [1; 4] }; unsafe { gadcu_array_dims(self.h, &mut d); } Dims(d.d) }
}
impl HasDims for Dims { type Dims = Dims; fn dims(&self) -> Dims { *self } }

/// Status code -> gad's own `Error` variant (error.rs:9-40); dims payload comes from gadcu_last_error().
fn to_result(st: gadcu_status, name: &'static str) -> Result<()> {
    match st {
        GADCU_OK => Ok(()),
        GADCU_ERR_EMPTY => Err(Error::empty(name)),
        GADCU_ERR_MISSING_ID | GADCU_ERR_MISSING_GRADIENT | GADCU_ERR_MISSING_NODE => Err(Error::missing_gradient(name)),
        _ => Err(Error::dimensions(name, unsafe { std::ffi::CStr::from_ptr(gadcu_last_error()) }.to_string_lossy())),
    }
}

macro_rules! binary { ($name:ident, $ffi:ident) => {
    fn $name(&mut self, v0: &CuArray<T>, v1: &CuArray<T>) -> Result<CuArray<T>> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        // the dimension check of `Check` runs inside the call, on the host, before any launch (arith.rs:66)
        to_result(unsafe { $ffi(backend().eval, &v0.as_value(), &v1.as_value(), &mut out) }, stringify!($name))?;
        Ok(CuArray::from_value(out))
    }
}}
macro_rules! unary { ($name:ident, $ffi:ident) => {
    fn $name(&mut self, v: &CuArray<T>) -> CuArray<T> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        let st = unsafe { $ffi(backend().eval, &v.as_value(), &mut out) };
        debug_assert_eq!(st, GADCU_OK);
        CuArray::from_value(out)
    }
}}
macro_rules! check_equal { ($name:ident) => {
    fn $name(&mut self, v0: &Dims, v1: &Dims) -> Result<Dims> {
        let (a, b) = (v0.c(), v1.c());
        let ptrs = [&a as *const _, &b as *const _];
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_equal(ptrs.as_ptr(), 2, &mut out) }, stringify!($name))?;
        Ok(Dims(out.d))
    }
}}

// ---- CoreAlgebra (core.rs:162-183) ----------------------------------------------------------------
impl<T: CuFloat> CoreAlgebra<CuArray<T>> for Eval {
    type Value = CuArray<T>;
    fn variable(&mut self, a: CuArray<T>) -> CuArray<T> { a }
    fn constant(&mut self, a: CuArray<T>) -> CuArray<T> { a }
    binary!(add, gadcu_add);
}
impl<T: CuFloat> CoreAlgebra<CuArray<T>> for Check {
    type Value = Dims;
    fn variable(&mut self, a: CuArray<T>) -> Dims { a.dims() }
    fn constant(&mut self, a: CuArray<T>) -> Dims { a.dims() }
    check_equal!(add);
}
// ---- ArithAlgebra (arith.rs:40-102) ---------------------------------------------------------------
impl<T: CuFloat> ArithAlgebra<CuArray<T>> for Eval {
    unary!(zeros, gadcu_zeros);
    unary!(ones, gadcu_ones);
    unary!(neg, gadcu_neg);
    binary!(sub, gadcu_sub);
    binary!(mul, gadcu_mul);
}
impl ArithAlgebra<Dims> for Check {
    fn zeros(&mut self, v: &Dims) -> Dims { *v }
    fn ones(&mut self, v: &Dims) -> Dims { *v }
    fn neg(&mut self, v: &Dims) -> Dims { *v }
    check_equal!(sub);
    check_equal!(mul);
}
// ---- AnalyticAlgebra (analytic.rs:64-128) ---------------------------------------------------------
impl<T: CuFloat> AnalyticAlgebra<CuArray<T>> for Eval {
    unary!(exp, gadcu_exp);
    unary!(log, gadcu_log);
    unary!(log1p, gadcu_log1p);
    unary!(sin, gadcu_sin);
    unary!(cos, gadcu_cos);
    unary!(tanh, gadcu_tanh);
    unary!(sigmoid, gadcu_sigmoid);
    unary!(reciprocal, gadcu_reciprocal);
    unary!(sqrt, gadcu_sqrt);
    binary!(div, gadcu_div);
    binary!(pow, gadcu_pow);
}
// ---- ConstArithAlgebra<_, T> and <_, i16> (const_arith.rs:33-56) ------------------------------------
macro_rules! constop { ($name:ident, $ffi:ident, $is_int:expr) => {
    fn $name(&mut self, v: &CuArray<T>, c: Self::C) -> CuArray<T> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        unsafe { $ffi(backend().eval, &v.as_value(), c.into(), $is_int, &mut out); }
        CuArray::from_value(out)
    }
}}
// (impl blocks for C = T and C = i16 follow the same pattern: setc/addc/mulc/powc -> gadcu_setc/addc/mulc/powc)

// ---- ArrayAlgebra (array.rs:57-127) ---------------------------------------------------------------
impl<T: CuFloat> ArrayAlgebra<CuArray<T>> for Eval {
    type Dims = Dims;
    type Scalar = T;
    unary!(flat, gadcu_flat);
    fn moddims(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_moddims(backend().eval, &v.as_value(), &d.c(), &mut out) }, "moddims")?;
        Ok(CuArray::from_value(out))
    }
    fn tile_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_tile_as(backend().eval, &v.as_value(), &d.c(), &mut out) }, "tile_as")?;
        Ok(CuArray::from_value(out))
    }
    fn sum_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_sum_as(backend().eval, &v.as_value(), &d.c(), &mut out) }, "sum_as")?;
        Ok(CuArray::from_value(out))
    }
    fn constant_as(&mut self, s: &T, d: Dims) -> CuArray<T> {
        let (mut sc, mut out) = (std::ptr::null_mut(), gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 });
        unsafe {
            gadcu_scalar_from_host(backend().ctx, T::DTYPE, (*s).into(), &mut sc);
            gadcu_constant_as(backend().eval, &gadcu_value { data: sc, arena: 0, index: 0 }, &d.c(), &mut out);
            gadcu_array_release(sc);
        }
        CuArray::from_value(out)
    }
    fn as_scalar(&mut self, v: &CuArray<T>) -> Result<T> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_as_scalar(backend().eval, &v.as_value(), &mut out) }, "as_scalar")?;
        let mut x = 0f64;
        unsafe { gadcu_scalar_to_host(out.data, &mut x); gadcu_array_release(out.data); }  // the D2H sync point of array.rs:106-111
        Ok(T::from_f64(x))
    }
    fn scale(&mut self, lambda: &T, v: &CuArray<T>) -> CuArray<T> {
        let mut out = std::ptr::null_mut();
        unsafe { gadcu_array_scale(v.h, (*lambda).into(), &mut out); }
        CuArray { h: out, _t: PhantomData }
    }
    fn dot(&mut self, a: &CuArray<T>, b: &CuArray<T>) -> Result<T> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_dot(backend().eval, &a.as_value(), &b.as_value(), &mut out) }, "dot")?;
        let mut x = 0f64;
        unsafe { gadcu_scalar_to_host(out.data, &mut x); gadcu_array_release(out.data); }  // array.rs:119-126
        Ok(T::from_f64(x))
    }
}
// ---- MatrixAlgebra (matrix.rs:40-61) --------------------------------------------------------------
impl<T: CuFloat> MatrixAlgebra<CuArray<T>> for Eval {
    fn matmul(&mut self, a: &CuArray<T>, b: &CuArray<T>, p1: MatProp, p2: MatProp) -> Result<CuArray<T>> {
        let c = |p: MatProp| gadcu_matprop { transposed: p.transposed as u8, conjugated: p.conjugated as u8 };
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_matmul(backend().eval, &a.as_value(), &b.as_value(), c(p1), c(p2), &mut out) }, "matmul")?;
        Ok(CuArray::from_value(out))
    }
    fn transpose(&mut self, v: &CuArray<T>, conjugate: bool) -> Result<CuArray<T>> {
        let mut out = gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 };
        to_result(unsafe { gadcu_transpose(backend().eval, &v.as_value(), conjugate as i32, &mut out) }, "transpose")?;
        Ok(CuArray::from_value(out))
    }
}
// CompareAlgebra::select_argmax and ArrayCompareAlgebra::{max_as, argmax_as, softmax_as} follow the same
// pattern over gadcu_select_argmax / gadcu_max_as / gadcu_argmax_as / gadcu_softmax_as; the `Check` impls call
// gadcu_check_{select_argmax, moddims, tile_as, sum_as, as_scalar, dot, reduced, transpose, matmul}.

/// WeightOps (net.rs:161-180): the only in-place mutation in gad.
impl<T: CuFloat> CuArray<T> {
    pub fn add_assign(&mut self, other: &CuArray<T>) -> Result<()> { to_result(unsafe { gadcu_array_add_assign(self.h, other.h) }, "add_assign") }
}

/// The native device tape behind the same trait surface: `CuGraph1` owns a GADCU_GRAPH1 algebra; its values are
/// `gadcu_value`s (data + id). `evaluate_gradients_once` runs the fused reverse sweep inside libgadcu.
pub struct CuGraph1 { pub h: *mut gadcu_algebra }
impl CuGraph1 {
    pub fn new() -> Self { let mut h = std::ptr::null_mut(); unsafe { gadcu_algebra_create(backend().ctx, GADCU_GRAPH1, &mut h); } CuGraph1 { h } }
}
impl Drop for CuGraph1 { fn drop(&mut self) { unsafe { gadcu_algebra_destroy(self.h); } } }
