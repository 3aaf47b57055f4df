//! gad-cuda: `CuArray<T>` + `impl XAlgebra<CuArray<T>> for gad::{Eval, Check}`.
//!
//! With these impls in scope, gad's `Graph1`/`GraphN` exist for `CuArray<T>` through gad's own generic graph impls
//! (`Graph<C>: XAlgebra<Value<D>>` whenever `Eval: XAlgebra<D>`, e.g. gad/src/arith.rs:153-234) — exactly how
//! gad/tests/symbolic.rs:40-93 derives `SymGraph1` from `SymEval`. A gad program generic over `A: CuAlgebra<T>` (the
//! analogue of `AfAlgebra<T>`, gad/src/arrayfire.rs:9-37) switches backends by changing the array type parameter only.
//!
//! `CuGraph1` / `CuGraphN` (bottom) additionally expose the NATIVE device tape (fused VJP kernels, deferred elementwise
//! regions, device store, data-parallel sweep) for callers who want the sweep itself to run inside libgadcu.
//!
//! SOURCE ONLY: never compiled (no Rust toolchain in the build image); `tests/test_check_and_abi.py` keeps `ffi.rs` in
//! step with `include/gadcu.h`. Orphan-rule note: these impls must live either in the gad crate behind a `cuda` feature
//! (as `arrayfire` does, gad/src/lib.rs:515) or on local newtypes of `Eval`/`Check`; INTEGRATION.md shows the in-crate form.
pub mod ffi;

use ffi::*;
use gad::error::{Error, Result};
use gad::prelude::*;
use std::marker::PhantomData;

/// f32 / f64 — the element types `AfAlgebra<T>` is instantiated with (gad/src/arrayfire.rs:64-87).
pub trait CuFloat: Copy + Default + Into<f64> + From<i16> + 'static + Send + Sync {
    const DTYPE: i32;
    fn from_f64(v: f64) -> Self;
}
impl CuFloat for f32 {
    const DTYPE: i32 = GADCU_F32;
    fn from_f64(v: f64) -> f32 { v as f32 }
}
impl CuFloat for f64 {
    const DTYPE: i32 = GADCU_F64;
    fn from_f64(v: f64) -> f64 { v }
}

/// Column-major Dim4 (af::Dim4 convention). `Default` is all ones = "reduce to a scalar" (array.rs:283).
#[derive(Clone, Copy, PartialEq, Eq, Debug, serde::Serialize, serde::Deserialize)]
pub struct Dims(pub [u64; 4]);
impl Default for Dims {
    fn default() -> Self { Dims([1, 1, 1, 1]) }
}
impl Dims {
    fn c(&self) -> gadcu_dim4 { gadcu_dim4 { d: self.0 } }
    pub fn elements(&self) -> u64 { self.0.iter().product() }
}

/// Process-wide context + Eval algebra handle (`Eval` itself is a zero-sized type in gad, so the device context cannot
/// travel inside it). libgadcu serialises API calls on one context, so sharing it across threads is sound.
pub struct Backend {
    pub ctx: *mut gadcu_ctx,
    pub eval: *mut gadcu_algebra,
}
unsafe impl Send for Backend {}
unsafe impl Sync for Backend {}
pub fn backend() -> &'static Backend {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut B: Option<Backend> = None;
    unsafe {
        INIT.call_once(|| {
            let mut ctx = std::ptr::null_mut();
            assert_eq!(gadcu_ctx_create(0, &mut ctx), GADCU_OK, "no B200 / libgadcu: there is no CPU fallback");
            let mut eval = std::ptr::null_mut();
            assert_eq!(gadcu_algebra_create(ctx, GADCU_EVAL, &mut eval), GADCU_OK);
            B = Some(Backend { ctx, eval });
        });
        B.as_ref().unwrap()
    }
}

fn null_value() -> gadcu_value { gadcu_value { data: std::ptr::null_mut(), arena: 0, index: 0 } }
fn last_error() -> String { unsafe { std::ffi::CStr::from_ptr(gadcu_last_error()) }.to_string_lossy().into_owned() }

/// Status code -> gad's own `Error` variant (error.rs:9-40, same order); the payload (dims / lengths) is the text
/// libgadcu recorded. Device failures have no variant in gad and are not recoverable: they panic with the message.
fn to_result(st: gadcu_status, name: &str) -> Result<()> {
    match st {
        GADCU_OK => Ok(()),
        GADCU_ERR_DIMENSIONS | GADCU_ERR_TYPE => Err(Error::dimensions(name, last_error())),
        GADCU_ERR_REDUCED_DIMENSIONS => Err(Error::reduced_dimensions(name, last_error(), String::new())),
        GADCU_ERR_LENGTHS => Err(Error::lengths(name, last_error())),
        GADCU_ERR_EMPTY => Err(Error::empty(name)),
        GADCU_ERR_MISSING_ID => Err(Error::missing_id(name)),
        GADCU_ERR_MISSING_GRADIENT => Err(Error::missing_gradient(name)),
        GADCU_ERR_MISSING_NODE => Err(Error::missing_node(name)),
        _ => panic!("libgadcu: {} ({})", last_error(), name),
    }
}

/// Device-resident, immutable, refcounted array: `Clone` is O(1) (gad clones values into every VJP closure).
pub struct CuArray<T: CuFloat> {
    h: *mut gadcu_array,
    _t: PhantomData<T>,
}
unsafe impl<T: CuFloat> Send for CuArray<T> {}
unsafe impl<T: CuFloat> Sync for CuArray<T> {}
impl<T: CuFloat> Clone for CuArray<T> {
    /// af::Array::clone: a NEW handle on the same buffer, so that `+=` on one clone (WeightOps::add_assign,
    /// net.rs:171-175, copy-on-write in the backend) leaves the others alone — get_weights() snapshots stay snapshots.
    fn clone(&self) -> Self {
        let mut h = std::ptr::null_mut();
        assert_eq!(unsafe { gadcu_array_clone(self.h, &mut h) }, GADCU_OK);
        CuArray { h, _t: PhantomData }
    }
}
impl<T: CuFloat> Drop for CuArray<T> {
    fn drop(&mut self) { unsafe { gadcu_array_release(self.h); } }
}
impl<T: CuFloat> CuArray<T> {
    /// af::Array::new: host values in column-major order.
    pub fn new(values: &[T], dims: Dims) -> Self {
        assert_eq!(values.len() as u64, dims.elements());
        let mut h = std::ptr::null_mut();
        let st = unsafe { gadcu_array_from_host(backend().ctx, T::DTYPE, &dims.c(), values.as_ptr() as *const _, &mut h) };
        assert_eq!(st, GADCU_OK, "{}", last_error());
        CuArray { h, _t: PhantomData }
    }
    /// af::Array::host (synchronises).
    pub fn host(&self, out: &mut [T]) {
        assert_eq!(out.len() as u64, self.dims().elements());
        unsafe { gadcu_array_to_host(self.h, out.as_mut_ptr() as *mut _); }
    }
    /// Queue the device -> host copy into `out` (pinned memory that outlives the copy) behind the work that produces this
    /// array and return a ticket; the bytes are there after `wait_download(ticket)`. Unlike `host`, the caller is not held
    /// up: a training loop reads step i's loss while step i + 1 is already issued.
    ///
    /// # Safety
    /// `out` must stay valid and untouched until `wait_download` has returned for the ticket.
    pub unsafe fn download_async(&self, out: *mut T) -> u64 {
        let mut ticket = 0u64;
        let st = gadcu_array_download_async(self.h, out as *mut _, &mut ticket);
        assert_eq!(st, GADCU_OK, "{}", last_error());
        ticket
    }
    fn from_raw(h: *mut gadcu_array) -> Self { CuArray { h, _t: PhantomData } }
    fn from_value(v: gadcu_value) -> Self { CuArray { h: v.data, _t: PhantomData } }
    fn as_value(&self) -> gadcu_value { gadcu_value { data: self.h, arena: 0, index: 0 } }
    fn scalar(s: T) -> *mut gadcu_array {
        let mut sc = std::ptr::null_mut();
        unsafe { gadcu_scalar_from_host(backend().ctx, T::DTYPE, s.into(), &mut sc); }
        sc
    }
}
impl<T: CuFloat> HasDims for CuArray<T> {
    type Dims = Dims;
    fn dims(&self) -> Dims {
        let mut d = gadcu_dim4 { d: [1; 4] };
        unsafe { gadcu_array_dims(self.h, &mut d); }
        Dims(d.d)
    }
}
impl HasDims for Dims {
    type Dims = Dims;
    fn dims(&self) -> Dims { *self }
}
// net.rs:143-160
impl<T: CuFloat> HasGradientId for CuArray<T> {
    type GradientId = ();
    fn gid(&self) -> Result<()> { Ok(()) }
}
impl HasGradientId for Dims {
    type GradientId = ();
    fn gid(&self) -> Result<()> { Ok(()) }
}

// ---- the three call shapes of the C ABI ------------------------------------------------------------------
// The dimension check of `Check` runs inside every call, on the host, before any launch (arith.rs:66, matrix.rs:52).
fn call1<T: CuFloat>(f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, *mut gadcu_value) -> gadcu_status,
                     v: &CuArray<T>, name: &str) -> Result<CuArray<T>> {
    let mut out = null_value();
    to_result(unsafe { f(backend().eval, &v.as_value(), &mut out) }, name)?;
    Ok(CuArray::from_value(out))
}
fn call2<T: CuFloat>(f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, *const gadcu_value, *mut gadcu_value) -> gadcu_status,
                     a: &CuArray<T>, b: &CuArray<T>, name: &str) -> Result<CuArray<T>> {
    let mut out = null_value();
    to_result(unsafe { f(backend().eval, &a.as_value(), &b.as_value(), &mut out) }, name)?;
    Ok(CuArray::from_value(out))
}
fn call_dims<T: CuFloat>(f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, *const gadcu_dim4, *mut gadcu_value) -> gadcu_status,
                         v: &CuArray<T>, d: Dims, name: &str) -> Result<CuArray<T>> {
    let mut out = null_value();
    to_result(unsafe { f(backend().eval, &v.as_value(), &d.c(), &mut out) }, name)?;
    Ok(CuArray::from_value(out))
}
fn check_equal(name: &str, dims: &[&Dims]) -> Result<Dims> {
    let c: Vec<gadcu_dim4> = dims.iter().map(|d| d.c()).collect();
    let ptrs: Vec<*const gadcu_dim4> = c.iter().map(|d| d as *const _).collect();
    let mut out = gadcu_dim4 { d: [1; 4] };
    to_result(unsafe { gadcu_check_equal(ptrs.as_ptr(), ptrs.len(), &mut out) }, name)?;
    Ok(Dims(out.d))
}

// ---- CoreAlgebra (core.rs:138-183) -----------------------------------------------------------------------
impl<T: CuFloat> CoreAlgebra<CuArray<T>> for Eval {
    type Value = CuArray<T>;
    fn variable(&mut self, a: CuArray<T>) -> CuArray<T> { a }
    fn constant(&mut self, a: CuArray<T>) -> CuArray<T> { a }
    fn add(&mut self, v0: &CuArray<T>, v1: &CuArray<T>) -> Result<CuArray<T>> { call2(gadcu_add, v0, v1, "add") }
}
impl<T: CuFloat> CoreAlgebra<CuArray<T>> for Check {
    type Value = Dims;
    fn variable(&mut self, a: CuArray<T>) -> Dims { a.dims() }
    fn constant(&mut self, a: CuArray<T>) -> Dims { a.dims() }
    fn add(&mut self, v0: &Dims, v1: &Dims) -> Result<Dims> { check_equal("add", &[v0, v1]) }
    fn add_all(&mut self, values: &[&Dims]) -> Result<Dims> { check_equal("add_all", values) }
}

// ---- ArithAlgebra (arith.rs:40-102) ----------------------------------------------------------------------
impl<T: CuFloat> ArithAlgebra<CuArray<T>> for Eval {
    fn zeros(&mut self, v: &CuArray<T>) -> CuArray<T> { call1(gadcu_zeros, v, "zeros").expect("zeros cannot fail") }
    fn ones(&mut self, v: &CuArray<T>) -> CuArray<T> { call1(gadcu_ones, v, "ones").expect("ones cannot fail") }
    fn neg(&mut self, v: &CuArray<T>) -> CuArray<T> { call1(gadcu_neg, v, "neg").expect("neg cannot fail") }
    fn sub(&mut self, v0: &CuArray<T>, v1: &CuArray<T>) -> Result<CuArray<T>> { call2(gadcu_sub, v0, v1, "sub") }
    fn mul(&mut self, v0: &CuArray<T>, v1: &CuArray<T>) -> Result<CuArray<T>> { call2(gadcu_mul, v0, v1, "mul") }
}
impl ArithAlgebra<Dims> for Check {
    fn zeros(&mut self, v: &Dims) -> Dims { *v }
    fn ones(&mut self, v: &Dims) -> Dims { *v }
    fn neg(&mut self, v: &Dims) -> Dims { *v }
    fn sub(&mut self, v0: &Dims, v1: &Dims) -> Result<Dims> { check_equal("sub", &[v0, v1]) }
    fn mul(&mut self, v0: &Dims, v1: &Dims) -> Result<Dims> { check_equal("mul", &[v0, v1]) }
}

// ---- ConstArithAlgebra<_, T> and <_, i16> (const_arith.rs:33-79): C = T or i16 via `T: From<C>` -----------------
impl<T: CuFloat + From<C>, C: 'static> ConstArithAlgebra<CuArray<T>, C> for Eval {
    fn setc(&mut self, v: &CuArray<T>, c: C) -> CuArray<T> { constop(gadcu_setc, v, c) }
    fn addc(&mut self, v: &CuArray<T>, c: C) -> CuArray<T> { constop(gadcu_addc, v, c) }
    fn mulc(&mut self, v: &CuArray<T>, c: C) -> CuArray<T> { constop(gadcu_mulc, v, c) }
    fn powc(&mut self, v: &CuArray<T>, c: C) -> CuArray<T> { constop(gadcu_powc, v, c) }
}
fn constop<T: CuFloat + From<C>, C: 'static>(
    f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, f64, std::os::raw::c_int, *mut gadcu_value) -> gadcu_status,
    v: &CuArray<T>, c: C) -> CuArray<T> {
    // an i16 constant keeps the integer-power kernel (repeated squaring, like num::pow::Pow<i16>); a T constant uses powf
    let is_int = (std::any::TypeId::of::<C>() == std::any::TypeId::of::<i16>()) as std::os::raw::c_int;
    let mut out = null_value();
    let st = unsafe { f(backend().eval, &v.as_value(), T::from(c).into(), is_int, &mut out) };
    assert_eq!(st, GADCU_OK, "{}", last_error());
    CuArray::from_value(out)
}
impl<C> ConstArithAlgebra<Dims, C> for Check {
    fn setc(&mut self, v: &Dims, _c: C) -> Dims { *v }
    fn addc(&mut self, v: &Dims, _c: C) -> Dims { *v }
    fn mulc(&mut self, v: &Dims, _c: C) -> Dims { *v }
    fn powc(&mut self, v: &Dims, _c: C) -> Dims { *v }
}

// ---- AnalyticAlgebra (analytic.rs:64-186) ----------------------------------------------------------------
macro_rules! unary_infallible {
    ($($name:ident => $ffi:ident),*) => { $(
        fn $name(&mut self, v: &CuArray<T>) -> CuArray<T> { call1($ffi, v, stringify!($name)).expect("elementwise op cannot fail") }
    )* }
}
impl<T: CuFloat> AnalyticAlgebra<CuArray<T>> for Eval {
    unary_infallible!(exp => gadcu_exp, log => gadcu_log, log1p => gadcu_log1p, sin => gadcu_sin, cos => gadcu_cos, tanh => gadcu_tanh,
                      sigmoid => gadcu_sigmoid, reciprocal => gadcu_reciprocal, sqrt => gadcu_sqrt);
    fn div(&mut self, v0: &CuArray<T>, v1: &CuArray<T>) -> Result<CuArray<T>> { call2(gadcu_div, v0, v1, "div") }
    fn pow(&mut self, v: &CuArray<T>, p: &CuArray<T>) -> Result<CuArray<T>> { call2(gadcu_pow, v, p, "pow") }
}
impl AnalyticAlgebra<Dims> for Check {
    fn exp(&mut self, v: &Dims) -> Dims { *v }
    fn log(&mut self, v: &Dims) -> Dims { *v }
    fn log1p(&mut self, v: &Dims) -> Dims { *v }
    fn sin(&mut self, v: &Dims) -> Dims { *v }
    fn cos(&mut self, v: &Dims) -> Dims { *v }
    fn tanh(&mut self, v: &Dims) -> Dims { *v }
    fn sigmoid(&mut self, v: &Dims) -> Dims { *v }
    fn reciprocal(&mut self, v: &Dims) -> Dims { *v }
    fn sqrt(&mut self, v: &Dims) -> Dims { *v }
    fn div(&mut self, v0: &Dims, v1: &Dims) -> Result<Dims> { check_equal("div", &[v0, v1]) }
    fn pow(&mut self, v: &Dims, p: &Dims) -> Result<Dims> { check_equal("pow", &[v, p]) }
}

// ---- CompareAlgebra (compare.rs:15-146): select_argmax is the one primitive; min/max/abs/sign/relu derive from it ------
impl<T: CuFloat> CompareAlgebra<CuArray<T>> for Eval {
    fn select_argmax(&mut self, v0: &CuArray<T>, v1: &CuArray<T>, r0: Option<&CuArray<T>>, r1: Option<&CuArray<T>>) -> Result<CuArray<T>> {
        let (a0, a1) = (r0.map(|r| r.as_value()), r1.map(|r| r.as_value()));
        let p = |o: &Option<gadcu_value>| o.as_ref().map_or(std::ptr::null(), |v| v as *const gadcu_value);
        let mut out = null_value();
        to_result(unsafe { gadcu_select_argmax(backend().eval, &v0.as_value(), &v1.as_value(), p(&a0), p(&a1), &mut out) }, "select_argmax")?;
        Ok(CuArray::from_value(out))
    }
}
impl CompareAlgebra<Dims> for Check {
    fn select_argmax(&mut self, v0: &Dims, v1: &Dims, r0: Option<&Dims>, r1: Option<&Dims>) -> Result<Dims> {
        let (c0, c1, q0, q1) = (v0.c(), v1.c(), r0.map(|d| d.c()), r1.map(|d| d.c()));
        let p = |o: &Option<gadcu_dim4>| o.as_ref().map_or(std::ptr::null(), |d| d as *const gadcu_dim4);
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_select_argmax(&c0, &c1, p(&q0), p(&q1), &mut out) }, "select_argmax")?;
        Ok(Dims(out.d))
    }
}

// ---- ArrayAlgebra (array.rs:57-195) ----------------------------------------------------------------------
impl<T: CuFloat> ArrayAlgebra<CuArray<T>> for Eval {
    type Dims = Dims;
    type Scalar = T;
    fn flat(&mut self, v: &CuArray<T>) -> CuArray<T> { call1(gadcu_flat, v, "flat").expect("flat cannot fail") }
    fn moddims(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> { call_dims(gadcu_moddims, v, d, "moddims") }
    fn tile_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> { call_dims(gadcu_tile_as, v, d, "tile_as") }
    fn sum_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> { call_dims(gadcu_sum_as, v, d, "sum_as") }
    fn constant_as(&mut self, s: &T, d: Dims) -> CuArray<T> {
        let (sc, mut out) = (CuArray::<T>::scalar(*s), null_value());
        unsafe {
            gadcu_constant_as(backend().eval, &gadcu_value { data: sc, arena: 0, index: 0 }, &d.c(), &mut out);
            gadcu_array_release(sc);
        }
        CuArray::from_value(out)
    }
    fn as_scalar(&mut self, v: &CuArray<T>) -> Result<T> {
        let mut out = null_value();
        to_result(unsafe { gadcu_as_scalar(backend().eval, &v.as_value(), &mut out) }, "as_scalar")?;
        let mut x = 0f64;
        unsafe { gadcu_scalar_to_host(out.data, &mut x); gadcu_array_release(out.data); } // the D2H sync point of array.rs:106-111
        Ok(T::from_f64(x))
    }
    fn scale(&mut self, lambda: &T, v: &CuArray<T>) -> CuArray<T> {
        let mut out = std::ptr::null_mut();
        unsafe { gadcu_array_scale(v.h, (*lambda).into(), &mut out); }
        CuArray::from_raw(out)
    }
    fn dot(&mut self, a: &CuArray<T>, b: &CuArray<T>) -> Result<T> {
        let mut out = null_value();
        to_result(unsafe { gadcu_dot(backend().eval, &a.as_value(), &b.as_value(), &mut out) }, "dot")?;
        let mut x = 0f64;
        unsafe { gadcu_scalar_to_host(out.data, &mut x); gadcu_array_release(out.data); } // array.rs:119-126
        Ok(T::from_f64(x))
    }
}
impl ArrayAlgebra<Dims> for Check {
    type Dims = Dims;
    type Scalar = ();
    fn flat(&mut self, v: &Dims) -> Dims {
        let mut out = gadcu_dim4 { d: [1; 4] };
        unsafe { gadcu_check_flat(&v.c(), &mut out); }
        Dims(out.d)
    }
    fn moddims(&mut self, v: &Dims, d: Dims) -> Result<Dims> {
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_moddims(&v.c(), &d.c(), &mut out) }, "moddims")?;
        Ok(Dims(out.d))
    }
    fn tile_as(&mut self, v: &Dims, d: Dims) -> Result<Dims> {
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_tile_as(&v.c(), &d.c(), &mut out) }, "tile_as")?;
        Ok(Dims(out.d))
    }
    fn sum_as(&mut self, v: &Dims, d: Dims) -> Result<Dims> {
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_sum_as(&v.c(), &d.c(), &mut out) }, "sum_as")?;
        Ok(Dims(out.d))
    }
    fn constant_as(&mut self, _s: &(), d: Dims) -> Dims { d }
    fn as_scalar(&mut self, v: &Dims) -> Result<()> { to_result(unsafe { gadcu_check_as_scalar(&v.c()) }, "as_scalar") }
    fn scale(&mut self, _lambda: &(), v: &Dims) -> Dims { *v }
    fn dot(&mut self, a: &Dims, b: &Dims) -> Result<()> { to_result(unsafe { gadcu_check_dot(&a.c(), &b.c()) }, "dot") }
}

// ---- ArrayCompareAlgebra (array_compare.rs:30-82) --------------------------------------------------------
impl<T: CuFloat> ArrayCompareAlgebra<CuArray<T>> for Eval {
    fn max_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> { call_dims(gadcu_max_as, v, d, "max_as") }
    fn argmax_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> { call_dims(gadcu_argmax_as, v, d, "argmax_as") }
    fn softmax_as(&mut self, v: &CuArray<T>, d: Dims) -> Result<CuArray<T>> { call_dims(gadcu_softmax_as, v, d, "softmax_as") }
}
impl ArrayCompareAlgebra<Dims> for Check {
    // max_as reduces to `rdims`; argmax_as and softmax_as keep the input's shape (array_compare.rs:84-118)
    fn max_as(&mut self, v: &Dims, d: Dims) -> Result<Dims> { check_reduced(v, d, 0, "max_as") }
    fn argmax_as(&mut self, v: &Dims, d: Dims) -> Result<Dims> { check_reduced(v, d, 1, "argmax_as") }
    fn softmax_as(&mut self, v: &Dims, d: Dims) -> Result<Dims> { check_reduced(v, d, 1, "softmax_as") }
}
fn check_reduced(v: &Dims, d: Dims, keeps_shape: i32, name: &str) -> Result<Dims> {
    let mut out = gadcu_dim4 { d: [1; 4] };
    to_result(unsafe { gadcu_check_reduced(&v.c(), &d.c(), keeps_shape, &mut out) }, name)?;
    Ok(Dims(out.d))
}

// ---- MatrixAlgebra (matrix.rs:40-126) --------------------------------------------------------------------
fn matprop(p: MatProp) -> gadcu_matprop { gadcu_matprop { transposed: p.transposed as u8, conjugated: p.conjugated as u8 } }
impl<T: CuFloat> MatrixAlgebra<CuArray<T>> for Eval {
    fn matmul(&mut self, a: &CuArray<T>, b: &CuArray<T>, p1: MatProp, p2: MatProp) -> Result<CuArray<T>> {
        let mut out = null_value();
        to_result(unsafe { gadcu_matmul(backend().eval, &a.as_value(), &b.as_value(), matprop(p1), matprop(p2), &mut out) }, "matmul")?;
        Ok(CuArray::from_value(out))
    }
    fn transpose(&mut self, v: &CuArray<T>, conjugate: bool) -> Result<CuArray<T>> {
        let mut out = null_value();
        to_result(unsafe { gadcu_transpose(backend().eval, &v.as_value(), conjugate as i32, &mut out) }, "transpose")?;
        Ok(CuArray::from_value(out))
    }
}
impl MatrixAlgebra<Dims> for Check {
    fn matmul(&mut self, a: &Dims, b: &Dims, p1: MatProp, p2: MatProp) -> Result<Dims> {
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_matmul(&a.c(), &b.c(), matprop(p1), matprop(p2), &mut out) }, "matmul")?;
        Ok(Dims(out.d))
    }
    fn transpose(&mut self, v: &Dims, _conjugate: bool) -> Result<Dims> {
        let mut out = gadcu_dim4 { d: [1; 4] };
        to_result(unsafe { gadcu_check_transpose(&v.c(), &mut out) }, "transpose")?;
        Ok(Dims(out.d))
    }
}

// ---- WeightOps (net.rs:100-104,161-180): the only in-place mutation in gad; serde through a host copy ----------
impl<T: CuFloat> WeightOps<T> for CuArray<T>
where
    T: serde::Serialize + serde::de::DeserializeOwned,
{
    fn add_assign(&mut self, other: Self) -> Result<()> { to_result(unsafe { gadcu_array_add_assign(self.h, other.h) }, "add_assign") }
    fn scale(&self, lambda: T) -> Self {
        let mut out = std::ptr::null_mut();
        unsafe { gadcu_array_scale(self.h, lambda.into(), &mut out); }
        CuArray::from_raw(out)
    }
}
impl<T: CuFloat + serde::Serialize> serde::Serialize for CuArray<T> {
    fn serialize<S: serde::Serializer>(&self, s: S) -> std::result::Result<S::Ok, S::Error> {
        let dims = self.dims();
        let mut host = vec![T::default(); dims.elements() as usize];
        self.host(&mut host);
        (dims, host).serialize(s)
    }
}
impl<'de, T: CuFloat + serde::Deserialize<'de>> serde::Deserialize<'de> for CuArray<T> {
    fn deserialize<D: serde::Deserializer<'de>>(d: D) -> std::result::Result<Self, D::Error> {
        let (dims, host): (Dims, Vec<T>) = serde::Deserialize::deserialize(d)?;
        Ok(CuArray::new(&host, dims))
    }
}

/// The analogue of `AfAlgebra<T>` (arrayfire.rs:9-37) restricted to what `Eval`/`Check`/`Graph1`/`GraphN` provide for
/// `CuArray<T>`: programs written against this bound run on either backend.
pub trait CuAlgebra<T: CuFloat>:
    CoreAlgebra<CuArray<T>, Value = <Self as CuAlgebra<T>>::Value>
    + ArithAlgebra<<Self as CuAlgebra<T>>::Value>
    + ConstArithAlgebra<<Self as CuAlgebra<T>>::Value, T>
    + ConstArithAlgebra<<Self as CuAlgebra<T>>::Value, i16>
    + AnalyticAlgebra<<Self as CuAlgebra<T>>::Value>
    + CompareAlgebra<<Self as CuAlgebra<T>>::Value>
    + ArrayAlgebra<<Self as CuAlgebra<T>>::Value, Dims = Dims>
    + ArrayCompareAlgebra<<Self as CuAlgebra<T>>::Value>
    + MatrixAlgebra<<Self as CuAlgebra<T>>::Value>
{
    type Value: HasGradientId;
}
impl<T: CuFloat> CuAlgebra<T> for Eval {
    type Value = CuArray<T>;
}
impl<T: CuFloat> CuAlgebra<T> for Check {
    type Value = Dims;
}

// =========================================================================================================
// The native device tape behind the same method names: `CuGraph1` / `CuGraphN` own a GADCU_GRAPH1 / GRAPHN algebra;
// their values are `CuValue<T>` = gad's `Value<D>` (graph.rs:35-43: data + optional id). Every op is ONE FFI call that
// records the node natively; `evaluate_gradients_once` runs the fused reverse sweep (and, with a communicator, the
// data-parallel exchange) inside libgadcu.
// =========================================================================================================
pub struct CuValue<T: CuFloat> {
    v: gadcu_value,
    _t: PhantomData<T>,
}
impl<T: CuFloat> CuValue<T> {
    fn new(v: gadcu_value) -> Self { CuValue { v, _t: PhantomData } }
    /// graph.rs:329-331
    pub fn data(&self) -> CuArray<T> {
        unsafe { gadcu_array_retain(self.v.data); }
        CuArray::from_raw(self.v.data)
    }
    /// graph.rs:334-336: `None` for constants
    pub fn id(&self) -> Option<(u32, u32)> { if self.v.index == 0 { None } else { Some((self.v.arena, self.v.index)) } }
    /// net.rs HasGradientId for Value<D>: constants have no id
    pub fn gid(&self) -> Result<(u32, u32)> { self.id().ok_or_else(|| Error::missing_id("gid")) }
}
impl<T: CuFloat> Clone for CuValue<T> {
    fn clone(&self) -> Self {
        unsafe { gadcu_array_retain(self.v.data); }
        CuValue::new(self.v)
    }
}
impl<T: CuFloat> Drop for CuValue<T> {
    fn drop(&mut self) { unsafe { gadcu_value_release(&mut self.v); } }
}

/// Gradients of a sweep (store.rs:27-140). `get` is `GradientReader::read`: absent is `None`, never zero.
pub struct CuStore<T: CuFloat> {
    h: *mut gadcu_store,
    _t: PhantomData<T>,
}
impl<T: CuFloat> CuStore<T> {
    pub fn get(&self, id: (u32, u32)) -> Option<CuValue<T>> {
        let (mut out, mut found) = (null_value(), 0);
        let st = unsafe { gadcu_store_get(self.h, id.0, id.1, &mut out, &mut found) };
        if st == GADCU_OK && found != 0 { Some(CuValue::new(out)) } else { None }
    }
}
impl<T: CuFloat> Drop for CuStore<T> {
    fn drop(&mut self) { unsafe { gadcu_store_destroy(self.h); } }
}

macro_rules! native_graph {
    ($name:ident, $kind:expr) => {
        pub struct $name { pub h: *mut gadcu_algebra }
        unsafe impl Send for $name {}
        unsafe impl Sync for $name {}
        impl $name {
            pub fn new() -> Self {
                let mut h = std::ptr::null_mut();
                assert_eq!(unsafe { gadcu_algebra_create(backend().ctx, $kind, &mut h) }, GADCU_OK);
                $name { h }
            }
            pub fn variable<T: CuFloat>(&mut self, data: CuArray<T>) -> CuValue<T> {
                let mut out = null_value();
                unsafe { gadcu_variable(self.h, data.h, &mut out); }
                CuValue::new(out)
            }
            pub fn constant<T: CuFloat>(&mut self, data: CuArray<T>) -> CuValue<T> {
                let mut out = null_value();
                unsafe { gadcu_constant(self.h, data.h, &mut out); }
                CuValue::new(out)
            }
            fn op1<T: CuFloat>(&mut self, f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, *mut gadcu_value) -> gadcu_status,
                               v: &CuValue<T>, name: &str) -> Result<CuValue<T>> {
                let mut out = null_value();
                to_result(unsafe { f(self.h, &v.v, &mut out) }, name)?;
                Ok(CuValue::new(out))
            }
            fn op2<T: CuFloat>(&mut self, f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, *const gadcu_value, *mut gadcu_value) -> gadcu_status,
                               a: &CuValue<T>, b: &CuValue<T>, name: &str) -> Result<CuValue<T>> {
                let mut out = null_value();
                to_result(unsafe { f(self.h, &a.v, &b.v, &mut out) }, name)?;
                Ok(CuValue::new(out))
            }
            fn opd<T: CuFloat>(&mut self, f: unsafe extern "C" fn(*mut gadcu_algebra, *const gadcu_value, *const gadcu_dim4, *mut gadcu_value) -> gadcu_status,
                               v: &CuValue<T>, d: Dims, name: &str) -> Result<CuValue<T>> {
                let mut out = null_value();
                to_result(unsafe { f(self.h, &v.v, &d.c(), &mut out) }, name)?;
                Ok(CuValue::new(out))
            }
            // the op entry points are algebra-polymorphic: the same functions the Eval impls above call, with a graph handle
            pub fn add<T: CuFloat>(&mut self, a: &CuValue<T>, b: &CuValue<T>) -> Result<CuValue<T>> { self.op2(gadcu_add, a, b, "add") }
            pub fn sub<T: CuFloat>(&mut self, a: &CuValue<T>, b: &CuValue<T>) -> Result<CuValue<T>> { self.op2(gadcu_sub, a, b, "sub") }
            pub fn mul<T: CuFloat>(&mut self, a: &CuValue<T>, b: &CuValue<T>) -> Result<CuValue<T>> { self.op2(gadcu_mul, a, b, "mul") }
            pub fn div<T: CuFloat>(&mut self, a: &CuValue<T>, b: &CuValue<T>) -> Result<CuValue<T>> { self.op2(gadcu_div, a, b, "div") }
            pub fn neg<T: CuFloat>(&mut self, v: &CuValue<T>) -> CuValue<T> { self.op1(gadcu_neg, v, "neg").expect("neg cannot fail") }
            pub fn exp<T: CuFloat>(&mut self, v: &CuValue<T>) -> CuValue<T> { self.op1(gadcu_exp, v, "exp").expect("exp cannot fail") }
            pub fn log<T: CuFloat>(&mut self, v: &CuValue<T>) -> CuValue<T> { self.op1(gadcu_log, v, "log").expect("log cannot fail") }
            pub fn tanh<T: CuFloat>(&mut self, v: &CuValue<T>) -> CuValue<T> { self.op1(gadcu_tanh, v, "tanh").expect("tanh cannot fail") }
            pub fn relu<T: CuFloat>(&mut self, v: &CuValue<T>) -> CuValue<T> { self.op1(gadcu_relu, v, "relu").expect("relu cannot fail") }
            pub fn norm2<T: CuFloat>(&mut self, v: &CuValue<T>) -> CuValue<T> { self.op1(gadcu_norm2, v, "norm2").expect("norm2 cannot fail") }
            pub fn as_scalar<T: CuFloat>(&mut self, v: &CuValue<T>) -> Result<CuValue<T>> { self.op1(gadcu_as_scalar, v, "as_scalar") }
            pub fn tile_as<T: CuFloat>(&mut self, v: &CuValue<T>, d: Dims) -> Result<CuValue<T>> { self.opd(gadcu_tile_as, v, d, "tile_as") }
            pub fn sum_as<T: CuFloat>(&mut self, v: &CuValue<T>, d: Dims) -> Result<CuValue<T>> { self.opd(gadcu_sum_as, v, d, "sum_as") }
            pub fn softmax_as<T: CuFloat>(&mut self, v: &CuValue<T>, d: Dims) -> Result<CuValue<T>> { self.opd(gadcu_softmax_as, v, d, "softmax_as") }
            pub fn matmul<T: CuFloat>(&mut self, a: &CuValue<T>, b: &CuValue<T>, p1: MatProp, p2: MatProp) -> Result<CuValue<T>> {
                let mut out = null_value();
                to_result(unsafe { gadcu_matmul(self.h, &a.v, &b.v, matprop(p1), matprop(p2), &mut out) }, "matmul")?;
                Ok(CuValue::new(out))
            }
            pub fn matmul_nn<T: CuFloat>(&mut self, a: &CuValue<T>, b: &CuValue<T>) -> Result<CuValue<T>> { self.op2(gadcu_matmul_nn, a, b, "matmul_nn") }
            pub fn num_nodes(&self) -> usize { unsafe { gadcu_algebra_num_nodes(self.h) } }
        }
        impl Clone for $name {
            /// graph.rs:353-360
            fn clone(&self) -> Self {
                let mut h = std::ptr::null_mut();
                assert_eq!(unsafe { gadcu_algebra_clone(self.h, &mut h) }, GADCU_OK);
                $name { h }
            }
        }
        impl Drop for $name {
            fn drop(&mut self) { unsafe { gadcu_algebra_destroy(self.h); } }
        }
    };
}
native_graph!(CuGraph1, GADCU_GRAPH1);
native_graph!(CuGraphN, GADCU_GRAPHN);

impl CuGraph1 {
    /// graph.rs:263-274 (`&self`: may run concurrently from many threads on one tape)
    pub fn evaluate_gradients<T: CuFloat>(&self, id: (u32, u32), seed: CuArray<T>) -> Result<CuStore<T>> { self.sweep(id, seed, 0, None) }
    /// graph.rs:277-290: consumes the nodes it visits, releasing captured forward buffers early
    pub fn evaluate_gradients_once<T: CuFloat>(&mut self, id: (u32, u32), seed: CuArray<T>) -> Result<CuStore<T>> { self.sweep(id, seed, 1, None) }
    /// One batch shard of a data-parallel step (net_ext.rs:47-73 sums per-example gradients): the weight gradients are
    /// summed over the ranks of `comm` while the sweep runs; `also` (e.g. the loss) is summed with them.
    pub fn evaluate_gradients_once_dp<T: CuFloat>(&mut self, id: (u32, u32), seed: CuArray<T>, comm: &CuComm, also: &[&CuArray<T>]) -> Result<CuStore<T>> {
        self.sweep(id, seed, 1, Some((comm, also)))
    }
    fn sweep<T: CuFloat>(&self, id: (u32, u32), seed: CuArray<T>, once: i32, dp: Option<(&CuComm, &[&CuArray<T>])>) -> Result<CuStore<T>> {
        let mut s = std::ptr::null_mut();
        let st = match dp {
            None => unsafe { gadcu_evaluate_gradients(self.h, id.0, id.1, seed.h, once, &mut s) },
            Some((comm, also)) => {
                let raw: Vec<*mut gadcu_array> = also.iter().map(|a| a.h).collect();
                unsafe { gadcu_evaluate_gradients_dp_with(self.h, id.0, id.1, seed.h, once, comm.h, raw.as_ptr(), raw.len(), &mut s) }
            }
        };
        to_result(st, "evaluate_gradients")?;
        Ok(CuStore { h: s, _t: PhantomData })
    }
}
impl CuGraphN {
    /// graph.rs:307-318: the backward pass is recorded on the tape, so the gradients can be differentiated again
    pub fn compute_gradients<T: CuFloat>(&mut self, id: (u32, u32), seed: &CuValue<T>) -> Result<CuStore<T>> {
        let mut s = std::ptr::null_mut();
        to_result(unsafe { gadcu_compute_gradients(self.h, id.0, id.1, &seed.v, &mut s) }, "compute_gradients")?;
        Ok(CuStore { h: s, _t: PhantomData })
    }
    /// The same sweep over one batch shard: every variable's gradient is summed over the ranks as soon as the sweep has
    /// completed it (overlapping the rest of the sweep) and comes back as a constant value.
    pub fn compute_gradients_dp<T: CuFloat>(&mut self, id: (u32, u32), seed: &CuValue<T>, comm: &CuComm) -> Result<CuStore<T>> {
        let mut s = std::ptr::null_mut();
        to_result(unsafe { gadcu_compute_gradients_dp(self.h, id.0, id.1, &seed.v, comm.h, &mut s) }, "compute_gradients_dp")?;
        Ok(CuStore { h: s, _t: PhantomData })
    }
}

/// One scalar `Graph1` tape per LANE (include/gadcu.h "batched scalar tapes"): what gad does by looping a scalar tape over
/// the examples (net_ext.rs:50-66) runs as one kernel, one lane per thread. Variables and constants are `[lanes]` columns;
/// every op only records (`Deref` to `CuGraph1`: the same methods); `evaluate_gradients_once` with a scalar seed compiles
/// forward + reverse sweep into one per-lane program and launches it. Values handed out are lazy (`data()` holds no
/// array): `force` computes a forward value's column; gradients come out of the store as `[lanes]` columns. The
/// derivation has a precedent in the reference: tests/symbolic.rs:93 builds a tape over another algebra the same way.
pub struct CuBatchedGraph1<T: CuFloat> {
    tape: CuGraph1,
    pub lanes: u64,
    _t: PhantomData<T>,
}
impl<T: CuFloat> CuBatchedGraph1<T> {
    pub fn new(lanes: u64) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        to_result(unsafe { gadcu_batched_create(backend().ctx, T::DTYPE, lanes, &mut h) }, "batched_create")?;
        Ok(CuBatchedGraph1 { tape: CuGraph1 { h }, lanes, _t: PhantomData })
    }
    /// The column of a forward value over all lanes.
    pub fn force(&mut self, v: &CuValue<T>) -> Result<CuArray<T>> {
        let mut out = std::ptr::null_mut();
        to_result(unsafe { gadcu_batched_force(self.tape.h, &v.v, &mut out) }, "batched_force")?;
        Ok(CuArray::from_raw(out))
    }
    /// `evaluate_gradients_once(id, T::one())` of every lane's tape at once (graph.rs:277-290).
    pub fn evaluate_gradients_once(&mut self, id: (u32, u32), seed: T) -> Result<CuStore<T>> {
        self.tape.evaluate_gradients_once(id, CuArray::from_raw(CuArray::<T>::scalar(seed)))
    }
}
impl<T: CuFloat> std::ops::Deref for CuBatchedGraph1<T> {
    type Target = CuGraph1;
    fn deref(&self) -> &CuGraph1 { &self.tape }
}
impl<T: CuFloat> std::ops::DerefMut for CuBatchedGraph1<T> {
    fn deref_mut(&mut self) -> &mut CuGraph1 { &mut self.tape }
}
impl<T: CuFloat> CuStore<T> {
    /// Batched sweeps only: the same tapes at new points — `columns[i]` replaces the data of the i-th variable (creation
    /// order) and the compiled lane program runs again, one launch, nothing re-recorded. The variables' gradient columns,
    /// `None` where the sweep produced none.
    pub fn rerun(&self, columns: &[&CuArray<T>]) -> Result<Vec<Option<CuArray<T>>>> {
        let raw: Vec<*mut gadcu_array> = columns.iter().map(|c| c.h).collect();
        let mut out: Vec<*mut gadcu_array> = vec![std::ptr::null_mut(); raw.len()];
        to_result(unsafe { gadcu_batched_rerun(self.h, raw.as_ptr(), raw.len(), out.as_mut_ptr()) }, "batched_rerun")?;
        Ok(out.into_iter().map(|h| if h.is_null() { None } else { Some(CuArray::from_raw(h)) }).collect())
    }
    /// (instructions, value slots) of the lane program behind a batched sweep.
    pub fn program_stats(&self) -> (u64, u64) {
        let (mut a, mut b) = (0u64, 0u64);
        unsafe { gadcu_batched_program_stats(self.h, &mut a, &mut b); }
        (a, b)
    }
}

/// Backend switches with no counterpart in gad (include/gadcu.h): fused GEMM epilogues (0 never, 1 where it pays, 2 always)
/// and automatic CUDA-graph replay keyed by the recorded kernel sequence (the tape is still rebuilt every step).
pub fn set_gemm_epilogue(mode: i32) { unsafe { gadcu_ctx_set_gemm_epilogue(backend().ctx, mode); } }
pub fn set_auto_replay(on: bool) { unsafe { gadcu_ctx_set_auto_replay(backend().ctx, on as i32); } }
pub fn flush() { unsafe { gadcu_ctx_flush(backend().ctx); } }
/// Block until the `CuArray::download_async` that returned `ticket` (and every earlier one) has landed in host memory.
pub fn wait_download(ticket: u64) {
    let st = unsafe { gadcu_ctx_wait_download(backend().ctx, ticket) };
    assert_eq!(st, GADCU_OK, "{}", last_error());
}
/// Batched sweeps (replayed from the structure table, walked node by node, structures held) since process start.
pub fn sweep_memo_stats() -> (u64, u64, u64) {
    let (mut a, mut b, mut c) = (0u64, 0u64, 0u64);
    unsafe { gadcu_batched_sweep_memo_stats(&mut a, &mut b, &mut c); }
    (a, b, c)
}

/// One communicator per rank (one process per GPU); the unique id travels through whatever launched the ranks.
pub struct CuComm { h: *mut gadcu_comm }
unsafe impl Send for CuComm {}
impl CuComm {
    pub fn unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        unsafe { gadcu_comm_unique_id(id.as_mut_ptr() as *mut _); }
        id
    }
    pub fn new(nranks: i32, rank: i32, id: &[u8; 128]) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        to_result(unsafe { gadcu_comm_init(backend().ctx, nranks, rank, id.as_ptr() as *const _, &mut h) }, "comm_init")?;
        Ok(CuComm { h })
    }
    pub fn allreduce_sum<T: CuFloat>(&self, arrays: &[&CuArray<T>]) -> Result<()> {
        let raw: Vec<*mut gadcu_array> = arrays.iter().map(|a| a.h).collect();
        to_result(unsafe { gadcu_comm_allreduce_sum(self.h, raw.as_ptr(), raw.len()) }, "allreduce_sum")
    }
}
impl Drop for CuComm {
    fn drop(&mut self) { unsafe { gadcu_comm_destroy(self.h); } }
}
