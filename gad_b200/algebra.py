"""Host-side mirror of gad's operator interface over the CUDA backend (through the C ABI only).

Names, argument meaning and error behaviour follow the reference traits so a gad program reads the
same here: `CoreAlgebra` (gad/src/core.rs:12-38), `ArithAlgebra` (arith.rs:14-32),
`ConstArithAlgebra` (const_arith.rs:14-26), `AnalyticAlgebra` (analytic.rs:16-56), `CompareAlgebra`
(compare.rs:15-66), `ArrayAlgebra` (array.rs:13-45), `MatrixAlgebra` (matrix.rs:20-32),
`ArrayCompareAlgebra` (array_compare.rs:16-22), and the four default algebras `Check`, `Eval`,
`Graph1`, `GraphN` (lib.rs:519-532). A program written against one algebra object runs unchanged on
the others — gad's "change only a type parameter".
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Dim4, MatProp as _MatPropC, ValueC

F32, F64 = 0, 1
_NP = {F32: np.float32, F64: np.float64}
_DT = {np.dtype(np.float32): F32, np.dtype(np.float64): F64}


# ---- errors: gad/src/error.rs:9-40 ---------------------------------------------------------------
class GadError(Exception):
    kind = "Error"

    def __init__(self, status: int, message: str = ""):
        super().__init__(f"{self.kind}({status}): {message}")
        self.status = status
        self.message = message


class DimensionsError(GadError):
    kind = "Dimensions"


class ReducedDimensionsError(GadError):
    kind = "ReducedDimensions"


class LengthsError(GadError):
    kind = "Lengths"


class EmptyError(GadError):
    kind = "Empty"


class MissingIdError(GadError):
    kind = "MissingId"


class MissingGradientError(GadError):
    kind = "MissingGradient"


class MissingNodeError(GadError):
    kind = "MissingNode"


class CudaError(GadError):
    kind = "Cuda"


class NcclError(GadError):
    kind = "Nccl"


class UnsupportedError(GadError):
    kind = "Unsupported"


class InvalidError(GadError):
    kind = "Invalid"


class TypeMismatchError(GadError):
    kind = "Type"


_ERRORS = {1: DimensionsError, 2: ReducedDimensionsError, 3: LengthsError, 4: EmptyError, 5: MissingIdError,
           6: MissingGradientError, 7: MissingNodeError, 8: CudaError, 9: NcclError, 10: UnsupportedError,
           11: InvalidError, 12: TypeMismatchError}


def _check(status: int) -> None:
    if status != 0:
        msg = _lib.load().gadcu_last_error()
        raise _ERRORS.get(status, GadError)(status, msg.decode() if msg else "")


class MatProp:
    """gad/src/matrix.rs:13-17, 202-221."""

    def __init__(self, transposed: bool = False, conjugated: bool = False):
        self.transposed, self.conjugated = bool(transposed), bool(conjugated)

    def transpose(self) -> "MatProp":
        return MatProp(not self.transposed, self.conjugated)

    def conjugate(self) -> "MatProp":
        return MatProp(self.transposed, not self.conjugated)

    def __eq__(self, other):
        return (self.transposed, self.conjugated) == (other.transposed, other.conjugated)

    def _c(self) -> _MatPropC:
        return _MatPropC(int(self.transposed), int(self.conjugated))


def _dims4(dims) -> tuple:
    if isinstance(dims, int):
        dims = (dims,)
    d = tuple(int(x) for x in dims)
    return d + (1,) * (4 - len(d))


# ---- context and arrays --------------------------------------------------------------------------
class Context:
    """Device, stream, caching allocator (gadcu_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _check(self.lib.gadcu_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device

    def synchronize(self) -> None:
        _check(self.lib.gadcu_ctx_synchronize(self.handle))

    def wait_download(self, ticket: int) -> None:
        """Block until the download_async that returned `ticket` (and every earlier one) has landed in host memory."""
        _check(self.lib.gadcu_ctx_wait_download(self.handle, C.c_uint64(ticket)))

    def wait_uploads(self) -> None:
        """The compute stream waits (on the device) for every upload_async issued so far."""
        _check(self.lib.gadcu_ctx_wait_uploads(self.handle))

    def set_fusion(self, fuse: bool) -> None:
        _check(self.lib.gadcu_ctx_set_fusion(self.handle, int(fuse)))

    def set_lazy(self, enable: bool = True, min_elements: int = 32768) -> None:
        """Deferred elementwise regions (gadcu_ctx_set_lazy): on by default for arrays of >= 32768 elements."""
        _check(self.lib.gadcu_ctx_set_lazy(self.handle, int(enable), int(min_elements)))

    def set_auto_replay(self, enable: bool) -> None:
        """Automatic CUDA-graph replay keyed by tape structure (gadcu_ctx_set_auto_replay): launches are recorded and run as
        one graph at the next read / synchronise / end of a reverse sweep, re-using the executable graph of the last step
        with the same kernel sequence. Off by default."""
        _check(self.lib.gadcu_ctx_set_auto_replay(self.handle, int(enable)))

    def flush(self) -> None:
        _check(self.lib.gadcu_ctx_flush(self.handle))

    def replay_stats(self) -> dict:
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.lib.gadcu_ctx_replay_stats(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"segments": a.value, "reused": b.value, "instantiated": c.value}

    def set_gemm_epilogue(self, enable) -> None:
        """Fused GEMM epilogues (gadcu_ctx_set_gemm_epilogue): 1 / True = where it pays (default), 2 = always, 0 / False = every
        product launched where it is recorded."""
        _check(self.lib.gadcu_ctx_set_gemm_epilogue(self.handle, int(enable)))

    def set_strict_reference(self, strict: bool) -> None:
        """True: reproduce the reference's matmul VJP for transposed operands as is."""
        _check(self.lib.gadcu_ctx_set_strict_reference(self.handle, int(strict)))

    @property
    def stream(self) -> int:
        return int(self.lib.gadcu_ctx_stream(self.handle) or 0)

    def stats(self) -> dict:
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.lib.gadcu_ctx_memory_stats(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"bytes_in_use": a.value, "bytes_reserved": b.value, "kernel_launches": c.value}

    def sweep_memo_stats(self) -> dict:
        """Batched sweeps served from the structure table / walked node by node / structures held (process-wide)."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self.lib.gadcu_batched_sweep_memo_stats(C.byref(a), C.byref(b), C.byref(c)))
        return {"replayed": a.value, "walked": b.value, "structures": c.value}

    def launches(self) -> int:
        return self.stats()["kernel_launches"]

    def profile_begin(self) -> None:
        _check(self.lib.gadcu_ctx_profile_begin(self.handle))

    def profile_end(self) -> dict:
        """ms and launch counts per kernel category since profile_begin (CUDA events on the stream)."""
        ms, cnt = (C.c_double * 4)(), (C.c_uint64 * 4)()
        _check(self.lib.gadcu_ctx_profile_end(self.handle, ms, cnt))
        names = ["matmul", "elementwise", "reduce", "tile_transpose"]
        return {n: {"ms": ms[i], "launches": int(cnt[i])} for i, n in enumerate(names)}

    # af::Array::new / af::constant / af::randu equivalents
    def array(self, data, dtype=None) -> "CuArray":
        arr = np.asarray(data)
        if dtype is not None:
            arr = arr.astype(dtype)
        if arr.dtype not in _DT:
            arr = arr.astype(np.float32)
        if arr.ndim > 4:
            raise ValueError("at most 4 dimensions")
        dims = _dims4(arr.shape if arr.ndim else (1,))
        flat = np.ascontiguousarray(arr.reshape(-1, order="F"))
        h = C.c_void_p()
        d = Dim4.of(dims)
        _check(self.lib.gadcu_array_from_host(self.handle, _DT[flat.dtype], C.byref(d), flat.ctypes.data_as(C.c_void_p), C.byref(h)))
        return CuArray(self, h)

    def scalar(self, value: float, dtype=F32) -> "CuArray":
        h = C.c_void_p()
        _check(self.lib.gadcu_scalar_from_host(self.handle, dtype, float(value), C.byref(h)))
        return CuArray(self, h)

    def random(self, dims, seed: int, dtype=F32, normal: bool = False, a: float = 0.0, b: float = 1.0) -> "CuArray":
        h = C.c_void_p()
        d = Dim4.of(_dims4(dims))
        _check(self.lib.gadcu_array_random(self.handle, dtype, C.byref(d), seed, int(normal), a, b, C.byref(h)))
        return CuArray(self, h)

    def empty(self, dims, dtype=F32) -> "CuArray":
        h = C.c_void_p()
        d = Dim4.of(_dims4(dims))
        _check(self.lib.gadcu_array_alloc(self.handle, dtype, C.byref(d), C.byref(h)))
        return CuArray(self, h)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.gadcu_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class CuArray:
    """A refcounted, immutable, column-major device array (gadcu_array); clone is O(1)."""

    def __init__(self, ctx: Context, handle: C.c_void_p):
        self.ctx = ctx
        self.handle = handle

    def clone(self) -> "CuArray":
        """af::Array::clone: a new handle on the same buffer (O(1)); `add_assign` on either leaves the other alone."""
        h = C.c_void_p()
        _check(self.ctx.lib.gadcu_array_clone(self.handle, C.byref(h)))
        return CuArray(self.ctx, h)

    def dims(self) -> tuple:
        d = Dim4()
        _check(self.ctx.lib.gadcu_array_dims(self.handle, C.byref(d)))
        return d.tuple()

    @property
    def dtype(self) -> int:
        return int(self.ctx.lib.gadcu_array_dtype(self.handle))

    @property
    def is_scalar(self) -> bool:
        return bool(self.ctx.lib.gadcu_array_is_scalar(self.handle))

    def numpy(self) -> np.ndarray:
        """af::Array::host: a host copy shaped by dims (column-major); synchronises."""
        dims = self.dims()
        out = np.empty(int(np.prod(dims)), dtype=_NP[self.dtype])
        _check(self.ctx.lib.gadcu_array_to_host(self.handle, out.ctypes.data_as(C.c_void_p)))
        return out.reshape(dims, order="F")

    def item(self) -> float:
        """Read a device scalar (the deferred half of `as_scalar` / `dot`); synchronises."""
        v = C.c_double()
        _check(self.ctx.lib.gadcu_scalar_to_host(self.handle, C.byref(v)))
        return v.value

    # WeightOps (gad/src/net.rs:161-180)
    def add_assign(self, other: "CuArray") -> None:
        _check(self.ctx.lib.gadcu_array_add_assign(self.handle, other.handle))

    def scale(self, lam: float) -> "CuArray":
        h = C.c_void_p()
        _check(self.ctx.lib.gadcu_array_scale(self.handle, float(lam), C.byref(h)))
        return CuArray(self.ctx, h)

    def upload(self, pinned: np.ndarray, stream: int = 0) -> None:
        """Overwrite the device buffer with `pinned`'s bytes as they lie: the host data must already be in the device
        layout (column-major, i.e. `np.asfortranarray(x).ravel(order="K")` or `x.reshape(-1, order="F")`)."""
        _check(self.ctx.lib.gadcu_array_upload(self.handle, pinned.ctypes.data_as(C.c_void_p), C.c_void_p(stream or None)))

    def upload_async(self, pinned: np.ndarray) -> None:
        """Prefetch from pinned host memory on the copy stream; readers go after ctx.wait_uploads()."""
        _check(self.ctx.lib.gadcu_array_upload_async(self.handle, pinned.ctypes.data_as(C.c_void_p)))

    def download_async(self, pinned: np.ndarray) -> int:
        """Queue the device -> host copy of this array into `pinned` behind the work that produces it and return a ticket;
        the bytes are there after ctx.wait_download(ticket). The host is not held up (unlike numpy() / item())."""
        if not pinned.flags["C_CONTIGUOUS"] and not pinned.flags["F_CONTIGUOUS"]:
            raise ValueError("download_async needs a contiguous host buffer")
        n = 1
        for d in self.dims():
            n *= d
        if pinned.nbytes < n * (4 if self.dtype == F32 else 8):
            raise ValueError("download_async: host buffer too small")
        t = C.c_uint64()
        _check(self.ctx.lib.gadcu_array_download_async(self.handle, pinned.ctypes.data_as(C.c_void_p), C.byref(t)))
        return int(t.value)

    def device_ptr(self) -> int:
        return int(self.ctx.lib.gadcu_array_device_ptr(self.handle) or 0)

    def force(self) -> "CuArray":
        """Make sure the bytes exist: a deferred value (an elementwise region, a product that has been recorded but not
        launched) is computed now, on the context's stream, without a host synchronisation."""
        self.ctx.lib.gadcu_array_device_ptr(self.handle)
        return self

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.gadcu_array_release(self.handle)
                self.handle = None
        except Exception:
            pass


class GradientId:
    """store.rs:20-23 — (arena, index)."""

    __slots__ = ("arena", "index")

    def __init__(self, arena: int, index: int):
        self.arena, self.index = arena, index

    def __eq__(self, other):
        return isinstance(other, GradientId) and (self.arena, self.index) == (other.arena, other.index)

    def __hash__(self):
        return hash((self.arena, self.index))

    def __repr__(self):
        return f"{self.index} @ {self.arena}"


class Value:
    """graph.rs:35-43 — forward data + optional node id."""

    def __init__(self, data: Optional[CuArray], id: Optional[GradientId] = None):
        self._data, self._id = data, id

    def data(self) -> CuArray:
        return self._data

    def id(self) -> Optional[GradientId]:
        return self._id

    def gid(self) -> GradientId:  # net.rs:183-190
        if self._id is None:
            raise MissingIdError(5, "Trying to obtain `id` of a constant value.")
        return self._id

    def dims(self) -> tuple:
        return self._data.dims()

    def _c(self) -> ValueC:
        return ValueC(self._data.handle if self._data is not None else None, self._id.arena if self._id else 0,
                      self._id.index if self._id else 0)


def _wrap(ctx: Context, vc: ValueC) -> Value:
    data = CuArray(ctx, C.c_void_p(vc.data)) if vc.data else None
    gid = GradientId(vc.arena, vc.index) if vc.index else None
    return Value(data, gid)


class Store:
    """GradientReader over a finished sweep (store.rs:27-29): `get` returns None when absent."""

    def __init__(self, ctx: Context, handle: C.c_void_p, values: bool):
        self.ctx, self.handle, self._values = ctx, handle, values

    def get(self, gid: GradientId):
        vc, found = ValueC(), C.c_int()
        _check(self.ctx.lib.gadcu_store_get(self.handle, gid.arena, gid.index, C.byref(vc), C.byref(found)))
        if not found.value:
            return None
        v = _wrap(self.ctx, vc)
        return v if self._values else v.data()

    def ids(self) -> list:
        n = self.ctx.lib.gadcu_store_ids(self.handle, None, 0)
        buf = (C.c_uint32 * max(1, n))()
        self.ctx.lib.gadcu_store_ids(self.handle, buf, n)
        return list(buf[:n])

    def rerun(self, columns: Sequence["CuArray"]) -> list:
        """Batched tapes only: the same tapes at new points (one column per variable, creation order) through the
        already compiled lane program; returns the variables' gradient columns."""
        n = len(columns)
        arr = (C.c_void_p * n)(*[c.handle for c in columns])
        out = (C.c_void_p * n)()
        _check(self.ctx.lib.gadcu_batched_rerun(self.handle, arr, n, out))
        return [CuArray(self.ctx, C.c_void_p(h)) if h else None for h in out]

    def program_stats(self) -> dict:
        a, b = C.c_uint64(), C.c_uint64()
        _check(self.ctx.lib.gadcu_batched_program_stats(self.handle, C.byref(a), C.byref(b)))
        return {"instructions": a.value, "slots": b.value}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.gadcu_store_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ---- Check: dims only, no device (lib.rs:519-520) -------------------------------------------------
class Check:
    """Values are dims: a 4-tuple for arrays, `()` for scalars."""

    def __init__(self):
        self.lib = _lib.load()

    @staticmethod
    def _d(v):
        return Dim4.of(v)

    def _eq(self, *vs):
        if all(v == () for v in vs):
            return ()
        ds = [self._d(v) for v in vs]
        arr = (C.POINTER(Dim4) * len(ds))(*[C.pointer(d) for d in ds])
        out = Dim4()
        _check(self.lib.gadcu_check_equal(arr, len(ds), C.byref(out)))
        return out.tuple()

    def variable(self, data):
        return self._dims_of(data)

    constant = variable

    @staticmethod
    def _dims_of(data):
        if isinstance(data, tuple):
            return data
        if isinstance(data, CuArray):
            return () if data.is_scalar else data.dims()
        arr = np.asarray(data)
        return () if arr.ndim == 0 else _dims4(arr.shape)

    def add(self, a, b): return self._eq(a, b)
    def add_all(self, vs):
        if not vs:
            raise EmptyError(4, "Unexpected empty input for add_all")
        return self._eq(*vs)
    def sub(self, a, b): return self._eq(a, b)
    def mul(self, a, b): return self._eq(a, b)
    def div(self, a, b): return self._eq(a, b)
    def pow(self, a, b): return self._eq(a, b)
    def min(self, a, b): return self._eq(a, b)
    def max(self, a, b): return self._eq(a, b)
    def zeros(self, v): return v
    ones = neg = exp = log = log1p = sin = cos = tanh = sigmoid = reciprocal = sqrt = zeros
    def setc(self, v, c): return v
    addc = mulc = powc = setc
    def abs(self, v): return self.select_argmax(v, v, v, v)
    sign = relu = abs

    def select_argmax(self, v0, v1, r0=None, r1=None):
        if v0 == ():
            return ()
        out = Dim4()
        a, b = self._d(v0), self._d(v1)
        c = self._d(r0) if r0 is not None else None
        d = self._d(r1) if r1 is not None else None
        _check(self.lib.gadcu_check_select_argmax(C.byref(a), C.byref(b), C.byref(c) if c else None, C.byref(d) if d else None, C.byref(out)))
        return out.tuple()

    def _one(self, fn, v, *extra):
        out = Dim4()
        a = self._d(v)
        ex = [self._d(e) for e in extra]
        _check(fn(C.byref(a), *[C.byref(e) for e in ex], C.byref(out)))
        return out.tuple()

    def flat(self, v): return self._one(self.lib.gadcu_check_flat, v)
    def moddims(self, v, d): return self._one(self.lib.gadcu_check_moddims, v, _dims4(d))
    def tile_as(self, v, d): return self._one(self.lib.gadcu_check_tile_as, v, _dims4(d))
    def sum_as(self, v, d): return self._one(self.lib.gadcu_check_sum_as, v, _dims4(d))
    def constant_as(self, s, d): return _dims4(d)
    def scale(self, lam, v): return v

    def as_scalar(self, v):
        a = self._d(v)
        _check(self.lib.gadcu_check_as_scalar(C.byref(a)))
        return ()

    def dot(self, a, b):
        x, y = self._d(a), self._d(b)
        _check(self.lib.gadcu_check_dot(C.byref(x), C.byref(y)))
        return ()

    def norm2(self, v): return self.dot(v, v)

    def _reduced(self, v, d, keeps):
        out = Dim4()
        a, b = self._d(v), self._d(_dims4(d))
        _check(self.lib.gadcu_check_reduced(C.byref(a), C.byref(b), int(keeps), C.byref(out)))
        return out.tuple()

    def max_as(self, v, d): return self._reduced(v, d, False)
    def argmax_as(self, v, d): return self._reduced(v, d, True)
    def softmax_as(self, v, d): return self._reduced(v, d, True)
    def transpose(self, v, conjugate=False): return self._one(self.lib.gadcu_check_transpose, v)

    def matmul(self, a, b, p1=None, p2=None):
        p1, p2 = p1 or MatProp(), p2 or MatProp()
        out = Dim4()
        x, y = self._d(a), self._d(b)
        _check(self.lib.gadcu_check_matmul(C.byref(x), C.byref(y), p1._c(), p2._c(), C.byref(out)))
        return out.tuple()

    def matmul_nn(self, a, b): return self.matmul(a, b)


# ---- Eval / Graph1 / GraphN over device arrays ----------------------------------------------------
class _DeviceAlgebra:
    KIND = 0

    def __init__(self, ctx: Context, _handle: Optional[C.c_void_p] = None):
        self.ctx, self.lib = ctx, ctx.lib
        if _handle is None:
            _handle = C.c_void_p()
            _check(self.lib.gadcu_algebra_create(ctx.handle, self.KIND, C.byref(_handle)))
        self.handle = _handle

    def clone(self):
        h = C.c_void_p()
        _check(self.lib.gadcu_algebra_clone(self.handle, C.byref(h)))
        return type(self)(self.ctx, h)

    def num_nodes(self) -> int:
        return int(self.lib.gadcu_algebra_num_nodes(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.gadcu_algebra_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- helpers
    def _array(self, data) -> CuArray:
        return data if isinstance(data, CuArray) else self.ctx.array(data)

    def _call(self, fn, *args) -> Value:
        out = ValueC()
        _check(fn(self.handle, *args, C.byref(out)))
        return _wrap(self.ctx, out)

    def _u(self, name, v: Value) -> Value:
        return self._call(getattr(self.lib, "gadcu_" + name), C.byref(v._c()))

    def _b(self, name, a: Value, b: Value) -> Value:
        return self._call(getattr(self.lib, "gadcu_" + name), C.byref(a._c()), C.byref(b._c()))

    def _c(self, name, v: Value, c) -> Value:
        is_int = isinstance(c, (int, np.integer)) and not isinstance(c, bool)
        return self._call(getattr(self.lib, "gadcu_" + name), C.byref(v._c()), float(c), int(is_int))

    def _d(self, name, v: Value, dims) -> Value:
        d = Dim4.of(_dims4(dims))
        return self._call(getattr(self.lib, "gadcu_" + name), C.byref(v._c()), C.byref(d))

    # -- CoreAlgebra
    def variable(self, data) -> Value:
        arr = self._array(data)  # keep the handle alive across the call
        return self._call(self.lib.gadcu_variable, arr.handle)

    def constant(self, data) -> Value:
        arr = self._array(data)
        return self._call(self.lib.gadcu_constant, arr.handle)

    def add(self, a, b): return self._b("add", a, b)

    def add_all(self, vs: Sequence[Value]) -> Value:
        cs = [v._c() for v in vs]
        arr = (C.POINTER(ValueC) * max(1, len(cs)))(*[C.pointer(c) for c in cs])
        return self._call(self.lib.gadcu_add_all, arr, len(cs))

    # -- ArithAlgebra
    def sub(self, a, b): return self._b("sub", a, b)
    def mul(self, a, b): return self._b("mul", a, b)
    def zeros(self, v): return self._u("zeros", v)
    def ones(self, v): return self._u("ones", v)
    def neg(self, v): return self._u("neg", v)
    # -- ConstArithAlgebra (a Python int selects the i16 instantiation, a float the T one)
    def setc(self, v, c): return self._c("setc", v, c)
    def addc(self, v, c): return self._c("addc", v, c)
    def mulc(self, v, c): return self._c("mulc", v, c)
    def powc(self, v, c): return self._c("powc", v, c)
    # -- AnalyticAlgebra
    def exp(self, v): return self._u("exp", v)
    def log(self, v): return self._u("log", v)
    def log1p(self, v): return self._u("log1p", v)
    def sin(self, v): return self._u("sin", v)
    def cos(self, v): return self._u("cos", v)
    def tanh(self, v): return self._u("tanh", v)
    def sigmoid(self, v): return self._u("sigmoid", v)
    def reciprocal(self, v): return self._u("reciprocal", v)
    def sqrt(self, v): return self._u("sqrt", v)
    def div(self, a, b): return self._b("div", a, b)
    def pow(self, v, p): return self._b("pow", v, p)
    # -- CompareAlgebra
    def select_argmax(self, v0, v1, r0=None, r1=None):
        c0 = r0._c() if r0 is not None else None
        c1 = r1._c() if r1 is not None else None
        return self._call(self.lib.gadcu_select_argmax, C.byref(v0._c()), C.byref(v1._c()),
                          C.byref(c0) if c0 is not None else None, C.byref(c1) if c1 is not None else None)
    def min(self, a, b): return self._b("min", a, b)
    def max(self, a, b): return self._b("max", a, b)
    def abs(self, v): return self._u("abs", v)
    def sign(self, v): return self._u("sign", v)
    def relu(self, v): return self._u("relu", v)
    # -- ArrayAlgebra
    def flat(self, v): return self._u("flat", v)
    def moddims(self, v, dims): return self._d("moddims", v, dims)
    def tile_as(self, v, dims): return self._d("tile_as", v, dims)
    def sum_as(self, v, dims): return self._d("sum_as", v, dims)
    def constant_as(self, s, dims): return self._d("constant_as", s, dims)
    def as_scalar(self, v): return self._u("as_scalar", v)
    def scale(self, lam, v): return self._b("scale", lam, v)
    def dot(self, a, b): return self._b("dot", a, b)
    def norm2(self, v): return self._u("norm2", v)
    # -- MatrixAlgebra
    def matmul(self, a, b, p1=None, p2=None):
        p1, p2 = p1 or MatProp(), p2 or MatProp()
        return self._call(self.lib.gadcu_matmul, C.byref(a._c()), C.byref(b._c()), p1._c(), p2._c())
    def matmul_nn(self, a, b): return self._b("matmul_nn", a, b)
    def transpose(self, v, conjugate=False):
        return self._call(self.lib.gadcu_transpose, C.byref(v._c()), int(conjugate))
    # -- ArrayCompareAlgebra
    def max_as(self, v, dims): return self._d("max_as", v, dims)
    def argmax_as(self, v, dims): return self._d("argmax_as", v, dims)
    def softmax_as(self, v, dims): return self._d("softmax_as", v, dims)

    def scalar(self, value: float, dtype=F32) -> CuArray:
        return self.ctx.scalar(value, dtype)


class Eval(_DeviceAlgebra):
    """lib.rs:523-526 — forward values only."""
    KIND = 1


class Graph1(_DeviceAlgebra):
    """lib.rs:529 — first-order tape; gradients are data."""
    KIND = 2

    def _seed(self, gradient) -> CuArray:
        if isinstance(gradient, CuArray):
            return gradient
        if np.isscalar(gradient):
            return self.ctx.scalar(gradient, F64 if isinstance(gradient, np.float64) else F32)
        return self.ctx.array(gradient)

    def evaluate_gradients(self, gid: GradientId, gradient) -> Store:  # graph.rs:263-274
        h = C.c_void_p()
        seed = self._seed(gradient)  # keep the handle alive across the call
        _check(self.lib.gadcu_evaluate_gradients(self.handle, gid.arena, gid.index, seed.handle, 0, C.byref(h)))
        return Store(self.ctx, h, values=False)

    def evaluate_gradients_once(self, gid: GradientId, gradient, comm: Optional["Comm"] = None,
                                also_reduce: Sequence[CuArray] = ()) -> Store:  # graph.rs:279-290
        """With `comm`: the sweep of one batch shard — every variable's gradient is summed over the ranks as soon as
        the sweep has completed it (gadcu_evaluate_gradients_dp), overlapping the rest of the sweep; `also_reduce`
        (the step's loss) rides in the same small exchange as the bias gradients."""
        h = C.c_void_p()
        seed = self._seed(gradient)
        if comm is not None and comm.nranks > 1:
            also = (C.c_void_p * max(1, len(also_reduce)))(*[a.handle for a in also_reduce])
            _check(self.lib.gadcu_evaluate_gradients_dp_with(self.handle, gid.arena, gid.index, seed.handle, 1, comm.handle, also,
                                                             len(also_reduce), C.byref(h)))
        else:
            _check(self.lib.gadcu_evaluate_gradients(self.handle, gid.arena, gid.index, seed.handle, 1, C.byref(h)))
        return Store(self.ctx, h, values=False)

    def make_node(self, data: CuArray, inputs: Sequence[Value], vjp) -> Value:
        """Graph::make_node (graph.rs:106-124). vjp(grad_algebra, store_add, gradient: Value)."""
        return _make_node(self, data, inputs, vjp)


class GraphN(_DeviceAlgebra):
    """lib.rs:532 — higher-order tape; gradients are values recorded on the tape."""
    KIND = 3

    def compute_gradients(self, gid: GradientId, gradient: Value, comm: Optional["Comm"] = None) -> Store:  # graph.rs:307-318
        """With `comm`: the sweep of one batch shard — every variable's gradient is summed over the ranks as soon as the
        sweep has completed it (gadcu_compute_gradients_dp) and comes back as a constant value."""
        h = C.c_void_p()
        if comm is not None and comm.nranks > 1:
            _check(self.lib.gadcu_compute_gradients_dp(self.handle, gid.arena, gid.index, C.byref(gradient._c()), comm.handle, C.byref(h)))
        else:
            _check(self.lib.gadcu_compute_gradients(self.handle, gid.arena, gid.index, C.byref(gradient._c()), C.byref(h)))
        return Store(self.ctx, h, values=True)

    def make_node(self, data: CuArray, inputs: Sequence[Value], vjp) -> Value:
        return _make_node(self, data, inputs, vjp)


class BatchedGraph1(Graph1):
    """One scalar Graph1 tape per lane, SoA columns; ops only record, the sweep runs one kernel."""
    KIND = 4

    def __init__(self, ctx: Context, lanes: int, dtype=F64, _handle=None):
        if _handle is None:
            _handle = C.c_void_p()
            _check(ctx.lib.gadcu_batched_create(ctx.handle, dtype, lanes, C.byref(_handle)))
        super().__init__(ctx, _handle)
        self.lanes, self.dtype = lanes, dtype

    def _seed(self, gradient) -> CuArray:
        if isinstance(gradient, CuArray):
            return gradient
        if np.isscalar(gradient):
            return self.ctx.scalar(gradient, self.dtype)
        return self.ctx.array(gradient)

    def force(self, v: Value) -> CuArray:
        h = C.c_void_p()
        _check(self.lib.gadcu_batched_force(self.handle, C.byref(v._c()), C.byref(h)))
        return CuArray(self.ctx, h)


# user-defined operators (tests/extension.rs): callbacks must stay alive as long as the tape
_KEEPALIVE = []


class _GradAlgebraView(_DeviceAlgebra):
    """The gradient algebra handed to a VJP callback (borrowed handle: never destroyed here)."""

    def __del__(self):
        self.handle = None


def _make_node(g: _DeviceAlgebra, data: CuArray, inputs: Sequence[Value], vjp) -> Value:
    ctx = g.ctx

    def trampoline(_user, alg_handle, store_handle, grad_ptr):
        try:
            ga = _GradAlgebraView(ctx, C.c_void_p(alg_handle))
            gc = grad_ptr.contents
            ctx.lib.gadcu_array_retain(C.c_void_p(gc.data))
            gradient = _wrap(ctx, gc)

            def add_gradient(gid: GradientId, value: Value) -> None:
                _check(ctx.lib.gadcu_store_add_gradient(C.c_void_p(store_handle), C.c_void_p(alg_handle), gid.arena, gid.index,
                                                         C.byref(value._c())))

            vjp(ga, add_gradient, gradient)
            return 0
        except GadError as e:
            return e.status
        except Exception:
            return 11

    cb = _lib.VJP_FN(trampoline)
    _KEEPALIVE.append(cb)
    cs = [v._c() for v in inputs]
    arr = (C.POINTER(ValueC) * max(1, len(cs)))(*[C.pointer(c) for c in cs])
    out = ValueC()
    _check(ctx.lib.gadcu_make_node(g.handle, data.handle, arr, len(cs), cb, None, None, C.byref(out)))
    return _wrap(ctx, out)


class Capture:
    """CUDA-graph capture of a whole step; handles created inside stay valid and are refilled by
    every launch()."""

    def __init__(self, ctx: Context):
        self.ctx, self.handle = ctx, None

    def __enter__(self):
        _check(self.ctx.lib.gadcu_capture_begin(self.ctx.handle))
        return self

    def __exit__(self, et, ev, tb):
        h = C.c_void_p()
        st = self.ctx.lib.gadcu_capture_end(self.ctx.handle, C.byref(h))
        if et is None:
            _check(st)
            self.handle = h
        return False

    def launch(self) -> None:
        _check(self.ctx.lib.gadcu_capture_launch(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.gadcu_capture_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Comm:
    """Batch-dimension data parallelism: NCCL all-reduce (sum) of gradient arrays, in place."""

    def __init__(self, ctx: Context, nranks: int, rank: int, unique_id: bytes):
        self.ctx = ctx
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        _check(ctx.lib.gadcu_comm_init(ctx.handle, nranks, rank, buf, C.byref(h)))
        self.handle, self.nranks, self.rank = h, nranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(_lib.load().gadcu_comm_unique_id(buf))
        return buf.raw

    def allreduce_sum(self, arrays: Sequence[CuArray]) -> None:
        arr = (C.c_void_p * len(arrays))(*[a.handle for a in arrays])
        _check(self.ctx.lib.gadcu_comm_allreduce_sum(self.handle, arr, len(arrays)))

    def allreduce_sum_async(self, arrays: Sequence[CuArray]) -> None:
        """On the communicator's stream, after what is already issued on the compute stream; see join()."""
        arr = (C.c_void_p * len(arrays))(*[a.handle for a in arrays])
        _check(self.ctx.lib.gadcu_comm_allreduce_sum_async(self.handle, arr, len(arrays)))

    def join(self) -> None:
        _check(self.ctx.lib.gadcu_comm_join(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.ctx.lib.gadcu_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
