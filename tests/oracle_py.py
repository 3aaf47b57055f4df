"""ctypes wrapper over oracle/liboracle.so (TEST INFRASTRUCTURE: the checker, never the product).

Presents the CPU oracle with the same Python interface as gad_b200.algebra so one generic gad
program can be run on both and compared.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

OPS = ["variable", "constant", "add", "add_all", "sub", "mul", "zeros", "ones", "neg", "setc", "addc", "mulc", "powc", "exp",
       "log", "log1p", "sin", "cos", "tanh", "sigmoid", "reciprocal", "sqrt", "div", "pow", "select_argmax", "min", "max", "abs",
       "sign", "relu", "flat", "moddims", "tile_as", "sum_as", "constant_as", "as_scalar", "scale", "dot", "norm2", "matmul",
       "transpose", "matmul_nn", "max_as", "argmax_as", "softmax_as"]
OP = {n: i for i, n in enumerate(OPS)}
CHECK, EVAL, GRAPH1, GRAPHN = 0, 1, 2, 3
ERR = {1: "Dimensions", 2: "ReducedDimensions", 3: "Lengths", 4: "Empty", 5: "MissingId", 6: "MissingGradient", 7: "MissingNode"}

_lib = None


def build() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_alg_new.restype = C.c_void_p
        L.oracle_alg_new.argtypes = [C.c_int]
        L.oracle_alg_clone.restype = C.c_void_p
        L.oracle_alg_clone.argtypes = [C.c_void_p]
        L.oracle_alg_free.argtypes = [C.c_void_p]
        L.oracle_alg_num_nodes.restype = C.c_size_t
        L.oracle_alg_num_nodes.argtypes = [C.c_void_p]
        L.oracle_scalar.restype = C.c_void_p
        L.oracle_scalar.argtypes = [C.c_int, C.c_double]
        L.oracle_array.restype = C.c_void_p
        L.oracle_array.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
        L.oracle_check_dims.restype = C.c_void_p
        L.oracle_check_dims.argtypes = [C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_val_free.argtypes = [C.c_void_p]
        L.oracle_val_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_val_read.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_op.restype = C.c_int
        L.oracle_op.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_double),
                                C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
        L.oracle_gradients.restype = C.c_int
        L.oracle_gradients.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.oracle_store_free.argtypes = [C.c_void_p]
        L.oracle_store_get.restype = C.c_void_p
        L.oracle_store_get.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.oracle_store_ids.restype = C.c_size_t
        L.oracle_store_ids.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_size_t]
        L.oracle_rosenbrock_batch.restype = C.c_double
        L.oracle_rosenbrock_batch.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_c3.restype = C.c_double
        L.oracle_c3.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_mlp_step.restype = C.c_double
        L.oracle_mlp_step.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
        L.oracle_set_corrected_adjoint.restype = C.c_int
        L.oracle_set_corrected_adjoint.argtypes = [C.c_int]
        L.oracle_set_blas.restype = C.c_int
        L.oracle_set_blas.argtypes = [C.c_char_p, C.c_int]
        L.oracle_mlp_hvp_corrected.restype = C.c_double
        L.oracle_mlp_hvp_corrected.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                               C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_void_p), C.c_int,
                                               C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def set_corrected_adjoint(on: bool) -> bool:
    """False (default): the reference's matmul VJP as written (matrix.rs:164-176), quirk for transposed operands
    included. True: the true adjoint there — the device library's default mode. Returns the previous setting."""
    return bool(lib().oracle_set_corrected_adjoint(int(on)))


def find_blas():
    """Path of a CBLAS this host has (the OpenBLAS numpy bundles), or None."""
    import glob
    d = os.path.join(os.path.dirname(os.path.dirname(np.__file__)), "numpy.libs")
    hits = sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so*"))) + sorted(glob.glob(os.path.join(d, "libopenblas*.so*")))
    return hits[0] if hits else None


def set_blas(path, threads: int = 0) -> bool:
    """CPU BASELINE only: the oracle's matmul through a BLAS sgemm, as ArrayFire's CPU backend does it (never used by
    parity checks: summation order differs from the documented loop). path=None switches back."""
    return bool(lib().oracle_set_blas(path.encode() if path else None, int(threads)))


class OracleError(Exception):
    def __init__(self, status, msg):
        super().__init__(f"{ERR.get(status, status)}: {msg}")
        self.status = status
        self.kind = ERR.get(status, "Other")


def _dims4(dims):
    if isinstance(dims, int):
        dims = (dims,)
    d = tuple(int(x) for x in dims)
    return d + (1,) * (4 - len(d))


class OValue:
    def __init__(self, handle):
        self.handle = handle
        dt, sc, ar, ix = C.c_int(), C.c_int(), C.c_uint32(), C.c_uint32()
        dims = (C.c_uint64 * 4)()
        lib().oracle_val_info(handle, C.byref(dt), C.byref(sc), dims, C.byref(ar), C.byref(ix))
        self.dtype, self.is_scalar = dt.value, bool(sc.value)
        self._dims = tuple(int(x) for x in dims)
        self._id = (ar.value, ix.value) if ix.value else None

    def id(self):
        return self._id

    def gid(self):
        if self._id is None:
            raise OracleError(5, "Trying to obtain `id` of a constant value.")
        return self._id

    def dims(self):
        return () if self.is_scalar else self._dims

    def data(self):
        """Host data: float for scalars, column-major numpy array (shape = dims) for arrays."""
        if self.dtype < 0:
            return self.dims()
        npdt = np.float32 if self.dtype == 0 else np.float64
        if self.is_scalar:
            out = np.zeros(1, dtype=npdt)
            lib().oracle_val_read(self.handle, out.ctypes.data_as(C.c_void_p))
            return out[0]
        out = np.empty(int(np.prod(self._dims)), dtype=npdt)
        lib().oracle_val_read(self.handle, out.ctypes.data_as(C.c_void_p))
        return out.reshape(self._dims, order="F")

    def __del__(self):
        try:
            lib().oracle_val_free(self.handle)
        except Exception:
            pass


def _raw(data, dtype=None) -> OValue:
    if isinstance(data, OValue):
        return data
    if np.isscalar(data) or (isinstance(data, np.ndarray) and data.ndim == 0):
        dt = 1 if isinstance(data, np.float64) else 0  # plain Python numbers default to f32
        if dtype is not None:
            dt = 0 if np.dtype(dtype) == np.float32 else 1
        return OValue(lib().oracle_scalar(dt, float(data)))
    arr = np.asarray(data)
    if dtype is not None:
        arr = arr.astype(dtype)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float32)
    dims = _dims4(arr.shape)
    flat = np.ascontiguousarray(arr.reshape(-1, order="F"))
    d = (C.c_uint64 * 4)(*dims)
    return OValue(lib().oracle_array(0 if flat.dtype == np.float32 else 1, d, flat.ctypes.data_as(C.c_void_p)))


class OStore:
    def __init__(self, handle):
        self.handle = handle

    def get(self, gid, like: OValue = None, dtype=0, is_scalar=False):
        """Typed read (the reference's store is typed by GradientId<T>): pass `like` = the value
        whose gradient is wanted, or dtype / is_scalar explicitly."""
        if like is not None:
            dtype, is_scalar = like.dtype, like.is_scalar
        h = lib().oracle_store_get(self.handle, gid[0], gid[1], dtype, int(is_scalar))
        return OValue(h) if h else None

    def ids(self):
        n = lib().oracle_store_ids(self.handle, None, 0)
        buf = (C.c_uint32 * max(1, n))()
        lib().oracle_store_ids(self.handle, buf, n)
        return list(buf[:n])

    def __del__(self):
        try:
            lib().oracle_store_free(self.handle)
        except Exception:
            pass


class OAlgebra:
    KIND = EVAL

    def __init__(self, _handle=None):
        self.handle = _handle or lib().oracle_alg_new(self.KIND)

    def clone(self):
        return type(self)(lib().oracle_alg_clone(self.handle))

    def num_nodes(self):
        return lib().oracle_alg_num_nodes(self.handle)

    def __del__(self):
        try:
            lib().oracle_alg_free(self.handle)
        except Exception:
            pass

    def _op(self, name, ins, sp=None, dp=None) -> OValue:
        arr = (C.c_void_p * max(1, len(ins)))(*[(v.handle if v is not None else None) for v in ins])
        spa = (C.c_double * 4)(*((list(sp) + [0.0] * 4)[:4])) if sp is not None else None
        dpa = (C.c_uint64 * 4)(*_dims4(dp)) if dp is not None else None
        out = C.c_void_p()
        st = lib().oracle_op(self.handle, OP[name], arr, len(ins), spa, dpa, C.byref(out))
        if st != 0:
            raise OracleError(st, lib().oracle_last_error().decode())
        return OValue(out.value)

    def variable(self, data, dtype=None): return self._op("variable", [_raw(data, dtype)])
    def constant(self, data, dtype=None): return self._op("constant", [_raw(data, dtype)])
    def add(self, a, b): return self._op("add", [a, b])
    def add_all(self, vs): return self._op("add_all", list(vs))
    def sub(self, a, b): return self._op("sub", [a, b])
    def mul(self, a, b): return self._op("mul", [a, b])
    def zeros(self, v): return self._op("zeros", [v])
    def ones(self, v): return self._op("ones", [v])
    def neg(self, v): return self._op("neg", [v])

    def _c(self, name, v, c):
        is_int = isinstance(c, (int, np.integer)) and not isinstance(c, bool)
        return self._op(name, [v], sp=[float(c), 1.0 if is_int else 0.0])

    def setc(self, v, c): return self._c("setc", v, c)
    def addc(self, v, c): return self._c("addc", v, c)
    def mulc(self, v, c): return self._c("mulc", v, c)
    def powc(self, v, c): return self._c("powc", v, c)
    def exp(self, v): return self._op("exp", [v])
    def log(self, v): return self._op("log", [v])
    def log1p(self, v): return self._op("log1p", [v])
    def sin(self, v): return self._op("sin", [v])
    def cos(self, v): return self._op("cos", [v])
    def tanh(self, v): return self._op("tanh", [v])
    def sigmoid(self, v): return self._op("sigmoid", [v])
    def reciprocal(self, v): return self._op("reciprocal", [v])
    def sqrt(self, v): return self._op("sqrt", [v])
    def div(self, a, b): return self._op("div", [a, b])
    def pow(self, a, b): return self._op("pow", [a, b])
    def select_argmax(self, v0, v1, r0=None, r1=None): return self._op("select_argmax", [v0, v1, r0, r1])
    def min(self, a, b): return self._op("min", [a, b])
    def max(self, a, b): return self._op("max", [a, b])
    def abs(self, v): return self._op("abs", [v])
    def sign(self, v): return self._op("sign", [v])
    def relu(self, v): return self._op("relu", [v])
    def flat(self, v): return self._op("flat", [v])
    def moddims(self, v, d): return self._op("moddims", [v], dp=d)
    def tile_as(self, v, d): return self._op("tile_as", [v], dp=d)
    def sum_as(self, v, d): return self._op("sum_as", [v], dp=d)
    def constant_as(self, s, d): return self._op("constant_as", [s], dp=d)
    def as_scalar(self, v): return self._op("as_scalar", [v])
    def scale(self, lam, v): return self._op("scale", [lam, v])
    def dot(self, a, b): return self._op("dot", [a, b])
    def norm2(self, v): return self._op("norm2", [v])

    def matmul(self, a, b, p1=None, p2=None):
        t1 = (p1.transposed, p1.conjugated) if p1 is not None else (False, False)
        t2 = (p2.transposed, p2.conjugated) if p2 is not None else (False, False)
        return self._op("matmul", [a, b], sp=[float(t1[0]), float(t1[1]), float(t2[0]), float(t2[1])])

    def matmul_nn(self, a, b): return self._op("matmul_nn", [a, b])
    def transpose(self, v, conjugate=False): return self._op("transpose", [v], sp=[float(conjugate)])
    def max_as(self, v, d): return self._op("max_as", [v], dp=d)
    def argmax_as(self, v, d): return self._op("argmax_as", [v], dp=d)
    def softmax_as(self, v, d): return self._op("softmax_as", [v], dp=d)


class OEval(OAlgebra):
    KIND = EVAL


class OCheck(OAlgebra):
    KIND = CHECK

    @staticmethod
    def dims_value(dims) -> OValue:
        if dims == ():
            return OValue(lib().oracle_check_dims(1, (C.c_uint64 * 4)(1, 1, 1, 1)))
        return OValue(lib().oracle_check_dims(0, (C.c_uint64 * 4)(*_dims4(dims))))

    def variable(self, data, dtype=None):
        return data if isinstance(data, OValue) else self.dims_value(data)

    constant = variable


class OGraph1(OAlgebra):
    KIND = GRAPH1

    def _grad(self, gid, gradient, mode, dtype=None):
        out = C.c_void_p()
        seed = _raw(gradient, dtype)
        st = lib().oracle_gradients(self.handle, gid[0], gid[1], seed.handle, mode, C.byref(out))
        if st != 0:
            raise OracleError(st, lib().oracle_last_error().decode())
        return OStore(out.value)

    def evaluate_gradients(self, gid, gradient, dtype=None): return self._grad(gid, gradient, 0, dtype)
    def evaluate_gradients_once(self, gid, gradient, dtype=None): return self._grad(gid, gradient, 1, dtype)


class OGraphN(OAlgebra):
    KIND = GRAPHN

    def compute_gradients(self, gid, gradient: OValue):
        out = C.c_void_p()
        st = lib().oracle_gradients(self.handle, gid[0], gid[1], gradient.handle, 2, C.byref(out))
        if st != 0:
            raise OracleError(st, lib().oracle_last_error().decode())
        return OStore(out.value)


# ---- bulk workloads (parity at scale + CPU baseline) --------------------------------------------
def rosenbrock_batch(x: np.ndarray, nthreads: int = 1):
    """x: [n, lanes] float64 (row i = variable i, SoA). Returns (f[lanes], grad[n, lanes], seconds)."""
    n, lanes = x.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = np.empty(lanes, dtype=np.float64)
    g = np.empty_like(x)
    secs = lib().oracle_rosenbrock_batch(n, lanes, x.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
                                         g.ctypes.data_as(C.c_void_p), nthreads)
    return f, g, secs


def c3(x: np.ndarray, y: np.ndarray):
    n = x.size
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    z = np.zeros(1, dtype=np.float32)
    gx, gy = np.empty_like(x), np.empty_like(y)
    sweep = C.c_double()
    secs = lib().oracle_c3(n, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                           gx.ctypes.data_as(C.c_void_p), gy.ctypes.data_as(C.c_void_p), C.byref(sweep))
    return float(z[0]), gx, gy, secs, sweep.value


def mlp_step(X, Ws, bs, target, nthreads: int = 1):
    """Column-major [B, D] inputs as numpy arrays of shape (B, D). Returns (loss, gWs, gbs, seconds)."""
    B, D = X.shape
    L = len(Ws)
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, order="F"))
    Xf, Tf = f(X), f(target)
    Wf, bf = [f(w) for w in Ws], [f(b) for b in bs]
    gW = [np.empty(D * D, dtype=np.float32) for _ in range(L)]
    gb = [np.empty(D, dtype=np.float32) for _ in range(L)]
    loss = np.zeros(1, dtype=np.float32)
    ptrs = lambda arrs: (C.c_void_p * L)(*[a.ctypes.data_as(C.c_void_p).value for a in arrs])
    secs = lib().oracle_mlp_step(B, D, L, Xf.ctypes.data_as(C.c_void_p), ptrs(Wf), ptrs(bf), Tf.ctypes.data_as(C.c_void_p),
                                 loss.ctypes.data_as(C.c_void_p), ptrs(gW), ptrs(gb), nthreads)
    return float(loss[0]), [g.reshape((D, D), order="F") for g in gW], [g.reshape((1, D), order="F") for g in gb], secs


def mlp_hvp(X, Ws, bs, target, Vs, nthreads: int = 1):
    """C5 on the oracle's GraphN with the CORRECTED matmul adjoint (the reference's own VJP cannot differentiate
    through its backward GEMMs for B != D: SURVEY App. C.2). Returns (loss, [H v]_l, seconds, tape nodes)."""
    B, D = X.shape
    L = len(Ws)
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, order="F"))
    Xf, Tf = f(X), f(target)
    Wf, bf, Vf = [f(w) for w in Ws], [f(b) for b in bs], [f(v) for v in Vs]
    hv = [np.empty(D * D, dtype=np.float32) for _ in range(L)]
    loss = np.zeros(1, dtype=np.float32)
    nodes = C.c_uint64()
    ptrs = lambda arrs: (C.c_void_p * L)(*[a.ctypes.data_as(C.c_void_p).value for a in arrs])
    secs = lib().oracle_mlp_hvp_corrected(B, D, L, Xf.ctypes.data_as(C.c_void_p), ptrs(Wf), ptrs(bf), Tf.ctypes.data_as(C.c_void_p),
                                          ptrs(Vf), loss.ctypes.data_as(C.c_void_p), ptrs(hv), nthreads, C.byref(nodes))
    return float(loss[0]), [h.reshape((D, D), order="F") for h in hv], secs, int(nodes.value)
