"""gad's `Net` combinators and the training step over any algebra of this backend — SURVEY §8(f) row 1.

Mirror of gad/src/net.rs (Net, WeightOps, InputData / ConstantData / WeightData, Map / Then / Using, tuples and
Vec of nets) and gad/src/net_ext.rs (SquareLoss, DiffNet::apply_gradient_step), with the reference's method names
and semantics. A net is generic over the algebra it is evaluated with, exactly as in the reference: the same net
runs on `Check` (dims only), `Eval`, `Graph1` and `GraphN` of gad_b200.algebra — and, because nothing here touches
a device, on any other object with the same operator surface (the tests run it on the CPU oracle too).

`Weights` are nested tuples / lists of arrays mirroring the net's structure (`()` for a net without weights,
net.rs:273-301); WeightOps (net.rs:100-104, 161-180) is `weights_scale` / `weights_add_assign` over that structure.
`apply_gradient_step` follows net_ext.rs:47-73 statement by statement; with a communicator the examples of the batch
are split over the ranks and the weight update (and the loss) summed before `update_weights` — the data-parallel seam
of SURVEY §8e (the reference sums per-example gradients, net_ext.rs:62-65).
"""
from __future__ import annotations

import io
import struct
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import dp as _dp
from .algebra import CuArray, DimensionsError, EmptyError, LengthsError, MissingGradientError


# ---- data helpers: a net's data is whatever the algebra's `variable` / `constant` accept --------------------------
def dims_of(data) -> tuple:
    """HasDims::dims (core.rs:41-73): Dim4 of an array, `()` of a scalar. Metadata only — nothing here computes."""
    if isinstance(data, CuArray):
        return () if data.is_scalar else data.dims()
    if isinstance(data, tuple):  # Check's values are dims
        return data
    if hasattr(data, "dims") and callable(data.dims):
        return data.dims()
    shape = getattr(data, "shape", None)  # a host array handed in as input / target
    if shape is None:
        return ()
    if len(shape) == 0:
        return ()
    shape = tuple(int(s) for s in shape)
    return shape + (1,) * (4 - len(shape))


def _check_equal_dimensions(name: str, a, b) -> None:  # error.rs:139-153
    if tuple(a) != tuple(b):
        raise DimensionsError(1, f"Incompatible dimensions for {name}: {a} {b}")


def _check_equal_lengths(name: str, a: int, b: int) -> None:  # error.rs:156-167
    if a != b:
        raise LengthsError(3, f"Incompatible lengths for {name}: {a} {b}")


# ---- WeightOps over the nested weight structure (net.rs:100-104,161-180,483-493,559-569,600-616,686-704) ---------
# A leaf is any array type with the reference's WeightOps surface — `scale(lambda) -> new`, `add_assign(other)` (in
# place, copy-on-write), `clone()` — i.e. gad_b200's CuArray. No arithmetic happens in this module: the CPU tests run
# the same nets on the oracle through an adapter of their own (tests/host_weights.py) that supplies such a type.
def weights_scale(w, lam: float):
    if isinstance(w, (tuple, list)):
        return type(w)(weights_scale(x, lam) for x in w)
    return w.scale(lam)


def weights_add_assign(w, other):
    """`self += other`, structure-wise (net.rs:171-175 per leaf); returns the structure."""
    if isinstance(w, (tuple, list)):
        _check_equal_lengths("add_assign", len(w), len(other))
        return type(w)(weights_add_assign(a, b) for a, b in zip(w, other))
    _check_equal_dimensions("add_assign", dims_of(other), dims_of(w))
    w.add_assign(other)
    return w


def weights_clone(w):
    if isinstance(w, (tuple, list)):
        return type(w)(weights_clone(x) for x in w)
    return w.clone()


def weights_leaves(w) -> list:
    if isinstance(w, (tuple, list)):
        out = []
        for x in w:
            out.extend(weights_leaves(x))
        return out
    return [w]


# ---- (de)serialisation of weights: checkpoint / resume (net.rs:100, tests/net.rs:136-149) ----------------------------
_MAGIC = b"GADW1\0"
_DT_CODE = {"float32": 0, "float64": 1}


def _host(x):
    """Host copy of a leaf as a numpy array shaped by its dims (column-major indexing preserved)."""
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def serialize_weights(w) -> bytes:
    """Nested structure of arrays -> bytes, self-describing. Layout: magic, then recursively: b'T'/b'L' + u32 count for
    tuples/lists, b'A' + u8 dtype (0 f32, 1 f64) + 4 x u64 dims + raw column-major little-endian data for arrays."""
    buf = io.BytesIO()
    buf.write(_MAGIC)

    def put(x):
        if isinstance(x, (tuple, list)):
            buf.write(b"T" if isinstance(x, tuple) else b"L")
            buf.write(struct.pack("<I", len(x)))
            for y in x:
                put(y)
            return
        host = _host(x)
        if host.dtype.name not in _DT_CODE:
            raise TypeError(f"weights must be f32 or f64 arrays, got {host.dtype}")
        dims = dims_of(x) or (1, 1, 1, 1)
        buf.write(b"A")
        buf.write(struct.pack("<B4Q", _DT_CODE[host.dtype.name], *dims))
        buf.write(np.ascontiguousarray(host.reshape(-1, order="F")).astype(host.dtype.newbyteorder("<")).tobytes())

    put(w)
    return buf.getvalue()


def deserialize_weights(data: bytes, ctx=None):
    """Inverse of serialize_weights; arrays come back on `ctx`'s device (or as numpy arrays when ctx is None)."""
    buf = io.BytesIO(data)
    if buf.read(len(_MAGIC)) != _MAGIC:
        raise ValueError("not a gad_b200 weight blob")

    def get():
        tag = buf.read(1)
        if tag in (b"T", b"L"):
            (n,) = struct.unpack("<I", buf.read(4))
            items = [get() for _ in range(n)]
            return tuple(items) if tag == b"T" else items
        if tag != b"A":
            raise ValueError("corrupt weight blob")
        dt, *dims = struct.unpack("<B4Q", buf.read(33))
        npdt = np.float32 if dt == 0 else np.float64
        n = int(np.prod(dims))
        arr = np.frombuffer(buf.read(n * np.dtype(npdt).itemsize), dtype=np.dtype(npdt).newbyteorder("<")).astype(npdt)
        arr = arr.reshape(tuple(dims), order="F")
        return ctx.array(arr) if ctx is not None else arr

    return get()


# The reference's own checkpoint format: `bincode::serialize(&net.get_weights())` (gad/tests/net.rs:136-149; bincode 1.3.1
# per Cargo.lock, default options = little-endian fixed-width integers). Weights are nested tuples / `Vec`s of
# `af::Array<T>`, whose serde impl is the `afserde` feature of the arrayfire crate 3.8.0 (gad/Cargo.toml:20): it
# serialises `struct ArrayOnHost { dtype: DType, shape: Dim4, data: Vec<T> }`. Under bincode that is
#     u32  variant index of af::DType (F32 = 0, F64 = 2)
#     4 x u64  Dim4.dims
#     u64  element count, then the elements in column-major order
# tuples are the concatenation of their fields, `()` is zero bytes, a Vec is a u64 length followed by its items.
# bincode is not self-describing, so reading needs the weight structure (`like`), exactly as Rust needs the type.
# The arrayfire crate is not vendored under /root/reference and cannot be built here, so this layout is restated from
# the published crate and is NOT checked against a blob written by the reference (unpinned); the round trip and the
# byte layout above are what the tests pin.
_AF_DTYPE = {"float32": 0, "float64": 2}


def serialize_weights_bincode(w) -> bytes:
    buf = io.BytesIO()

    def put(x):
        if isinstance(x, tuple):
            for y in x:
                put(y)
            return
        if isinstance(x, list):
            buf.write(struct.pack("<Q", len(x)))
            for y in x:
                put(y)
            return
        host = _host(x)
        if host.dtype.name not in _AF_DTYPE:
            raise TypeError(f"weights must be f32 or f64 arrays, got {host.dtype}")
        dims = dims_of(x) or (1, 1, 1, 1)
        flat = np.ascontiguousarray(host.reshape(-1, order="F")).astype(host.dtype.newbyteorder("<"))
        buf.write(struct.pack("<I4QQ", _AF_DTYPE[host.dtype.name], *dims, flat.size))
        buf.write(flat.tobytes())

    put(w)
    return buf.getvalue()


def deserialize_weights_bincode(data: bytes, like, ctx=None):
    """`like`: a weight structure of the expected shape (e.g. `net.get_weights()` of a freshly built net) — the stand-in
    for the Rust type that drives bincode's deserializer. Dtype, dims and length prefixes in the blob are checked."""
    buf = io.BytesIO(data)

    def take(n):
        b = buf.read(n)
        if len(b) != n:
            raise ValueError("truncated bincode weight blob")
        return b

    def get(shape):
        if isinstance(shape, tuple):
            return tuple(get(y) for y in shape)
        if isinstance(shape, list):
            (n,) = struct.unpack("<Q", take(8))
            _check_equal_lengths("deserialize", n, len(shape))
            return [get(y) for y in shape]
        dt, d0, d1, d2, d3, count = struct.unpack("<I4QQ", take(44))
        if dt not in (0, 2):
            raise ValueError(f"unsupported af::DType index {dt} in weight blob")
        npdt = np.float32 if dt == 0 else np.float64
        if count != d0 * d1 * d2 * d3:
            raise ValueError("element count does not match dims in weight blob")
        arr = np.frombuffer(take(count * np.dtype(npdt).itemsize), dtype=np.dtype(npdt).newbyteorder("<")).astype(npdt)
        arr = arr.reshape((d0, d1, d2, d3), order="F")
        return ctx.array(arr) if ctx is not None else arr

    out = get(like)
    if buf.read(1):
        raise ValueError("trailing bytes in bincode weight blob")
    return out


# ---- Net (net.rs:33-96) ----------------------------------------------------------------------------------------
class Net:
    """A network over an algebra. Subclasses implement the five required methods of the reference trait."""

    def eval_with_gradient_info(self, graph, input):  # -> (output, gradient info)
        raise NotImplementedError

    def get_weights(self):
        raise NotImplementedError

    def update_weights(self, delta) -> None:
        raise NotImplementedError

    def set_weights(self, weights) -> None:
        raise NotImplementedError

    def read_weight_gradients(self, info, reader):
        raise NotImplementedError

    # provided methods
    def eval(self, graph, input):  # net.rs:58-60
        return self.eval_with_gradient_info(graph, input)[0]

    def map(self, f: Callable) -> "Map":  # net.rs:62-68
        return Map(self, f)

    def using(self, net: "Net") -> "Using":  # net.rs:70-76
        return Using(self, net)

    def then(self, net: "Net") -> "Then":  # net.rs:78-84
        return Then(self, net)

    def and_(self, net: "Net") -> "TupleNet":  # net.rs:86-92 (`and` is a Python keyword)
        return TupleNet((self, net))

    def add_square_loss(self) -> "SquareLoss":  # net_ext.rs:21-27
        return SquareLoss(self)

    def evaluate(self, eval_algebra, input):  # EvalNet::evaluate (net.rs:118-123): the caller supplies Eval(ctx)
        return self.eval(eval_algebra, input)

    def check(self, check_algebra, input):  # CheckNet::check (net.rs:129-134)
        return self.eval(check_algebra, input)

    # DiffNet::apply_gradient_step (net_ext.rs:47-73)
    def apply_gradient_step(self, lam: float, batch: Sequence, new_graph: Callable[[], Any], comm=None) -> float:
        """One gradient step over `batch`: per example a fresh Graph1, forward, reverse sweep with seed 1, then
        `delta += lambda * gradients`; finally `update_weights(delta)`. Returns the cumulated output (the loss).
        With `comm` (a gad_b200 Comm) every rank passes the same batch, works on its share of the examples, and the
        weight update and the loss are summed over the ranks before the update."""
        if comm is not None and comm.nranks > 1:
            a, b = _dp.shard_bounds(len(batch), comm.nranks, comm.rank)
            examples = list(batch)[a:b]
        else:
            examples = list(batch)
        if len(batch) == 0:
            raise EmptyError(4, "Unexpected empty input for apply_gradient_step")
        delta = None
        cumulated = None
        for example in examples:
            g = new_graph()  # Graph1::new()
            output, info = self.eval_with_gradient_info(g, example)
            data = output.data()
            value = data.item() if hasattr(data, "item") and not isinstance(data, float) else float(data)
            cumulated = value if cumulated is None else cumulated + value
            store = g.evaluate_gradients_once(output.gid(), 1.0)
            gradients = self.read_weight_gradients(info, store)
            scaled = weights_scale(gradients, lam)
            delta = scaled if delta is None else weights_add_assign(delta, scaled)
        if comm is not None and comm.nranks > 1:
            delta, cumulated = _allreduce_update(comm, self, delta, cumulated)
        if delta is not None:
            self.update_weights(delta)
        return cumulated


def _allreduce_update(comm, net: Net, delta, cumulated):
    """Sum the weight update and the loss over the ranks (a rank without examples contributes zeros)."""
    ctx = comm.ctx
    if delta is None:
        delta = weights_scale(net.get_weights(), 0.0)
    loss = ctx.array([0.0 if cumulated is None else cumulated])
    leaves = [x for x in weights_leaves(delta) if isinstance(x, CuArray)]
    comm.allreduce_sum(leaves + [loss])
    return delta, float(loss.numpy().reshape(-1)[0])


# ---- leaf nets (net.rs:193-380) -----------------------------------------------------------------------------------
class InputData(Net):
    """net.rs:252-301: the input of the network, checked against the declared dims, recorded as a constant."""

    def __init__(self, dims):
        self.dims = tuple(dims) + (1,) * (4 - len(tuple(dims))) if tuple(dims) != () else ()

    def eval_with_gradient_info(self, graph, input):
        _check_equal_dimensions("eval_with_gradient_info", dims_of(input), self.dims)
        return graph.constant(input), ()

    def get_weights(self):
        return ()

    def set_weights(self, weights) -> None:
        pass

    def update_weights(self, delta) -> None:
        pass

    def read_weight_gradients(self, info, reader):
        return ()


class ConstantData(Net):
    """net.rs:303-338."""

    def __init__(self, data):
        self.data = data

    def get(self):
        return self.data

    def eval_with_gradient_info(self, graph, input=()):
        return graph.constant(self.data), ()

    def get_weights(self):
        return ()

    def set_weights(self, weights) -> None:
        pass

    def update_weights(self, delta) -> None:
        pass

    def read_weight_gradients(self, info, reader):
        return ()


class WeightData(Net):
    """net.rs:340-380: a trainable array, recorded as a variable; its gradient info is the variable's id."""

    def __init__(self, data):
        self.data = data

    def get(self):
        return self.data

    def eval_with_gradient_info(self, graph, input=()):
        value = graph.variable(self.data)
        if isinstance(value, tuple):  # Check: values are dims and carry no id
            return value, ()
        return value, value.gid() if _has_id(value) else ()

    def get_weights(self):
        return self.data.clone()  # net.rs:357-359 (an O(1) handle copy; `+=` on the net's own handle is copy-on-write)

    def set_weights(self, weights) -> None:
        _check_equal_dimensions("set_weights", dims_of(weights), dims_of(self.data))
        self.data = weights

    def update_weights(self, delta) -> None:
        _check_equal_dimensions("update_weights", dims_of(delta), dims_of(self.data))
        self.data = weights_add_assign(self.data, delta)  # `self.data += delta`

    def read_weight_gradients(self, info, reader):
        grad = None if isinstance(info, tuple) and info == () else reader.get(info)
        if grad is None:
            raise MissingGradientError(6, "No gradient of the expected type could be found in gradient store.")
        return grad.data() if hasattr(grad, "data") and callable(grad.data) and not isinstance(grad, CuArray) else grad


def _has_id(value) -> bool:
    try:
        return value.id() is not None
    except Exception:
        return False


# ---- combinators (net.rs:382-704) ----------------------------------------------------------------------------------
class Map(Net):
    """net.rs:382-426: apply `f(graph, output)` to the output; weights are the inner net's."""

    def __init__(self, net: Net, f: Callable):
        self.net, self.f = net, f

    def eval_with_gradient_info(self, graph, input):
        output, info = self.net.eval_with_gradient_info(graph, input)
        return self.f(graph, output), info

    def get_weights(self):
        return self.net.get_weights()

    def set_weights(self, weights) -> None:
        self.net.set_weights(weights)

    def update_weights(self, delta) -> None:
        self.net.update_weights(delta)

    def read_weight_gradients(self, info, reader):
        return self.net.read_weight_gradients(info, reader)


class _Pair(Net):
    def __init__(self, first: Net, second: Net):
        self.first, self.second = first, second

    def get_weights(self):
        return (self.first.get_weights(), self.second.get_weights())

    def set_weights(self, weights) -> None:
        self.first.set_weights(weights[0])
        self.second.set_weights(weights[1])

    def update_weights(self, delta) -> None:
        self.first.update_weights(delta[0])
        self.second.update_weights(delta[1])

    def read_weight_gradients(self, info, reader):
        return (self.first.read_weight_gradients(info[0], reader), self.second.read_weight_gradients(info[1], reader))


class Then(_Pair):
    """net.rs:428-493: feed the first net's output to the second."""

    def eval_with_gradient_info(self, graph, input):
        out0, info0 = self.first.eval_with_gradient_info(graph, input)
        out1, info1 = self.second.eval_with_gradient_info(graph, out0)
        return out1, (info0, info1)


class Using(_Pair):
    """net.rs:495-569: pair the first net's output with the output of an input-less net (weights, constants)."""

    def eval_with_gradient_info(self, graph, input):
        out0, info0 = self.first.eval_with_gradient_info(graph, input)
        out1, info1 = self.second.eval_with_gradient_info(graph, ())
        return (out0, out1), (info0, info1)


class TupleNet(Net):
    """net.rs:571-637: a tuple of nets maps a tuple of inputs to a tuple of outputs."""

    def __init__(self, nets: Sequence[Net]):
        self.nets = tuple(nets)

    def eval_with_gradient_info(self, graph, input):
        _check_equal_lengths("eval_with_gradient_info", len(self.nets), len(input))
        results = [n.eval_with_gradient_info(graph, i) for n, i in zip(self.nets, input)]
        return tuple(r[0] for r in results), tuple(r[1] for r in results)

    def get_weights(self):
        return tuple(n.get_weights() for n in self.nets)

    def set_weights(self, weights) -> None:
        _check_equal_lengths("set_weights", len(self.nets), len(weights))
        for n, w in zip(self.nets, weights):
            n.set_weights(w)

    def update_weights(self, delta) -> None:
        _check_equal_lengths("update_weights", len(self.nets), len(delta))
        for n, d in zip(self.nets, delta):
            n.update_weights(d)

    def read_weight_gradients(self, info, reader):
        _check_equal_lengths("read_weight_gradients", len(self.nets), len(info))
        return tuple(n.read_weight_gradients(i, reader) for n, i in zip(self.nets, info))


class VecNet(TupleNet):
    """net.rs:639-704: `Vec<N>`; weights and outputs are lists, length mismatches are `Error::Lengths`."""

    def __init__(self, nets: Sequence[Net]):
        super().__init__(nets)

    def eval_with_gradient_info(self, graph, input):
        out, info = super().eval_with_gradient_info(graph, input)
        return list(out), list(info)

    def get_weights(self):
        return [n.get_weights() for n in self.nets]

    def read_weight_gradients(self, info, reader):
        return list(super().read_weight_gradients(info, reader))


class SquareLoss(Net):
    """net_ext.rs:85-140: input = (net input, target); output = norm2(target - net(input)) — the SQUARED L2 norm,
    no 1/2 and no mean (array.rs:42-44)."""

    def __init__(self, net: Net):
        self.net = net

    def eval_with_gradient_info(self, graph, input):
        x, target = input
        output, info = self.net.eval_with_gradient_info(graph, x)
        _check_equal_dimensions("eval_with_gradient_info", dims_of(output), dims_of(target))
        t = graph.constant(target)
        delta = graph.sub(t, output)
        return graph.norm2(delta), info

    def get_weights(self):
        return self.net.get_weights()

    def set_weights(self, weights) -> None:
        self.net.set_weights(weights)

    def update_weights(self, delta) -> None:
        self.net.update_weights(delta)

    def read_weight_gradients(self, info, reader):
        return self.net.read_weight_gradients(info, reader)
