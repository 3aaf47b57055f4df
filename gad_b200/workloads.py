"""The BASELINE.json workloads as gad programs over the CUDA backend (used by bench.py, smoke() and
the parity tests). Each is written against the algebra interface exactly as a gad user would
(SURVEY §8d): nothing here knows about kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from .algebra import F32, F64, BatchedGraph1, Comm, Context, CuArray, Graph1, GraphN, Value


# ---- C2: batched scalar tapes, Rosenbrock-N per lane -----------------------------
def rosenbrock(g, xs: Sequence[Value]) -> Value:
    terms = []
    for i in range(len(xs) - 1):
        sq = g.mul(xs[i], xs[i])
        d = g.sub(xs[i + 1], sq)
        d2 = g.mul(d, d)
        t1 = g.mulc(d2, 100)
        nx = g.neg(xs[i])
        e = g.addc(nx, 1)
        e2 = g.mul(e, e)
        terms.append(g.add(t1, e2))
    return g.add_all(terms)


def rosenbrock_batched_gradients(ctx: Context, columns: Sequence[CuArray], dtype=F64):
    """columns[i] = x_i over all lanes ([lanes] device column). Returns (tape, f value, [grad columns])."""
    lanes = columns[0].dims()[0]
    g = BatchedGraph1(ctx, lanes, dtype)
    xs = [g.variable(c) for c in columns]
    f = rosenbrock(g, xs)
    store = g.evaluate_gradients_once(f.gid(), 1.0)
    grads = [store.get(x.gid()) for x in xs]
    return g, f, grads, store


# ---- C3: elementwise array tape ------------------------------------------------------------------
def c3_forward(g: Graph1, x: CuArray, y: CuArray):
    vx, vy = g.variable(x), g.variable(y)
    a = g.exp(vx)
    b = g.log(vy)
    c = g.mul(a, b)
    d = g.add(c, vx)
    s = g.sum_as(d, (1,))
    z = g.as_scalar(s)
    return vx, vy, z


def c3_step(ctx: Context, x: CuArray, y: CuArray):
    """Forward + reverse sweep from z with seed 1.0. Returns (z, dz/dx, dz/dy) as device arrays."""
    g = Graph1(ctx)
    vx, vy, z = c3_forward(g, x, y)
    store = g.evaluate_gradients_once(z.gid(), 1.0)
    return z.data(), store.get(vx.gid()), store.get(vy.gid())


# ---- C4: tanh MLP with square loss, batch-sharded data parallel -----------------------------------
def mlp_forward(g, X: CuArray, Ws: Sequence[CuArray], bs: Sequence[CuArray], target: CuArray):
    B = X.dims()[0]
    h = g.constant(X)
    vW = [g.variable(w) for w in Ws]
    vb = [g.variable(b) for b in bs]
    L = len(Ws)
    for l in range(L):
        z = g.matmul_nn(h, vW[l])
        bt = g.tile_as(vb[l], (B, z.dims()[1]))
        zb = g.add(z, bt)
        h = g.tanh(zb) if l + 1 < L else zb
    t = g.constant(target)
    delta = g.sub(t, h)       # SquareLoss: net_ext.rs:110-118
    loss = g.norm2(delta)     # squared L2 norm, no 1/2, no mean (array.rs:42-44)
    return loss, vW, vb


def mlp_step(ctx: Context, X: CuArray, Ws: Sequence[CuArray], bs: Sequence[CuArray], target: CuArray,
             comm: Optional[Comm] = None, overlap: bool = True) -> Tuple[CuArray, List[CuArray], List[CuArray]]:
    """One forward + backward over this rank's rows of the batch; with `comm`, the weight gradients
    and the loss are summed over ranks (the reference sums per-example gradients: net_ext.rs:62-65).
    With `overlap` the sweep is the data-parallel one (gadcu_evaluate_gradients_dp): the dW GEMMs reduce-scatter
    their tiles over NVLink peer memory from the epilogue and the gather runs on the communicator's stream while the
    layers below are still being differentiated; otherwise one grouped NCCL all-reduce after the sweep.
    Returns (loss scalar, dL/dW_l, dL/db_l) as device arrays; nothing is read back to the host."""
    g = Graph1(ctx)
    loss, vW, vb = mlp_forward(g, X, Ws, bs, target)
    lossd = loss.data()
    dp = comm is not None and comm.nranks > 1
    store = g.evaluate_gradients_once(loss.gid(), 1.0, comm if overlap else None, [lossd] if dp and overlap else ())
    gW = [store.get(v.gid()) for v in vW]
    gb = [store.get(v.gid()) for v in vb]
    if dp and not overlap:
        comm.allreduce_sum(gW + gb + [lossd])
    return lossd, gW, gb


def mlp_hvp(ctx: Context, X: CuArray, Ws, bs, target: CuArray, vs: Sequence[CuArray], comm: Optional[Comm] = None, overlap: bool = True):
    """C5: Hessian-vector product of the MLP loss by reverse-over-reverse on GraphN:
    g = grad L; s = sum_l dot(g_{W_l}, v_l); hv = grad s. The loss is a sum over batch rows, so with `comm` each rank
    differentiates its rows twice and the products (and the loss) are summed over the ranks."""
    g = GraphN(ctx)
    loss, vW, vb = mlp_forward(g, X, Ws, bs, target)
    one = g.constant(ctx.scalar(1.0))
    first = g.compute_gradients(loss.gid(), one)
    terms = [g.dot(first.get(w.gid()), g.constant(v)) for w, v in zip(vW, vs)]
    s = g.add_all(terms)
    dp = comm is not None and comm.nranks > 1
    # the second sweep's outputs are what is summed over the ranks: each product travels as soon as the sweep has completed
    # it, while the layers below are still being differentiated (gadcu_compute_gradients_dp)
    second = g.compute_gradients(s.gid(), one, comm if dp and overlap else None)
    # GraphN gradients are VALUES — further ops may be recorded on them — so nothing forces them at the end of the sweep the
    # way Graph1 forces its variables' gradients; the products are this workload's result, so they are computed here
    hv, lossd = [second.get(w.gid()).data().force() for w in vW], loss.data()
    if dp:
        comm.allreduce_sum(([] if overlap else hv) + [lossd])
    return hv, lossd, g.num_nodes()
