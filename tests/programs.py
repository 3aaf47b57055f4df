"""Generic gad programs, written once against the algebra interface and run on any backend
(gad_b200's Eval/Graph1/GraphN over CUDA, or the oracle's) — the reference's programming style
(gad/src/lib.rs:154-161)."""
import numpy as np


def run_expr(g, expr: str, **values):
    """Evaluate e.g. "exp(mulc(a,2))" with the algebra's methods; variables come from `values`."""
    ns = {n: getattr(g, n) for n in dir(g) if not n.startswith("_") and callable(getattr(g, n))}
    ns.update(values)
    return eval(expr, {"__builtins__": {}}, ns)


def rosenbrock(g, xs):
    """f(x) = sum_i 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2, 8 nodes per term."""
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


def rosenbrock_grad_closed_form(x):
    """x: [n, lanes]; closed-form gradient (f64)."""
    n = x.shape[0]
    g = np.zeros_like(x)
    g[:-1] += -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (1.0 - x[:-1])
    g[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
    return g


def c3_tape(g, x, y):
    """BASELINE config 3: z = sum(exp(x) * log(y) + x); returns (vx, vy, z)."""
    vx, vy = g.variable(x), g.variable(y)
    a = g.exp(vx)
    b = g.log(vy)
    c = g.mul(a, b)
    d = g.add(c, vx)
    s = g.sum_as(d, (1,))
    z = g.as_scalar(s)
    return vx, vy, z


def mlp_loss(g, X, Ws, bs, target):
    """BASELINE config 4: tanh MLP with square loss; X and target are constants, W_l/b_l variables."""
    B, D = X.dims()[:2] if hasattr(X, "dims") else X.shape[:2]
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
    delta = g.sub(t, h)
    loss = g.norm2(delta)
    return loss, vW, vb
