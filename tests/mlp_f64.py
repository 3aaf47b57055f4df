"""f64 closed forms of the C4 / C5 workloads in numpy (test infrastructure): the second reference beside the oracle.

The network is SURVEY §8d's: h_0 = X, z_l = h_{l-1} W_l + b_l, h_l = tanh(z_l) for hidden layers and h_L = z_L,
loss = |T - h_L|^2 (gad's norm2 is the SQUARED norm, array.rs:42-44). Everything row-major numpy of shape (B, D);
the device's column-major [B, D] arrays read back with .numpy() have the same indexing.
"""
import numpy as np


def mlp_inputs(B, D, L=3, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, D)).astype(np.float32)
    Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(L)]
    bs = [np.full((1, D), 0.01, np.float32) for _ in range(L)]
    T = rng.standard_normal((B, D)).astype(np.float32)
    return X, Ws, bs, T


def mlp_forward_f64(X, Ws, bs):
    """Activations h_0 .. h_L in f64."""
    h = X.astype(np.float64)
    acts = [h]
    for l, (W, b) in enumerate(zip(Ws, bs)):
        z = h @ W.astype(np.float64) + b.astype(np.float64)
        h = np.tanh(z) if l + 1 < len(Ws) else z
        acts.append(h)
    return acts


def mlp_deltas_f64(acts, Ws, T):
    """delta_l = dLoss/dz_l for l = 1..L (index l-1), and the loss."""
    L = len(Ws)
    d = T.astype(np.float64) - acts[L]
    loss = float(np.sum(d * d))
    g = -2.0 * d
    deltas = [None] * L
    for l in reversed(range(L)):
        if l + 1 < L:
            g = g * (1 - acts[l + 1] ** 2)
        deltas[l] = g
        if l > 0:
            g = g @ Ws[l].astype(np.float64).T
    return deltas, loss


def mlp_grad_f64(X, Ws, bs, T):
    """(loss, [dLoss/dW_l], [dLoss/db_l]) in f64."""
    acts = mlp_forward_f64(X, Ws, bs)
    deltas, loss = mlp_deltas_f64(acts, Ws, T)
    gW = [acts[l].T @ deltas[l] for l in range(len(Ws))]
    gb = [deltas[l].sum(axis=0, keepdims=True) for l in range(len(Ws))]
    return loss, gW, gb


def mlp_hvp_f64(X, Ws, bs, T, Vs):
    """Hessian-vector product of the loss w.r.t. the weights in direction Vs (biases held fixed), by the R-operator
    (forward-over-reverse; Pearlmutter 1994) — independent of how the device gets it (reverse over reverse on GraphN).
    Returns [R{dLoss/dW_l}]."""
    L = len(Ws)
    W = [w.astype(np.float64) for w in Ws]
    V = [v.astype(np.float64) for v in Vs]
    acts = mlp_forward_f64(X, Ws, bs)
    deltas, _ = mlp_deltas_f64(acts, Ws, T)
    # forward R pass
    Rh = [np.zeros_like(acts[0])]
    for l in range(L):
        Rz = Rh[l] @ W[l] + acts[l] @ V[l]
        Rh.append((1 - acts[l + 1] ** 2) * Rz if l + 1 < L else Rz)
    # backward R pass
    Rdelta = [None] * L
    Rg = 2.0 * Rh[L]  # R{dLoss/dh_L}
    for l in reversed(range(L)):
        if l + 1 < L:
            # delta_l = u_l * (1 - h_l^2) with u_l = delta_{l+1} W_{l+1}^T ; Rg holds R{u_l}
            u = deltas[l + 1] @ W[l + 1].T
            Rdelta[l] = Rg * (1 - acts[l + 1] ** 2) + u * (-2.0 * acts[l + 1] * Rh[l + 1])
        else:
            Rdelta[l] = Rg
        if l > 0:
            Rg = Rdelta[l] @ W[l].T + deltas[l] @ V[l].T
    return [Rh[l].T @ deltas[l] + acts[l].T @ Rdelta[l] for l in range(L)]
