"""CPU: the two things the oracle carries beyond the reference's behaviour, both off by default.

* `corrected_matmul_adjoint`: the true adjoint for transposed matmul operands. With it the oracle's GraphN can run C5
  (Hessian-vector products of the MLP loss), which the reference's own VJP (gad/src/matrix.rs:164-176) cannot for B != D;
  checked here against an independent f64 R-operator closed form, itself checked against central differences of the f64
  gradient. Without it the oracle raises what the reference would.
* `set_blas`: the CPU BASELINE's matmul through a BLAS sgemm (what ArrayFire's CPU backend calls). Same values as the
  documented loop up to summation order; never on in parity checks (the switch is restored here).
"""
import numpy as np
import pytest

import mlp_f64
import oracle_py


def _vs(seed, D):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((D, D)).astype(np.float32) for _ in range(3)]


def test_f64_hvp_closed_form_matches_central_differences():
    B, D = 48, 24
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    Vs = _vs(10, D)
    hv = mlp_f64.mlp_hvp_f64(X, Ws, bs, T, Vs)
    eps = 1e-5
    _, gp, _ = mlp_f64.mlp_grad_f64(X, [w.astype(np.float64) + eps * v for w, v in zip(Ws, Vs)], bs, T)
    _, gm, _ = mlp_f64.mlp_grad_f64(X, [w.astype(np.float64) - eps * v for w, v in zip(Ws, Vs)], bs, T)
    for l in range(3):
        fd = (gp[l] - gm[l]) / (2 * eps)
        assert np.abs(fd - hv[l]).max() <= 1e-6 * np.abs(hv[l]).max()


@pytest.mark.parametrize("B,D", [(64, 32), (40, 56)])
def test_oracle_c5_with_the_corrected_adjoint_matches_the_f64_closed_form(B, D):
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    Vs = _vs(11, D)
    loss, hv, _, nodes = oracle_py.mlp_hvp(X, Ws, bs, T, Vs, nthreads=2)
    want = mlp_f64.mlp_hvp_f64(X, Ws, bs, T, Vs)
    wl, _, _ = mlp_f64.mlp_grad_f64(X, Ws, bs, T)
    assert abs(loss - wl) <= 1e-5 * abs(wl)
    assert nodes > 19  # the two recorded backward passes grew the tape
    for l in range(3):
        assert np.abs(hv[l] - want[l]).max() <= 2e-5 * np.abs(want[l]).max(), l
    assert oracle_py.set_corrected_adjoint(False) is False  # the workload restored the reference's mode


def test_reference_mode_cannot_differentiate_through_a_transposed_operand():
    """matrix.rs:164-176 on C = A^T B with A [3, 4]: dA = matmul(G, B, p1, p2^T) has incompatible dimensions."""
    from gad_b200 import MatProp
    og = oracle_py.OGraphN()
    a = og.variable(np.ones((3, 4), np.float32))
    b = og.variable(np.ones((3, 5), np.float32))
    c = og.matmul(a, b, MatProp(True), MatProp())
    with pytest.raises(oracle_py.OracleError) as e:
        og.compute_gradients(c.gid(), og.constant(np.ones((4, 5), np.float32)))
    assert e.value.kind == "Dimensions"
    before = oracle_py.set_corrected_adjoint(True)
    try:
        og2 = oracle_py.OGraphN()
        a2, b2 = og2.variable(np.ones((3, 4), np.float32)), og2.variable(2 * np.ones((3, 5), np.float32))
        c2 = og2.matmul(a2, b2, MatProp(True), MatProp())
        store = og2.compute_gradients(c2.gid(), og2.constant(np.ones((4, 5), np.float32)))
        ga = store.get(a2.gid(), like=a2).data()[:, :, 0, 0]
        assert ga.shape == (3, 4) and np.array_equal(ga, np.full((3, 4), 10.0, np.float32))  # dA = B G^T: 5 columns of 2
    finally:
        oracle_py.set_corrected_adjoint(before)


def test_blas_baseline_agrees_with_the_documented_loop():
    path = oracle_py.find_blas()
    if path is None:
        pytest.skip("no BLAS on this host (the baseline then times the port's own loop and says so)")
    B, D = 96, 128
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    l0, gW0, gb0, _ = oracle_py.mlp_step(X, Ws, bs, T, nthreads=2)
    assert oracle_py.set_blas(path, 2)
    try:
        l1, gW1, gb1, _ = oracle_py.mlp_step(X, Ws, bs, T, nthreads=2)
    finally:
        oracle_py.set_blas(None)
    assert abs(l0 - l1) <= 1e-5 * abs(l0)
    for a, b in zip(gW0 + gb0, gW1 + gb1):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(a).max()
    l2, gW2, _, _ = oracle_py.mlp_step(X, Ws, bs, T, nthreads=2)
    assert l2 == l0 and all(np.array_equal(a, b) for a, b in zip(gW0, gW2))  # switched back: the loop's bits again
