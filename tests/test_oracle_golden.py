"""CPU: the oracle against the reference's golden vectors / known-answer tests."""
import json
import os
import subprocess

import numpy as np
import pytest

import programs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")) as f:
    GOLDEN = json.load(f)


def test_oracle_acceptance_suite_replays_reference_tests():
    """oracle/test_oracle.cc: gad/tests/{graph,symbolic,core,arith,const_arith,analytic,compare,
    extension}.rs + lib.rs doctests, with the reference's exact expected values (i32 exact, symbolic
    strings exact, GraphN Hessian exact incl. absent entries)."""
    res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "test"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failures" in res.stdout


def test_golden_file_is_consistent_with_its_provenance():
    res = subprocess.run(["python", os.path.join(ROOT, "tests", "golden", "make_golden.py")], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout


@pytest.mark.parametrize("case", GOLDEN["scalar_f32_direction_1p7"], ids=lambda c: c["name"])
def test_scalar_analytic_known_answers(oracle, case):
    g = oracle.OGraph1()
    names = ["a", "b"]
    vs = {n: g.variable(np.float32(v)) for n, v in zip(names, case["inputs"])}
    out = programs.run_expr(g, case["program"], **vs)
    assert abs(float(out.data()) - case["value"]) < case["tol"]
    store = g.evaluate_gradients_once(out.gid(), np.float32(1.7))
    for n, want in zip(names, case["grads"]):
        got = store.get(vs[n].gid(), like=vs[n])
        assert got is not None
        assert abs(float(got.data()) - want) < case["tol"]


def test_doctest_mul_exact(oracle):
    g = oracle.OGraph1()
    a, b = g.variable(np.float32(1.0)), g.variable(np.float32(2.0))
    c = g.mul(a, b)
    store = g.evaluate_gradients(c.gid(), np.float32(1.0))
    assert store.get(a.gid(), like=a).data() == np.float32(2.0)
    assert store.get(b.gid(), like=b).data() == np.float32(1.0)


@pytest.mark.parametrize("through_arrays", [False, True])
def test_hessian_and_more(oracle, through_arrays):
    """gad/tests/graph.rs:28-69 and :107-153 — exact f32 equality, absent entries stay absent."""
    h = GOLDEN["hessian_f32"]
    g = oracle.OGraphN()
    x, y = g.variable(np.float32(h["x"])), g.variable(np.float32(h["y"]))
    if through_arrays:
        xa, ya = g.constant_as(x, (1,)), g.constant_as(y, (1,))
        z = g.as_scalar(g.matmul_nn(g.matmul_nn(xa, ya), ya))
    else:
        z = g.mul(g.mul(x, y), y)
    dz = g.compute_gradients(z.gid(), g.constant(np.float32(1.0)))
    dz_dx, dz_dy = dz.get(x.gid(), like=x), dz.get(y.gid(), like=y)
    one = g.constant(np.float32(1.0))
    ddx = g.compute_gradients(dz_dx.gid(), one)
    ddy = g.compute_gradients(dz_dy.gid(), one)
    assert ddx.get(x.gid(), like=x) is None
    assert ddx.get(y.gid(), like=y).data() == np.float32(h["ddz_dxdy"])
    assert ddy.get(x.gid(), like=x).data() == np.float32(h["ddz_dydx"])
    assert ddy.get(y.gid(), like=y).data() == np.float32(h["ddz_dydy"])
    dddx = g.compute_gradients(ddx.get(y.gid(), like=y).gid(), one)
    assert dddx.get(x.gid(), like=x) is None
    assert dddx.get(y.gid(), like=y).data() == np.float32(h["dddz_dxdydy"])


def test_rosenbrock_closed_form(oracle):
    r = GOLDEN["rosenbrock_f64"]
    x = np.array(r["n2"]["x"], dtype=np.float64).reshape(2, 1)
    f, grad, _ = oracle.rosenbrock_batch(x)
    assert abs(f[0] - r["n2"]["f"]) < r["n2"]["tol"]
    assert np.allclose(grad[:, 0], r["n2"]["grad"], atol=r["n2"]["tol"], rtol=0)
    ones = np.ones((64, 3))
    f, grad, _ = oracle.rosenbrock_batch(ones)
    assert np.all(f == 0.0) and np.all(grad == 0.0)
    # generic tape through the C API == the bulk entry point, and both match the closed form
    rng = np.random.default_rng(1)
    x = -1.2 + 2.4 * rng.random((8, 5))
    _, grad, _ = oracle.rosenbrock_batch(x)
    assert np.allclose(grad, programs.rosenbrock_grad_closed_form(x), rtol=1e-12, atol=1e-10)
    g = oracle.OGraph1()
    xs = [g.variable(np.float64(v)) for v in x[:, 0]]
    fv = programs.rosenbrock(g, xs)
    assert g.num_nodes() == 8 + 8 * 7 + 1
    store = g.evaluate_gradients_once(fv.gid(), np.float64(1.0))
    got = np.array([store.get(v.gid(), like=v).data() for v in xs])
    assert np.array_equal(got, grad[:, 0])


def test_structure_pins(oracle):
    """zeros/ones/sign/argmax_as are constants; select_argmax leaves a, b absent from the store."""
    g = oracle.OGraph1()
    a = g.variable(np.float32(3.0))
    assert g.zeros(a).id() is None and g.ones(a).id() is None and g.sign(a).id() is None
    arr = g.variable(np.arange(12, dtype=np.float32).reshape(4, 3))
    assert g.argmax_as(arr, (1, 3)).id() is None
    b, c, d = g.variable(np.float32(2.0)), g.variable(np.float32(3.0)), g.variable(np.float32(4.0))
    a1 = g.variable(np.float32(1.0))
    e = g.select_argmax(a1, b, c, d)
    assert e.data() == np.float32(4.0)
    store = g.evaluate_gradients_once(e.gid(), np.float32(1.0))
    assert store.get(a1.gid(), like=a1) is None and store.get(b.gid(), like=b) is None
    assert store.get(c.gid(), like=c).data() == 0.0 and store.get(d.gid(), like=d).data() == 1.0


def test_oracle_reproduces_its_array_fixtures(oracle):
    """tests/golden/oracle_array_fixtures.npz (made by tests/golden/make_array_fixtures.py): the oracle's outputs for the
    array workloads at toy sizes. Recomputed bit for bit (no drift of the checker), and the stored values themselves are
    checked against f64 closed forms."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fx = np.load(os.path.join(here, "oracle_array_fixtures.npz"))
    spec = importlib.util.spec_from_file_location("make_array_fixtures", os.path.join(here, "make_array_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    inp = mod.inputs()
    for k, v in inp.items():
        assert np.array_equal(fx[k], v), k  # the seeded inputs are what the file says
    out = mod.outputs(inp)
    for k, v in out.items():
        assert np.array_equal(fx[k], v), k
    # closed forms in f64
    x, y = fx["c3_x"].astype(np.float64), fx["c3_y"].astype(np.float64)
    eps = np.finfo(np.float32).eps
    assert np.all(np.abs(fx["c3_gx"] - (np.exp(x) * np.log(y) + 1)) <= 6 * eps * (np.abs(np.exp(x) * np.log(y)) + 1))
    assert np.allclose(fx["c3_gy"], np.exp(x) / y, rtol=8 * eps, atol=0)
    assert abs(float(fx["c3_z"]) - float(np.sum(np.exp(x) * np.log(y) + x))) <= 64 * x.size * eps
    X, W, b, T = (fx[k].astype(np.float64) for k in ["mlp_X", "mlp_W", "mlp_b", "mlp_T"])
    h, hs = X, []
    for l in range(3):
        z = h @ W[l] + b[l]
        hs.append(h)
        h = np.tanh(z) if l < 2 else z
    delta = T - h
    assert abs(float(fx["mlp_loss"]) - float(np.sum(delta * delta))) <= 1e-5 * float(np.sum(delta * delta))
    g = -2 * delta
    for l in (2, 1, 0):
        assert np.allclose(fx["mlp_gW"][l], hs[l].T @ g, rtol=1e-4, atol=1e-5), l
        assert np.allclose(fx["mlp_gb"][l].reshape(-1), g.sum(axis=0), rtol=1e-4, atol=1e-5), l
        if l:
            g = (g @ W[l].T) * (1 - np.tanh(hs[l - 1] @ W[l - 1] + b[l - 1]) ** 2)
    rx = fx["ros_x"]
    import programs
    assert np.allclose(fx["ros_g"], programs.rosenbrock_grad_closed_form(rx), rtol=1e-11, atol=1e-9)
