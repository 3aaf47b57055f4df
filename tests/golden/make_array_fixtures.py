"""Generates tests/golden/oracle_array_fixtures.npz: seeded inputs of the array workloads at toy sizes and the ORACLE's
outputs for them. The reference itself (Rust + the arrayfire crate) cannot run in this image and ships no golden vectors
for array arithmetic, so these fixtures pin the oracle — the checker every GPU parity test compares against — against
drift: tests/test_oracle_golden.py::test_oracle_reproduces_its_array_fixtures recomputes them bit for bit and checks them
against f64 closed forms.   python tests/golden/make_array_fixtures.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py  # noqa: E402


def inputs():
    rng = np.random.default_rng(20261020)
    n = 257
    c3_x = rng.uniform(-1, 1, n).astype(np.float32)
    c3_y = rng.uniform(0.5, 2, n).astype(np.float32)
    B, D = 8, 4
    mlp_X = rng.standard_normal((B, D)).astype(np.float32)
    mlp_W = np.stack([(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(3)])
    mlp_b = np.full((3, 1, D), 0.01, np.float32)
    mlp_T = rng.standard_normal((B, D)).astype(np.float32)
    ros_x = -1.2 + 2.4 * rng.random((8, 5))
    return dict(c3_x=c3_x, c3_y=c3_y, mlp_X=mlp_X, mlp_W=mlp_W, mlp_b=mlp_b, mlp_T=mlp_T, ros_x=ros_x)


def outputs(i):
    z, gx, gy, _, _ = oracle_py.c3(i["c3_x"], i["c3_y"])
    loss, gW, gb, _ = oracle_py.mlp_step(i["mlp_X"], list(i["mlp_W"]), list(i["mlp_b"]), i["mlp_T"], nthreads=1)
    f, g, _ = oracle_py.rosenbrock_batch(i["ros_x"], nthreads=1)
    return dict(c3_z=np.float32(z), c3_gx=gx, c3_gy=gy, mlp_loss=np.float32(loss), mlp_gW=np.stack(gW), mlp_gb=np.stack(gb), ros_f=f, ros_g=g)


if __name__ == "__main__":
    i = inputs()
    o = outputs(i)
    np.savez(os.path.join(HERE, "oracle_array_fixtures.npz"), **i, **o)
    print({k: (v.shape, v.dtype) for k, v in {**i, **o}.items()})
