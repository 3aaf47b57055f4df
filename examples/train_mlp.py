"""Train a small tanh MLP with gad's `Net` combinators on the CUDA backend (gad/src/net.rs, net_ext.rs):

    python examples/train_mlp.py            # needs a B200

The net is written once against the algebra interface and runs on Check (dims), Eval (inference) and Graph1 (training).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gad_b200 as gb
from gad_b200 import net as gn


def layer(ctx, n_in, n_out, rng, activation=True):
    w = gn.WeightData(ctx.array((rng.standard_normal((n_in, n_out)) / np.sqrt(n_in)).astype(np.float32)))
    b = gn.WeightData(ctx.array(np.zeros((1, n_out), np.float32)))

    class Layer(gn.Net):  # x -> tanh(x W + b)
        def eval_with_gradient_info(self, g, x):
            (wv, bv), info = w.and_(b).eval_with_gradient_info(g, ((), ()))
            z = g.matmul_nn(x, wv)
            z = g.add(z, g.tile_as(bv, gn.dims_of(z)))
            return (g.tanh(z) if activation else z), info

        def get_weights(self): return (w.get_weights(), b.get_weights())
        def set_weights(self, ws): w.set_weights(ws[0]); b.set_weights(ws[1])
        def update_weights(self, d): w.update_weights(d[0]); b.update_weights(d[1])
        def read_weight_gradients(self, info, reader): return (w.read_weight_gradients(info[0], reader), b.read_weight_gradients(info[1], reader))

    return Layer()


def main():
    ctx = gb.Context(0)
    rng = np.random.default_rng(0)
    B, D, H = 256, 8, 32
    X = rng.uniform(-1, 1, (B, D)).astype(np.float32)
    Y = np.sin(X.sum(axis=1, keepdims=True)).astype(np.float32)
    model = gn.InputData((B, D)).then(layer(ctx, D, H, rng)).then(layer(ctx, H, 1, rng, activation=False))
    print("check():", model.check(gb.Check(), (B, D, 1, 1)))
    train = model.add_square_loss()
    batch = [(ctx.array(X), ctx.array(Y))]
    for step in range(301):
        loss = train.apply_gradient_step(-2e-4, batch, lambda: gb.Graph1(ctx))
        if step % 50 == 0:
            print(f"step {step:4d}  loss {loss:.5f}")
    pred = model.evaluate(gb.Eval(ctx), ctx.array(X)).data().numpy()[:, :, 0, 0]
    print("rms error:", float(np.sqrt(np.mean((pred - Y) ** 2))))


if __name__ == "__main__":
    main()
