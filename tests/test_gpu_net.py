"""GPU: gad's Net combinators + DiffNet::apply_gradient_step over the CUDA backend —
the reference's tests/net.rs replayed (3x3 linear map trained to the identity), weight (de)serialisation
round trip, Check over the same net, and parity of the training trajectory with the CPU oracle."""
import numpy as np
import pytest

import gad_b200 as gb
import host_weights as hw
from gad_b200 import net as gn
from helpers import to_np

pytestmark = pytest.mark.gpu

A = np.array([1.0, 2.0, 1.0, 1.0, 0.0, 1.0, 0.0, -2.0, -1.0], np.float32).reshape((3, 3), order="F")  # tests/net.rs:94-97
I3 = np.eye(3, dtype=np.float32)


class TestNet(gn.Net):
    """tests/net.rs:9-78: a hand-written net, output = input . weights."""
    __test__ = False

    def __init__(self, dims, weights):
        self.dims, self.weights = dims, weights

    def eval_with_gradient_info(self, g, input):
        assert gn.dims_of(input) == self.dims
        x = g.constant(input)
        w = g.variable(self.weights)
        out = g.matmul_nn(x, w)
        return out, (w.gid() if not isinstance(w, tuple) and w.id() is not None else ())

    def get_weights(self):
        return self.weights.clone()

    def update_weights(self, delta):
        self.weights = gn.weights_add_assign(self.weights, delta)

    def set_weights(self, weights):
        assert gn.dims_of(weights) == gn.dims_of(self.weights)
        self.weights = weights

    def read_weight_gradients(self, info, reader):
        g = reader.get(info)
        assert g is not None
        return g


def make_net(w0):  # tests/net.rs:80-89
    return gn.InputData((3, 3)).using(gn.WeightData(w0)).map(lambda g, iw: g.matmul_nn(iw[0], iw[1]))


def train(net, samples, new_graph, max_steps=20000):
    for step in range(max_steps):
        loss = net.apply_gradient_step(-0.01, samples, new_graph)
        assert np.isfinite(loss)
        if loss < 1e-6:
            return step
    raise AssertionError(f"did not converge: loss {loss}")


def test_testnet_trains_to_the_identity(ctx):
    """tests/net.rs:91-114."""
    rng = np.random.default_rng(0)
    w0 = rng.standard_normal((3, 3)).astype(np.float32)
    train_net = TestNet((3, 3, 1, 1), ctx.array(w0)).add_square_loss()
    samples = [(ctx.array(A), ctx.array(I3))]
    train(train_net, samples, lambda: gb.Graph1(ctx))
    net = TestNet((3, 3, 1, 1), ctx.array(rng.standard_normal((3, 3)).astype(np.float32)))
    net.set_weights(train_net.get_weights())
    i2 = to_np(net.evaluate(gb.Eval(ctx), ctx.array(A)))[:, :, 0, 0]
    assert np.max(np.abs(i2 - I3)) < 0.01  # testing::assert_almost_all_equal(&i, &i2, 0.01)


def test_combinator_net_serialises_and_checks(ctx):
    """tests/net.rs:116-154: combinator net, weights through a byte round trip, Check returns the evaluated dims."""
    rng = np.random.default_rng(1)
    train_net = make_net(ctx.array(rng.standard_normal((3, 3)).astype(np.float32))).add_square_loss()
    samples = [(ctx.array(A), ctx.array(I3))]
    train(train_net, samples, lambda: gb.Graph1(ctx))
    blob = gn.serialize_weights(train_net.get_weights())
    weights = gn.deserialize_weights(blob, ctx)
    net = make_net(ctx.array(np.zeros((3, 3), np.float32)))
    net.set_weights(weights)
    i2 = to_np(net.evaluate(gb.Eval(ctx), ctx.array(A)))[:, :, 0, 0]
    assert np.max(np.abs(i2 - I3)) < 0.01
    # Check dimensions (net.rs:129-134): same net, dims only
    net2 = make_net(ctx.array(np.zeros((3, 3), np.float32)))
    net2.set_weights(gn.deserialize_weights(blob, ctx))
    assert net2.check(gb.Check(), (3, 3, 1, 1)) == (3, 3, 1, 1)
    # the weight structure mirrors the net: ((), W) for input.using(weight)
    w = train_net.get_weights()
    assert w[0] == () and isinstance(w[1], gb.CuArray)
    # errors are values: wrong input dims, wrong weight dims
    with pytest.raises(gb.DimensionsError):
        net.evaluate(gb.Eval(ctx), ctx.array(np.zeros((4, 3), np.float32)))
    with pytest.raises(gb.DimensionsError):
        net.set_weights(((), ctx.array(np.zeros((4, 3), np.float32))))


def test_training_trajectory_matches_the_oracle(ctx, oracle):
    """The same generic net, same initial weights and batch, stepped on the CUDA Graph1 and on the CPU oracle's
    Graph1: losses and weights agree to f32 rounding after 25 steps (two examples per batch: the per-example
    gradients are accumulated in example order, net_ext.rs:62-65)."""
    rng = np.random.default_rng(2)
    w0 = rng.standard_normal((3, 3)).astype(np.float32)
    B = rng.standard_normal((3, 3)).astype(np.float32)
    dev = make_net(ctx.array(w0)).add_square_loss()
    cpu = make_net(hw.HostWeight(w0)).add_square_loss()
    dev_batch = [(ctx.array(A), ctx.array(I3)), (ctx.array(B), ctx.array(np.float32(0.5) * B))]
    cpu_batch = [(A, I3), (B, np.float32(0.5) * B)]
    for _ in range(25):
        ld = dev.apply_gradient_step(-0.01, dev_batch, lambda: gb.Graph1(ctx))
        lc = cpu.apply_gradient_step(-0.01, cpu_batch, lambda: hw.OGraph1())
        assert abs(ld - lc) <= 1e-5 * max(1.0, abs(lc))
    wd = dev.get_weights()[1].numpy()[:, :, 0, 0]
    wc = np.asarray(cpu.get_weights()[1]).reshape(3, 3)
    # the reference's checkpoint layout (bincode + afserde) round-trips through the device
    blob = gn.serialize_weights_bincode(dev.get_weights())
    back = gn.deserialize_weights_bincode(blob, like=dev.get_weights(), ctx=ctx)
    assert np.array_equal(back[1].numpy(), dev.get_weights()[1].numpy())
    assert np.max(np.abs(wd - wc)) <= 1e-5


def test_tuple_vec_then_and_weight_ops(ctx):
    """Then / tuple / Vec combinators and WeightOps over the nested weight structure (net.rs:428-704)."""
    rng = np.random.default_rng(3)
    w1, w2 = rng.standard_normal((3, 3)).astype(np.float32), rng.standard_normal((3, 3)).astype(np.float32)
    layer = lambda w: gn.InputData((3, 3)).using(gn.WeightData(ctx.array(w))).map(lambda g, iw: g.tanh(g.matmul_nn(iw[0], iw[1])))

    class Feed(gn.Net):  # a net whose input is already a value (the output of the previous one)
        def __init__(self, w):
            self.w = gn.WeightData(ctx.array(w))

        def eval_with_gradient_info(self, g, x):
            wv, info = self.w.eval_with_gradient_info(g, ())
            return g.matmul_nn(x, wv), info

        def get_weights(self): return self.w.get_weights()
        def set_weights(self, w): self.w.set_weights(w)
        def update_weights(self, d): self.w.update_weights(d)
        def read_weight_gradients(self, info, reader): return self.w.read_weight_gradients(info, reader)

    two = layer(w1).then(Feed(w2)).add_square_loss()
    x = ctx.array(A)
    g = gb.Graph1(ctx)
    loss, info = two.eval_with_gradient_info(g, (x, ctx.array(I3)))
    store = g.evaluate_gradients_once(loss.gid(), 1.0)
    grads = two.read_weight_gradients(info, store)
    assert grads[0][0] == () and isinstance(grads[0][1], gb.CuArray) and isinstance(grads[1], gb.CuArray)
    # against f64 numpy
    h = np.tanh(A.astype(np.float64) @ w1)
    out = h @ w2
    d = I3 - out
    gw2 = h.T @ (-2 * d)
    gh = (-2 * d) @ w2.T.astype(np.float64)
    gw1 = A.astype(np.float64).T @ (gh * (1 - h * h))
    assert np.allclose(grads[1].numpy()[:, :, 0, 0], gw2, rtol=1e-4, atol=1e-5)
    assert np.allclose(grads[0][1].numpy()[:, :, 0, 0], gw1, rtol=1e-4, atol=1e-5)
    scaled = gn.weights_scale(grads, -0.5)
    assert np.allclose(scaled[1].numpy(), -0.5 * grads[1].numpy())
    summed = gn.weights_add_assign(gn.weights_scale(grads, 1.0), grads)
    assert np.allclose(summed[0][1].numpy(), 2 * grads[0][1].numpy())
    # tuple / Vec of nets; Lengths error on a mismatch (net.rs:646)
    pair = gn.ConstantData(ctx.array(w1)).and_(gn.WeightData(ctx.array(w2)))
    (c, w), (ic, iw) = pair.eval_with_gradient_info(gb.Graph1(ctx), ((), ()))
    assert c.id() is None and w.id() is not None and ic == ()
    vec = gn.VecNet([gn.WeightData(ctx.array(w1)), gn.WeightData(ctx.array(w2))])
    assert isinstance(vec.get_weights(), list) and len(vec.get_weights()) == 2
    with pytest.raises(gb.LengthsError):
        vec.eval_with_gradient_info(gb.Graph1(ctx), [()])
    blob = gn.serialize_weights(vec.get_weights())
    back = gn.deserialize_weights(blob)
    assert isinstance(back, list) and np.array_equal(back[0][:, :, 0, 0], w1)
