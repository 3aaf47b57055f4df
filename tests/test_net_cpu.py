"""CPU: the host logic of gad_b200/net.py (Net combinators, apply_gradient_step, weight (de)serialisation) with
the oracle's algebras standing in for the device ones — the same generic net code runs on both (tests/test_gpu_net.py
is the device half). Replays gad/tests/net.rs:116-154 on the oracle."""
import numpy as np
import pytest

import host_weights as hw
import oracle_py
from gad_b200 import net as gn
from gad_b200.algebra import Check, DimensionsError, LengthsError

A = np.array([1.0, 2.0, 1.0, 1.0, 0.0, 1.0, 0.0, -2.0, -1.0], np.float32).reshape((3, 3), order="F")
I3 = np.eye(3, dtype=np.float32)


def make_net(w0):
    return gn.InputData((3, 3)).using(gn.WeightData(w0)).map(lambda g, iw: g.matmul_nn(iw[0], iw[1]))


def test_combinator_net_trains_to_the_identity_on_the_oracle():
    rng = np.random.default_rng(1)
    train = make_net(hw.HostWeight(rng.standard_normal((3, 3)).astype(np.float32))).add_square_loss()
    for _ in range(20000):
        loss = train.apply_gradient_step(-0.01, [(A, I3)], lambda: hw.OGraph1())
        assert np.isfinite(loss)
        if loss < 1e-6:
            break
    assert loss < 1e-6
    blob = gn.serialize_weights(train.get_weights())
    net = make_net(hw.HostWeight(np.zeros((3, 3), np.float32)))
    back = gn.deserialize_weights(blob)
    net.set_weights(((), hw.HostWeight(back[1])))
    out = net.evaluate(hw.OEval(), A).data().reshape(3, 3)
    assert np.max(np.abs(out - I3)) < 0.01
    # Check is pure host code of the product (dims only, no device)
    assert net.check(Check(), (3, 3, 1, 1)) == (3, 3, 1, 1)


def test_weight_blob_round_trip_and_errors():
    w = ((), [np.arange(6, dtype=np.float32).reshape(2, 3), np.float64([[1.5], [2.5]])], (np.float32([7.0]),))
    back = gn.deserialize_weights(gn.serialize_weights(w))
    assert back[0] == () and isinstance(back[1], list) and isinstance(back[2], tuple)
    assert back[1][0].dtype == np.float32 and np.array_equal(back[1][0][:, :, 0, 0], w[1][0])
    assert back[1][1].dtype == np.float64 and np.array_equal(back[1][1][:, :, 0, 0], w[1][1])
    with pytest.raises(ValueError):
        gn.deserialize_weights(b"nonsense")
    with pytest.raises(LengthsError):
        gn.weights_add_assign((hw.HostWeight(np.zeros(2)),), (hw.HostWeight(np.zeros(2)), hw.HostWeight(np.zeros(2))))
    with pytest.raises(DimensionsError):
        gn.InputData((3, 3)).eval(oracle_py.OEval(), np.zeros((4, 3), np.float32))
    hwts = ((), [hw.HostWeight(w[1][0]), hw.HostWeight(w[1][1])], (hw.HostWeight(w[2][0]),))
    scaled = gn.weights_scale(hwts, 2.0)
    assert np.array_equal(scaled[1][0].a, 2 * w[1][0]) and scaled[0] == ()
    # get_weights() is a snapshot: a later update leaves it alone (net.rs:171-175 `+=` is copy-on-write on af::Array)
    layer = gn.WeightData(hw.HostWeight(w[1][0]))
    snap = layer.get_weights()
    layer.update_weights(hw.HostWeight(np.ones_like(w[1][0])))
    assert np.array_equal(snap.a, w[1][0]) and np.array_equal(layer.get_weights().a, w[1][0] + 1)


def test_bincode_layout_of_weights():
    """`bincode::serialize(&net.get_weights())` (gad/tests/net.rs:136-149) for the net of tests/net.rs:80-89, whose
    weights are `((), af::Array<f32>)`: no bytes for `()`, then u32 DType index, Dim4 as 4 x u64, u64 length, data."""
    import struct
    a = np.arange(6, dtype=np.float32).reshape((2, 3), order="F")
    blob = gn.serialize_weights_bincode(((), a))
    assert len(blob) == 4 + 32 + 8 + 24
    assert struct.unpack("<I4QQ", blob[:44]) == (0, 2, 3, 1, 1, 6)
    assert np.array_equal(np.frombuffer(blob[44:], "<f4"), np.arange(6, dtype=np.float32))  # column-major order
    back = gn.deserialize_weights_bincode(blob, like=((), a))
    assert back[0] == () and np.array_equal(back[1][:, :, 0, 0], a)
    # f64 is DType index 2 (F32, C32, F64, ...); a Vec of nets carries a u64 length
    d = np.float64([[1.5], [2.5]])
    blob2 = gn.serialize_weights_bincode([d, d])
    assert struct.unpack("<Q", blob2[:8]) == (2,) and struct.unpack("<I", blob2[8:12]) == (2,)
    back2 = gn.deserialize_weights_bincode(blob2, like=[d, d])
    assert isinstance(back2, list) and back2[1].dtype == np.float64 and np.array_equal(back2[1][:, :, 0, 0], d)
    with pytest.raises(LengthsError):
        gn.deserialize_weights_bincode(blob2, like=[d])
    with pytest.raises(ValueError):
        gn.deserialize_weights_bincode(blob[:-1], like=((), a))
