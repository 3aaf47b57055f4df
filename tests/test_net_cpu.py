"""CPU: the host logic of gad_b200/net.py (Net combinators, apply_gradient_step, weight (de)serialisation) with
the oracle's algebras standing in for the device ones — the same generic net code runs on both (tests/test_gpu_net.py
is the device half). Replays gad/tests/net.rs:116-154 on the oracle."""
import numpy as np
import pytest

import oracle_py
from gad_b200 import net as gn
from gad_b200.algebra import Check, DimensionsError, LengthsError

A = np.array([1.0, 2.0, 1.0, 1.0, 0.0, 1.0, 0.0, -2.0, -1.0], np.float32).reshape((3, 3), order="F")
I3 = np.eye(3, dtype=np.float32)


def make_net(w0):
    return gn.InputData((3, 3)).using(gn.WeightData(w0)).map(lambda g, iw: g.matmul_nn(iw[0], iw[1]))


def test_combinator_net_trains_to_the_identity_on_the_oracle():
    rng = np.random.default_rng(1)
    train = make_net(rng.standard_normal((3, 3)).astype(np.float32)).add_square_loss()
    for _ in range(20000):
        loss = train.apply_gradient_step(-0.01, [(A, I3)], lambda: oracle_py.OGraph1())
        assert np.isfinite(loss)
        if loss < 1e-6:
            break
    assert loss < 1e-6
    blob = gn.serialize_weights(train.get_weights())
    net = make_net(np.zeros((3, 3), np.float32))
    net.set_weights(gn.deserialize_weights(blob))
    out = net.evaluate(oracle_py.OEval(), A).data().reshape(3, 3)
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
        gn.weights_add_assign((np.zeros(2),), (np.zeros(2), np.zeros(2)))
    with pytest.raises(DimensionsError):
        gn.InputData((3, 3)).eval(oracle_py.OEval(), np.zeros((4, 3), np.float32))
    scaled = gn.weights_scale(w, 2.0)
    assert np.array_equal(scaled[1][0], 2 * w[1][0]) and scaled[0] == ()
