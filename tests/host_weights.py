"""TEST ADAPTER: lets gad_b200/net.py's generic Net code run on the CPU oracle.

net.py itself does no arithmetic: weight leaves must bring the reference's WeightOps surface (`scale`, `add_assign`,
`clone`: gad/src/net.rs:100-104,161-180). On the device that is CuArray. For the oracle this module supplies the host
counterpart — a numpy array behind the same three methods — and wraps the oracle's Graph1 so that the gradients it
hands back are of that type. All host arithmetic of the CPU net tests lives here, in tests/.
"""
import numpy as np

import oracle_py


class HostWeight:
    """A host array as a WeightOps value; column-major Dim4 like the oracle's HostArray."""

    def __init__(self, data):
        self.a = np.array(data, copy=True)

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    @property
    def shape(self):
        return self.a.shape

    def dims(self):
        s = tuple(int(x) for x in self.a.shape)
        return s + (1,) * (4 - len(s))

    def numpy(self):
        return self.a.reshape(self.dims(), order="F") if self.a.ndim < 4 else self.a

    def clone(self):
        return HostWeight(self.a)

    def scale(self, lam):
        return HostWeight(self.a * np.asarray(lam, dtype=self.a.dtype))

    def add_assign(self, other):
        o = np.asarray(other.a if isinstance(other, HostWeight) else other)
        self.a = self.a + o.reshape(self.a.shape)  # rebinding: earlier clones keep their values (copy-on-write)


def _host(data):
    return data if isinstance(data, oracle_py.OValue) else np.asarray(data)


class _Store:
    def __init__(self, store):
        self.store = store

    def get(self, gid):
        v = self.store.get(gid)
        return None if v is None else HostWeight(v.data())


class OGraph1(oracle_py.OGraph1):
    """The oracle's Graph1 taking HostWeight data and returning HostWeight gradients."""

    def variable(self, data, dtype=None):
        return super().variable(_host(data), dtype)

    def constant(self, data, dtype=None):
        return super().constant(_host(data), dtype)

    def evaluate_gradients_once(self, gid, gradient, dtype=None):
        return _Store(super().evaluate_gradients_once(gid, gradient, dtype))

    def evaluate_gradients(self, gid, gradient, dtype=None):
        return _Store(super().evaluate_gradients(gid, gradient, dtype))


class OEval(oracle_py.OEval):
    def variable(self, data, dtype=None):
        return super().variable(_host(data), dtype)

    def constant(self, data, dtype=None):
        return super().constant(_host(data), dtype)
