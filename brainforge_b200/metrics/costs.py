"""metrics/costs.py of the reference: cost value + the `outputs - targets` shortcut derivative.

Values are reduced on the device (bf_cost_value) and come back as one float; derivative stays in HBM.
"""
import ctypes

import numpy as np

from .. import _lib
from ..device import DeviceArray, as_device, call, stream_ptr

L = _lib.lib
_NULL = ctypes.c_void_p(0)


class CostFunction:
    kind = None

    def value_into(self, outputs, targets, result):
        """result[0] = cost, no host sync (used by the learner's fused step)."""
        o, t = as_device(outputs), as_device(targets)
        m = o.shape[0]
        n = o.size // max(m, 1)
        call(L.bf_cost_value, _lib.COST[self.kind], o.ptr, t.ptr, m, n, result.ptr, stream_ptr())

    def derivative_and_value(self, outputs, targets, w, result):
        """delta (device) and result[0] = cost in ONE kernel: what the training step needs of a forward pass."""
        o, t = as_device(outputs), as_device(targets)
        m = o.shape[0]
        n = o.size // max(m, 1)
        delta = DeviceArray.empty(*o.shape)
        wd = as_device(w) if w is not None else None
        call(L.bf_cost_delta_value, _lib.COST[self.kind], o.ptr, t.ptr, wd.ptr if wd is not None else _NULL, delta.ptr,
             result.ptr, m, n, stream_ptr())
        return delta

    def __call__(self, outputs, targets):
        res = DeviceArray.empty(1)
        self.value_into(outputs, targets, res)
        return float(res.numpy()[0])

    def __str__(self):
        return self.__class__.__name__

    @staticmethod
    def derivative(outputs, targets, w=None):
        """metrics/costs.py:19-21 (+ the per-sample weights of learner/backpropagation.py:20-21)."""
        o, t = as_device(outputs), as_device(targets)
        m = o.shape[0]
        n = o.size // max(m, 1)
        delta = DeviceArray.empty(*o.shape)
        wd = as_device(w) if w is not None else None
        call(L.bf_cost_delta, o.ptr, t.ptr, wd.ptr if wd is not None else _NULL, delta.ptr, m, n, stream_ptr())
        return delta if isinstance(outputs, DeviceArray) else delta.numpy()


class _MeanSquaredError(CostFunction):
    kind = "mse"           # metrics/costs.py:24-31


class _CategoricalCrossEntropy(CostFunction):
    kind = "cxent"         # metrics/costs.py:34-43


class _BinaryCrossEntropy(CostFunction):
    kind = "bxent"         # metrics/costs.py:46-53


class _Hinge(CostFunction):
    kind = "hinge"         # metrics/costs.py:55-66: value sum(max(0, 1 - t*o)); sub-derivative -t where o <= 1, else 0

    def derivative(self, outputs, targets, w=None):
        scratch = DeviceArray.empty(1)
        delta = self.derivative_and_value(outputs, targets, w, scratch)
        return delta if isinstance(outputs, DeviceArray) else delta.numpy()


mean_squared_error = _MeanSquaredError()
categorical_crossentropy = _CategoricalCrossEntropy()
binary_crossentropy = _BinaryCrossEntropy()
hinge = _Hinge()
mse = mean_squared_error
cxent = categorical_crossentropy
bxent = binary_crossentropy

_costs = {"mean_squared_error": mse, "categorical_crossentropy": cxent, "binary_crossentropy": bxent, "hinge": hinge,
          "mse": mse, "cxent": cxent, "bxent": bxent}


def get(cost_function):
    if isinstance(cost_function, CostFunction):
        return cost_function
    cost = _costs.get(cost_function)
    if cost is None:
        raise ValueError("No such cost function: {}".format(cost_function))
    return cost
