"""metrics/metrics.py of the reference: accuracy (argmax match count, bit-exact)."""
from .. import _lib
from ..device import DeviceArray, as_device, call, stream_ptr


class Metric:

    def __call__(self, outputs, targets):
        raise NotImplementedError

    def __str__(self):
        return self.__class__.__name__.lower().replace("_", "")


class _Accuracy(Metric):

    def value_into(self, outputs, targets, result):
        o, t = as_device(outputs), as_device(targets)
        call(_lib.lib.bf_accuracy, o.ptr, t.ptr, o.shape[0], o.shape[1], result.ptr, stream_ptr())

    def __call__(self, outputs, targets):
        res = DeviceArray.empty(1)
        self.value_into(outputs, targets, res)
        return int(round(float(res.numpy()[0])))


accuracy = _Accuracy()
acc = accuracy

_metrics = {"accuracy": accuracy, "acc": accuracy}


def get(metric):
    if isinstance(metric, Metric):
        return metric
    m = _metrics.get(metric)
    if m is None:
        raise ValueError("No such metric: {}".format(metric))
    return m
