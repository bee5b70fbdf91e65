"""layers/abstract_layer.py of the reference, with device-resident state underneath.

Public surface kept verbatim: attributes weights, biases, nabla_w, nabla_b, inputs, output,
trainable, brain, inshape, outshape, num_params; methods connect, feedforward, backpropagate,
get_weights, set_weights, get_gradients, set_gradients, shuffle.  The array attributes are
`DeviceArray`s (FP32 in HBM); `get_*` return host float64 NumPy arrays like the reference.
"""
import abc

import numpy as np

from .. import atomic
from ..device import DeviceArray, as_device
from ..util import white


class LayerBase(abc.ABC):

    trainable = True

    def __init__(self, activation="linear", **kw):
        self.brain = None
        self.inputs = None
        self.output = None
        self.inshape = None
        self.op = None
        self.opb = None
        self.weights = None
        self.biases = None
        self.nabla_w = None
        self.nabla_b = None
        self.compiled = kw.get("compiled", True)   # kept for signature compatibility; there is one back-end
        self.trainable = kw.get("trainable", self.trainable)
        self.activation = atomic.activations[activation]() if isinstance(activation, str) else activation

    def connect(self, brain):
        self.brain = brain
        self.inshape = brain.outshape

    # ---- parameters (host float64 in / out; layers/abstract_layer.py:36-55)
    def _init_params(self, w_host, b_host):
        self.weights = DeviceArray.from_host(w_host)
        self.biases = DeviceArray.from_host(b_host)
        self.nabla_w = DeviceArray.zeros(*self.weights.shape)
        self.nabla_b = DeviceArray.zeros(*self.biases.shape)

    def shuffle(self):
        self.weights.copy_from_host(white(*self.weights.shape))
        self.biases.fill_zero()

    def get_weights(self, unfold=True):
        if unfold:
            return np.concatenate((self.weights.numpy().ravel(), self.biases.numpy().ravel()))
        return [self.weights.numpy(), self.biases.numpy()]

    def set_weights(self, w, fold=True):
        if fold:
            w = np.asarray(w)
            nW = self.weights.size
            self.weights.copy_from_host(w[:nW].reshape(self.weights.shape))
            self.biases.copy_from_host(w[nW:nW + self.biases.size].reshape(self.biases.shape))
        else:
            self.weights.copy_from_host(np.asarray(w[0]))
            self.biases.copy_from_host(np.asarray(w[1]))

    def get_gradients(self, unfold=True):
        nabla = [self.nabla_w.numpy(), self.nabla_b.numpy()]
        return nabla if not unfold else np.concatenate([grad.ravel() for grad in nabla])

    def set_gradients(self, nabla, fold=True):
        if fold:
            nabla = np.asarray(nabla)
            nW = self.nabla_w.size
            self.nabla_w.copy_from_host(nabla[:nW].reshape(self.nabla_w.shape))
            self.nabla_b.copy_from_host(nabla[nW:nW + self.nabla_b.size].reshape(self.nabla_b.shape))
        else:
            self.nabla_w.copy_from_host(np.asarray(nabla[0]))
            self.nabla_b.copy_from_host(np.asarray(nabla[1]))

    @property
    def gradients(self):
        return self.get_gradients(unfold=True)

    @property
    def num_params(self):
        return self.weights.size + self.biases.size

    # ---- compute: host arrays in -> host float64 out; DeviceArray in -> DeviceArray out
    _standalone = False    # True while a layer is driven on its own through feedforward(): no cross-layer fusion

    def feedforward(self, X):
        self._standalone = True
        try:
            out = self._fwd(as_device(X))
        finally:
            self._standalone = False
        return out if isinstance(X, DeviceArray) else out.numpy()

    def backpropagate(self, delta):
        d = as_device(delta) if isinstance(delta, DeviceArray) else DeviceArray.from_host(delta)
        self._standalone = True
        try:
            out = self._bwd(d)
        finally:
            self._standalone = False
        return out if isinstance(delta, DeviceArray) else out.numpy()

    @abc.abstractmethod
    def _fwd(self, x): raise NotImplementedError

    @abc.abstractmethod
    def _bwd(self, delta): raise NotImplementedError

    @property
    @abc.abstractmethod
    def outshape(self): raise NotImplementedError

    def __str__(self):
        return self.__class__.__name__


class NoParamMixin(abc.ABC):

    trainable = False

    def shuffle(self): pass

    def get_weights(self, unfold=True): pass

    def set_weights(self, w, fold=True): pass

    def get_gradients(self): pass

    def set_gradients(self): pass

    @property
    def nparams(self):
        return None


class FFBase(LayerBase):
    """Base class for the fully connected layer types (layers/abstract_layer.py:106-123)."""

    def __init__(self, neurons, activation="linear", **kw):
        super().__init__(activation, **kw)
        if not isinstance(neurons, int):
            neurons = np.prod(neurons)
        self.neurons = int(neurons)

    @property
    def outshape(self):
        return self.neurons,
