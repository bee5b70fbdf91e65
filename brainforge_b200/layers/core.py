"""layers/core.py of the reference: Dense, Activation, InputLayer, Reshape, Flatten."""
import numpy as np

from .. import atomic
from ..util import white
from .abstract_layer import LayerBase, NoParamMixin, FFBase


class Dense(FFBase):
    """layers/core.py:9-42."""

    def connect(self, brain):
        inshape = brain.outshape
        if len(inshape) != 1:
            err = "Dense only accepts input shapes with 1 dimension!\n"
            err += "Maybe you should consider placing <Flatten> before <Dense>?"
            raise RuntimeError(err)
        self._init_params(white(inshape[0], self.neurons), np.zeros(self.neurons))
        super().connect(brain)
        self.op = atomic.DenseOp()

    def _fwd(self, x):
        self.inputs = x
        # op + activation fused in one launch (softmax rows finished by a second in-place kernel)
        self.output = self.op.forward(x, self.weights, self.biases, activation=str(self.activation))
        return self.output

    _can_skip_dx = True

    def _bwd(self, delta, need_dX=True):
        delta = self.activation.backward_mul(self.output, delta)             # core.py:35, in place
        _, _, dX = self.op.backward(self.inputs, delta, self.weights, dW=self.nabla_w, db=self.nabla_b,
                                    accumulate=True, need_dX=need_dX)        # core.py:37-38 `+=`
        return dX

    def __str__(self):
        return "Dense-{}-{}".format(self.neurons, str(self.activation)[:5])


class Activation(NoParamMixin, LayerBase):
    """layers/core.py:45-59."""

    def _fwd(self, x):
        if str(self.activation) == "linear":
            self.output = x
        else:
            self.output = self.activation.forward(x)
        return self.output

    def _bwd(self, delta):
        return self.activation.backward_mul(self.output, delta)

    @property
    def outshape(self):
        return self.inshape

    def __str__(self):
        return "Activation-{}".format(str(self.activation))


class InputLayer(Activation):
    """layers/core.py:62-68."""

    def __init__(self, inshape):
        super().__init__(activation="linear")
        if isinstance(inshape, (int, np.integer)):
            inshape = int(inshape),
        self.inshape = tuple(inshape)


class Reshape(NoParamMixin, LayerBase):
    """layers/core.py:70-96.  A view: no data moves."""

    def __init__(self, shape=None):
        super().__init__(activation="linear")
        self.shape = shape

    def connect(self, brain):
        if self.shape is None:
            self.shape = int(np.prod(brain.outshape)),
        super().connect(brain)

    def _fwd(self, x):
        return atomic.ReshapeOp.forward(x, self.shape)

    def _bwd(self, delta):
        return atomic.ReshapeOp.forward(delta, self.inshape)

    @property
    def outshape(self):
        return self.shape

    def __str__(self):
        return self.__class__.__name__


class Flatten(Reshape):
    pass
