"""layers/core.py of the reference: Dense, Activation, InputLayer, Reshape, Flatten."""
import numpy as np

from .. import atomic, _lib
from ..util import white
from ..device import DeviceArray
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

    # ---- fused output head (learn_batch only): forward + cost + accuracy + cost' + backward of the LAST layer in one kernel
    def head_supported(self, cost):
        k = self.weights.shape[0]
        return (type(self) is Dense and self.neurons <= 32 and cost.kind in ("mse", "cxent", "bxent") and
                bool(_lib.lib.bf_dense_head_supported(k, self.neurons, _lib.ACT[str(self.activation)], _lib.COST[cost.kind])))

    def learn_head(self, x, y, w, cost, cost_out, hits_out, need_dX=True):
        """What `_fwd` -> cost.derivative_and_value -> [accuracy] -> `_bwd` compute for this layer, in one pass over x
        (bf_dense_head): sets inputs / output, accumulates nabla_w / nabla_b, fills the step scalars, returns dX."""
        from ..device import call, stream_ptr
        m, n = x.shape[0], self.neurons
        cl = x.cl_flat is not None
        if cl:
            c, yy, xx = x.cl_flat
            chan, hw = c, yy * xx
        else:
            chan, hw = x.shape[1], 1
        self.inputs = x
        self.output = DeviceArray.empty(m, n)
        dX = None
        if need_dX:
            dX = DeviceArray(DeviceArray.empty(m, chan * hw).t, cl_flat=x.cl_flat) if cl else DeviceArray.empty(m, chan * hw)
        call(_lib.lib.bf_dense_head, x.pptr, self.weights.ptr, self.biases.ptr, y.ptr, w.ptr if w is not None else None,
             self.output.ptr, self.nabla_w.ptr, self.nabla_b.ptr, dX.pptr if dX is not None else None, cost_out.ptr,
             hits_out.ptr if hits_out is not None else None, m, chan, hw, n, _lib.ACT[str(self.activation)],
             _lib.COST[cost.kind], 1 if cl else 0, 1, stream_ptr())
        return dX

    def __str__(self):
        return "Dense-{}-{}".format(self.neurons, str(self.activation)[:5])


class Activation(NoParamMixin, LayerBase):
    """layers/core.py:45-59."""

    _next_layer = None     # set by LayerStack.add: lets relu hand its pass to a following PoolLayer
    _lazy_src = None       # callable -> the pre-activation array, when this layer moved no bytes

    def _fwd(self, x):
        self._lazy_src = None
        if type(x).__name__ == "FusedConvPool":
            # the convolution kernel upstream already applied max(0,.) and pooled (layers/tensor.py ConvLayer._fwd)
            conv = x.conv
            self._lazy_src, self._output = (lambda: conv.output), None
            return x
        if str(self.activation) == "linear":
            self._output = x
        elif (str(self.activation) == "relu" and x.nhwc and not self._standalone
              and type(self._next_layer).__name__ == "PoolLayer"):
            # channels-last conv block: the pooling kernel applies max(0,.) while it reads (and relu' while it
            # scatters), so this layer moves no bytes.  `self.output` is materialised only if somebody reads it.
            self._lazy_src, self._output = (lambda: x), None
            y = DeviceArray(x.t, nhwc=True)
            y.pending_relu = True
            return y
        else:
            self._output = self.activation.forward(x)
        return self._output

    @property
    def output(self):
        if self._output is None and self._lazy_src is not None:
            self._output = self.activation.forward(self._lazy_src())
        return self._output

    @output.setter
    def output(self, value):
        self._output = value

    def _bwd(self, delta):
        if self._lazy_src is not None and getattr(delta, "relu_applied", False):
            return delta                 # relu' was applied by the pooling scatter (or will be, by the fused un-pool)
        if type(delta).__name__ == "FusedUnpool":
            delta = delta.materialize()
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

    _next_layer = None

    def _fwd(self, x):
        nxt = self._next_layer
        if (x.nhwc and not self._standalone and len(self.shape) == 1 and type(nxt).__name__ == "Dense"
                and nxt.neurons <= 32 and _lib.lib.bf_dense_cl_supported(int(np.prod(x.shape[1:])), nxt.neurons)):
            # channels-last block -> skinny Dense: no data moves; the Dense kernels index W through the (c,y,x) <-> (y,x,c)
            # permutation (bf_dense_forward_cl), so the layout-change pass disappears
            im, c, y, xx = x.shape
            return DeviceArray(x.t.view(im, c * y * xx), cl_flat=(c, y, xx))
        return atomic.ReshapeOp.forward(x, self.shape)

    def _bwd(self, delta):
        if getattr(delta, "cl_flat", None) is not None:
            c, y, xx = delta.cl_flat
            return DeviceArray(delta.t.view(delta.t.shape[0], y, xx, c), nhwc=True)
        return atomic.ReshapeOp.forward(delta, self.inshape)

    @property
    def outshape(self):
        return self.shape

    def __str__(self):
        return self.__class__.__name__


class Flatten(Reshape):
    pass
