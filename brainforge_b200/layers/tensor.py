"""layers/tensor.py of the reference: PoolLayer, ConvLayer."""
import numpy as np

from .. import atomic
from ..util import white
from .abstract_layer import LayerBase, NoParamMixin


class FusedConvPool:
    """What a ConvLayer hands downstream when its kernel has already done the [ReLU ->] 2x2 pooling that follows it:
    the pooled tensor and tie mask, plus the way back to the (never materialised) convolution output."""

    def __init__(self, conv, pooled, mask, relu_in):
        self.conv, self.pooled, self.mask, self.relu_in = conv, pooled, mask, relu_in


class FusedUnpool:
    """The delta a PoolLayer sends upstream when the convolution right below it can un-pool inside its own
    filter-gradient kernel: the pooled-resolution delta plus the pooling state, materialised only if somebody
    needs the full-resolution array."""

    def __init__(self, pool, delta):
        self.pool, self.delta = pool, delta
        self.gate = pool.output if pool._fused_relu else None
        self.relu_applied = pool._fused_relu

    def materialize(self):
        dX = self.pool.op.backward(self.delta, self.pool.filter, gate=self.gate, want_db=True)
        if self.relu_applied:
            dX.relu_applied = True
        return dX


class PoolLayer(NoParamMixin, LayerBase):
    """layers/tensor.py:7-40."""

    def __init__(self, filter_size, compiled=True):
        LayerBase.__init__(self, activation="linear", trainable=False)
        self.fdim = filter_size
        self.filter = None
        self.op = atomic.MaxPoolOp()
        self._outshape = None

    def connect(self, brain):
        ic, iy, ix = brain.outshape[-3:]
        if any((iy % self.fdim, ix % self.fdim)):
            raise RuntimeError("Incompatible shapes: {} % {}".format((ix, iy), self.fdim))
        LayerBase.connect(self, brain)
        self._outshape = (ic, iy // self.fdim, ix // self.fdim)

    def _fwd(self, x):
        self._conv_src = None
        if isinstance(x, FusedConvPool):      # the convolution kernel pooled its own tile (ConvLayer._fwd)
            self._fused_relu = x.relu_in
            self._conv_src = x.conv
            self.output, self.filter = x.pooled, x.mask
            return self.output
        # a channels-last input flagged `pending_relu` comes from an Activation("relu") layer that left its
        # max(0,.) to this layer's kernel (layers/core.py Activation._fwd): one pass instead of two
        self._fused_relu = bool(getattr(x, "pending_relu", False))
        self.output, self.filter = self.op.forward(x, self.fdim, relu_in=self._fused_relu)
        return self.output

    def _bwd(self, delta):
        if getattr(self, "_conv_src", None) is not None and not self._standalone:
            return FusedUnpool(self, delta)   # the convolution's filter-gradient kernel un-pools on the fly
        if getattr(self, "_fused_relu", False):
            dX = self.op.backward(delta, self.filter, gate=self.output, want_db=not self._standalone)   # pool' and relu' in one scatter
            dX.relu_applied = True
            return dX
        return self.op.backward(delta, self.filter, want_db=not self._standalone)

    @property
    def outshape(self):
        return self._outshape

    def __str__(self):
        return "Pool-{}x{}".format(self.fdim, self.fdim)


class ConvLayer(LayerBase):
    """layers/tensor.py:43-92.  Quirks kept on purpose: the bias is added AFTER the activation (:77-78),
    gradients are ASSIGNED not accumulated (:83), weights are white(nfilters, depth, fx, fy) (:70) and the
    op reads dims 2,3 of the filter as (rows, cols).  `outshape` reports the true output shape (the
    reference swaps the two spatial entries, :86-89; identical for square inputs and filters)."""

    def __init__(self, nfilters, filterx=3, filtery=3, compiled=True, **kw):
        super().__init__(compiled=compiled, **kw)
        self.nfilters = nfilters
        self.fx = filterx
        self.fy = filtery
        self.depth = 0
        self.stride = 1
        self.inshape = None
        self.op = None

    def connect(self, brain):
        depth, iy, ix = brain.outshape[-3:]
        if any((iy < self.fy, ix < self.fx)):
            raise RuntimeError("Incompatible shapes: iy ({}) < fy ({}) OR ix ({}) < fx ({})"
                               .format(iy, self.fy, ix, self.fx))
        super().connect(brain)
        self.op = atomic.ConvolutionOp()
        self.inshape = brain.outshape
        self.depth = depth
        self._init_params(white(self.nfilters, self.depth, self.fx, self.fy),
                          np.zeros((1, self.nfilters, 1, 1)))

    fuse_pool = True     # class-wide switch: let the first-layer kernel absorb a following [ReLU ->] PoolLayer(2)

    def _pool_after(self):
        """(PoolLayer, relu_in) when this layer's output is consumed only by [Activation("relu") ->] PoolLayer(2)."""
        nxt = getattr(self, "_next_layer", None)
        if isinstance(nxt, PoolLayer) and nxt.fdim == 2:
            return nxt, False
        if type(nxt).__name__ == "Activation" and str(nxt.activation) == "relu":
            nn = getattr(nxt, "_next_layer", None)
            if isinstance(nn, PoolLayer) and nn.fdim == 2:
                return nn, True
        return None, False

    def _fwd(self, x):
        im, ic, iy, ix = x.shape
        route = self.op._route(x.nhwc, ic, iy, ix, self.nfilters, self.fy, self.fx, "valid", str(self.activation))
        if not x.nhwc and route == "nhwc":
            x = x.to_nhwc()          # converted once here, reused by the filter gradient
        self.inputs = x
        self._output = None
        if getattr(self, "_pack", None) is None:
            self._pack = atomic.FilterPack()
        if self.fuse_pool and not self._standalone and route in ("c3", "nhwc") and str(self.activation) == "linear":
            pool, relu_in = self._pool_after()
            if pool is not None:
                fused = self.op.forward_pool(x, self.weights, bias=self.biases, relu_in=relu_in, pack=self._pack)
                if fused is not None:
                    # the convolution output is not materialised; `self.output` re-runs the plain kernel on demand
                    return FusedConvPool(self, fused[0], fused[1], relu_in)
        self._output = self.op.forward(x, self.weights, "valid", bias=self.biases, activation=str(self.activation),
                                       pack=self._pack if not self._standalone else None)
        return self._output

    @property
    def output(self):
        if self._output is None and self.inputs is not None:
            # fused path: recomputed from the cached inputs and the CURRENT weights (exact until the next update)
            self._output = self.op.forward(self.inputs, self.weights, "valid", bias=self.biases,
                                           activation=str(self.activation))
        return self._output

    @output.setter
    def output(self, value):
        self._output = value

    _can_skip_dx = True

    def _bwd(self, delta, need_dX=True):
        if isinstance(delta, FusedUnpool):
            if not need_dX and str(self.activation) == "linear" and self.op.backward_filter_unpool(
                    self.inputs, delta.delta, delta.pool.filter, delta.gate, self.weights.shape, self.nabla_w, self.nabla_b):
                return None
            delta = delta.materialize()
        if str(self.activation) != "linear":
            delta = self.activation.backward_mul(self.output, delta)
            delta.db_partial = None          # channel sums left by a pooling layer describe the delta BEFORE act'
        _, _, dX = self.op.backward(X=self.inputs, E=delta, F=self.weights, dF=self.nabla_w, db=self.nabla_b,
                                    need_dX=need_dX, pack=getattr(self, "_pack", None))
        return dX

    @property
    def outshape(self):
        return self.op.outshape(self.inshape, self.weights.shape, "valid")

    def __str__(self):
        return "Conv({}x{}x{})-{}".format(self.nfilters, self.fy, self.fx, str(self.activation)[:4])


class GlobalAveragePooling(NoParamMixin, LayerBase):
    """layers/tensor.py:95-114: (im, c, y, x) -> (im, c) mean over the plane; backward repeats delta / (y*x).
    Takes either storage layout of the incoming activation and answers in the same one."""

    def __init__(self):
        LayerBase.__init__(self, activation="linear", trainable=False)
        self.repeats = 0
        self._nhwc = False

    def _fwd(self, x):
        from ..device import DeviceArray, call, stream_ptr
        from .. import _lib
        im, c, y, xx = x.shape
        self.repeats = y * xx
        self._nhwc = x.nhwc
        out = DeviceArray.empty(im, c)
        call(_lib.lib.bf_gap_forward, x.pptr, out.ptr, im, c, y * xx, 1 if x.nhwc else 0, stream_ptr())
        self.output = out
        return out

    def _bwd(self, delta):
        from ..device import DeviceArray, call, stream_ptr
        from .. import _lib
        c, y, xx = self.inshape[-3:]
        im = delta.shape[0]
        dX = DeviceArray.empty_nhwc(im, c, y, xx) if self._nhwc else DeviceArray.empty(im, c, y, xx)
        call(_lib.lib.bf_gap_backward, delta.to_nchw().ptr, dX.pptr, im, c, y * xx, 1 if self._nhwc else 0, stream_ptr())
        return dX

    @property
    def outshape(self):
        return self.inshape[0],

    def __str__(self):
        return "GlobalAveragePooling"
