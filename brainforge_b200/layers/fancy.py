"""layers/fancy.py of the reference: DropOut (the other fancy layers are out of scope)."""
import numpy as np

from .. import _lib
from ..device import DeviceArray, call, stream_ptr
from .abstract_layer import LayerBase, NoParamMixin

_NULL = __import__("ctypes").c_void_p(0)


class DropOut(NoParamMixin, LayerBase):
    """layers/fancy.py:60-90, quirks kept: the mask has the shape of ONE sample (drawn from the global NumPy RNG on the
    host exactly like the reference, so the same seed gives the same mask) and is shared by the whole batch; outside
    `brain.learning` the input is scaled by the keep probability; backpropagate multiplies by the mask and then RESETS
    the mask to the keep probability (fancy.py:80-82)."""

    def __init__(self, dropchance):
        super().__init__(activation="linear", trainable=False)
        self.dropchance = 1. - dropchance
        self.mask = None
        self.inshape = None
        self.training = True

    def connect(self, brain):
        self.inshape = brain.outshape
        super().connect(brain)

    def _fwd(self, x):
        x = x.to_nchw()
        self.inputs = x
        self.mask = (np.random.uniform(0, 1, self.inshape) < self.dropchance)
        cols = int(np.prod(self.inshape))
        out = DeviceArray.empty(*x.shape)
        if self.brain.learning:
            md = DeviceArray.from_host(self.mask.astype(np.float32))
            call(_lib.lib.bf_mul_rowmask, x.ptr, md.ptr, out.ptr, x.size // cols, cols, 1.0, stream_ptr())
        else:
            call(_lib.lib.bf_mul_rowmask, x.ptr, _NULL, out.ptr, x.size // cols, cols, float(self.dropchance), stream_ptr())
        self.output = out
        return out

    def _bwd(self, delta):
        delta = delta.to_nchw()
        cols = int(np.prod(self.inshape))
        md = DeviceArray.from_host(np.asarray(self.mask, dtype=np.float32))
        out = DeviceArray.empty(*delta.shape)
        call(_lib.lib.bf_mul_rowmask, delta.ptr, md.ptr, out.ptr, delta.size // cols, cols, 1.0, stream_ptr())
        self.mask = np.ones_like(self.mask, dtype=np.float64) * self.dropchance
        return out

    @property
    def outshape(self):
        return self.inshape

    def __str__(self):
        return "DropOut({})".format(self.dropchance)
