"""layers/recurrent.py of the reference: RLayer, LSTM, GRU and ClockworkLayer (Reservoir is out of scope)."""
import ctypes

import numpy as np

from .. import atomic, _lib
from ..device import DeviceArray, call, stream_ptr
from ..util import white
from .abstract_layer import FFBase


def _transpose01(x):
    d0, d1, d2 = x.shape
    out = DeviceArray.empty(d1, d0, d2)
    call(_lib.lib.bf_transpose01, x.ptr, out.ptr, d0, d1, d2, stream_ptr())
    return out


class RLayer(FFBase):
    """layers/recurrent.py:12-47,50-77.  Input (B,T,D); weights white(D+H, H), zero biases; gradients ASSIGNED (:74)."""

    def __init__(self, neurons, activation, return_seq=False, **kw):
        super().__init__(neurons, activation, **kw)
        self.Z = 0
        self.time = 0
        self.return_seq = return_seq
        self.op = atomic.RecurrentOp(activation if isinstance(activation, str) else str(activation))

    def connect(self, brain):
        self.Z = brain.outshape[-1] + self.neurons
        self._init_params(white(self.Z, self.neurons), np.zeros(self.neurons))
        super().connect(brain)

    def _fwd(self, x):
        self.inputs = _transpose01(x)
        self.time = self.inputs.shape[0]
        self.output, self.Z = self.op.forward(self.inputs, self.weights, self.biases)
        return _transpose01(self.output) if self.return_seq else self.output[-1]

    def _bwd(self, delta):
        if self.return_seq:
            E = _transpose01(delta)
        else:
            E = DeviceArray.zeros(self.time, len(delta), self.neurons)
            last = E[-1]
            call(_lib.lib.bf_act_forward, _lib.ACT["linear"], delta.ptr, last.ptr, last.size, stream_ptr())
        dX, _, _ = self.op.backward(Z=self.Z, O=self.output, E=E, W=self.weights, nablaW=self.nabla_w,
                                    nablab=self.nabla_b)
        return _transpose01(dX)

    @property
    def outshape(self):
        return self.neurons,

    def __str__(self):
        return self.__class__.__name__ + "-{}-{}".format(self.neurons, str(self.activation)[:4])


class GRU(FFBase):
    """layers/recurrent.py:12-47,116-190.  Input (B,T,D); weights white(D+H, 3H) with columns [U|R|O], zero biases."""

    def __init__(self, neurons, activation, return_seq=False, **kw):
        super().__init__(neurons, activation, **kw)
        self.time = 0
        self.return_seq = return_seq
        self.Zs = self.gates = None
        self.op = atomic.GRUOp(activation if isinstance(activation, str) else str(activation))

    def connect(self, brain):
        self._init_params(white(brain.outshape[-1] + self.neurons, self.neurons * 3), np.zeros(self.neurons * 3))
        super().connect(brain)

    def _fwd(self, x):
        self.inputs = _transpose01(x)
        self.time = self.inputs.shape[0]
        hs, Z, Zk, self.gates = self.op.forward(self.inputs, self.weights, self.biases)
        self.Zs = (Z, Zk)
        self.output = _transpose01(hs) if self.return_seq else hs[-1]      # recurrent.py:148-151
        return self.output

    def _bwd(self, delta):
        if self.return_seq:
            E = _transpose01(delta)
        else:
            E = DeviceArray.zeros(self.time, len(delta), self.neurons)
            last = E[-1]
            call(_lib.lib.bf_act_forward, _lib.ACT["linear"], delta.ptr, last.ptr, last.size, stream_ptr())
        dX, _, _ = self.op.backward(Z=self.Zs[0], Zk=self.Zs[1], gates=self.gates, E=E, W=self.weights,
                                    nablaW=self.nabla_w, nablab=self.nabla_b)
        return _transpose01(dX)

    @property
    def outshape(self):
        return self.neurons,

    def __str__(self):
        return self.__class__.__name__ + "-{}-{}".format(self.neurons, str(self.activation)[:4])


def clockwork_init(indim, neurons, blocksizes, ticks):
    """layers/recurrent.py:224-237, with its RNG consumption (U first, then one draw per block) and its block offsets
    `start = i * bls` kept literally.  Returns (weights (H+D, H), tick_array (H))."""
    W = np.zeros((neurons, neurons))
    U = white(indim, neurons)
    tick_array = np.zeros(neurons)
    for i, bls in enumerate(blocksizes):
        start = i * bls
        end = start + bls
        W[start:end, start:] = white(*W[start:end, start:].shape)
        tick_array[start:end] = ticks[i]
    return np.concatenate((W, U), axis=0), tick_array


class ClockworkLayer(FFBase):
    """layers/recurrent.py:193-283.  Input (B,T,D); the recurrent block W is block upper-triangular, block i updates
    every ticktimes[i] steps."""

    def __init__(self, neurons, activation, blocksizes=None, ticktimes=None, return_seq=False, **kw):
        super().__init__(neurons, activation, **kw)
        self.time = 0
        self.return_seq = return_seq
        if blocksizes is None:                                  # recurrent.py:198-201
            block = neurons // 5
            blocksizes = [block] * 5
            blocksizes[0] += (neurons % block)
        elif sum(blocksizes) != self.neurons:
            raise RuntimeError("Please specify blocksizes so that they sum up to the number of neurons specified for "
                               "this layer! ({})".format(neurons))
        self.blocksizes = blocksizes
        if ticktimes is None:
            ticktimes = [2 ** i for i in range(len(self.blocksizes))]
        elif min(ticktimes) < 0 or len(ticktimes) != len(self.blocksizes):
            raise RuntimeError("Please specify the <ticks> parameter so, that its length is equal to the number of "
                               "blocks specified (defaults to 5); timesteps < 0 are invalid!")
        self.ticks = np.array(ticktimes)
        self.tick_array = None
        self.Z = 0
        self.op = atomic.ClockworkOp(activation if isinstance(activation, str) else str(activation))

    def connect(self, brain):
        self.Z = brain.outshape[-1] + self.neurons
        weights, ticks = clockwork_init(brain.outshape[-1], self.neurons, self.blocksizes, self.ticks)
        self._init_params(weights, np.zeros(self.neurons))
        self.tick_array = DeviceArray.from_host(ticks)
        super().connect(brain)

    def _fwd(self, x):
        self.inputs = _transpose01(x)
        self.time = self.inputs.shape[0]
        self.cache, self.Zs = self.op.forward(self.inputs, self.weights, self.biases, self.tick_array)
        self.output = _transpose01(self.cache) if self.return_seq else self.cache[-1]      # recurrent.py:252-255
        return self.output

    def _bwd(self, delta):
        if self.return_seq:
            E = _transpose01(delta)
        else:
            E = DeviceArray.zeros(self.time, len(delta), self.neurons)
            last = E[-1]
            call(_lib.lib.bf_act_forward, _lib.ACT["linear"], delta.ptr, last.ptr, last.size, stream_ptr())
        dX, _, _ = self.op.backward(Z=self.Zs, O=self.cache, E=E, W=self.weights, tick=self.tick_array,
                                    nablaW=self.nabla_w, nablab=self.nabla_b)
        return _transpose01(dX)

    @property
    def outshape(self):
        return self.neurons,

    def __str__(self):
        return self.__class__.__name__ + "-{}-{}".format(self.neurons, str(self.activation)[:4])


class LSTM(FFBase):
    """layers/recurrent.py:12-47,80-113.  Input (B,T,D) batch-major like the reference; weights
    white(D+H, 4H), biases +7.0 (bias_init_factor, :82,:98); gradients are ASSIGNED (:110)."""

    def __init__(self, neurons, activation, bias_init_factor=7., return_seq=False, **kw):
        super().__init__(neurons, activation, **kw)
        self.G = neurons * 3
        self.Z = 0
        self.cache = None
        self.time = 0
        self.return_seq = return_seq
        self.bias_init_factor = bias_init_factor
        self.op = atomic.LSTMOp(activation if isinstance(activation, str) else str(activation))

    def connect(self, brain):
        self.Z = brain.outshape[-1] + self.neurons
        self._init_params(white(self.Z, self.neurons * 4), np.zeros(self.neurons * 4) + self.bias_init_factor)
        super().connect(brain)

    _can_skip_dx = True

    @property
    def inputs(self):
        """(T,B,D) like the reference's `X.transpose(1, 0, 2)` (recurrent.py:26); materialised on demand -- the kernel
        reads the batch-major array directly."""
        if self._inputs_tm is None and self._inputs_bm is not None:
            self._inputs_tm = _transpose01(self._inputs_bm)
        return self._inputs_tm

    @inputs.setter
    def inputs(self, value):
        self._inputs_tm, self._inputs_bm = value, None

    def _fwd(self, x):
        self._inputs_tm, self._inputs_bm = None, x
        self.time = x.shape[1]
        self.output, self.Z, self.cache = self.op.forward(x, self.weights, self.biases, batch_major=True)
        return _transpose01(self.output) if self.return_seq else self.output[-1]

    def _bwd(self, delta, need_dX=True):
        if self.return_seq:
            E = _transpose01(delta)                         # recurrent.py:35-36
        else:
            E = DeviceArray.zeros(self.time, len(delta), self.neurons)   # recurrent.py:38-40
            last = E[-1]
            call(_lib.lib.bf_act_forward, _lib.ACT["linear"], delta.ptr, last.ptr, last.size, stream_ptr())
        dX, _, _ = self.op.backward(Z=self.Z, O=self.output, E=E, W=self.weights, cache=self.cache,
                                    nablaW=self.nabla_w, nablab=self.nabla_b, need_dX=need_dX)
        return _transpose01(dX) if dX is not None else None

    @property
    def outshape(self):
        return self.neurons,

    def __str__(self):
        return self.__class__.__name__ + "-{}-{}".format(self.neurons, str(self.activation)[:4])
