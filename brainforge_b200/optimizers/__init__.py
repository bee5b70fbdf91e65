"""optimizers/ of the reference: flat-vector update rules, one fused in-place kernel per step.

`optimize(W, gW, m)` keeps the reference's contract (optimizers/abstract_optimizer.py:19-30): with host
arrays it is a pure function returning the new flat vector (what reinforcement/policygradient.py:56 uses);
with `DeviceArray`s it updates W in place in HBM and optionally clears gW in the same pass.
Hyper-parameters and defaults are the reference's, including its non-standard Adam (adaptive_gd.py:94-99:
raw sum gradient in both moments, no bias correction, epsilon inside the square root).
"""
import ctypes

import numpy as np

from .. import _lib
from ..device import DeviceArray, call, stream_ptr, note_mutation

_NULL = ctypes.c_void_p(0)


class Optimizer:
    pass


class GradientDescent(Optimizer):
    kind = "sgd"
    _n_state = (False, False)   # uses velocity (s1), memory (s2)

    def __init__(self, eta=0.01):
        self.eta = eta
        self.nparams = None
        self.velocity = None
        self.memory = None

    def initialize(self, nparams=None, velocity=None, memory=None, **kw):
        self.nparams = nparams
        self.velocity = DeviceArray.from_host(velocity) if velocity is not None else None
        self.memory = DeviceArray.from_host(memory) if memory is not None else None

    def _hp(self):
        return 0.0, 0.0, 0.0

    def _state(self, n):
        use1, use2 = self._n_state
        if use1 and (self.velocity is None or self.velocity.size != n):
            self.velocity = DeviceArray.zeros(n)
        if use2 and (self.memory is None or self.memory.size != n):
            self.memory = DeviceArray.zeros(n)
        return (self.velocity.ptr if use1 else _NULL), (self.memory.ptr if use2 else _NULL)

    def optimize(self, W, gW, m, zero_grad=False):
        host = not isinstance(W, DeviceArray)
        Wd = DeviceArray.from_host(np.asarray(W).ravel()) if host else W
        gd = DeviceArray.from_host(np.asarray(gW).ravel()) if not isinstance(gW, DeviceArray) else gW
        s1, s2 = self._state(Wd.size)
        hp0, hp1, hp2 = self._hp()
        note_mutation()
        call(_lib.lib.bf_opt_step, _lib.OPT[self.kind], Wd.ptr, gd.ptr, s1, s2, Wd.size, float(self.eta), float(m),
             float(hp0), float(hp1), float(hp2), 1 if zero_grad else 0, stream_ptr())
        return Wd.numpy().reshape(np.shape(W)) if host else Wd

    def optimize_allreduce(self, W, G, n_params, m, symm, m_local):
        """Data-parallel step: G (this rank's gradient vector, step scalars at its tail) is pushed into every rank's
        symmetric receive buffer, the world copies are summed and the update applied in one kernel; G's parameter part
        is cleared, its tail holds the summed scalars (cost sum, hits, sample count)."""
        s1, s2 = self._state(n_params)
        hp0, hp1, hp2 = self._hp()
        note_mutation()
        call(_lib.lib.bf_opt_step_allreduce, _lib.OPT[self.kind], W.ptr, G.ptr, s1, s2, n_params, G.size, float(self.eta),
             float(m), float(hp0), float(hp1), float(hp2), 1, symm.peers_dev, symm.pads_dev, symm.epoch.data_ptr(),
             symm.timeout.data_ptr(), float(m_local), symm.rank, symm.world, symm.blocks, stream_ptr())
        return W

    def __str__(self):
        return self.__class__.__name__


class SGD(GradientDescent):
    """optimizers/gradient_descent.py:6-12."""
    kind = "sgd"


class Momentum(SGD):
    """optimizers/gradient_descent.py:15-31."""
    kind = "momentum"
    _n_state = (True, False)

    def __init__(self, eta=0.1, mu=0.9):
        super().__init__(eta)
        self.mu = mu

    def _hp(self):
        return self.mu, 0.0, 0.0


class Nesterov(Momentum):
    """optimizers/gradient_descent.py:34-53 (literal, including its quirks)."""
    kind = "nesterov"
    _n_state = (True, True)


class Adagrad(SGD):
    """optimizers/adaptive_gd.py:5-23."""
    kind = "adagrad"
    _n_state = (False, True)

    def __init__(self, eta=0.01, epsilon=1e-8):
        super().__init__(eta)
        self.epsilon = epsilon

    def _hp(self):
        return self.epsilon, 0.0, 0.0


class RMSprop(Adagrad):
    """optimizers/adaptive_gd.py:65-79."""
    kind = "rmsprop"

    def __init__(self, eta=0.1, decay=0.9, epsilon=1e-8):
        super().__init__(eta, epsilon)
        self.decay = decay

    def _hp(self):
        return self.decay, self.epsilon, 0.0


class Adam(SGD):
    """optimizers/adaptive_gd.py:81-102."""
    kind = "adam"
    _n_state = (True, True)

    def __init__(self, eta=0.001, decay_memory=0.999, decay_velocity=0.9, epsilon=1e-5):
        super().__init__(eta)
        self.decay_velocity = decay_velocity
        self.decay_memory = decay_memory
        self.epsilon = epsilon

    def _hp(self):
        return self.decay_velocity, self.decay_memory, self.epsilon


optimizers = {"sgd": SGD, "momentum": Momentum, "nesterov": Nesterov,
              "adagrad": Adagrad, "rmsprop": RMSprop, "adam": Adam}
