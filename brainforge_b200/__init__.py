"""brainforge_b200: a B200-native (sm_100a) back-end for the brainforge layer forward/backward +
optimizer hot path, behind the reference's own Layer / LayerStack / Backpropagation API.

    from brainforge_b200 import Backpropagation, LayerStack, NeuroEvolution
    from brainforge_b200.layers import Dense, ConvLayer, PoolLayer, Flatten, Activation, LSTM

Importing this package loads libbf_b200.so (hand-written CUDA behind a C ABI, include/bf_b200.h) and fails
loudly if it is missing; there is no CPU or PyTorch fallback for any operator.
"""
from . import _lib                      # noqa: F401  (raises ImportError when the library is not built)
from .device import DeviceArray, set_precision, precision_name, pinned_empty, synchronize
from . import atomic
from .model import LayerStack
from .learner import Backpropagation, NeuroEvolution, Learner
from . import layers, optimizers, metrics, evolution, dist, gradientcheck

BackpropNetwork = Backpropagation       # the stale name the README / BASELINE.json still use
