from .abstract_layer import LayerBase, NoParamMixin, FFBase
from .core import Dense, Activation, InputLayer, Reshape, Flatten
from .tensor import PoolLayer, ConvLayer, GlobalAveragePooling
from .recurrent import LSTM, RLayer, GRU, ClockworkLayer
from .fancy import DropOut
