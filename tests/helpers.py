"""Shared test helpers: build the same network on the B200 back-end and on the CPU oracle."""
import numpy as np


def build_b200(input_shape, spec, cost, optimizer, seed=None, **kw):
    import brainforge_b200 as bf
    from brainforge_b200.layers import Dense, ConvLayer, PoolLayer, Flatten, Activation, LSTM
    if seed is not None:
        np.random.seed(seed)
    layers = []
    for s in spec:
        k = s[0]
        if k == "dense":
            layers.append(Dense(s[1], activation=s[2] if len(s) > 2 else "linear"))
        elif k == "act":
            layers.append(Activation(s[1]))
        elif k == "flatten":
            layers.append(Flatten())
        elif k == "conv":
            layers.append(ConvLayer(s[1], s[2], s[3], activation=s[4] if len(s) > 4 else "linear"))
        elif k == "pool":
            layers.append(PoolLayer(s[1]))
        elif k == "dropout":
            from brainforge_b200.layers import DropOut
            layers.append(DropOut(s[1]))
        elif k == "gap":
            from brainforge_b200.layers import GlobalAveragePooling
            layers.append(GlobalAveragePooling())
        elif k == "lstm":
            layers.append(LSTM(s[1], activation=s[2], return_seq=s[3] if len(s) > 3 else False))
        elif k == "clockwork":
            from brainforge_b200.layers import ClockworkLayer
            layers.append(ClockworkLayer(s[1], s[2], blocksizes=s[4] if len(s) > 4 else None,
                                         ticktimes=s[5] if len(s) > 5 else None, return_seq=s[3] if len(s) > 3 else False))
        elif k == "gru":
            from brainforge_b200.layers import GRU
            layers.append(GRU(s[1], activation=s[2], return_seq=s[3] if len(s) > 3 else False))
        elif k == "rnn":
            from brainforge_b200.layers import RLayer
            layers.append(RLayer(s[1], activation=s[2], return_seq=s[3] if len(s) > 3 else False))
        else:
            raise ValueError(k)
    return bf.Backpropagation(input_shape=input_shape, layerstack=layers, cost=cost, optimizer=optimizer, **kw)


def build_oracle(input_shape, spec, cost, optimizer, seed=None):
    from oracle import bf_oracle as O
    if seed is not None:
        np.random.seed(seed)
    return O.Stack(input_shape, spec, cost=cost, optimizer=optimizer)


def onehot(rs, n, classes):
    return np.eye(classes)[rs.randint(0, classes, n)]


CFG = {
    # name: (input_shape, spec, cost, optimizer)   -- SURVEY 8d / BASELINE.md section 2
    "cfg1": ((784,), [("dense", 60, "sigmoid"), ("dense", 10, "softmax")], "cxent", "sgd"),
    "cfg2": ((1, 28, 28), [("conv", 8, 3, 3, "linear"), ("pool", 2), ("act", "relu"), ("flatten",),
                           ("dense", 10, "softmax")], "cxent", "sgd"),
    "cfg3": ((3, 30, 30), [("conv", 32, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                           ("conv", 64, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                           ("conv", 128, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                           ("flatten",), ("dense", 10, "softmax")], "cxent", "adam"),
    "cfg4": ((64, 128), [("lstm", 256, "tanh", False), ("dense", 10, "softmax")], "cxent", "sgd"),
}
