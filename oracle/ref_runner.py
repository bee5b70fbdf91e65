"""Runs the UNMODIFIED reference (csxeba/brainforge installed under oracle/_ref by oracle/build_ref.sh) through its own
public API: test / bench infrastructure only (bench.py --impl reference, bench.py's cpu_baseline leg).

The two environment shims SURVEY 8c documents are applied before the import: `np.asscalar` (removed from NumPy >= 1.23,
used at util/typing.py:11) and HOME pointed at a scratch directory (the import writes ~/.brainforgerc,
config/parser.py:18-23).  Back-end choice per layer follows SURVEY 8d: Dense and LSTM on the NumPy `atomic` ops
(`compiled=False`: BLAS dgemm; the Numba DenseOp.backward is NotImplemented and the Numba LSTM is slower), ConvLayer /
PoolLayer on the Numba `llatomic` ops (`compiled=True`, ~15x faster than the NumPy patch loops).
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "brainforge"))


def load():
    """-> the reference's top-level package, imported from oracle/_ref."""
    if not available():
        raise ImportError("oracle/_ref is empty: run `bash oracle/build_ref.sh` where /root/reference exists")
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()          # noqa: E731
    if "brainforge" not in sys.modules:
        os.environ["HOME"] = tempfile.mkdtemp(prefix="bfhome_")
        sys.path.insert(0, REF)
    import brainforge
    assert os.path.dirname(os.path.abspath(brainforge.__file__)).startswith(REF), brainforge.__file__
    return brainforge


def build(input_shape, spec, cost, optimizer, seed=1234):
    """The reference's Backpropagation for one of bench.py's / tests' layer specs."""
    load()
    from brainforge.learner import Backpropagation
    from brainforge.layers import Dense, ConvLayer, PoolLayer, Flatten, Activation, LSTM
    np.random.seed(seed)
    layers = []
    for s in spec:
        kind = s[0]
        if kind == "dense":
            layers.append(Dense(s[1], activation=s[2], compiled=False))
        elif kind == "act":
            layers.append(Activation(s[1]))
        elif kind == "flatten":
            layers.append(Flatten())
        elif kind == "conv":
            layers.append(ConvLayer(s[1], s[2], s[3], activation=s[4], compiled=True))
        elif kind == "pool":
            layers.append(PoolLayer(s[1], compiled=True))
        elif kind == "lstm":
            layers.append(LSTM(s[1], activation=s[2], return_seq=s[3], compiled=False))
        else:
            raise ValueError(kind)
    return Backpropagation(input_shape=input_shape, layerstack=layers, cost=cost, optimizer=optimizer)


def time_learn_batch(net, X, Y, steps, warmup):
    """-> (samples/s, seconds/step) of `learn_batch` (fwd + bwd + update), JIT warm-up excluded."""
    for _ in range(max(warmup, 1)):
        net.learn_batch(X, Y)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.learn_batch(X, Y)
    dt = time.perf_counter() - t0
    return len(X) * steps / dt, dt / steps


def fitness_individuals_per_s(nin, nh, nout, P, X, Y, steps, warmup):
    """cfg5 through the reference's own objects: Population.update's serial loop (evolution/_evolution.py:83-97) with
    the fitness NeuroEvolution intends (set_weights(as_weights(g)) -> feedforward -> cost/m; its shipped default raises
    TypeError)."""
    load()
    from brainforge.learner import NeuroEvolution
    from brainforge.layers import Dense
    np.random.seed(1234)
    holder = {}

    def fitness(genome, X, Y):
        ne = holder["ne"]
        ne.layers.set_weights(ne.as_weights(genome))
        with np.errstate(divide="ignore", invalid="ignore"):
            return (ne.cost(ne.layers.feedforward(X), Y) / len(X),)

    ne = NeuroEvolution(input_shape=(nin,), layerstack=[Dense(nh, activation="sigmoid", compiled=False),
                                                        Dense(nout, activation="softmax", compiled=False)],
                        cost="cxent", population_size=P, fitness_function=fitness)
    holder["ne"] = ne
    for _ in range(max(warmup, 1)):
        ne.population.update(np.arange(min(P, 4)), 0, X=X, Y=Y)
    t0 = time.perf_counter()
    for _ in range(steps):
        ne.population.update(None, 0, X=X, Y=Y)
    dt = time.perf_counter() - t0
    return P * steps / dt, dt / steps
