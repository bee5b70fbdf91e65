"""Golden fixtures for the plain recurrent layer (atomic/recurrent_op.py:9-43, layers/recurrent.py:50-77), generated from
the REAL reference like make_golden.py (authoring container only):

    python tests/golden/make_golden_rnn.py
"""
import os
import sys
import tempfile

os.environ["HOME"] = tempfile.mkdtemp(prefix="bfhome_")
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

np.asscalar = lambda a: a.item()  # noqa: E731  (util/typing.py:11 needs it)
from brainforge.atomic import RecurrentOp  # noqa: E402
from brainforge.layers import RLayer, GRU, ClockworkLayer  # noqa: E402
from brainforge.util import testing  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rs = np.random.RandomState(777)
out = {}
# operator level: D > H, D < H (the reference's dX slice then has only H columns) and D == H
for tag, (T, B, D, H) in {"a": (5, 7, 6, 4), "b": (4, 3, 2, 5), "c": (3, 9, 4, 4)}.items():
    X = rs.randn(T, B, D)
    W = rs.randn(D + H, H) * np.sqrt(2.0 / (D + 2 * H))
    b = rs.randn(H) * 0.1
    E = rs.randn(T, B, H)
    out.update({"%s_X" % tag: X, "%s_W" % tag: W, "%s_b" % tag: b, "%s_E" % tag: E})
    for act in ("tanh", "sigmoid", "relu"):
        op = RecurrentOp(act)
        O, Z = op.forward(X, W, b)
        dX, gW, gb = op.backward(Z=Z, O=O, E=E.copy(), W=W)
        out.update({"%s_%s_O" % (tag, act): O, "%s_%s_Z" % (tag, act): Z, "%s_%s_dX" % (tag, act): np.array(dX),
                    "%s_%s_gW" % (tag, act): gW, "%s_%s_gb" % (tag, act): gb})
# layer level: batch-major input, both return_seq settings
for seq in (False, True):
    brain = testing.NoBrainer(outshape=(6, 5))
    np.random.seed(31)
    layer = RLayer(4, "tanh", return_seq=seq)
    layer.connect(brain)
    LX = rs.randn(8, 6, 5)
    y = np.array(layer.feedforward(LX))
    LD = rs.randn(*y.shape)
    dx = np.array(layer.backpropagate(LD.copy()))
    k = "layer_seq%d_" % int(seq)
    out.update({k + "X": LX, k + "W": layer.weights.copy(), k + "b": layer.biases.copy(), k + "Y": y, k + "D": LD,
                k + "dX": dx, k + "gW": layer.nabla_w.copy(), k + "gb": layer.nabla_b.copy()})
# GRU (layers/recurrent.py:116-190): layer level only (the reference has no atomic op for it)
for seq in (False, True):
    for act in ("tanh", "relu"):
        brain = testing.NoBrainer(outshape=(6, 5))
        np.random.seed(47)
        layer = GRU(4, act, return_seq=seq)
        layer.connect(brain)
        layer.biases = rs.randn(12) * 0.1
        GX = rs.randn(8, 6, 5)
        y = np.array(layer.feedforward(GX))
        GD = rs.randn(*y.shape)
        dx = np.array(layer.backpropagate(GD.copy()))
        k = "gru_%s_seq%d_" % (act, int(seq))
        out.update({k + "X": GX, k + "W": layer.weights.copy(), k + "b": layer.biases.copy(), k + "Y": y, k + "D": GD,
                    k + "dX": dx, k + "gW": layer.nabla_w.copy(), k + "gb": layer.nabla_b.copy()})
# ClockworkLayer (layers/recurrent.py:193-283): default blocks (5 x 2, ticks 1,2,4,8,16) and a custom split
import contextlib  # noqa: E402
import io  # noqa: E402
for tag, (neurons, kw) in {"def": (10, {}), "cus": (7, dict(blocksizes=[3, 2, 2], ticktimes=[1, 3, 2]))}.items():
    for seq in (False, True):
        brain = testing.NoBrainer(outshape=(9, 5))
        np.random.seed(53)
        with contextlib.redirect_stdout(io.StringIO()):
            layer = ClockworkLayer(neurons, "tanh", return_seq=seq, **kw)
            layer.connect(brain)
        layer.biases = rs.randn(neurons) * 0.1
        CX = rs.randn(6, 9, 5)
        y = np.array(layer.feedforward(CX))
        CD = rs.randn(*y.shape)
        dx = np.array(layer.backpropagate(CD.copy()))
        k = "cw_%s_seq%d_" % (tag, int(seq))
        out.update({k + "X": CX, k + "W": layer.weights.copy(), k + "b": layer.biases.copy(), k + "tick": layer.tick_array.copy(),
                    k + "Y": y, k + "D": CD, k + "dX": dx, k + "gW": layer.nabla_w.copy(), k + "gb": layer.nabla_b.copy()})
path = os.path.join(HERE, "rnn.npz")
np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), len(out), "arrays")
