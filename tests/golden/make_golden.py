"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Runs only in the authoring container (needs /root/reference; the GPU box does
not have it).  Usage:

    python tests/golden/make_golden.py

It imports csxeba/brainforge unmodified from /root/reference with the two
environment shims the survey documents (np.asscalar for NumPy>=1.23, HOME
pointed at a scratch dir because the import rewrites ~/.brainforgerc), runs
seeded cases through the reference's NumPy ``atomic`` path (cross-checked
against the Numba ``llatomic`` path where both exist and agree) and writes small
.npz files holding inputs and outputs.  Nothing from the reference is copied
into the repository; only numbers it produced.
"""
import os
import sys
import tempfile

os.environ["HOME"] = tempfile.mkdtemp(prefix="bfhome_")
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

np.asscalar = lambda a: a.item()  # noqa: E731  (util/typing.py:11 needs it)

from brainforge import atomic, llatomic  # noqa: E402
from brainforge.learner import Backpropagation  # noqa: E402
from brainforge.layers import Dense, ConvLayer, PoolLayer, Flatten, Activation, LSTM  # noqa: E402
from brainforge.metrics import costs as ref_costs, metrics as ref_metrics  # noqa: E402
from brainforge.optimizers import optimizers as ref_optimizers  # noqa: E402
from brainforge.learner.neuroevolution import NeuroEvolution  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


# ------------------------------------------------------------------ op level
def ops():
    rs = np.random.RandomState(1234)
    out = {}
    # dense  (atomic/core_op.py:12-20)
    X, W, b = rs.randn(7, 12), rs.randn(12, 5), rs.randn(5)
    E = rs.randn(7, 5)
    out.update(dense_X=X, dense_W=W, dense_b=b, dense_E=E, dense_Z=atomic.DenseOp.forward(X, W, b))
    dW, db, dX = atomic.DenseOp.backward(X, E, W)
    out.update(dense_dW=dW, dense_db=db, dense_dX=dX)
    # the reference's own dense test shape (tests/test_numba_ops.py:61-72), numpy vs numba
    A12, W12, b12 = rs.uniform(size=(12, 12)), rs.uniform(size=(12, 12)), rs.uniform(size=(12,))
    z_np = atomic.DenseOp.forward(A12, W12, b12)
    assert np.allclose(z_np, llatomic.DenseOp.forward(A12, W12, b12))
    out.update(dense12_X=A12, dense12_W=W12, dense12_b=b12, dense12_Z=z_np)

    # convolution (atomic/tensor_op.py:9-61); numpy vs numba cross-check
    A, F = rs.randn(3, 2, 7, 6), rs.randn(4, 2, 3, 2)
    npc, nbc = atomic.ConvolutionOp(), llatomic.ConvolutionOp()
    yv, yf = npc.forward(A, F, "valid"), npc.forward(A, F, "full")
    assert np.allclose(yv, nbc.forward(A, F, "valid")) and np.allclose(yf, nbc.forward(A, F, "full"))
    Ec = rs.randn(*yv.shape)
    dF, dbc, dXc = npc.backward(X=A, E=Ec, F=F)
    dF2, dbc2, dXc2 = nbc.backward(X=A, E=Ec, F=F)
    assert np.allclose(dF, dF2) and np.allclose(dbc.ravel(), dbc2.ravel()) and np.allclose(dXc, dXc2)
    out.update(conv_A=A, conv_F=F, conv_valid=yv, conv_full=yf, conv_E=Ec, conv_dF=dF, conv_db=dbc, conv_dX=dXc)
    # the reference's own conv test shape (tests/test_numba_ops.py:27-37): full mode, 1x1x12x12 * 1x1x3x3
    A1, F1 = rs.uniform(size=(1, 1, 12, 12)), rs.uniform(size=(1, 1, 3, 3))
    y1 = npc.forward(A1, F1, mode="full")
    assert np.allclose(y1, nbc.forward(A1, F1, mode="full"))
    out.update(conv12_A=A1, conv12_F=F1, conv12_full=y1)

    # max pool with forced ties (atomic/tensor_op.py:93-119); numpy vs numba incl. mask and backward
    P = np.round(rs.randn(2, 3, 8, 6) * 2.0) / 2.0      # coarse grid -> many exact ties
    P[0, 0, :2, :2] = 0.0                               # an all-equal window
    npp, nbp = atomic.MaxPoolOp(), llatomic.MaxPoolOp()
    po, pf = npp.forward(P, 2)
    po2, pf2 = nbp.forward(P, 2)
    assert np.array_equal(po, po2) and np.array_equal(pf, pf2)
    Ep = rs.randn(*po.shape)
    pb = npp.backward(Ep, pf.copy())
    assert np.array_equal(pb, nbp.backward(Ep, pf2.copy()))
    out.update(pool_A=P, pool_out=po, pool_filt=pf, pool_E=Ep, pool_dX=pb)
    P3 = rs.randn(2, 2, 9, 6)
    po3, pf3 = npp.forward(P3, 3)
    Ep3 = rs.randn(*po3.shape)
    out.update(pool3_A=P3, pool3_out=po3, pool3_filt=pf3, pool3_E=Ep3, pool3_dX=npp.backward(Ep3, pf3.copy()))
    # reference test shape (tests/test_numba_ops.py:42-56)
    P12 = rs.uniform(0., 1., (1, 1, 12, 12))
    o12, f12 = npp.forward(P12, 2)
    out.update(pool12_A=P12, pool12_out=o12, pool12_filt=f12, pool12_dX=npp.backward(o12, f12.copy()))

    # activations fwd / bwd-from-output (atomic/activation_op.py) ; tests/test_numba_activations.py uses randn(100)
    Zs = np.concatenate([rs.randn(100), [0.0, -0.0, 30.0, -30.0, 1e-8]])
    for kind in ("sigmoid", "tanh", "relu", "linear"):
        act = atomic.activations[kind]()
        a = act.forward(Zs)
        out["act_%s_fwd" % kind] = a
        out["act_%s_bwd" % kind] = np.asarray(act.backward(a), dtype=float) * np.ones_like(Zs)
    out["act_Z"] = Zs
    Zm = rs.randn(9, 10) * 3
    out.update(softmax_Z=Zm, softmax_A=atomic.activations["softmax"]().forward(Zm))

    # costs / metrics (metrics/costs.py, metrics/metrics.py)
    O = atomic.activations["softmax"]().forward(rs.randn(11, 6))
    T = np.eye(6)[rs.randint(0, 6, 11)]
    out.update(cost_O=O, cost_T=T, cost_mse=ref_costs.mse(O, T), cost_cxent=ref_costs.cxent(O, T),
               cost_delta=ref_costs.mse.derivative(O, T), cost_acc=ref_metrics.accuracy(O, T))

    # LSTM op (atomic/recurrent_op.py:51-112); reference test shape T=2,B=20,D=10,H=15 (tests/test_numba_ops.py:91-115)
    Xl, Wl, bl = rs.randn(2, 20, 10), rs.randn(25, 60), rs.randn(60)
    Ol, Zl, Cl = atomic.LSTMOp("tanh").forward(Xl, Wl, bl)
    O2, Z2, C2 = llatomic.LSTMOp("tanh").forward(Xl, Wl, bl)
    assert np.allclose(Ol, O2) and np.allclose(Zl, Z2) and np.allclose(Cl, C2)
    out.update(lstm20_X=Xl, lstm20_W=Wl, lstm20_b=bl, lstm20_O=Ol, lstm20_Z=Zl, lstm20_cache=Cl)
    Xl, Wl, bl = rs.randn(5, 4, 3), rs.randn(9, 24) * 0.5, rs.randn(24) * 0.5
    for act in ("tanh", "relu"):
        op = atomic.LSTMOp(act)
        Ol, Zl, Cl = op.forward(Xl, Wl, bl)
        El = rs.randn(*Ol.shape)
        dXl, gWl, gbl = op.backward(Z=Zl, O=Ol, E=El.copy(), W=Wl, cache=Cl)
        out.update({"lstm_%s_O" % act: Ol, "lstm_%s_Z" % act: Zl, "lstm_%s_cache" % act: Cl,
                    "lstm_%s_E" % act: El, "lstm_%s_dX" % act: dXl, "lstm_%s_gW" % act: gWl,
                    "lstm_%s_gb" % act: gbl})
    out.update(lstm_X=Xl, lstm_W=Wl, lstm_b=bl)

    # optimizers, 3 steps each on the same (W, g) stream (optimizers/*.py)
    n = 50
    W0 = rs.randn(n)
    gs = rs.randn(3, n) * 4
    out.update(opt_W0=W0, opt_g=gs)
    for name in ("sgd", "momentum", "nesterov", "adagrad", "rmsprop", "adam"):
        opt = ref_optimizers[name]()
        opt.initialize(nparams=n)
        W = W0.copy()
        traj = []
        for g in gs:
            W = opt.optimize(W, g.copy(), 7)
            traj.append(W.copy())
        out["opt_" + name] = np.array(traj)
    save("ops", **out)


# ----------------------------------------------------------------- net level
def build(seed, input_shape, layers, cost, optimizer):
    np.random.seed(seed)
    return Backpropagation(input_shape=input_shape, layerstack=layers, cost=cost, optimizer=optimizer)


sys.path.insert(0, os.path.dirname(HERE))
from golden_nets import NETS  # noqa: E402  (shared with the tests)


def ref_layers(spec, compiled):
    out = []
    for s in spec:
        k = s[0]
        if k == "dense":
            out.append(Dense(s[1], activation=s[2]))
        elif k == "act":
            out.append(Activation(s[1]))
        elif k == "flatten":
            out.append(Flatten())
        elif k == "conv":
            out.append(ConvLayer(s[1], s[2], s[3], compiled=compiled, activation=s[4]))
        elif k == "pool":
            out.append(PoolLayer(s[1], compiled=compiled))
        elif k == "lstm":
            out.append(LSTM(s[1], activation=s[2], return_seq=s[3], compiled=False))
    return out


def nets():
    for name, (seed, ishape, spec, cost, optimizer, batch) in NETS.items():
        results = []
        for compiled in (False, True):
            net = build(seed, ishape, ref_layers(spec, compiled), cost, optimizer)
            rs = np.random.RandomState(seed + 1000)
            nout = net.layers.outshape[-1]
            X = rs.randn(batch, *ishape)
            if cost == "cxent":
                Y = np.eye(nout)[rs.randint(0, nout, batch)]
            else:
                Y = rs.randn(batch, nout)
            W0 = net.layers.get_weights(unfold=True).copy()
            r = dict(X=X, Y=Y, W0=W0)
            # step 0 without update: record per-layer outputs, gradient, input delta
            m0 = net.learn_batch(X.copy(), Y.copy(), metrics=[ref_metrics.accuracy], update=False)
            r["cost0"] = m0["cost"]
            r["acc0"] = m0["accuracy"]
            r["grad0"] = net.get_gradients(unfold=True).copy()
            for li, layer in enumerate(net.layers.layers[1:]):
                if layer.output is not None and not isinstance(layer, Flatten):
                    o = np.asarray(layer.output)
                    if isinstance(layer, LSTM):
                        o = o.transpose(1, 0, 2) if layer.return_seq else o[-1]
                    r["out%d" % li] = o.copy()
                if isinstance(layer, PoolLayer):
                    pass  # the cached mask was consumed by backward (tensor_op.py:118-119)
            preds = net.predict(X.copy())
            r["input_delta"] = np.asarray(net.backpropagate(net.cost.derivative(preds, Y))).copy()
            # Dense accumulates, Conv/LSTM assign: the second backprop above doubles only Dense grads
            r["grad0_twice"] = net.get_gradients(unfold=True).copy()
            net.zero_gradients()
            costs, weights = [], []
            for step in range(3):
                mm = net.learn_batch(X.copy(), Y.copy())
                costs.append(mm["cost"])
                weights.append(net.layers.get_weights(unfold=True).copy())
            r["costs"] = np.array(costs)
            r["weights"] = np.array(weights)
            results.append(r)
        a, b = results
        for k in a:  # numpy and numba conv/pool back-ends must agree before we pin anything
            assert np.allclose(a[k], b[k], rtol=1e-9, atol=1e-11), (name, k)
        save("net_" + name, **a)


# ------------------------------------------------------------ neuro-evolution
def evolution():
    np.random.seed(21)
    nin, nh, nout, P, N = 7, 5, 3, 6, 9
    ne = NeuroEvolution(input_shape=(nin,), layerstack=[Dense(nh, activation="sigmoid"), Dense(nout, activation="softmax")],
                        cost="cxent", population_size=P)
    rs = np.random.RandomState(22)
    X = rs.randn(N, nin)
    Y = np.eye(nout)[rs.randint(0, nout, N)]
    # scale genomes towards 0.5 so the literal float64 -sum(t*log o) stays finite for the pin
    genomes = 0.5 + (ne.population.individuals - 0.5) * 0.2
    fit = []
    for g in genomes:
        ne.layers.set_weights(ne.as_weights(g))          # learner/neuroevolution.py:35
        out = ne.layers.feedforward(X)
        fit.append(ne.cost(out, Y) / len(X))             # what evaluate()["cost"] computes (abstract_learner.py:91-94)
    genomes_wide = ne.population.individuals             # full +-10 weight range (underflow territory)
    fit_wide = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for g in genomes_wide:
            ne.layers.set_weights(ne.as_weights(g))
            out = ne.layers.feedforward(X)
            fit_wide.append(ne.cost(out, Y) / len(X))
    save("evolution", X=X, Y=Y, genomes=genomes, fitness=np.array(fit), genomes_wide=genomes_wide,
         fitness_wide=np.array(fit_wide), dims=np.array([nin, nh, nout]))


if __name__ == "__main__":
    ops()
    nets()
    evolution()
