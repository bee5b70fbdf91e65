"""Pin the CPU oracle (oracle/bf_oracle.py) against golden vectors produced by
the real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, relerr
from golden_nets import NETS
from oracle import bf_oracle as O

TOL = 1e-12  # float64 vs float64; only BLAS summation order differs


def close(a, b, tol=TOL):
    assert np.shape(a) == np.shape(b), (np.shape(a), np.shape(b))
    assert relerr(a, b) <= tol, relerr(a, b)


def test_dense(golden_ops):
    g = golden_ops
    close(O.dense_forward(g["dense_X"], g["dense_W"], g["dense_b"]), g["dense_Z"])
    dW, db, dX = O.dense_backward(g["dense_X"], g["dense_E"], g["dense_W"])
    close(dW, g["dense_dW"]); close(db, g["dense_db"]); close(dX, g["dense_dX"])
    close(O.dense_forward(g["dense12_X"], g["dense12_W"], g["dense12_b"]), g["dense12_Z"])


def test_conv(golden_ops):
    g = golden_ops
    close(O.conv_forward(g["conv_A"], g["conv_F"], "valid"), g["conv_valid"])
    close(O.conv_forward(g["conv_A"], g["conv_F"], "full"), g["conv_full"])
    dF, db, dX = O.conv_backward(g["conv_A"], g["conv_E"], g["conv_F"])
    close(dF, g["conv_dF"]); close(db, g["conv_db"]); close(dX, g["conv_dX"])
    close(O.conv_forward(g["conv12_A"], g["conv12_F"], "full"), g["conv12_full"])
    with pytest.raises(ValueError):
        O.conv_valid(np.zeros((1, 2, 5, 5)), np.zeros((1, 3, 3, 3)))


def test_pool_bit_exact(golden_ops):
    g = golden_ops
    for p, fd in (("pool", 2), ("pool3", 3), ("pool12", 2)):
        out, filt = O.maxpool_forward(g[p + "_A"], fd)
        assert np.array_equal(out, g[p + "_out"])
        assert np.array_equal(filt, g[p + "_filt"])
        E = g[p + "_E"] if p + "_E" in g else g[p + "_out"]
        assert np.array_equal(O.maxpool_backward(E, filt), g[p + "_dX"])
    assert g["pool_filt"][0, 0, :2, :2].sum() == 4.0  # the all-equal window is multi-hot


def test_activations(golden_ops):
    g = golden_ops
    for kind in ("sigmoid", "tanh", "relu", "linear"):
        a = O.act_forward(kind, g["act_Z"])
        close(a, g["act_%s_fwd" % kind])
        close(O.act_backward(kind, a) * np.ones_like(a), g["act_%s_bwd" % kind])
    close(O.act_forward("softmax", g["softmax_Z"]), g["softmax_A"])
    assert O.act_backward("softmax", g["softmax_A"]) == 1.0
    assert O.act_backward("relu", np.array([0.0]))[0] == 0.0  # NumPy path: relu'(0) = 0


def test_global_average_pooling_known_answer():
    """The reference's own known-answer test for GlobalAveragePooling (tests/test_layers.py:9-30) on the oracle."""
    inputs = np.empty([3, 5, 5])
    for ch in range(3):
        inputs[ch] = np.full([5, 5], fill_value=ch + 1.0)
    st = O.Stack((3, 5, 5), [("gap",)])
    np.testing.assert_equal(st.feedforward(inputs[None, ...])[0], np.array([1., 2., 3.]))
    np.testing.assert_allclose(st.backpropagate(np.array([[1., 2., 3.]]))[0], inputs / 25)


def test_dropout_against_reference():
    """DropOut (layers/fancy.py:60-90) with the reference's RNG consumption: same seed, same mask."""
    g = load_golden("costs2")
    st = O.Stack((4, 3), [("dropout", 0.3)])
    st.learning = True
    np.random.seed(99)
    assert np.array_equal(st.feedforward(g["drop_X"]), g["drop_train"])
    assert np.array_equal(st.backpropagate(g["drop_D"]), g["drop_back"])
    assert np.array_equal(st.L[0]["mask"], g["drop_mask_after"])
    st.learning = False
    np.random.seed(99)
    close(st.feedforward(g["drop_X"]), g["drop_eval"])


def test_rnn_against_reference():
    """RecurrentOp / RLayer (atomic/recurrent_op.py:9-43, layers/recurrent.py:50-77) against arrays produced by the
    reference (tests/golden/make_golden_rnn.py) -- including its `dX = E[:, :, :indim]` slice."""
    g = load_golden("rnn")
    for tag in "abc":
        X, W, b, E = (g["%s_%s" % (tag, k)] for k in "XWbE")
        for act in ("tanh", "sigmoid", "relu"):
            o, z = O.rnn_forward(X, W, b, act)
            close(o, g["%s_%s_O" % (tag, act)]); close(z, g["%s_%s_Z" % (tag, act)])
            dX, gW, gb = O.rnn_backward(z, o, E, W, act)
            assert dX.shape == g["%s_%s_dX" % (tag, act)].shape
            close(dX, g["%s_%s_dX" % (tag, act)]); close(gW, g["%s_%s_gW" % (tag, act)]); close(gb, g["%s_%s_gb" % (tag, act)])
    for seq in (0, 1):
        k = "layer_seq%d_" % seq
        st = O.Stack((6, 5), [("rnn", 4, "tanh", bool(seq))])
        st.L[0]["W"], st.L[0]["b"] = g[k + "W"].copy(), g[k + "b"].copy()
        close(st.feedforward(g[k + "X"]), g[k + "Y"])
        close(st.backpropagate(g[k + "D"]), g[k + "dX"])
        close(st.L[0]["gW"], g[k + "gW"]); close(st.L[0]["gb"], g[k + "gb"])


def test_gru_against_reference():
    """GRU layer (layers/recurrent.py:116-190) against arrays produced by the reference (make_golden_rnn.py)."""
    g = load_golden("rnn")
    for seq in (0, 1):
        for act in ("tanh", "relu"):
            k = "gru_%s_seq%d_" % (act, seq)
            st = O.Stack((6, 5), [("gru", 4, act, bool(seq))])
            st.L[0]["W"], st.L[0]["b"] = g[k + "W"].copy(), g[k + "b"].copy()
            close(st.feedforward(g[k + "X"]), g[k + "Y"])
            close(st.backpropagate(g[k + "D"]), g[k + "dX"])
            close(st.L[0]["gW"], g[k + "gW"]); close(st.L[0]["gb"], g[k + "gb"])


CW_CASES = {"def": (10, None, None), "cus": (7, [3, 2, 2], [1, 3, 2])}


def test_clockwork_against_reference():
    """ClockworkLayer (layers/recurrent.py:193-283) against arrays produced by the reference: the initialiser's RNG
    consumption and block offsets (one neuron of the custom case never ticks), forward and backward."""
    g = load_golden("rnn")
    for tag, (n, blocks, ticks) in CW_CASES.items():
        for seq in (0, 1):
            k = "cw_%s_seq%d_" % (tag, seq)
            np.random.seed(53)
            st = O.Stack((9, 5), [("clockwork", n, "tanh", bool(seq), blocks, ticks)])
            assert np.array_equal(st.L[0]["W"], g[k + "W"]) and np.array_equal(st.L[0]["tick"], g[k + "tick"])
            st.L[0]["b"] = g[k + "b"].copy()
            close(st.feedforward(g[k + "X"]), g[k + "Y"])
            close(st.backpropagate(g[k + "D"]), g[k + "dX"])
            close(st.L[0]["gW"], g[k + "gW"]); close(st.L[0]["gb"], g[k + "gb"])


def test_costs_bxent_hinge():
    """The remaining costs of the registry (metrics/costs.py:46-69) against values from the reference
    (tests/golden/make_golden_costs.py)."""
    g = load_golden("costs2")
    close(O.cost_value("bxent", g["O"], g["T"]), g["bxent"])
    close(O.cost_derivative(g["O"], g["T"], "bxent"), g["bxent_delta"])
    close(O.cost_value("hinge", g["S"], g["Y"]), g["hinge"])
    assert np.array_equal(O.cost_derivative(g["S"], g["Y"], "hinge"), g["hinge_delta"])


def test_costs(golden_ops):
    g = golden_ops
    close(O.cost_value("mse", g["cost_O"], g["cost_T"]), g["cost_mse"])
    close(O.cost_value("cxent", g["cost_O"], g["cost_T"]), g["cost_cxent"])
    close(O.cost_derivative(g["cost_O"], g["cost_T"]), g["cost_delta"])
    assert O.accuracy(g["cost_O"], g["cost_T"]) == g["cost_acc"]


def test_lstm(golden_ops):
    g = golden_ops
    Ol, Zl, Cl = O.lstm_forward(g["lstm20_X"], g["lstm20_W"], g["lstm20_b"], "tanh")
    close(Ol, g["lstm20_O"]); close(Zl, g["lstm20_Z"]); close(Cl, g["lstm20_cache"])
    for act in ("tanh", "relu"):
        Ol, Zl, Cl = O.lstm_forward(g["lstm_X"], g["lstm_W"], g["lstm_b"], act)
        close(Ol, g["lstm_%s_O" % act]); close(Zl, g["lstm_%s_Z" % act]); close(Cl, g["lstm_%s_cache" % act])
        dX, gW, gb = O.lstm_backward(Zl, Ol, g["lstm_%s_E" % act], g["lstm_W"], Cl, act)
        close(dX, g["lstm_%s_dX" % act]); close(gW, g["lstm_%s_gW" % act]); close(gb, g["lstm_%s_gb" % act])


@pytest.mark.parametrize("name", ["sgd", "momentum", "nesterov", "adagrad", "rmsprop", "adam"])
def test_optimizers(golden_ops, name):
    g = golden_ops
    opt = O.Optimizer(name, g["opt_W0"].size)
    W = g["opt_W0"].copy()
    for step, grad in enumerate(g["opt_g"]):
        W = opt.optimize(W, grad, 7)
        close(W, g["opt_" + name][step])


@pytest.mark.parametrize("name", sorted(NETS))
def test_nets(name):
    seed, ishape, spec, cost, optimizer, batch = NETS[name]
    g = load_golden("net_" + name)
    np.random.seed(seed)
    net = O.Stack(ishape, spec, cost=cost, optimizer=optimizer)
    close(net.get_weights(), g["W0"])          # same RNG consumption as the reference's connect()
    X, Y = g["X"], g["Y"]
    m0 = net.learn_batch(X, Y, update=False, metrics=("acc",))
    close(m0["cost"], g["cost0"], 1e-11)
    assert m0["accuracy"] == g["acc0"]
    close(net.get_gradients(), g["grad0"], 1e-10)
    for li, lay in enumerate(net.L):
        if "out%d" % li in g:
            close(lay["output"], g["out%d" % li], 1e-11)
    preds = net.feedforward(X)
    close(net.backpropagate(O.cost_derivative(preds, Y)), g["input_delta"], 1e-10)
    close(net.get_gradients(), g["grad0_twice"], 1e-10)  # Dense accumulates, Conv/LSTM assign
    net.zero_gradients()
    for step in range(3):
        mm = net.learn_batch(X, Y)
        close(mm["cost"], g["costs"][step], 1e-10)
        close(net.get_weights(), g["weights"][step], 1e-10)


def test_fitness():
    g = load_golden("evolution")
    nin, nh, nout = (int(v) for v in g["dims"])
    net = O.Stack((nin,), [("dense", nh, "sigmoid"), ("dense", nout, "softmax")], cost="cxent")
    fit = [O.fitness(net, gen, g["X"], g["Y"]) for gen in g["genomes"]]
    close(fit, g["fitness"], 1e-12)
    close(O.fitness_mlp_logsoftmax(g["genomes"], g["X"], g["Y"], nin, nh, nout), g["fitness"], 1e-12)
    # wide (+-10) genomes: literal float64 formula may hit log(0); where the reference stayed finite the
    # log-softmax form must agree with it
    wide = O.fitness_mlp_logsoftmax(g["genomes_wide"], g["X"], g["Y"], nin, nh, nout)
    fin = np.isfinite(g["fitness_wide"])
    assert fin.any()
    close(wide[fin], g["fitness_wide"][fin], 1e-10)
