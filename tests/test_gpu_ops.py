"""Op-level parity on the B200: libbf_b200 (through the `b200atomic` mirror of the reference's op API)
against the golden vectors produced by the reference and against the float64 CPU oracle.

Tolerances (stated): FP32 SIMT path -- relative error ||a-b||/max(||a||,||b||) <= 1e-5 for
every float result; pooling outputs / masks / pool gradients and argmax counts are bit-exact.
"""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def bf():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import brainforge_b200 as bf
    bf.set_precision("fp32")
    return bf


@pytest.fixture(scope="module")
def O():
    from oracle import bf_oracle
    return bf_oracle


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert relerr(a, b) <= tol, relerr(a, b)


# ---------------------------------------------------------------- golden vectors from the reference
def test_dense_golden(bf, golden_ops):
    g = golden_ops
    op = bf.atomic.DenseOp()
    close(op.forward(g["dense_X"], g["dense_W"], g["dense_b"]), g["dense_Z"])
    dW, db, dX = op.backward(g["dense_X"], g["dense_E"], g["dense_W"])
    close(dW, g["dense_dW"]); close(db, g["dense_db"]); close(dX, g["dense_dX"])
    close(op.forward(g["dense12_X"], g["dense12_W"], g["dense12_b"]), g["dense12_Z"])   # tests/test_numba_ops.py:61-72


def test_conv_golden(bf, golden_ops):
    g = golden_ops
    op = bf.atomic.ConvolutionOp()
    close(op.forward(g["conv_A"], g["conv_F"], "valid"), g["conv_valid"])
    close(op.forward(g["conv_A"], g["conv_F"], "full"), g["conv_full"])
    dF, db, dX = op.backward(X=g["conv_A"], E=g["conv_E"], F=g["conv_F"])
    close(dF, g["conv_dF"]); close(db, g["conv_db"]); close(dX, g["conv_dX"])
    close(op.forward(g["conv12_A"], g["conv12_F"], mode="full"), g["conv12_full"])      # tests/test_numba_ops.py:27-37
    with pytest.raises(ValueError, match="filter depth"):
        op.forward(np.zeros((1, 2, 5, 5)), np.zeros((1, 3, 3, 3)))


def test_pool_golden_bit_exact(bf, golden_ops):
    g = golden_ops
    op = bf.atomic.MaxPoolOp()
    for p, fd in (("pool", 2), ("pool3", 3), ("pool12", 2)):     # pool12: tests/test_numba_ops.py:42-56
        A32 = g[p + "_A"].astype(np.float32)
        assert np.array_equal(A32.astype(np.float64), g[p + "_A"]) or p != "pool"   # the tie fixture is FP32-exact
        out, filt = op.forward(g[p + "_A"], fd)
        if p == "pool":
            assert np.array_equal(out, g[p + "_out"])
            assert np.array_equal(filt, g[p + "_filt"])       # multi-hot tie mask, bit-exact
        else:
            close(out, g[p + "_out"], 1e-7)
            assert np.array_equal(filt, g[p + "_filt"])       # distinct random values: same argmax after rounding
        E = g[p + "_E"] if p + "_E" in g else g[p + "_out"]
        dX = op.backward(E, filt)
        close(dX, g[p + "_dX"], 1e-7)
        assert np.array_equal(dX != 0, g[p + "_dX"] != 0)


def test_activations_golden(bf, golden_ops):
    g = golden_ops
    for kind in ("sigmoid", "tanh", "relu", "linear"):
        act = bf.atomic.activations[kind]()
        a = act.forward(g["act_Z"])
        close(a, g["act_%s_fwd" % kind])
        bw = act.backward(g["act_%s_fwd" % kind])
        close(np.asarray(bw) * np.ones(len(g["act_Z"])), g["act_%s_bwd" % kind])
    close(bf.atomic.SoftMax().forward(g["softmax_Z"]), g["softmax_A"])
    assert bf.atomic.ReLU().backward(np.zeros(3))[0] == 0.0   # NumPy semantics: relu'(0) = 0


def test_costs_and_accuracy_golden(bf, golden_ops):
    g = golden_ops
    close(bf.metrics.mse(g["cost_O"], g["cost_T"]), g["cost_mse"])
    close(bf.metrics.cxent(g["cost_O"], g["cost_T"]), g["cost_cxent"])
    close(bf.metrics.mse.derivative(g["cost_O"], g["cost_T"]), g["cost_delta"])
    assert bf.metrics.accuracy(g["cost_O"], g["cost_T"]) == int(g["cost_acc"])


def test_lstm_golden(bf, golden_ops):
    g = golden_ops
    O_, Z_, C_ = bf.atomic.LSTMOp("tanh").forward(g["lstm20_X"], g["lstm20_W"], g["lstm20_b"])   # tests/test_numba_ops.py:91-115
    close(O_, g["lstm20_O"]); close(Z_, g["lstm20_Z"]); close(C_, g["lstm20_cache"])
    for act in ("tanh", "relu"):
        op = bf.atomic.LSTMOp(act)
        O_, Z_, C_ = op.forward(g["lstm_X"], g["lstm_W"], g["lstm_b"])
        close(O_, g["lstm_%s_O" % act]); close(Z_, g["lstm_%s_Z" % act]); close(C_, g["lstm_%s_cache" % act])
        dX, gW, gb = op.backward(Z=g["lstm_%s_Z" % act], O=g["lstm_%s_O" % act], E=g["lstm_%s_E" % act],
                                 W=g["lstm_W"], cache=g["lstm_%s_cache" % act])
        close(dX, g["lstm_%s_dX" % act]); close(gW, g["lstm_%s_gW" % act]); close(gb, g["lstm_%s_gb" % act])


@pytest.mark.parametrize("name", ["sgd", "momentum", "nesterov", "adagrad", "rmsprop", "adam"])
def test_optimizers_golden(bf, golden_ops, name):
    g = golden_ops
    opt = bf.optimizers.optimizers[name]()
    opt.initialize(nparams=g["opt_W0"].size)
    W = g["opt_W0"].copy()
    for step, grad in enumerate(g["opt_g"]):
        W = opt.optimize(W, grad, 7)
        close(W, g["opt_" + name][step], 2e-6)


@pytest.mark.parametrize("world", [1, 4])
@pytest.mark.parametrize("name", ["sgd", "adam"])
def test_opt_step_allreduce_simulated_ranks(bf, name, world):
    """bf_opt_step_allreduce with `world` simulated ranks on ONE device (one stream per rank, receive buffers and signal
    pads in ordinary device memory): every replica ends at optimize(W, sum of the ranks' gradients), gradients are
    cleared, the step scalars at the tail hold their sums, slot 2 of the tail the summed sample counts.  Three steps, so
    that both parity halves of the receive buffers are reused.  (Across real GPUs: tests/test_gpu_multi.py.)"""
    import torch
    from brainforge_b200 import _lib
    from brainforge_b200.device import DeviceArray, call
    rs = np.random.RandomState(world)
    n_params, n_total, blocks, m = 4096 + 64, 4096 + 128, 4, 37.0
    W0 = rs.randn(n_params).astype(np.float32)
    Gs = [rs.randn(n_total).astype(np.float32) for _ in range(world)]
    dev = torch.device("cuda")
    G = [torch.tensor(g, device=dev) for g in Gs]
    W = [torch.tensor(np.concatenate([W0, np.zeros(n_total - n_params, np.float32)]), device=dev) for _ in range(world)]
    recv = [torch.full((2 * world * n_total,), float("nan"), device=dev) for _ in range(world)]
    pads = [torch.zeros(blocks * world, dtype=torch.int32, device=dev) for _ in range(world)]
    recv_t = torch.tensor([g.data_ptr() for g in recv], dtype=torch.int64, device=dev)
    pads_t = torch.tensor([p.data_ptr() for p in pads], dtype=torch.int64, device=dev)
    epochs = [torch.zeros(blocks, dtype=torch.int32, device=dev) for _ in range(world)]
    timeouts = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(world)]
    opts = [bf.optimizers.optimizers[name]() for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for step in range(3):
        for r in range(world):
            s1, s2 = opts[r]._state(n_params)
            hp = opts[r]._hp()
            call(_lib.lib.bf_opt_step_allreduce, _lib.OPT[name], W[r].data_ptr(), G[r].data_ptr(), s1, s2, n_params, n_total,
                 float(opts[r].eta), m, float(hp[0]), float(hp[1]), float(hp[2]), 1, recv_t.data_ptr(), pads_t.data_ptr(),
                 epochs[r].data_ptr(), timeouts[r].data_ptr(), float(5 + r), r, world, blocks, streams[r].cuda_stream)
        torch.cuda.synchronize()
        assert all(int(t[0]) == 0 for t in timeouts), "a simulated rank timed out at the exchange"
        total = np.sum(np.stack(Gs).astype(np.float64), axis=0)
        total[n_params + 2] = sum(5 + r for r in range(world))          # the kernel posts each rank's sample count there
        if step == 0:
            ref = bf.optimizers.optimizers[name]()
            Wref = DeviceArray.from_host(W0)
        gref = DeviceArray.from_host(total[:n_params])
        ref.optimize(Wref, gref, m)
        for r in range(world):
            close(W[r][:n_params].cpu().numpy(), Wref.numpy(), 1e-6)
            assert not G[r][:n_params].any().item()
            close(G[r][n_params:].cpu().numpy(), total[n_params:], 1e-6)
            assert torch.equal(W[r], W[0])
        Gs = [rs.randn(n_total).astype(np.float32) for _ in range(world)]       # next step's local gradients
        for r in range(world):
            G[r].copy_(torch.tensor(Gs[r], device=dev))
        torch.cuda.synchronize()


def test_opt_step_allreduce_times_out_without_touching_the_replica(bf):
    """A peer that never arrives: the kernel gives up, raises the host-visible flag and leaves W and the optimizer state as
    they were (a straggler must not make the replicas diverge silently).  The wait is ~10 s of GPU time."""
    import torch
    from brainforge_b200 import _lib
    from brainforge_b200.device import call
    dev = torch.device("cuda")
    n_params, n_total, blocks, world = 256, 320, 2, 2
    W = torch.ones(n_total, device=dev)
    G = torch.ones(n_total, device=dev)
    recv = [torch.zeros(2 * world * n_total, device=dev) for _ in range(world)]
    pads = [torch.zeros(blocks * world, dtype=torch.int32, device=dev) for _ in range(world)]
    recv_t = torch.tensor([g.data_ptr() for g in recv], dtype=torch.int64, device=dev)
    pads_t = torch.tensor([p.data_ptr() for p in pads], dtype=torch.int64, device=dev)
    epoch = torch.zeros(blocks, dtype=torch.int32, device=dev)
    timeout = torch.zeros(4, dtype=torch.int32).pin_memory()
    call(_lib.lib.bf_opt_step_allreduce, _lib.OPT["sgd"], W.data_ptr(), G.data_ptr(), None, None, n_params, n_total, 0.1, 4.0,
         0.0, 0.0, 0.0, 1, recv_t.data_ptr(), pads_t.data_ptr(), epoch.data_ptr(), timeout.data_ptr(), 4.0, 0, world, blocks,
         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(timeout[0]) == 1
    assert torch.equal(W, torch.ones_like(W)), "the replica was updated from a partial sum"


# ---------------------------------------------------------------- oracle at ragged / awkward shapes
@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (20, 784, 60), (20, 60, 10), (33, 17, 5), (257, 130, 129), (4096, 512, 10)])
def test_dense_vs_oracle(bf, O, m, k, n):
    rs = np.random.RandomState(m * 131 + k)
    X, W, b, E = rs.randn(m, k), rs.randn(k, n) * 0.2, rs.randn(n), rs.randn(m, n)
    op = bf.atomic.DenseOp()
    close(op.forward(X, W, b), O.dense_forward(X, W, b))
    for act in ("sigmoid", "tanh", "relu", "softmax"):
        close(op.forward(X, W, b, activation=act), O.act_forward(act, O.dense_forward(X, W, b)))
    dW, db, dX = op.backward(X, E, W)
    rW, rb, rX = O.dense_backward(X, E, W)
    close(dW, rW); close(db, rb); close(dX, rX)


@pytest.mark.parametrize("shape", [
    (2, 1, 28, 28, 8, 3, 3), (3, 3, 30, 30, 32, 3, 3), (2, 32, 14, 14, 64, 3, 3), (5, 64, 6, 6, 128, 3, 3),
    (1, 1, 3, 3, 1, 3, 3), (2, 5, 9, 7, 3, 2, 4), (1, 2, 6, 6, 17, 5, 1), (3, 4, 8, 8, 6, 1, 1),
])
def test_conv_vs_oracle(bf, O, shape):
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape))
    A, F, bias = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.3, rs.randn(nf)
    op = bf.atomic.ConvolutionOp()
    yv = O.conv_forward(A, F, "valid")
    close(op.forward(A, F, "valid"), yv)
    close(op.forward(A, F, "full"), O.conv_forward(A, F, "full"))
    # layer epilogue: act(conv) + bias, bias AFTER the activation (layers/tensor.py:77-78)
    close(op.forward(A, F, "valid", bias=bias, activation="tanh"), np.tanh(yv) + bias[None, :, None, None])
    E = rs.randn(*yv.shape)
    dF, db, dX = op.backward(X=A, E=E, F=F)
    rF, rb, rX = O.conv_backward(A, E, F)
    close(dF, rF); close(db, rb); close(dX, rX)


@pytest.mark.parametrize("shape,fdim", [((2, 8, 26, 26), 2), ((3, 32, 28, 28), 2), ((2, 3, 9, 12), 3), ((1, 2, 8, 8), 4),
                                        ((1, 1, 16, 8), 8), ((2, 2, 5, 7), 1)])
def test_pool_vs_oracle_bit_exact(bf, O, shape, fdim):
    rs = np.random.RandomState(fdim)
    A = np.maximum(0, np.round(rs.randn(*shape) * 2) / 2).astype(np.float32).astype(np.float64)   # relu'd, many ties
    op = bf.atomic.MaxPoolOp()
    out, filt = op.forward(A, fdim)
    ro, rf = O.maxpool_forward(A, fdim)
    assert np.array_equal(out, ro) and np.array_equal(filt, rf)
    E = rs.randn(*ro.shape).astype(np.float32).astype(np.float64)
    assert np.array_equal(op.backward(E, filt), O.maxpool_backward(E, rf))
    # device-resident mask object behaves like the reference's filt array
    from brainforge_b200 import DeviceArray
    o2, mask = op.forward(DeviceArray.from_host(A), fdim)
    assert np.array_equal(np.asarray(mask), rf) and mask.shape == A.shape
    assert np.array_equal(op.backward(DeviceArray.from_host(E), mask).numpy(), O.maxpool_backward(E, rf))


def test_pool_rejects_non_divisible(bf):
    with pytest.raises(ValueError, match="Incompatible shapes"):
        bf.atomic.MaxPoolOp().forward(np.zeros((1, 1, 5, 4)), 2)


@pytest.mark.parametrize("T,B,D,H,act", [(1, 1, 1, 1, "tanh"), (7, 3, 5, 4, "tanh"), (4, 33, 20, 17, "relu"), (16, 8, 128, 256, "tanh")])
def test_lstm_vs_oracle(bf, O, T, B, D, H, act):
    rs = np.random.RandomState(T * 7 + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, 4 * H) * np.sqrt(2.0 / (D + 5 * H)), rs.randn(4 * H) * 0.1
    op = bf.atomic.LSTMOp(act)
    O_, Z_, C_ = op.forward(X, W, b)
    rO, rZ, rC = O.lstm_forward(X, W, b, act)
    close(O_, rO); close(Z_, rZ); close(C_, rC)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W, cache=rC)
    rX, rW, rb = O.lstm_backward(rZ, rO, E, W, rC, act)
    close(dX, rX); close(gW, rW); close(gb, rb)


def test_rnn_against_reference(bf):
    """RecurrentOp / RLayer against arrays produced by the reference (tests/golden/make_golden_rnn.py)."""
    from conftest import load_golden
    from brainforge_b200.layers import RLayer
    g = load_golden("rnn")
    for tag in "abc":
        X, W, b, E = (g["%s_%s" % (tag, k)] for k in "XWbE")
        for act in ("tanh", "sigmoid", "relu"):
            op = bf.atomic.RecurrentOp(act)
            o, z = op.forward(X, W, b)
            close(o, g["%s_%s_O" % (tag, act)]); close(z, g["%s_%s_Z" % (tag, act)])
            E0 = E.copy()
            dX, gW, gb = op.backward(Z=g["%s_%s_Z" % (tag, act)], O=g["%s_%s_O" % (tag, act)], E=E, W=W)
            assert np.array_equal(E, E0)                     # the caller's array is left alone
            assert dX.shape == g["%s_%s_dX" % (tag, act)].shape
            close(dX, g["%s_%s_dX" % (tag, act)]); close(gW, g["%s_%s_gW" % (tag, act)]); close(gb, g["%s_%s_gb" % (tag, act)])
    for seq in (0, 1):
        k = "layer_seq%d_" % seq
        stack = bf.LayerStack((6, 5), [RLayer(4, "tanh", return_seq=bool(seq))])
        layer = stack.layers[-1]
        layer.set_weights([g[k + "W"], g[k + "b"]], fold=False)
        close(layer.feedforward(g[k + "X"]), g[k + "Y"])
        close(layer.backpropagate(g[k + "D"]), g[k + "dX"])
        close(layer.nabla_w, g[k + "gW"]); close(layer.nabla_b, g[k + "gb"])


@pytest.mark.parametrize("T,B,D,H,act", [(1, 1, 1, 1, "tanh"), (9, 33, 20, 17, "relu"), (16, 200, 64, 128, "tanh")])
def test_rnn_vs_oracle(bf, O, T, B, D, H, act):
    rs = np.random.RandomState(T * 5 + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, H) * np.sqrt(2.0 / (D + 2 * H)), rs.randn(H) * 0.1
    op = bf.atomic.RecurrentOp(act)
    O_, Z_ = op.forward(X, W, b)
    rO, rZ = O.rnn_forward(X, W, b, act)
    close(O_, rO); close(Z_, rZ)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W)
    rX, rW, rb = O.rnn_backward(rZ, rO, E, W, act)
    close(dX, rX); close(gW, rW); close(gb, rb)


def test_rnn_network_trains_like_oracle(bf):
    """Stacked recurrent layers (return_seq=True feeding an LSTM, as in the reference's README examples) + Dense."""
    from helpers import build_b200, build_oracle, onehot
    spec = [("rnn", 12, "tanh", True), ("lstm", 9, "tanh", False), ("dense", 5, "softmax")]
    net = build_b200((7, 12), spec, "cxent", "adam", seed=5)
    ora = build_oracle((7, 12), spec, "cxent", "adam", seed=5)
    rs = np.random.RandomState(3)
    for _ in range(3):
        X, Y = rs.randn(10, 7, 12), onehot(rs, 10, 5)
        close(net.learn_batch(X, Y)["cost"], ora.learn_batch(X, Y)["cost"], 1e-4)
    close(net.layers.get_weights(), ora.get_weights(), 1e-4)


def test_gru_against_reference(bf):
    """GRU layer against arrays produced by the reference (tests/golden/make_golden_rnn.py)."""
    from conftest import load_golden
    from brainforge_b200.layers import GRU
    g = load_golden("rnn")
    for seq in (0, 1):
        for act in ("tanh", "relu"):
            k = "gru_%s_seq%d_" % (act, seq)
            stack = bf.LayerStack((6, 5), [GRU(4, act, return_seq=bool(seq))])
            layer = stack.layers[-1]
            layer.set_weights([g[k + "W"], g[k + "b"]], fold=False)
            close(layer.feedforward(g[k + "X"]), g[k + "Y"])
            close(layer.backpropagate(g[k + "D"]), g[k + "dX"])
            close(layer.nabla_w, g[k + "gW"]); close(layer.nabla_b, g[k + "gb"])


@pytest.mark.parametrize("T,B,D,H,act", [(1, 1, 1, 1, "tanh"), (9, 33, 20, 17, "relu"), (12, 150, 64, 96, "tanh")])
def test_gru_vs_oracle(bf, O, T, B, D, H, act):
    rs = np.random.RandomState(T * 3 + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, 3 * H) * np.sqrt(2.0 / (D + 4 * H)), rs.randn(3 * H) * 0.1
    op = bf.atomic.GRUOp(act)
    hs, Z, Zk, gates = op.forward(X, W, b)
    rh, rZ, rZk, rg = O.gru_forward(X, W, b, act)
    close(hs, rh); close(Z, rZ); close(Zk, rZk); close(gates, rg)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, Zk=rZk, gates=rg, E=E, W=W)
    rX, rW, rb = O.gru_backward(rZ, rZk, rg, E, W, act)
    close(dX, rX); close(gW, rW); close(gb, rb)


def test_gru_network_trains_like_oracle(bf):
    from helpers import build_b200, build_oracle, onehot
    spec = [("gru", 12, "tanh", True), ("gru", 8, "relu", False), ("dense", 5, "softmax")]
    net = build_b200((7, 10), spec, "cxent", "adam", seed=6)
    ora = build_oracle((7, 10), spec, "cxent", "adam", seed=6)
    rs = np.random.RandomState(4)
    for _ in range(3):
        X, Y = rs.randn(10, 7, 10), onehot(rs, 10, 5)
        close(net.learn_batch(X, Y)["cost"], ora.learn_batch(X, Y)["cost"], 1e-4)
    close(net.layers.get_weights(), ora.get_weights(), 1e-4)


def test_clockwork_against_reference(bf):
    """ClockworkLayer against arrays produced by the reference, initialiser included (same seed, same weights)."""
    from conftest import load_golden
    from brainforge_b200.layers import ClockworkLayer
    g = load_golden("rnn")
    for tag, (n, blocks, ticks) in {"def": (10, None, None), "cus": (7, [3, 2, 2], [1, 3, 2])}.items():
        for seq in (0, 1):
            k = "cw_%s_seq%d_" % (tag, seq)
            np.random.seed(53)
            stack = bf.LayerStack((9, 5), [ClockworkLayer(n, "tanh", blocksizes=blocks, ticktimes=ticks, return_seq=bool(seq))])
            layer = stack.layers[-1]
            close(layer.weights, g[k + "W"], 1e-7)
            assert np.array_equal(layer.tick_array.numpy(), g[k + "tick"])
            layer.set_weights([g[k + "W"], g[k + "b"]], fold=False)
            close(layer.feedforward(g[k + "X"]), g[k + "Y"])
            close(layer.backpropagate(g[k + "D"]), g[k + "dX"])
            close(layer.nabla_w, g[k + "gW"]); close(layer.nabla_b, g[k + "gb"])


@pytest.mark.parametrize("T,B,D,H,act", [(1, 1, 1, 5, "tanh"), (9, 33, 20, 17, "relu"), (12, 150, 64, 80, "tanh")])
def test_clockwork_vs_oracle(bf, O, T, B, D, H, act):
    rs = np.random.RandomState(T * 11 + H)
    np.random.seed(T)
    W, tick = O.clockwork_init(D, H)
    X, b = rs.randn(T, B, D), rs.randn(H) * 0.1
    op = bf.atomic.ClockworkOp(act)
    O_, Z_ = op.forward(X, W, b, tick)
    rO, rZ = O.clockwork_forward(X, W, b, tick, act)
    close(O_, rO); close(Z_, rZ)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W, tick=tick)
    rX, rW, rb = O.clockwork_backward(rZ, rO, E, W, tick, act)
    close(dX, rX); close(gW, rW); close(gb, rb)


def test_clockwork_network_trains_like_oracle(bf):
    from helpers import build_b200, build_oracle, onehot
    spec = [("clockwork", 10, "tanh", True), ("gru", 8, "tanh", False), ("dense", 5, "softmax")]
    net = build_b200((7, 6), spec, "cxent", "adam", seed=8)
    ora = build_oracle((7, 6), spec, "cxent", "adam", seed=8)
    rs = np.random.RandomState(4)
    for _ in range(3):
        X, Y = rs.randn(10, 7, 6), onehot(rs, 10, 5)
        close(net.learn_batch(X, Y)["cost"], ora.learn_batch(X, Y)["cost"], 1e-4)
    close(net.layers.get_weights(), ora.get_weights(), 1e-4)


def test_empty_batch(bf):
    op = bf.atomic.DenseOp()
    assert op.forward(np.zeros((0, 5)), np.ones((5, 3)), np.zeros(3)).shape == (0, 3)
    dW, db, dX = op.backward(np.zeros((0, 5)), np.zeros((0, 3)), np.ones((5, 3)))
    assert not dW.any() and not db.any() and dX.shape == (0, 5)
    assert bf.atomic.ConvolutionOp().forward(np.zeros((0, 2, 6, 6)), np.ones((3, 2, 3, 3))).shape == (0, 3, 4, 4)


def test_softmax_wide_rows_and_large_logits(bf, O):
    rs = np.random.RandomState(0)
    Z = rs.randn(37, 200) * 50
    close(bf.atomic.SoftMax().forward(Z), O.act_forward("softmax", Z))


def test_accuracy_first_max_wins(bf):
    o = np.array([[1., 1., 0.], [0., 2., 2.], [3., 1., 3.]])
    t = np.array([[1., 0., 0.], [0., 0., 1.], [0., 0., 1.]])
    assert bf.metrics.accuracy(o, t) == int(np.sum(o.argmax(1) == t.argmax(1))) == 1


def test_costs_bxent_hinge_golden(bf, O):
    """binary cross-entropy and hinge (metrics/costs.py:46-69) against values produced by the reference."""
    from conftest import load_golden
    g = load_golden("costs2")
    close(bf.metrics.bxent(g["O"], g["T"]), g["bxent"])
    close(bf.metrics.bxent.derivative(g["O"], g["T"]), g["bxent_delta"])
    close(bf.metrics.hinge(g["S"], g["Y"]), g["hinge"])
    d = bf.metrics.hinge.derivative(g["S"], g["Y"])
    assert np.array_equal(d, g["hinge_delta"])            # -y where output <= 1, exactly 0 elsewhere
    assert bf.metrics.costs.get("binary_crossentropy") is bf.metrics.bxent and bf.metrics.costs.get("hinge") is bf.metrics.hinge


@pytest.mark.parametrize("cost,act", [("bxent", "sigmoid"), ("hinge", "tanh")])
def test_training_step_with_bxent_hinge(bf, cost, act):
    """A learn_batch trajectory under the two costs equals the oracle's (the delta kernel is cost-specific for hinge)."""
    from helpers import build_b200, build_oracle
    spec = [("dense", 12, "tanh"), ("dense", 5, act)]
    net = build_b200((9,), spec, cost, "momentum", seed=3)
    ora = build_oracle((9,), spec, cost, "momentum", seed=3)
    rs = np.random.RandomState(8)
    X = rs.randn(17, 9)
    Y = (rs.uniform(size=(17, 5)) < 0.5).astype(np.float64) if cost == "bxent" else np.where(rs.uniform(size=(17, 5)) < 0.5, -1.0, 1.0)
    for _ in range(3):
        a, b = net.learn_batch(X, Y), ora.learn_batch(X, Y)
        close(a["cost"], b["cost"])
    close(net.layers.get_weights(), ora.get_weights(), 3e-5)


def test_dropout_against_reference(bf):
    """DropOut: the mask is drawn on the host from the global NumPy RNG exactly like the reference, so the same seed
    reproduces the reference's outputs bit for bit (up to FP32 storage)."""
    from conftest import load_golden
    from brainforge_b200.layers import DropOut
    g = load_golden("costs2")
    stack = bf.LayerStack((4, 3), [DropOut(0.3)])
    layer = stack.layers[-1]
    stack.learning = True
    np.random.seed(99)
    close(layer.feedforward(g["drop_X"]), g["drop_train"], 1e-7)
    close(layer.backpropagate(g["drop_D"]), g["drop_back"], 1e-7)
    assert np.array_equal(np.asarray(layer.mask), g["drop_mask_after"])
    stack.learning = False
    np.random.seed(99)
    close(layer.feedforward(g["drop_X"]), g["drop_eval"], 1e-7)


def test_global_average_pooling_known_answer(bf):
    """The reference's own known-answer test (tests/test_layers.py:9-30): planes filled with 1, 2, 3 -> means 1, 2, 3;
    the backward pass of [1, 2, 3] gives inputs / 25."""
    from brainforge_b200.layers import GlobalAveragePooling
    inputs = np.empty([3, 5, 5], dtype="float32")
    for ch in range(3):
        inputs[ch] = np.full([5, 5], fill_value=ch + 1, dtype="float32")
    stack = bf.LayerStack((3, 5, 5), [GlobalAveragePooling()])
    layer = stack.layers[-1]
    assert layer.outshape == (3,)
    out = layer.feedforward(inputs[None, ...])[0]
    np.testing.assert_equal(out, np.array([1., 2., 3.]))
    delta = layer.backpropagate(np.array([[1., 2., 3.]]))[0]
    np.testing.assert_allclose(delta, inputs / 25)


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_global_average_pooling_in_a_net(bf, O, prec):
    """conv -> relu -> GlobalAveragePooling -> dense against the oracle; on the TF32 path the layer is fed channels-last."""
    from helpers import build_b200, build_oracle
    bf.set_precision(prec)
    try:
        spec = [("conv", 32, 3, 3, "linear"), ("act", "relu"), ("gap",), ("dense", 4, "softmax")]
        net = build_b200((3, 12, 10), spec, "cxent", "sgd", seed=21)
        ora = build_oracle((3, 12, 10), spec, "cxent", "sgd", seed=21)
        rs = np.random.RandomState(2)
        X, Y = rs.randn(9, 3, 12, 10), np.eye(4)[rs.randint(0, 4, 9)]
        a, b = net.learn_batch(X, Y, update=False), ora.learn_batch(X, Y, update=False)
        tol = 1e-5 if prec == "fp32" else 5e-3
        close(a["cost"], b["cost"], tol)
        close(net.get_gradients(), ora.get_gradients(), tol if prec == "fp32" else 3e-2)
        if prec == "tf32":
            assert net.layers.layers[2].output.nhwc and net.layers.layers[3]._nhwc
    finally:
        bf.set_precision("fp32")


def test_large_cost_reduction_is_deterministic(bf, O):
    rs = np.random.RandomState(1)
    o = O.act_forward("softmax", rs.randn(50000, 10))
    t = np.eye(10)[rs.randint(0, 10, 50000)]
    a, b = bf.metrics.cxent(o, t), bf.metrics.cxent(o, t)
    assert a == b
    close(a, O.cost_value("cxent", o, t), 1e-6)
    close(bf.metrics.mse(o, t), O.cost_value("mse", o, t), 1e-6)


@pytest.mark.parametrize("n", [1 << 16, (4 << 20) + 3, (9 << 20) + 17])
def test_pageable_upload_through_the_host_thread_pool(n):
    """bf_upload_host: pageable float64 (the reference API's arrays) narrowed on the host by a thread pool into pinned staging
    chunks == the device-side cast bit for bit (round to nearest even), pageable float32 copied unchanged; chunk boundaries,
    back-to-back calls reusing the staging buffers, source overwritten right after the call."""
    import brainforge_b200 as bf
    rs = np.random.RandomState(n % 1000)
    a64 = rs.randn(n) * np.exp(rs.uniform(-20, 20, n))          # wide exponent range: the rounding is exercised
    want = a64.astype(np.float32)
    d = bf.DeviceArray.from_host(a64)
    a64[:] = 0.0                                                 # the call has read the source completely
    e = bf.DeviceArray.from_host(want.copy())                    # pageable float32
    assert np.array_equal(d.numpy(np.float32), want) and np.array_equal(e.numpy(np.float32), want)
    small = rs.randn(1000)                                       # below the pool threshold: device-side cast
    assert np.array_equal(bf.DeviceArray.from_host(small).numpy(np.float32), small.astype(np.float32))
