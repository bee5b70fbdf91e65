"""The tensor-core path (BF_PREC_TF32: tcgen05.mma kind::tf32, FP32 accumulate in TMEM) against the float64
oracle.  Stated bound for this path: norm-wise relative error <= 5e-3 per op output / gradient (operands are
rounded to TF32's 10-bit mantissa, ~5e-4 each; measured errors are ~3e-4), <= 3e-2 for whole-network
gradients of the ReLU/max-pool conv nets (gate flips) and <= 2e-2 on a 3-step Adam trajectory.  Pooling masks and argmax stay exact-by-construction only where the inputs are identical, so
they are checked on the FP32 path (test_gpu_ops / test_gpu_nets)."""
import numpy as np
import pytest

from conftest import relerr
from helpers import CFG, build_b200, build_oracle, onehot

pytestmark = pytest.mark.gpu
TOL = 5e-3


@pytest.fixture(scope="module")
def bf():
    import torch
    assert torch.cuda.is_available()
    import brainforge_b200 as bf
    bf.set_precision("tf32")
    yield bf
    bf.set_precision("fp32")


@pytest.fixture(scope="module")
def O():
    from oracle import bf_oracle
    return bf_oracle


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    e = relerr(a, b)
    assert e <= tol, e
    return e


def used_tensor_cores(bf):
    return bf._lib.lib.bf_last_path() in (1, 2)


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (128, 32, 16), (128, 64, 128), (20, 784, 60), (33, 17, 5), (257, 130, 129),
                                   (4096, 512, 10), (300, 1352, 200)])
def test_dense_tf32(bf, O, m, k, n):
    rs = np.random.RandomState(m + k + n)
    X, W, b, E = rs.randn(m, k), rs.randn(k, n) * 0.2, rs.randn(n), rs.randn(m, n)
    op = bf.atomic.DenseOp()
    e = close(op.forward(X, W, b), O.dense_forward(X, W, b))
    # n <= 32 output columns: HBM-bound skinny layer, exact-FP32 FMA kernel on both precision paths (dense.cu)
    assert used_tensor_cores(bf) == (n > 32)
    close(op.forward(X, W, b, activation="tanh"), np.tanh(O.dense_forward(X, W, b)))
    dW, db, dX = op.backward(X, E, W)
    rW, rb, rX = O.dense_backward(X, E, W)
    close(dW, rW); close(db, rb); close(dX, rX)
    # the tensor-core path must not be (accidentally) the FP32 path: operands are rounded to TF32
    if k >= 64 and n > 32:
        assert e > 1e-6, e


@pytest.mark.parametrize("shape", [
    (2, 1, 28, 28, 8, 3, 3), (3, 3, 30, 30, 32, 3, 3), (2, 32, 14, 14, 64, 3, 3), (5, 64, 6, 6, 128, 3, 3),
    (1, 1, 3, 3, 1, 3, 3), (2, 5, 9, 7, 3, 2, 4), (1, 2, 6, 6, 17, 5, 1), (40, 4, 8, 8, 6, 1, 1),
])
def test_conv_tf32(bf, O, shape):
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape))
    A, F, bias = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.3, rs.randn(nf)
    op = bf.atomic.ConvolutionOp()
    yv = O.conv_forward(A, F, "valid")
    close(op.forward(A, F, "valid"), yv)
    assert used_tensor_cores(bf)
    close(op.forward(A, F, "full"), O.conv_forward(A, F, "full"))
    close(op.forward(A, F, "valid", bias=bias, activation="tanh"), np.tanh(yv) + bias[None, :, None, None])
    E = rs.randn(*yv.shape)
    dF, db, dX = op.backward(X=A, E=E, F=F)
    rF, rb, rX = O.conv_backward(A, E, F)
    close(dF, rF); close(db, rb); close(dX, rX)


@pytest.mark.parametrize("T,B,D,H", [(3, 5, 4, 6), (8, 64, 128, 256), (20, 100, 60, 70), (7, 33, 3, 1)])
def test_lstm_tf32(bf, O, T, B, D, H):
    rs = np.random.RandomState(T + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, 4 * H) * np.sqrt(2.0 / (D + 5 * H)), rs.randn(4 * H) * 0.1
    op = bf.atomic.LSTMOp("tanh")
    O_, Z_, C_ = op.forward(X, W, b)
    rO, rZ, rC = O.lstm_forward(X, W, b, "tanh")
    close(O_, rO); close(Z_, rZ); close(C_, rC)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W, cache=rC)
    rX, rW, rb = O.lstm_backward(rZ, rO, E, W, rC, "tanh")
    close(dX, rX); close(gW, rW); close(gb, rb)


@pytest.mark.parametrize("env", [{"BF_LSTM_STEPWISE": "1"}, {"BF_LSTM_PERSIST_BWD": "1"}, {"BF_LSTM_PERSIST_BWD": "0"},
                                 {"BF_LSTM_UNFUSED": "1", "BF_LSTM_STEPWISE": "1"}])
@pytest.mark.parametrize("T,B,D,H", [(5, 70, 12, 40), (8, 200, 64, 96)])
def test_lstm_tf32_kernel_variants(bf, O, monkeypatch, env, T, B, D, H):
    """Every LSTM kernel family behind the same entry points: persistent forward / per-step fused forward / GEMM + gate
    kernel, persistent backward / per-step backward (the switches are read per call)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rs = np.random.RandomState(T + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, 4 * H) * np.sqrt(2.0 / (D + 5 * H)), rs.randn(4 * H) * 0.1
    op = bf.atomic.LSTMOp("tanh")
    O_, Z_, C_ = op.forward(X, W, b)                      # device arrays carry the coefficient rows (cache.bw)
    rO, rZ, rC = O.lstm_forward(X, W, b, "tanh")
    close(O_, rO); close(Z_, rZ); close(C_, rC)
    E = rs.randn(T, B, H)
    from brainforge_b200.device import DeviceArray
    Od, Zd, Cd = op.forward(DeviceArray.from_host(X), W, b)
    dX, gW, gb = op.backward(Z=Zd, O=Od, E=DeviceArray.from_host(E), W=DeviceArray.from_host(W), cache=Cd)
    rX, rW, rb = O.lstm_backward(rZ, rO, E, W, rC, "tanh")
    close(dX.numpy(), rX); close(gW.numpy(), rW); close(gb.numpy(), rb)
    # and from a host cache (coefficients formed on the host in float64)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W, cache=rC)
    close(dX, rX); close(gW, rW); close(gb, rb)


@pytest.mark.parametrize("T,B,D,H", [(6, 50, 12, 20), (10, 300, 64, 128)])
def test_rnn_tf32(bf, O, T, B, D, H):
    rs = np.random.RandomState(T + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, H) * np.sqrt(2.0 / (D + 2 * H)), rs.randn(H) * 0.1
    op = bf.atomic.RecurrentOp("tanh")
    O_, Z_ = op.forward(X, W, b)
    assert used_tensor_cores(bf)
    rO, rZ = O.rnn_forward(X, W, b, "tanh")
    close(O_, rO); close(Z_, rZ)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W)
    rX, rW, rb = O.rnn_backward(rZ, rO, E, W, "tanh")
    close(dX, rX); close(gW, rW); close(gb, rb)


@pytest.mark.parametrize("T,B,D,H", [(6, 50, 12, 20), (10, 300, 64, 128)])
def test_gru_tf32(bf, O, T, B, D, H):
    rs = np.random.RandomState(T + H)
    X, W, b = rs.randn(T, B, D), rs.randn(D + H, 3 * H) * np.sqrt(2.0 / (D + 4 * H)), rs.randn(3 * H) * 0.1
    op = bf.atomic.GRUOp("tanh")
    hs, Z, Zk, gates = op.forward(X, W, b)
    assert used_tensor_cores(bf)
    rh, rZ, rZk, rg = O.gru_forward(X, W, b, "tanh")
    close(hs, rh); close(Z, rZ); close(Zk, rZk); close(gates, rg)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, Zk=rZk, gates=rg, E=E, W=W)
    rX, rW, rb = O.gru_backward(rZ, rZk, rg, E, W, "tanh")
    close(dX, rX); close(gW, rW); close(gb, rb)


def test_clockwork_tf32(bf, O):
    T, B, D, H = 10, 300, 64, 120
    rs = np.random.RandomState(T + H)
    np.random.seed(2)
    W, tick = O.clockwork_init(D, H)
    X, b = rs.randn(T, B, D), rs.randn(H) * 0.1
    op = bf.atomic.ClockworkOp("tanh")
    O_, Z_ = op.forward(X, W, b, tick)
    assert used_tensor_cores(bf)
    rO, rZ = O.clockwork_forward(X, W, b, tick, "tanh")
    close(O_, rO); close(Z_, rZ)
    E = rs.randn(T, B, H)
    dX, gW, gb = op.backward(Z=rZ, O=rO, E=E, W=W, tick=tick)
    rX, rW, rb = O.clockwork_backward(rZ, rO, E, W, tick, "tanh")
    close(dX, rX); close(gW, rW); close(gb, rb)


@pytest.mark.parametrize("cfg,batch", [("cfg1", 20), ("cfg2", 32), ("cfg3", 16), ("cfg4", 4)])
def test_configs_tf32(bf, cfg, batch):
    ishape, spec, cost, optimizer = CFG[cfg]
    net = build_b200(ishape, spec, cost, optimizer, seed=1234)
    ora = build_oracle(ishape, spec, cost, optimizer, seed=1234)
    rs = np.random.RandomState(99)
    X, Y = rs.randn(batch, *ishape), onehot(rs, batch, 10)
    a = net.learn_batch(X, Y, update=False)
    b = ora.learn_batch(X, Y, update=False)
    close(a["cost"], b["cost"])
    # whole-network gradients: a TF32-sized perturbation of a pre-activation can flip a ReLU gate or a
    # max-pool argmax, which moves the gradient discontinuously -> looser bound for the gated conv nets
    close(net.get_gradients(), ora.get_gradients(), 3e-2 if cfg in ("cfg2", "cfg3") else TOL)
    net.zero_gradients(); ora.zero_gradients()
    for _ in range(3):
        a, b = net.learn_batch(X, Y), ora.learn_batch(X, Y)
        close(a["cost"], b["cost"], 2e-2)
    close(net.layers.get_weights(), ora.get_weights(), 2e-2)


def test_fitness_tf32(bf, O):
    from brainforge_b200.layers import Dense
    rs = np.random.RandomState(1)
    np.random.seed(1)
    ne = bf.NeuroEvolution(input_shape=(784,), layerstack=[Dense(60, activation="sigmoid"), Dense(10, activation="softmax")],
                           cost="cxent", population_size=24)
    X, Y = rs.randn(256, 784), onehot(rs, 256, 10)
    G = ne.population.individuals
    fit = ne.batch_fitness(G, X, Y).ravel()
    ref = O.fitness_mlp_logsoftmax(G, X, Y, 784, 60, 10)
    # +-10 weights make logits of several hundred: TF32 operand rounding moves the cost by ~1e-3 relative
    close(fit, ref, 1e-2)
