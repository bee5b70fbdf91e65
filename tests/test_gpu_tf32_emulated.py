"""The tensor-core path against the TF32-EMULATING oracle, at the benchmarked shapes.

`oracle.bf_oracle.tf32_emulation(policy)` narrows both operands of every contraction to TF32 the way the kernel
behind that site feeds the tensor core ("tma": rounded to nearest by the TMA unit of a TFLOAT32 tensor map; "rna":
cvt.rna.tf32 / `+0x1000` in the SIMT operand builders) and accumulates in float64, so what remains between
the oracle and the device is FP32 accumulation order.  That turns the 5e-3 / 3e-2 operand-rounding bounds of
test_gpu_tf32.py into EMU_TOL below -- tight enough that a 0.1 % kernel error on the benchmarked path fails.  The
float64 comparison stays as the second, looser assertion.

Shapes: cfg3 at its bench batch (4096), the cfg4 LSTM at T=64 / B=1024 / D=128 / H=256 (persistent cooperative
kernels, 8 batch tiles), the cfg5 fitness fan-out at P=1024 genomes x N=1024 samples (512 tiles > 3 waves).
"""
import json
import os

import numpy as np
import pytest

from conftest import relerr, ROOT
from helpers import CFG, build_b200, build_oracle, onehot

pytestmark = pytest.mark.gpu

EMU_TOL = 1e-4        # per op / per network gradient against the emulating oracle (measured: see profiles/r02_tf32_emulation.json)
EMU_TOL_TRAJ = 1e-3   # weights after 3 optimizer steps (Adam divides by sqrt(memory): amplifies gradient noise near 0)
F64_TOL = 5e-3        # per op against the plain float64 oracle (operand rounding)

_RECORD = {}


def record(name, err):
    _RECORD[name] = float(err)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "tf32_emulation_errors.json"), "w") as f:
            json.dump(_RECORD, f, indent=1, sort_keys=True)


def close(name, a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    e = relerr(a, b)
    record(name, e)
    assert e <= tol, (name, e, tol)
    return e


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


def test_operand_rounding_modes_on_hardware(bf, O):
    """Pins what the emulation assumes about the two ways operands reach `tcgen05.mma.kind::tf32`: probes whose FP32 bit
    patterns sit on, just below and just above a TF32 rounding tie go through a TMA-staged GEMM (Dense, k % 4 == 0) and a
    gather-staged one (k % 4 != 0) against a unit weight column; the products reveal the conversion bit for bit."""
    probes = np.array([1 + 2.0 ** -11, 1 + 3 * 2.0 ** -11, 1 + 2.0 ** -11 - 2.0 ** -23, 1 + 2.0 ** -11 + 2.0 ** -23,
                       -(1 + 2.0 ** -11), -(1 + 3 * 2.0 ** -11), 1 + 2.0 ** -12, 1 + 7 * 2.0 ** -13, 3.0, 0.1], dtype=np.float32)
    op = bf.atomic.DenseOp()
    found = {}
    for name, k in (("tma", 64), ("rna", 66)):
        assert bool(bf._lib.lib.bf_dense_tma_supported(len(probes), k, 64, 0)) == (name == "tma")
        X = np.zeros((len(probes), k), np.float32)
        X[:, 0] = probes
        W = np.zeros((k, 64), np.float32)
        W[0, :] = 1.0
        got = op.forward(X, W)[:, 3].astype(np.float32)
        match = [m for m in ("rne", "rna", "rz") if np.array_equal(got, O.tf32_round(probes, m).astype(np.float32))]
        found[name] = match
        record("rounding." + name + "." + "/".join(match or ["none"]), 0.0)
    assert "rna" in found["rna"] and "rne" not in found["rna"], found
    assert ("rne" if O.TMA_TIES == "even" else "rna") in found["tma"] and "rz" not in found["tma"], found


TMA, RNA = ("tma", "tma"), ("rna", "rna")


# ------------------------------------------------------------------------------------------------ single ops
@pytest.mark.parametrize("m,k,n", [(128, 64, 128), (20, 784, 60), (257, 130, 129), (4096, 512, 64)])
def test_dense_emulated(bf, O, m, k, n):
    """Dense with n > 32 columns: TMA-staged tcgen05 GEMMs where the shape allows, the gather-staged core otherwise."""
    rs = np.random.RandomState(m + k + n)
    X, W, b, E = rs.randn(m, k), rs.randn(k, n) * 0.2, rs.randn(n), rs.randn(m, n)
    op = bf.atomic.DenseOp()
    pol = bf.atomic.tf32_operand_modes("dense", m=m, k=k, n=n)
    with O.tf32_emulation(pol):
        r = O.dense_forward(X, W, b)
        rW, rb, rX = O.dense_backward(X, E, W)
    tag = "dense(%d,%d,%d)" % (m, k, n)
    close(tag + ".fwd", op.forward(X, W, b), r, EMU_TOL)
    dW, db, dX = op.backward(X, E, W)
    close(tag + ".dW", dW, rW, EMU_TOL); close(tag + ".dX", dX, rX, EMU_TOL); close(tag + ".db", db, rb, EMU_TOL)
    close(tag + ".fwd.f64", op.forward(X, W, b), O.dense_forward(X, W, b), F64_TOL)


@pytest.mark.parametrize("shape", [(300, 3, 30, 30, 32, 3, 3), (300, 32, 14, 14, 64, 3, 3), (300, 64, 6, 6, 128, 3, 3),
                                   (7, 1, 28, 28, 8, 3, 3), (5, 5, 9, 7, 3, 2, 4)])
def test_conv_emulated(bf, O, shape):
    """Each convolution family at the cfg3 layer shapes, > 1 wave of tiles (300 images), through atomic.ConvolutionOp."""
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape))
    A, F = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.3
    op = bf.atomic.ConvolutionOp()
    pol = bf.atomic.tf32_operand_modes("conv", ic=ic, iy=iy, ix=ix, nf=nf, fy=fy, fx=fx)
    tag = "conv" + str(shape)
    with O.tf32_emulation(pol):
        yv = O.conv_forward(A, F, "valid")
    close(tag + ".fwd", op.forward(A, F, "valid"), yv, EMU_TOL)
    E = rs.randn(*yv.shape)
    with O.tf32_emulation(pol):
        rF, rb, rX = O.conv_backward(A, E, F)
    dF, db, dX = op.backward(X=A, E=E, F=F)
    close(tag + ".dF", dF, rF, EMU_TOL); close(tag + ".db", db, rb, EMU_TOL); close(tag + ".dX", dX, rX, EMU_TOL)
    r64 = O.conv_backward(A, E, F)
    close(tag + ".dF.f64", dF, r64[0], F64_TOL); close(tag + ".dX.f64", dX, r64[2], F64_TOL)


def test_lstm_bench_shape_emulated(bf, O):
    """cfg4's LSTM at the benchmarked shape: T=64, B=1024, D=128, H=256 -- the persistent cooperative forward and
    backward kernels with 8 batch tiles of release/acquire flags (atomic/recurrent_op.py:51-112)."""
    from brainforge_b200.device import DeviceArray
    T, B, D, H = 64, 1024, 128, 256
    rs = np.random.RandomState(11)
    X = rs.randn(T, B, D)
    W = rs.randn(D + H, 4 * H) * np.sqrt(2.0 / (D + 5 * H))
    b = rs.randn(4 * H) * 0.1
    E = rs.randn(T, B, H) * (rs.rand(T, 1, 1) < 0.3)       # errors on some time steps only, like return_seq stacks
    op = bf.atomic.LSTMOp("tanh")
    Od, Zd, Cd = op.forward(DeviceArray.from_host(X), W, b)
    pol = bf.atomic.tf32_operand_modes("lstm", T=T, B=B, D=D, H=H)
    with O.tf32_emulation(pol):
        rO, rZ, rC = O.lstm_forward(X, W, b, "tanh")
    close("lstm64.O", Od.numpy(), rO, EMU_TOL); close("lstm64.Z", Zd.numpy(), rZ, EMU_TOL)
    close("lstm64.cache", Cd.numpy(), rC, EMU_TOL)
    dX, gW, gb = op.backward(Z=Zd, O=Od, E=DeviceArray.from_host(E), W=DeviceArray.from_host(W), cache=Cd)
    with O.tf32_emulation(pol):
        rX, rW, rb = O.lstm_backward(rZ, rO, E, W, rC, "tanh")
    # the device differentiates from the pre-activations (bw rows), the oracle from the FP32-exact outputs in
    # float64: same function, so the bound is accumulation order + the SFU-form activations of the persistent kernel
    close("lstm64.dX", dX.numpy(), rX, 5e-4); close("lstm64.gW", gW.numpy(), rW, 5e-4); close("lstm64.gb", gb.numpy(), rb, 5e-4)
    p64 = O.lstm_forward(X, W, b, "tanh")
    close("lstm64.O.f64", Od.numpy(), p64[0], F64_TOL)
    g64 = O.lstm_backward(p64[1], p64[0], E, W, p64[2], "tanh")
    close("lstm64.gW.f64", gW.numpy(), g64[1], F64_TOL); close("lstm64.dX.f64", dX.numpy(), g64[0], F64_TOL)


def test_fitness_bench_shape_emulated(bf, O):
    """cfg5 at P=1024 genomes x N=1024 samples (512 tiles of 2 genomes x 4 row passes: > 3 waves of the persistent
    fitness kernel), learner/neuroevolution.py:20-22,34-37 + evolution/_evolution.py:83-97."""
    from brainforge_b200.layers import Dense
    rs = np.random.RandomState(2)
    np.random.seed(2)
    P, N = 1024, 1024
    ne = bf.NeuroEvolution(input_shape=(784,), layerstack=[Dense(60, activation="sigmoid"), Dense(10, activation="softmax")],
                           cost="cxent", population_size=P)
    X, Y = rs.randn(N, 784), onehot(rs, N, 10)
    G = ne.population.individuals
    fit = ne.batch_fitness(G, X, Y).ravel()
    Gh = np.asarray(G, dtype=np.float32)
    with O.tf32_emulation({"fit.l1": ("tma", "rna")}):
        ref = O.fitness_mlp_logsoftmax(Gh, X.astype(np.float32), Y, 784, 60, 10)
    # the hidden layer's sigmoid uses the SFU forms (__expf, __fdividef) and the +-10 weights give logits of several
    # hundred: FP32 evaluation of the head moves a fitness by ~1e-5 relative
    close("fitness1024", fit, ref, EMU_TOL)
    assert np.max(np.abs(fit - ref) / np.abs(ref)) <= 10 * EMU_TOL       # every genome, not only the norm
    close("fitness1024.f64", fit, O.fitness_mlp_logsoftmax(G, X, Y, 784, 60, 10), 1e-2)


# ------------------------------------------------------------------------------------------------ whole networks
def _nets(cfg, **kw):
    ishape, spec, cost, optimizer = CFG[cfg]
    return (build_b200(ishape, spec, cost, optimizer, seed=1234, **kw), build_oracle(ishape, spec, cost, optimizer, seed=1234),
            build_oracle(ishape, spec, cost, optimizer, seed=1234))


def impose_device_gates(net, emu, tag, tie=1e-3):
    """After a forward pass of both: hand the device's discrete decisions (pooling masks, ReLU gates) to the oracle's
    backward pass, after checking that wherever they differ from the oracle's own the oracle's forward values are a
    NEAR-TIE (a pooling window whose two largest entries, or a pre-activation whose distance from zero, is within `tie` of
    the layer's scale: the TF32 forward pass legitimately differs from the float64 one by ~1e-5 per layer, the FP32
    one by ~1e-7).  One flipped decision moves the first layer's gradient by ~3e-4 (one of ~4e5 terms per filter).  Returns the
    number of differing decisions."""
    from brainforge_b200.layers import PoolLayer, Activation
    emu.forced = {}
    flips = 0
    layers = net.layers.layers[1:]
    for li, (layer, ol) in enumerate(zip(layers, emu.L)):
        if isinstance(layer, PoolLayer):
            dev = np.asarray(layer.filter)
            src = emu.L[li - 1]["output"]                      # what the pooling layer saw
            fd = ol["fdim"]
            im, c, iy, ix = src.shape
            win = src.reshape(im, c, iy // fd, fd, ix // fd, fd).transpose(0, 1, 2, 4, 3, 5).reshape(im, c, iy // fd, ix // fd, fd * fd)
            mdev = dev.reshape(im, c, iy // fd, fd, ix // fd, fd).transpose(0, 1, 2, 4, 3, 5).reshape(win.shape)
            mora = ol["filter"].reshape(im, c, iy // fd, fd, ix // fd, fd).transpose(0, 1, 2, 4, 3, 5).reshape(win.shape)
            differ = (mdev != mora).any(axis=-1)
            if differ.any():
                w = win[differ]
                top = np.sort(w, axis=-1)
                gap = top[:, -1] - (w * mdev[differ]).max(axis=-1)    # how far below the oracle's max the device's pick sits
                scale = np.abs(src).mean() + 1e-30
                assert np.all(gap <= tie * scale), (tag, li, "pooling decision differs on a window that is not a near-tie", gap.max() / scale)
            assert np.all(mdev.sum(axis=-1) >= 1)
            flips += int(differ.sum())
            emu.forced[li] = dev
        elif isinstance(layer, Activation) and str(layer.activation) == "relu":
            out = np.asarray(layer.output)
            gate = (out > 0).astype(np.float64)
            ref = ol["output"]
            differ = gate != (ref > 0)
            if differ.any():
                scale = np.abs(ref).mean() + 1e-30
                assert np.all(np.abs(ref[differ]) <= tie * scale) and np.all(np.abs(out[differ]) <= tie * scale), (tag, li, "ReLU gate differs away from zero")
            flips += int(differ.sum())
            emu.forced[li] = gate
    record(tag + ".gate_decisions_differing", flips)
    return flips


@pytest.mark.parametrize("cfg,batch", [("cfg1", 20), ("cfg2", 32), ("cfg3", 64), ("cfg4", 6)])
def test_configs_emulated(bf, O, cfg, batch):
    """cost, gradients and a 3-step trajectory of every training config against the emulating oracle."""
    net, emu, f64 = _nets(cfg)
    ishape = CFG[cfg][0]
    rs = np.random.RandomState(99)
    X, Y = rs.randn(batch, *ishape).astype(np.float32), onehot(rs, batch, 10)
    pol = bf.atomic.tf32_policy_for(net, learn=True)
    a = net.learn_batch(X, Y, update=False)
    with O.tf32_emulation(pol):
        emu.feedforward(X)
        impose_device_gates(net, emu, cfg)
        b = emu.learn_batch(X, Y, update=False)
    c = f64.learn_batch(X, Y, update=False)
    close(cfg + ".cost", a["cost"], b["cost"], EMU_TOL)
    close(cfg + ".grad", net.get_gradients(), emu.get_gradients(), 3 * EMU_TOL)
    close(cfg + ".grad.f64", net.get_gradients(), f64.get_gradients(), 3e-2)
    net.zero_gradients(); emu.zero_gradients()
    emu.forced = {}
    for _ in range(3):
        a = net.learn_batch(X, Y)
        with O.tf32_emulation(pol):
            b = emu.learn_batch(X, Y)
        close(cfg + ".traj.cost", a["cost"], b["cost"], EMU_TOL_TRAJ)
    # free-running (no imposed decisions): a flipped gate in step 1 moves a few weights of the gated conv nets
    close(cfg + ".traj.weights", net.layers.get_weights(), emu.get_weights(), EMU_TOL_TRAJ if cfg in ("cfg1", "cfg4") else 2e-2)


def test_cfg3_input_delta_emulated(bf, O):
    """backpropagate() down to the network input (three ReLU / max-pool gates, the first layer's dX and the
    materialising form of its filter gradient) against the emulating oracle: 1e-1 against float64 becomes 3 * EMU_TOL."""
    net, emu, _ = _nets("cfg3")
    rs = np.random.RandomState(5)
    X, Y = rs.randn(37, 3, 30, 30).astype(np.float32), onehot(rs, 37, 10)
    pol = bf.atomic.tf32_policy_for(net, learn=False)
    preds = net.predict(X)
    with O.tf32_emulation(pol):
        po = emu.feedforward(X)
        impose_device_gates(net, emu, "cfg3.backprop")
    d = net.backpropagate(net.cost.derivative(preds, Y))
    with O.tf32_emulation(pol):
        dref = emu.backpropagate(O.cost_derivative(po, Y))
    close("cfg3.input_delta", d, dref, 3 * EMU_TOL)
    close("cfg3.backprop.grad", net.get_gradients(), emu.get_gradients(), 3 * EMU_TOL)


def test_cfg3_bench_batch_emulated(bf, O):
    """cfg3 at the benchmarked batch of 4096 (TF32 path, and the FP32 path against float64): cost, accuracy count,
    every gradient segment."""
    net, emu, f64 = _nets("cfg3")
    rs = np.random.RandomState(7)
    X = rs.randn(4096, 3, 30, 30).astype(np.float32)
    Y = onehot(rs, 4096, 10).astype(np.float32)
    pol = bf.atomic.tf32_policy_for(net, learn=True)
    a = net.learn_batch(X, Y, metrics=["acc"], update=False)
    g = net.get_gradients()
    with O.tf32_emulation(pol):
        emu.feedforward(X)
        flips = impose_device_gates(net, emu, "cfg3@4096")
        b = emu.learn_batch(X, Y, metrics=["acc"], update=False)
    ge = emu.get_gradients()
    emu.forced = {}
    # pooling windows + ReLU elements of the net: 4096 * (6272 + 2304 + 512 + 25088 + 9216 + 2048); a few dozen near-ties flip
    assert flips <= 2e-5 * 4096 * 45440, flips
    close("cfg3@4096.cost", a["cost"], b["cost"], EMU_TOL)
    assert abs(a["accuracy"] - b["accuracy"]) * 4096 <= 2, (a, b)        # argmax count: at most a near-tie or two
    s = 0
    for i, l in enumerate(emu.trainable):       # per segment: a small layer must not hide behind a large one
        n = l["W"].size + l["b"].size
        close("cfg3@4096.grad.layer%d" % i, g[s:s + n], ge[s:s + n], 3 * EMU_TOL)
        s += n
    close("cfg3@4096.grad", g, ge, 3 * EMU_TOL)
    c = f64.learn_batch(X, Y, update=False)
    close("cfg3@4096.grad.f64", g, f64.get_gradients(), 3e-2)
    # the FP32 path against the plain float64 oracle at the same size: <= 1e-5 (north_star), decisions imposed likewise
    # (near-ties within 1e-5 of the layer scale only)
    bf.set_precision("fp32")
    try:
        net.zero_gradients(); f64.zero_gradients()
        a32 = net.learn_batch(X, Y, update=False)
        flips32 = impose_device_gates(net, f64, "cfg3@4096.fp32", tie=1e-5)
        assert flips32 <= 64, flips32
        c = f64.learn_batch(X, Y, update=False)
        close("cfg3@4096.fp32.cost", a32["cost"], c["cost"], 1e-5)
        close("cfg3@4096.fp32.grad", net.get_gradients(), f64.get_gradients(), 1e-5)
    finally:
        bf.set_precision("tf32")
