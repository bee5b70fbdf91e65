"""The channels-last (NHWC) block of the tensor-core path: layout changes, channels-last max pooling (bit-exact,
with and without the fused ReLU), the TMA + tcgen05 convolution kernels reached through the unchanged
`atomic.ConvolutionOp` / layer API, and whole networks whose conv blocks run channels-last under the hood while
every array a caller can see keeps the reference's NCHW index order.

Bounds: layout changes and pooling bit-exact (integer / compare work); convolutions <= 5e-3 (TF32 operands);
whole-network gradients of gated conv nets <= 3e-2."""
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


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (3, 4, 5, 7), (2, 32, 14, 14), (5, 128, 2, 2), (7, 3, 30, 30),
                                   (2, 64, 33, 17), (0, 32, 4, 4), (70000, 4, 1, 2)])
def test_layout_round_trip(bf, shape):
    rs = np.random.RandomState(sum(shape))
    A = rs.randn(*shape).astype(np.float32)
    d = bf.DeviceArray.from_host(A)
    n = d.to_nhwc()
    assert n.nhwc and n.shape == A.shape
    phys = n.t.cpu().numpy()
    assert phys.shape == (shape[0], shape[2], shape[3], shape[1])
    assert np.array_equal(phys, A.transpose(0, 2, 3, 1))
    assert np.array_equal(n.numpy(np.float32), A)                   # .numpy() speaks NCHW
    assert np.array_equal(n.to_nchw().numpy(np.float32), A)
    flat = shape[1] * shape[2] * shape[3]
    assert np.array_equal(n.reshape(shape[0], flat).numpy(np.float32), A.reshape(shape[0], flat))
    with pytest.raises(RuntimeError):
        n.ptr                                                       # layout-unaware access is refused


@pytest.mark.parametrize("shape,fdim", [((2, 4, 4, 4), 2), ((3, 32, 28, 28), 2), ((5, 128, 4, 4), 2), ((2, 8, 6, 9), 3),
                                        ((2, 4, 8, 8), 1), ((1, 12, 16, 8), 8), ((0, 4, 4, 4), 2)])
@pytest.mark.parametrize("relu", [False, True])
def test_pool_nhwc_bit_exact(bf, O, shape, fdim, relu):
    """MaxPoolOp on channels-last storage == the reference on NCHW, bit for bit, ties (multi-hot) included;
    relu_in / gate reproduce Activation("relu") -> PoolLayer and its backward exactly."""
    rs = np.random.RandomState(sum(shape) + fdim)
    A = rs.randn(*shape).astype(np.float32)
    A[rs.rand(*shape) < 0.3] = 0.0                                  # plenty of exact ties
    if shape[0]:
        A[0, :, :fdim, :fdim] = -1.0                                # an all-negative window: relu makes it a 4-way tie at 0
    op = bf.atomic.MaxPoolOp()
    a = bf.DeviceArray.from_host(A).to_nhwc()
    out, mask = op.forward(a, fdim, relu_in=relu)
    assert out.nhwc and mask.nhwc
    Ar = np.maximum(A, 0) if relu else A
    ro, rf = O.maxpool_forward(Ar.astype(np.float64), fdim)
    assert np.array_equal(out.numpy(np.float32), ro.astype(np.float32))
    assert np.array_equal(np.asarray(mask), rf)
    E = rs.randn(*ro.shape).astype(np.float32)
    dX = op.backward(bf.DeviceArray.from_host(E), mask, gate=out if relu else None)
    ref = O.maxpool_backward(E.astype(np.float64), rf.copy())
    if relu:
        ref = ref * (Ar > 0)                                        # layers/core.py:52 with relu'(0) = 0
    assert dX.nhwc
    assert np.array_equal(dX.numpy(np.float32), ref.astype(np.float32))


@pytest.mark.parametrize("shape", [
    (2, 32, 14, 14, 64, 3, 3), (9, 64, 6, 6, 128, 3, 3), (3, 32, 8, 8, 32, 3, 3), (5, 128, 7, 9, 64, 2, 3),
    (2, 3, 30, 30, 32, 3, 3), (5, 1, 28, 28, 32, 3, 3), (3, 4, 17, 40, 64, 2, 4), (130, 32, 14, 14, 64, 3, 3),
])
def test_conv_routes_to_tma_kernels(bf, O, shape):
    """atomic.ConvolutionOp (unchanged signature) picks the TMA-fed channels-last kernels when it can; the
    caller sees NCHW results."""
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape))
    A, F, bias = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.3, rs.randn(nf)
    op = bf.atomic.ConvolutionOp()
    yv = O.conv_forward(A, F, "valid")
    close(op.forward(A, F, "valid"), yv)
    assert bf._lib.lib.bf_last_path() == 2
    close(op.forward(A, F, "valid", bias=bias, activation="relu"), np.maximum(yv, 0) + bias[None, :, None, None])
    E = rs.randn(*yv.shape)
    dF, db, dX = op.backward(X=A, E=E, F=F)
    rF, rb, rX = O.conv_backward(A, E, F)
    close(dF, rF); close(db.ravel(), rb.ravel()); close(dX, rX)
    # device-resident: channels-last in, channels-last out
    if ic % 32 == 0:
        a = bf.DeviceArray.from_host(A).to_nhwc()
        y = op.forward(a, bf.DeviceArray.from_host(F), "valid")
        assert y.nhwc
        close(y, yv)
        dFd, dbd, dXd = op.backward(X=a, E=bf.DeviceArray.from_host(E), F=bf.DeviceArray.from_host(F))
        assert dXd.nhwc
        close(dFd, rF); close(dXd, rX)


@pytest.mark.parametrize("shape", [(3, 3, 30, 30, 32, 3, 3), (5, 1, 28, 28, 32, 3, 3), (2, 4, 18, 40, 64, 3, 3), (130, 3, 30, 30, 32, 3, 3),
                                   (2, 2, 13, 50, 128, 2, 3)])
@pytest.mark.parametrize("relu_in", [False, True])
def test_conv_pool_fused_epilogue_bit_exact(bf, O, shape, relu_in):
    """bf_conv_c3_forward_pool == MaxPoolOp.forward([relu](ConvolutionOp.forward(..)), 2) bit for bit (pooled values
    and multi-hot tie mask): same MMA sequence, pooling done on the staged tile instead of a second pass over HBM."""
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape))
    A, F, bias = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.3, rs.randn(nf)
    op, pool = bf.atomic.ConvolutionOp(), bf.atomic.MaxPoolOp()
    a, f, b = bf.DeviceArray.from_host(A), bf.DeviceArray.from_host(F), bf.DeviceArray.from_host(bias)
    fused = op.forward_pool(a, f, bias=b, relu_in=relu_in)
    assert fused is not None
    y = op.forward(a, f, "valid", bias=b)
    assert y.nhwc
    ref_out, ref_mask = pool.forward(y, 2, relu_in=relu_in)
    assert np.array_equal(fused[0].numpy(np.float32), ref_out.numpy(np.float32))
    assert np.array_equal(np.asarray(fused[1]), np.asarray(ref_mask))
    yo = O.conv_forward(A, F, "valid") + bias[None, :, None, None]
    ro, _ = O.maxpool_forward(np.maximum(yo, 0) if relu_in else yo, 2)
    close(fused[0], ro)


@pytest.mark.parametrize("shape", [(5, 32, 14, 14, 64, 3, 3), (301, 32, 14, 14, 64, 3, 3), (7, 64, 6, 6, 128, 3, 3), (1100, 64, 6, 6, 128, 3, 3),
                                   (3, 32, 10, 18, 32, 3, 3), (2, 64, 8, 8, 64, 3, 3), (4, 32, 5, 7, 128, 2, 2)])
@pytest.mark.parametrize("relu_in", [False, True])
def test_conv_pool_fused_epilogue_channels_last_bit_exact(bf, O, shape, relu_in):
    """bf_conv_nhwc_forward_pool_packed == MaxPoolOp.forward([relu](ConvolutionOp.forward(..)), 2) bit for bit on the
    channels-last implicit GEMM (the 32->64 and 64->128 layers of cfg3, more stages than CTAs, ragged last stage):
    same MMA sequence per position, pooling done on the staged accumulators instead of a second pass over HBM."""
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape))
    A, F, bias = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.1, rs.randn(nf)
    A[0, :, :4, :4] = 0.0            # all-zero windows after the ReLU: four-way ties, multi-hot mask
    op, pool = bf.atomic.ConvolutionOp(), bf.atomic.MaxPoolOp()
    a, f, b = bf.DeviceArray.from_host(A).to_nhwc(), bf.DeviceArray.from_host(F), bf.DeviceArray.from_host(bias * 0 if relu_in else bias)
    pack = bf.atomic.FilterPack()
    fused = op.forward_pool(a, f, bias=b, relu_in=relu_in, pack=pack)
    assert fused is not None and bf._lib.lib.bf_last_path() == 2
    y = op.forward(a, f, "valid", bias=b)
    assert y.nhwc
    ref_out, ref_mask = pool.forward(y, 2, relu_in=relu_in)
    assert np.array_equal(fused[0].numpy(np.float32), ref_out.numpy(np.float32))
    assert np.array_equal(np.asarray(fused[1]), np.asarray(ref_mask))
    yo = O.conv_forward(A, F, "valid") + (bias * 0 if relu_in else bias)[None, :, None, None]
    ro, _ = O.maxpool_forward(np.maximum(yo, 0) if relu_in else yo, 2)
    close(fused[0], ro)


@pytest.mark.parametrize("shape", [(3, 3, 30, 30, 32, 3, 3), (5, 1, 28, 28, 32, 3, 3), (2, 3, 18, 42, 64, 3, 3), (130, 3, 30, 30, 32, 3, 3),
                                   (2, 2, 13, 50, 128, 2, 3), (7, 3, 12, 12, 32, 3, 3)])
@pytest.mark.parametrize("relu", [False, True])
def test_filter_gradient_with_fused_unpool(bf, O, shape, relu):
    """bf_conv_c3_backward_filter_unpool(X, dP, mask, gate) == conv filter gradient of E = pool'(dP) [* relu'] without E
    ever existing in HBM: equal to the two-kernel route up to summation order, and to the oracle within the TF32 bound."""
    im, ic, iy, ix, nf, fy, fx = shape
    rs = np.random.RandomState(sum(shape) + relu)
    A, F, bias = rs.randn(im, ic, iy, ix), rs.randn(nf, ic, fy, fx) * 0.3, rs.randn(nf) * 0.1
    op, pool = bf.atomic.ConvolutionOp(), bf.atomic.MaxPoolOp()
    a, f, b = bf.DeviceArray.from_host(A), bf.DeviceArray.from_host(F), bf.DeviceArray.from_host(bias)
    pooled, mask = op.forward_pool(a, f, bias=b, relu_in=relu)
    dP = rs.randn(*pooled.shape)
    dPd = bf.DeviceArray.from_host(dP).to_nhwc()
    gate = pooled if relu else None
    dF, db = bf.DeviceArray.empty(*F.shape), bf.DeviceArray.empty(1, nf, 1, 1)
    done = op.backward_filter_unpool(a, dPd, mask, gate, F.shape, dF, db)
    E = pool.backward(dPd, mask, gate=gate)                       # the materialised route
    if nf == 128:            # a band of this plane does not fit the in-kernel un-pooling: the caller must fall back
        assert not done
        return
    assert done
    dF2, db2, _ = op.backward(X=a, E=E, F=f, need_dX=False)
    close(dF, dF2, 5e-5); close(db, db2, 5e-5)
    rF, rb, _ = O.conv_backward(A, np.asarray(E), F)              # oracle on the SAME (device-made) E: isolates the kernel
    close(dF, rF); close(db.numpy().ravel(), rb.ravel())


@pytest.mark.parametrize("m,c,y,x,n,act", [(37, 128, 2, 2, 10, "softmax"), (5, 32, 3, 5, 7, "linear"), (300, 64, 1, 1, 32, "tanh"),
                                           (1, 4, 2, 3, 1, "sigmoid")])
def test_dense_behind_channels_last_flatten(bf, O, m, c, y, x, n, act):
    """bf_dense_forward_cl / bf_dense_backward_cl: a Dense layer fed by the flattened view of a channels-last
    activation (features stored (y,x,c), weights indexed by the reference's (c,y,x) flatten) == the reference on the
    NCHW flatten; exact FP32 FMA kernels, so the FP32 bound applies even on the TF32 path."""
    rs = np.random.RandomState(m + c + n)
    A = rs.randn(m, c, y, x)
    W, b, E = rs.randn(c * y * x, n) * 0.2, rs.randn(n), rs.randn(m, n)
    a = bf.DeviceArray.from_host(A).to_nhwc()
    flat = bf.DeviceArray(a.t.view(m, c * y * x), cl_flat=(c, y, x))
    assert np.array_equal(flat.numpy(np.float32), A.reshape(m, -1).astype(np.float32))      # readers still see NCHW order
    op = bf.atomic.DenseOp()
    w, bd = bf.DeviceArray.from_host(W), bf.DeviceArray.from_host(b)
    out = op.forward(flat, w, bd, activation=act)
    ref = O.act_forward(act, O.dense_forward(A.reshape(m, -1), W, b))
    close(out, ref, 1e-5)
    dW, db, dX = op.backward(flat, bf.DeviceArray.from_host(E), w)
    rW, rb, rX = O.dense_backward(A.reshape(m, -1), E, W)
    close(dW, rW, 1e-5); close(db, rb, 1e-5)
    assert dX.cl_flat == (c, y, x)
    close(dX, rX, 1e-5)                                                                    # .numpy() converts back


def _cfg3(bf, batch, **kw):
    ishape, spec, cost, optimizer = CFG["cfg3"]
    return build_b200(ishape, spec, cost, optimizer, seed=1234, **kw), build_oracle(ishape, spec, cost, optimizer, seed=1234)


def test_cfg3_runs_channels_last_and_matches(bf, O):
    from brainforge_b200.layers import ConvLayer, PoolLayer, Activation
    net, ora = _cfg3(bf, 37)
    rs = np.random.RandomState(5)
    X, Y = rs.randn(37, 3, 30, 30), onehot(rs, 37, 10)
    a = net.learn_batch(X, Y, metrics=["acc"], update=False)
    b = ora.learn_batch(X, Y, update=False, metrics=["acc"])
    close(a["cost"], b["cost"])
    close(net.get_gradients(), ora.get_gradients(), 3e-2)
    # the conv blocks really ran channels-last, with the ReLU folded into the pooling kernels
    convs = [l for l in net.layers.layers if isinstance(l, ConvLayer)]
    pools = [l for l in net.layers.layers if isinstance(l, PoolLayer)]
    acts = [l for l in net.layers.layers[1:] if isinstance(l, Activation)]
    assert all(c.output.nhwc for c in convs) and all(p.output.nhwc and p._fused_relu for p in pools)
    assert net.layers.layers[-1].inputs.cl_flat == (128, 2, 2)      # Flatten moved no data; Dense reads channels-last
    assert not convs[0].inputs.nhwc and convs[1].inputs.nhwc
    # every array a caller can read keeps the reference's NCHW order and values (lazy ReLU outputs included)
    for layer, ol in zip(net.layers.layers[1:], ora.L):
        if layer.output is not None:                                 # Reshape/Flatten keep none (layers/core.py:81-82)
            close(np.asarray(layer.output), ol["output"], 1e-2)
    for act, conv in zip(acts, convs):
        assert np.array_equal(np.asarray(act.output), np.maximum(np.asarray(conv.output), 0))
    for p, act in zip(pools, acts):
        ro, rf = O.maxpool_forward(np.asarray(act.output), 2)
        assert np.array_equal(np.asarray(p.output), ro)              # bit-exact given the same inputs
        assert np.array_equal(np.asarray(p.filter), rf)
    # input delta through backpropagate() (first layer's dX: generic kernel on NCHW)
    preds = net.predict(X)
    d = net.backpropagate(net.cost.derivative(preds, Y))
    ora.zero_gradients()
    dref = ora.backpropagate(O.cost_derivative(ora.feedforward(X), Y))
    assert d.shape == X.shape
    # the input delta crosses all three ReLU/max-pool gates: a TF32-sized perturbation flips a gate in a few
    # samples (discontinuous), every other sample agrees to operand-rounding accuracy
    close(d, dref, 1e-1)
    # against TF32 arithmetic on the NCHW gather kernels (operands rounded by cvt.rna there, by the TMA unit here;
    # different summation order) gate flips are rare: the typical sample agrees to ~1e-4, two orders of magnitude
    # closer than to the float64 oracle
    bf.atomic.ConvolutionOp.channels_last = False
    try:
        net2, _ = _cfg3(bf, 37)
        d2 = net2.backpropagate(net2.cost.derivative(net2.predict(X), Y))
        assert not any(l.output.nhwc for l in net2.layers.layers[1:] if l.output is not None)
    finally:
        bf.atomic.ConvolutionOp.channels_last = True
    per_sample = np.array([relerr(d[i], d2[i]) for i in range(len(X))])
    assert np.median(per_sample) <= 5e-4, per_sample


def test_cfg3_trajectory_and_graph(bf):
    net, ora = _cfg3(bf, 64)
    netg, _ = _cfg3(bf, 64, use_graph=True)
    rs = np.random.RandomState(6)
    X, Y = rs.randn(64, 3, 30, 30), onehot(rs, 64, 10)
    for step in range(4):
        a, g, b = net.learn_batch(X, Y), netg.learn_batch(X, Y), ora.learn_batch(X, Y)
        close(a["cost"], b["cost"], 2e-2)
        assert a["cost"] == g["cost"]                                # graph replay == eager, bit for bit
    close(net.layers.get_weights(), ora.get_weights(), 2e-2)
    assert np.array_equal(net.layers.get_weights(), netg.layers.get_weights())


def test_cfg3_full_size_properties(bf):
    """Size-independent properties at BASELINE's full batch (4096): the gradient of the batch is the sum of the
    gradients of its halves, cost additivity, every pooling window has at least one mask hit."""
    from brainforge_b200.layers import PoolLayer
    net, _ = _cfg3(bf, 4096)
    rs = np.random.RandomState(7)
    X = rs.randn(4096, 3, 30, 30).astype(np.float32)
    Y = onehot(rs, 4096, 10).astype(np.float32)
    full = net.learn_batch(X, Y, update=False)
    g = net.get_gradients(); net.zero_gradients()
    pool = [l for l in net.layers.layers if isinstance(l, PoolLayer)][-1]
    words = pool.filter.words.cpu().numpy()
    assert words.min() >= 1 and words.max() <= 15
    h1 = net.learn_batch(X[:2048], Y[:2048], update=False)
    g1 = net.get_gradients(); net.zero_gradients()
    h2 = net.learn_batch(X[2048:], Y[2048:], update=False)
    g2 = net.get_gradients(); net.zero_gradients()
    close(g, g1 + g2, 1e-4)
    assert abs(full["cost"] * 4096 - (h1["cost"] + h2["cost"]) * 2048) <= 1e-4 * abs(full["cost"] * 4096)


def test_xp_conv_net_large_flatten_feeds_skinny_dense(bf, O):
    """The reference's xperiments/xp_conv.py network under TF32: Conv32 -> ReLU -> Conv64 -> Pool2 -> ReLU -> Flatten ->
    Dense10 on 1x28x28.  The flattened channels-last activation has k = 64*12*12 = 9216 features, more than the skinny
    Dense kernels hold in shared memory (k <= 2560 for n = 10): the channels-last shortcut must step aside for a layout
    change + the plain Dense path instead of failing."""
    assert not bf._lib.lib.bf_dense_cl_supported(9216, 10) and bf._lib.lib.bf_dense_cl_supported(512, 10)
    spec = [("conv", 32, 3, 3, "linear"), ("act", "relu"), ("conv", 64, 3, 3, "linear"), ("pool", 2), ("act", "relu"),
            ("flatten",), ("dense", 10, "softmax")]
    net = build_b200((1, 28, 28), spec, "cxent", "adam", seed=3)
    ora = build_oracle((1, 28, 28), spec, "cxent", "adam", seed=3)
    rs = np.random.RandomState(4)
    X, Y = rs.randn(9, 1, 28, 28), onehot(rs, 9, 10)
    a, b = net.learn_batch(X, Y, update=False), ora.learn_batch(X, Y, update=False)
    close(a["cost"], b["cost"])
    close(net.get_gradients(), ora.get_gradients(), 3e-2)
    assert net.layers.layers[-1].inputs.cl_flat is None
    for _ in range(2):
        a, b = net.learn_batch(X, Y), ora.learn_batch(X, Y)
        close(a["cost"], b["cost"], 2e-2)


def test_workspace_regrowth_recaptures_graphs(bf):
    """A CUDA graph has the scratch workspace's address baked in: when a later, larger call makes the workspace grow, the
    old block stays allocated and the graph is captured again before its next replay -- the trajectory keeps
    matching an eager twin bit for bit."""
    from brainforge_b200 import device
    net, _ = _cfg3(bf, 8, use_graph=True)
    twin, _ = _cfg3(bf, 8)
    rs = np.random.RandomState(9)
    X, Y = rs.randn(8, 3, 30, 30), onehot(rs, 8, 10)
    for _ in range(3):
        a, b = net.learn_batch(X, Y), twin.learn_batch(X, Y)
        assert a["cost"] == b["cost"]
    gen = device.workspace_generation()
    device._ensure_workspace(device._state["workspace"].numel() + (64 << 20))        # what a larger predict() would trigger
    assert device.workspace_generation() == gen + 1 and device._state["ws_retired"]
    device._state["ws_retired"][-1].fill_(255)                 # a replay through the stale address would now read garbage
    for _ in range(3):
        a, b = net.learn_batch(X, Y), twin.learn_batch(X, Y)
        assert a["cost"] == b["cost"]
    assert np.array_equal(net.layers.get_weights(), twin.layers.get_weights())


def test_fill_zero_clears_non_finite_values(bf):
    """zero_gradients / shuffle must clear NaN and Inf (0 * NaN = NaN: ADVICE r1)."""
    a = bf.DeviceArray.from_host(np.array([np.nan, np.inf, -np.inf, 1.0, 2.0]))
    a.fill_zero()
    assert np.array_equal(a.numpy(), np.zeros(5))
    a = bf.DeviceArray.from_host(np.array([np.nan, 3.0]))
    a *= 0
    assert np.array_equal(a.numpy(), np.zeros(2))
    assert np.array_equal(bf.DeviceArray.empty(7).fill(2.5).numpy(), np.full(7, 2.5))
