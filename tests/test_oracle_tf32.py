"""The oracle's TF32 operand emulation (test infrastructure for the tensor-core path): bit-level behaviour of the two
rounding modes, policy plumbing, and that an inactive policy changes nothing."""
import numpy as np
import pytest

from oracle import bf_oracle as O


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def test_tf32_round_modes_bit_level():
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.randn(1000) * 10.0 ** rs.randint(-20, 20, 1000), [0.0, -0.0, 1.0, -1.0, 1.0 + 2 ** -11, 1.0 + 3 * 2 ** -12]])
    rz, rna = O.tf32_round(x, "rz"), O.tf32_round(x, "rna")
    assert np.all(bits(rz) & 0x1FFF == 0) and np.all(bits(rna) & 0x1FFF == 0)         # 10 mantissa bits survive
    assert np.all(np.abs(rz) <= np.abs(np.float32(x))) and np.all(np.sign(rz) == np.sign(np.float32(x)))   # toward zero
    ulp = np.abs(np.float32(x)) * 2.0 ** -10
    assert np.all((np.abs(rz - np.float32(x)) < ulp) | (np.float32(x) == 0)) and np.all(np.abs(rna - np.float32(x)) <= ulp / 2 * (1 + 1e-6))
    assert O.tf32_round(1.0 + 2 ** -11, "rna") == 1.0 + 2 ** -10                         # ties away from zero (cvt.rna)
    assert O.tf32_round(-(1.0 + 2 ** -11), "rna") == -(1.0 + 2 ** -10)
    assert O.tf32_round(1.0 + 2 ** -11, "rz") == 1.0
    rne = O.tf32_round(x, "rne")
    assert np.all(bits(rne) & 0x1FFF == 0) and np.all(np.abs(rne - np.float32(x)) <= ulp / 2 * (1 + 1e-6))
    assert O.tf32_round(1.0 + 2 ** -11, "rne") == 1.0 and O.tf32_round(1.0 + 3 * 2 ** -11, "rne") == 1.0 + 2 ** -9   # ties to even
    assert O.tf32_round(1.0 + 2 ** -11 + 2 ** -23, "rne") == 1.0 + 2 ** -10
    assert np.array_equal(O.tf32_round(x, "tma"), rne if O.TMA_TIES == "even" else rna)
    assert np.array_equal(O.tf32_round(x, None), np.float32(x).astype(np.float64))       # None: FP32 storage only
    with pytest.raises(ValueError):
        O.tf32_round(x, "stochastic")


def test_policy_applies_per_site_and_per_layer():
    rs = np.random.RandomState(1)
    X, W, b, E = rs.randn(9, 16), rs.randn(16, 40), rs.randn(40), rs.randn(9, 40)
    ref = O.dense_forward(X, W, b)
    with O.tf32_emulation({"dense.fwd": ("rz", "rna")}):
        emu = O.dense_forward(X, W, b)
        dW, db, dX = O.dense_backward(X, E, W)            # sites without an entry: FP32 storage of the operands only
    assert np.array_equal(emu, O.tf32_round(X, "rz") @ O.tf32_round(W, "rna") + b)
    assert 1e-5 < np.linalg.norm(emu - ref) / np.linalg.norm(ref) < 2e-3
    assert np.array_equal(dW, np.float32(X).astype(np.float64).T @ np.float32(E).astype(np.float64))
    assert np.array_equal(O.dense_forward(X, W, b), ref)                                  # nothing leaks out of the context
    # per-layer callable: only the second dense layer of a stack is emulated
    np.random.seed(0)
    st = O.Stack((16,), [("dense", 40, "tanh"), ("dense", 40, "tanh")], cost="mse")
    seen = []

    def policy(site, layer):
        seen.append((site, layer))
        return ("rna", "rna") if layer == 1 else None

    with O.tf32_emulation(policy, fp32_storage=False):
        out = st.feedforward(X)
    h = np.tanh(X @ st.L[0]["W"] + st.L[0]["b"])
    assert np.array_equal(out, np.tanh(O.tf32_round(h, "rna") @ O.tf32_round(st.L[1]["W"], "rna") + st.L[1]["b"]))
    assert ("dense.fwd", 0) in seen and ("dense.fwd", 1) in seen


def test_conv_lstm_fitness_sites():
    rs = np.random.RandomState(2)
    A, F = rs.randn(2, 3, 6, 6), rs.randn(4, 3, 3, 3)
    with O.tf32_emulation({"conv.fwd": ("rz", "rz"), "conv.dF": ("rna", "rz"), "conv.dX": ("rz", "rz")}):
        y = O.conv_forward(A, F)
        dF, db, dX = O.conv_backward(A, y, F)
    assert np.array_equal(y, O.conv_valid(O.tf32_round(A, "rz"), O.tf32_round(F, "rz")))
    ref = O.conv_backward(A, y, F)
    for got, want in ((dF, ref[0]), (dX, ref[2])):
        e = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert 1e-6 < e < 3e-3
    # db comes from the all-ones operand of the dF MMA: it sees E as that site rounds it
    assert np.array_equal(db, O.tf32_round(y, "rz").sum(axis=(0, 2, 3), keepdims=True))
    with O.tf32_emulation({"conv.fwd": ("rz", "rz")}):
        assert np.array_equal(O.conv_backward(A, y, F)[1], ref[1])            # no tensor-core site: the plain sum
    X, W, b = rs.randn(4, 3, 5), rs.randn(13, 32) * 0.3, rs.randn(32) * 0.1
    with O.tf32_emulation({"lstm.fwd": ("rz", "rz")}):
        o1 = O.lstm_forward(X, W, b)[0]
    o0 = O.lstm_forward(X, W, b)[0]
    assert 1e-6 < np.linalg.norm(o1 - o0) / np.linalg.norm(o0) < 3e-3
    G = rs.uniform(size=(5, 4 * 5 + 5 + 5 * 3 + 3)).astype(np.float32)
    Xf, Yf = rs.randn(7, 4), np.eye(3)[rs.randint(0, 3, 7)]
    f0 = O.fitness_mlp_logsoftmax(G, Xf, Yf, 4, 5, 3)
    f0c = O.fitness_mlp_logsoftmax(G, Xf, Yf, 4, 5, 3, chunk=2)
    assert np.allclose(f0, f0c, rtol=1e-13)
    one = np.array([O.fitness(O.Stack((4,), [("dense", 5, "sigmoid"), ("dense", 3, "softmax")], cost="cxent"), g, Xf, Yf) for g in G])
    assert np.allclose(f0, one, rtol=1e-10)
    with O.tf32_emulation({"fit.l1": ("tma", "rna")}):
        f1 = O.fitness_mlp_logsoftmax(G, Xf, Yf, 4, 5, 3)
    assert 1e-7 < np.linalg.norm(f1 - f0) / np.linalg.norm(f0) < 1e-2
