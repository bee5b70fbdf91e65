"""CPU oracle for the brainforge layer forward/backward + optimizer hot path.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy (float64) restatement of the
reference's arithmetic.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``brainforge_b200`` never does (it fails loudly when the CUDA
library is missing instead of falling back to anything in here).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference`` (authoring container only), runs seeded
cases through the reference's NumPy ``atomic`` path (and the Numba ``llatomic``
path where the two agree) and commits inputs+outputs under ``tests/golden``;
``tests/test_oracle_golden.py`` checks every function below against them.

Every function cites the reference file:line it restates (paths relative to
``/root/reference/brainforge``).  Loops over patches / windows / individuals in
the reference are vectorised here (same arithmetic, BLAS summation order), the
loop over LSTM time steps is kept because it is a true recurrence.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

F64 = np.float64


# --------------------------------------------------------------------------
# TF32 operand emulation (checker for the tensor-core path)
# --------------------------------------------------------------------------
# `tcgen05.mma.kind::tf32` reads 19 bits of each FP32 operand (sign, 8 exponent, 10 mantissa bits) and accumulates in
# FP32.  With a policy active every contraction below narrows its operands to FP32 (the device's storage type) and
# then to TF32 the way the kernel that runs this site does it -- "tma": staged by a tensor map of data type TFLOAT32
# (the TMA unit rounds to nearest while it copies; tie rule TMA_TIES, pinned on the GPU by
# tests/test_gpu_tf32_emulated.py::test_operand_rounding_modes_on_hardware), "rna": round to nearest, ties away
# (cvt.rna.tf32.f32 / `+0x1000` in the SIMT operand builders), "rz": the low 13 bits dropped (what the MMA does to raw
# FP32 bits that nobody rounded) -- and accumulates in float64.  Bias gradients that the kernels take from an all-ones
# operand of the same MMA (db = ones^T E) see the rounded E as well.  What is left between this and the device is FP32 accumulation order only, so the
# tensor-core path can be held to ~1e-5 instead of the ~5e-4 of operand rounding.  A policy is a callable
# (site, layer_index) -> (mode_A, mode_B) | None, or a dict {site: (mode_A, mode_B)}; sites are "dense.fwd",
# "dense.dW", "dense.dX", "conv.fwd", "conv.dF", "conv.dX", "lstm.fwd", "lstm.dZ", "lstm.dW", "fit.l1".
_TF32 = {"policy": None, "layer": None, "fp32_storage": False}
TMA_TIES = "even"      # tie rule of the TMA unit's FP32 -> TF32 conversion ("even" | "away")


def tf32_round(x, mode):
    """x narrowed to FP32, then to TF32 by `mode` ("tma" | "rna" | "rne" | "rz" | None = FP32 only); returned as float64."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    if mode is None:
        return a.astype(F64)
    u = a.view(np.uint32)
    if mode == "tma":
        mode = "rne" if TMA_TIES == "even" else "rna"
    if mode == "rna":
        u = u + np.uint32(0x1000)
    elif mode == "rne":
        u = u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))
    elif mode != "rz":
        raise ValueError(mode)
    return (u & np.uint32(0xFFFFE000)).view(np.float32).astype(F64)


class tf32_emulation:
    """Context manager: `with tf32_emulation(policy): ...` (see above)."""

    def __init__(self, policy, fp32_storage=True):
        self.policy, self.fp32_storage = policy, fp32_storage

    def __enter__(self):
        self.saved = dict(_TF32)
        _TF32.update(policy=self.policy, fp32_storage=self.fp32_storage)
        return self

    def __exit__(self, *exc):
        _TF32.update(self.saved)


def _site_modes(site):
    pol = _TF32["policy"]
    if pol is None:
        return None
    return pol(site, _TF32["layer"]) if callable(pol) else pol.get(site)


def _db_site_declared(pol):
    """True when the policy speaks about the "conv.db" site for the current layer (a per-layer table marks it with the
    attribute `declares`, a dict by containing the key)."""
    if callable(pol):
        decl = getattr(pol, "declares", None)
        return bool(decl and decl("conv.db", _TF32["layer"]))
    return "conv.db" in pol


def _operands(site, A, B):
    """The two operands of a contraction as the kernel behind `site` feeds them to the tensor core."""
    if _TF32["policy"] is None:
        return A, B
    modes = _site_modes(site)
    if modes is None:
        return (tf32_round(A, None), tf32_round(B, None)) if _TF32["fp32_storage"] else (A, B)
    return tf32_round(A, modes[0]), tf32_round(B, modes[1])


# --------------------------------------------------------------------------
# activations                                   atomic/activation_op.py:26-135
# --------------------------------------------------------------------------
def act_forward(kind, Z):
    """atomic/activation_op.py: Sigmoid:30-31, Tanh:54-55, ReLU:87-88,
    Linear:76-77, SoftMax.t1:127-130 (row softmax with max subtraction)."""
    if kind == "sigmoid":
        return 1.0 / (1.0 + np.exp(-Z))
    if kind == "tanh":
        return np.tanh(Z)
    if kind == "relu":
        return np.maximum(0.0, Z)
    if kind == "linear":
        return Z
    if kind == "softmax":
        eZ = np.exp(Z - Z.max(axis=1, keepdims=True))
        return eZ / np.sum(eZ, axis=1, keepdims=True)
    raise ValueError("unknown activation: %r" % (kind,))


def act_backward(kind, A):
    """Derivative expressed from the OUTPUT A.  atomic/activation_op.py:
    Sigmoid:33-34 A(1-A); Tanh:57-58 1-A^2; ReLU:90-91 (A>0); Linear:79-80 and
    SoftMax:134-135 return the scalar 1 (softmax Jacobian deliberately skipped)."""
    if kind == "sigmoid":
        return A * (1.0 - A)
    if kind == "tanh":
        return 1.0 - A ** 2
    if kind == "relu":
        return (A > 0.0).astype(A.dtype)
    if kind in ("linear", "softmax"):
        return 1.0
    raise ValueError("unknown activation: %r" % (kind,))


# --------------------------------------------------------------------------
# dense                                               atomic/core_op.py:9-20
# --------------------------------------------------------------------------
def dense_forward(X, W, b=0.0):
    """atomic/core_op.py:12-13  Z = X@W + b."""
    X, W = _operands("dense.fwd", X, W)
    return np.dot(X, W) + b


def dense_backward(X, E, W):
    """atomic/core_op.py:16-20  returns (dW, db, dX); gradients are batch SUMS."""
    Xa, Ea = _operands("dense.dW", X, E)
    dW = Xa.T @ Ea
    Eb, Wb = _operands("dense.dX", E, W)
    dX = Eb @ Wb.T
    db = (Ea if _site_modes("dense.dW") else E).sum(axis=0)      # ones-row of the same MMA: sees the rounded E
    return dW, db, dX


# --------------------------------------------------------------------------
# convolution                                      atomic/tensor_op.py:6-75
# --------------------------------------------------------------------------
def conv_valid(A, F):
    """atomic/tensor_op.py:9-32.  Cross-correlation, stride 1, no dilation.
    The reference fills an im2col matrix patch by patch (order c,ky,kx) and
    multiplies by F.reshape(nf,-1).T; here the patches are a strided view."""
    im, ic, iy, ix = A.shape
    nf, fc, fy, fx = F.shape
    if fc != ic:
        raise ValueError(
            "Supplied filter (F) is incompatible with supplied input! (X)\n"
            "input depth: {} != {} :filter depth".format(ic, fc))
    win = sliding_window_view(A, (fy, fx), axis=(2, 3))  # (im, ic, oy, ox, fy, fx)
    out = np.tensordot(win, F, axes=([1, 4, 5], [1, 2, 3]))  # (im, oy, ox, nf)
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2))


def conv_full(A, F):
    """atomic/tensor_op.py:34-40.  Zero-pad by (f-1) on each side, then valid."""
    nf, fc, fy, fx = F.shape
    py, px = fy - 1, fx - 1
    pA = np.pad(A, ((0, 0), (0, 0), (py, py), (px, px)), mode="constant")
    return conv_valid(pA, F)


def conv_forward(A, F, mode="valid", site="conv.fwd"):
    """atomic/tensor_op.py:42-46."""
    if site is not None:
        A, F = _operands(site, A, F)
    return conv_valid(A, F) if mode == "valid" else conv_full(A, F)


def conv_backward(X, E, F):
    """atomic/tensor_op.py:49-61.  dF = valid(X^T(1023), E^T(1023))^T(1023);
    db = E.sum((0,2,3), keepdims) ; dX = full(E, flip_hw(F)^T(1023))."""
    Xa, Ea = _operands("conv.dF", X, E)
    dF = conv_forward(Xa.transpose(1, 0, 2, 3), Ea.transpose(1, 0, 2, 3), "valid", site=None).transpose(1, 0, 2, 3)
    # db: all-ones operand of the dF MMA (sees E as that operand is rounded) unless the policy says how db is formed
    # ("conv.db": a rounding mode, or None = summed in FP32 by the kernel that produced E)
    pol = _TF32["policy"]
    if pol is None:
        db = E.sum(axis=(0, 2, 3), keepdims=True)
    elif _db_site_declared(pol):
        m = _site_modes("conv.db")
        db = (tf32_round(E, m) if m is not None else tf32_round(E, None)).sum(axis=(0, 2, 3), keepdims=True)
    else:
        db = (Ea if _site_modes("conv.dF") else E).sum(axis=(0, 2, 3), keepdims=True)
    Eb, Fb = _operands("conv.dX", E, F)
    dX = conv_forward(Eb, Fb[:, :, ::-1, ::-1].transpose(1, 0, 2, 3), "full", site=None)
    return np.ascontiguousarray(dF), db, dX


def conv_outshape(inshape, fshape, mode="valid"):
    """llatomic/lltensor_op.py:111-119 (the consistent nf,fc,fy,fx unpacking)."""
    ic, iy, ix = inshape[-3:]
    nf, fc, fy, fx = fshape
    if mode == "valid":
        return nf, iy - fy + 1, ix - fx + 1
    return nf, iy + fy - 1, ix + fx - 1


# --------------------------------------------------------------------------
# max pooling                                     atomic/tensor_op.py:78-128
# --------------------------------------------------------------------------
def maxpool_forward(A, fdim):
    """atomic/tensor_op.py:93-107.  Non-overlapping fdim x fdim windows; returns
    (output, filt) where filt has A's shape and is 1.0 at EVERY element equal to
    its window's max (multi-hot on ties), 0.0 elsewhere."""
    im, ic, iy, ix = A.shape
    oy, ox = iy // fdim, ix // fdim
    W = A[:, :, :oy * fdim, :ox * fdim].reshape(im, ic, oy, fdim, ox, fdim)
    out = W.max(axis=(3, 5))
    filt = np.zeros_like(A, dtype=F64)
    eq = (W == out[:, :, :, None, :, None]).astype(F64)
    filt[:, :, :oy * fdim, :ox * fdim] = eq.reshape(im, ic, oy * fdim, ox * fdim)
    return out.astype(F64), filt


def maxpool_backward(E, filt):
    """atomic/tensor_op.py:110-119.  dX[window] = filt[window] * E[m,c,i,j].
    (The reference does it in place on the cached mask and returns it; the
    oracle returns a fresh array and leaves ``filt`` untouched.)"""
    em, ec, ey, ex = E.shape
    fm, fc, fy, fx = filt.shape
    fdim = fy // ey
    up = np.repeat(np.repeat(E, fdim, axis=2), fdim, axis=3)
    out = np.array(filt, dtype=F64, copy=True)
    out[:, :, :ey * fdim, :ex * fdim] *= up
    return out


# --------------------------------------------------------------------------
# costs and metrics                   metrics/costs.py:11-37, metrics.py:13-16
# --------------------------------------------------------------------------
def cost_derivative(outputs, targets, kind="mse"):
    """metrics/costs.py:19-21  delta = outputs - targets (for every cost), except hinge (:58-66):
    -targets, zeroed where outputs > 1."""
    if kind == "hinge":
        out = -np.asarray(targets, dtype=F64)
        out[np.asarray(outputs) > 1.0] = 0.0
        return out
    return outputs - targets


def cost_value(kind, outputs, targets):
    """metrics/costs.py:26-27 (mse = 0.5*||o-t||^2, a SUM) and :36-37
    (cxent = -sum(t*log o)).  The caller divides by m."""
    if kind in ("mse", "mean_squared_error"):
        return 0.5 * np.linalg.norm(outputs - targets) ** 2.0
    if kind in ("cxent", "categorical_crossentropy"):
        return -(targets * np.log(outputs)).sum()
    if kind in ("bxent", "binary_crossentropy"):          # metrics/costs.py:48-49
        return -(targets * np.log(outputs) + (1.0 - targets) * np.log(1.0 - outputs)).sum()
    if kind == "hinge":                                   # metrics/costs.py:54-55
        return np.maximum(0.0, 1.0 - targets * outputs).sum()
    raise ValueError("No such cost function: %r" % (kind,))


def accuracy(outputs, targets):
    """metrics/metrics.py:15-16  count of argmax matches (caller divides by m)."""
    return np.sum(outputs.argmax(axis=1) == targets.argmax(axis=1))


# --------------------------------------------------------------------------
# plain recurrence                              atomic/recurrent_op.py:9-43
# --------------------------------------------------------------------------
def rnn_forward(X, W, b, activation="tanh"):
    """atomic/recurrent_op.py:14-27.  X (T,B,D); W (D+H,H).  Returns O (T,B,H), Z (T,B,D+H)."""
    H = W.shape[-1]
    T, B, D = X.shape
    Z = np.zeros((T, B, D + H))
    O = np.zeros((T, B, H))
    for t in range(T):
        Z[t] = np.concatenate((X[t], O[t - 1]), axis=-1)
        O[t] = act_forward(activation, np.dot(Z[t], W) + b)
    return O, Z


def rnn_backward(Z, O, E, W, activation="tanh"):
    """atomic/recurrent_op.py:29-43 (E copied; the reference scribbles on the caller's array).  The returned dX is the
    reference's `E[:, :, :indim]` (:40) -- a slice of the scaled error tensor, not deltaZ; kept as is."""
    H = W.shape[-1]
    T, B, zdim = Z.shape
    D = zdim - H
    E = np.array(E, dtype=F64, copy=True)
    bwO = act_backward(activation, O)
    for t in range(T - 1, -1, -1):
        E[t] *= bwO[t]
        deltaZ = np.dot(E[t], W.T)
        if t:
            E[t - 1] += deltaZ[:, D:]
    nablaW = np.matmul(Z.transpose(0, 2, 1), E).sum(axis=0)
    nablab = E.sum(axis=(0, 1))
    return E[:, :, :D], nablaW, nablab


# --------------------------------------------------------------------------
# Clockwork layer                               layers/recurrent.py:193-283
# --------------------------------------------------------------------------
def clockwork_init(indim, neurons, blocksizes=None, ticktimes=None):
    """layers/recurrent.py:195-237: default blocks / ticks, then the weights with the reference's RNG consumption
    (U first, one draw per block) and its `start = i * bls` offsets.  Returns (W (H+D,H), tick_array (H))."""
    if blocksizes is None:
        block = neurons // 5
        blocksizes = [block] * 5
        blocksizes[0] += (neurons % block)
    if ticktimes is None:
        ticktimes = [2 ** i for i in range(len(blocksizes))]
    W = np.zeros((neurons, neurons))
    U = white(indim, neurons)
    tick_array = np.zeros(neurons)
    for i, bls in enumerate(blocksizes):
        start = i * bls
        end = start + bls
        W[start:end, start:] = white(*W[start:end, start:].shape)
        tick_array[start:end] = ticktimes[i]
    return np.concatenate((W, U), axis=0), tick_array


def clockwork_forward(X, W, b, tick, activation="tanh"):
    """layers/recurrent.py:239-250.  X (T,B,D) time-major.  Returns the outputs (T,B,H) and the operands Z (T,B,D+H)."""
    T, B, D = X.shape
    H = W.shape[-1]
    out = np.zeros((B, H))
    cache, Zs = [], []
    for t in range(1, T + 1):
        with np.errstate(divide="ignore", invalid="ignore"):
            gate = np.equal(t % tick, 0.0)
        Z = np.concatenate((X[t - 1], out), axis=-1)
        out = act_forward(activation, Z.dot(W * gate[None, :]) + b * gate)
        Zs.append(Z); cache.append(out)
    return np.array(cache).reshape(T, B, H), np.array(Zs).reshape(T, B, D + H)


def clockwork_backward(Z, O, delta, W, tick, activation="tanh"):
    """layers/recurrent.py:259-277.  Returns dX (T,B,D), nablaW (D+H,H), nablab (H)."""
    T, B, zdim = Z.shape
    H = W.shape[-1]
    dh = np.zeros((B, H))
    dX = np.zeros((T, B, zdim - H))
    nW, nb = np.zeros_like(W), np.zeros(H)
    for t in range(T - 1, -1, -1):
        with np.errstate(divide="ignore", invalid="ignore"):
            gate = np.equal((t + 1) % tick, 0.0)
        dh = (dh + delta[t]) * act_backward(activation, O[t])
        nW += (Z[t].T @ dh) * gate[None, :]
        nb += dh.sum(axis=0) * gate
        deltaZ = dh @ (W * gate[None, :]).T
        dX[t] = deltaZ[:, :-H]
        dh = deltaZ[:, -H:]
    return dX, nW, nb


# --------------------------------------------------------------------------
# GRU                                           layers/recurrent.py:116-190
# --------------------------------------------------------------------------
def gru_forward(X, W, b, activation="tanh"):
    """layers/recurrent.py:123-153.  X (T,B,D) time-major; W (D+H,3H), columns [U|R|O].  Returns the hidden states
    hs (T,B,H), the operands Z = [X[t], h] and Zk = [X[t], R*h] (T,B,D+H), gates (3,T,B,H) = U, R, O."""
    H = W.shape[-1] // 3
    T, B, D = X.shape
    Wur, Wo = W[:, :-H], W[:, -H:]
    bur, bo = b[:-H], b[-H:]
    h = np.zeros((B, H))
    hs, Zs, Zks, gates = [], [], [], []
    for t in range(T):
        Z = np.concatenate((X[t], h), axis=1)
        U, R = np.split(act_forward("sigmoid", Z @ Wur + bur), 2, axis=1)
        Zk = np.concatenate((X[t], R * h), axis=1)
        O = act_forward(activation, Zk @ Wo + bo)
        h = U * h + (1.0 - U) * O
        hs.append(h); Zs.append(Z); Zks.append(Zk); gates.append((U, R, O))
    g = np.array(gates).transpose(1, 0, 2, 3) if T else np.zeros((3, 0, B, H))
    return np.array(hs), np.array(Zs), np.array(Zks), g


def gru_backward(Z, Zk, gates, delta, W, activation="tanh"):
    """layers/recurrent.py:155-190.  delta (T,B,H).  Returns dX (T,B,D), nablaW (D+H,3H), nablab (3H)."""
    H = W.shape[-1] // 3
    T, B, zdim = Z.shape
    D = zdim - H
    Wu, Wr, Wo = np.split(W, 3, axis=1)
    Wur = np.concatenate((Wu, Wr), axis=1)
    dh = np.zeros((B, H))
    dX = np.zeros((T, B, D))
    nW, nb = np.zeros_like(W), np.zeros(3 * H)
    for t in range(T - 1, -1, -1):
        U, R, O = gates[0][t], gates[1][t], gates[2][t]
        prevout = Z[t][:, -H:]
        dh = dh + delta[t]
        dU = (prevout - O) * act_backward("sigmoid", U) * dh
        dO = (1.0 - U) * act_backward(activation, O) * dh
        dZk = dO @ Wo.T
        dK = dZk[:, -H:]
        dR = prevout * act_backward("sigmoid", R) * dK
        dUR = np.concatenate((dU, dR), axis=1)
        dZ = dUR @ Wur.T
        nW += np.concatenate((Z[t].T @ dUR, Zk[t].T @ dO), axis=1)
        nb += np.concatenate((dU, dR, dO), axis=1).sum(axis=0)
        dh = dZ[:, -H:] + R * dK + U * dh
        dX[t] = dZ[:, :-H] + dZk[:, :-H]
    return dX, nW, nb


# --------------------------------------------------------------------------
# LSTM                                        atomic/recurrent_op.py:46-112
# --------------------------------------------------------------------------
def lstm_forward(X, W, b, activation="tanh"):
    """atomic/recurrent_op.py:51-77.  X (T,B,D); W (D+H,4H); gate order along 4H
    is [cand | f | i | o].  Returns O (T,B,H), Z (T,B,D+H), cache (6,T,B,H)
    ordered C, Ca, cand, f, i, o.  At t=0, O[t-1] is the still-zero last row."""
    H = W.shape[-1] // 4
    T, B, D = X.shape
    Z = np.zeros((T, B, D + H))
    O = np.zeros((T, B, H))
    C, f, i, o, cand, Ca = np.zeros((6, T, B, H))
    for t in range(T):
        Z[t] = np.concatenate((X[t], O[t - 1]), axis=-1)
        Zt, Wt = _operands("lstm.fwd", Z[t], W)
        p = np.dot(Zt, Wt) + b
        p[:, :H] = act_forward(activation, p[:, :H])
        p[:, H:] = act_forward("sigmoid", p[:, H:])
        cand[t] = p[:, :H]
        f[t] = p[:, H:2 * H]
        i[t] = p[:, 2 * H:3 * H]
        o[t] = p[:, 3 * H:]
        C[t] = C[t - 1] * f[t] + cand[t] * i[t]
        Ca[t] = act_forward(activation, C[t])
        O[t] = Ca[t] * o[t]
    return O, Z, np.stack((C, Ca, cand, f, i, o))


def lstm_backward(Z, O, E, W, cache, activation="tanh"):
    """atomic/recurrent_op.py:79-112.  E (T,B,H) is the error on O (copied, the
    reference scribbles on the caller's array).  Returns (dX (T,B,D),
    nablaW (D+H,4H), nablab (4H,))."""
    H = W.shape[-1] // 4
    T, B, zdim = Z.shape
    D = zdim - H
    E = np.array(E, dtype=F64, copy=True)
    C, Ca, cand, f, i, o = cache
    bwgates = np.concatenate(cache[2:], axis=-1)
    bwgates[..., H:] = act_backward("sigmoid", bwgates[..., H:])
    bwgates[..., :H] = act_backward(activation, bwgates[..., :H])
    bwCa = np.atleast_2d(act_backward(activation, Ca))
    deltaC = np.zeros_like(O[-1])
    deltaZ = np.zeros_like(Z)
    dgates = np.zeros((T, B, 4 * H))
    for t in range(T - 1, -1, -1):
        deltaC += E[t] * o[t] * bwCa[t]
        dcand = deltaC * i[t]
        df = deltaC * (C[t - 1] if t else 0.0)
        di = deltaC * cand[t]
        do = Ca[t] * E[t]
        dgates[t] = np.concatenate((dcand, df, di, do), axis=-1) * bwgates[t]
        deltaC *= f[t]
        dg, Wb = _operands("lstm.dZ", dgates[t], W)
        deltaZ[t] = np.dot(dg, Wb.T)
        if t:
            E[t - 1] += deltaZ[t, :, D:]
    Za, dga = _operands("lstm.dW", Z, dgates)
    nablaW = np.matmul(Za.transpose(0, 2, 1), dga).sum(axis=0)
    nablab = np.sum(dga if _site_modes("lstm.dW") else dgates, axis=(0, 1))
    return deltaZ[:, :, :D], nablaW, nablab


# --------------------------------------------------------------------------
# optimizers          optimizers/gradient_descent.py, optimizers/adaptive_gd.py
# --------------------------------------------------------------------------
class Optimizer:
    """Flat-vector update rules; ``optimize(W, gW, m) -> W'``.  gW is the batch
    SUM gradient; the 1/m lives here (gradient_descent.py:9)."""

    def __init__(self, kind, nparams, **hp):
        self.kind = kind
        d = dict(
            sgd=dict(eta=0.01),                                   # abstract_optimizer.py:21
            momentum=dict(eta=0.1, mu=0.9),                       # gradient_descent.py:17
            nesterov=dict(eta=0.1, mu=0.9),                       # gradient_descent.py:36
            adagrad=dict(eta=0.01, epsilon=1e-8),                 # adaptive_gd.py:7
            rmsprop=dict(eta=0.1, decay=0.9, epsilon=1e-8),       # adaptive_gd.py:67
            adam=dict(eta=0.001, decay_memory=0.999, decay_velocity=0.9, epsilon=1e-5),  # adaptive_gd.py:83
        )[kind]
        d.update(hp)
        self.__dict__.update(d)
        self.velocity = np.zeros(nparams)
        self.memory = np.zeros(nparams)

    def optimize(self, W, gW, m):
        k = self.kind
        if k == "sgd":                       # gradient_descent.py:8-9
            return W - gW * (self.eta / m)
        if k == "momentum":                  # gradient_descent.py:25-28
            self.velocity = self.velocity * self.mu + gW * (self.eta / m)
            return W - self.velocity
        if k == "nesterov":                  # gradient_descent.py:44-50 (literal, incl. its quirks)
            nabla = gW * (self.eta / m)
            W_ = self.memory - self.velocity + nabla
            self.memory = W
            self.velocity = self.velocity * self.mu + nabla
            return W_
        if k == "adagrad":                   # adaptive_gd.py:16-20
            nabla = gW / m
            self.memory = self.memory + nabla ** 2
            return W - (self.eta / np.sqrt(self.memory + self.epsilon)) * nabla
        if k == "rmsprop":                   # adaptive_gd.py:71-76
            nabla = gW / m
            self.memory = self.memory * self.decay + (1.0 - self.decay) * nabla ** 2
            return W - (self.eta * nabla) / np.sqrt(self.memory + self.epsilon)
        if k == "adam":                      # adaptive_gd.py:94-99 (non-standard: raw sum gradient in
            eta = self.eta / m               #  the moments, no bias correction, eps inside the sqrt)
            self.velocity = self.decay_velocity * self.velocity + (1.0 - self.decay_velocity) * gW
            self.memory = self.decay_memory * self.memory + (1.0 - self.decay_memory) * gW ** 2
            return W - (eta * self.velocity) / np.sqrt(self.memory + self.epsilon)
        raise ValueError(k)


# --------------------------------------------------------------------------
# weight initialisers                                  util/typing.py:22-33
# --------------------------------------------------------------------------
def white(*dims):
    """util/typing.py:22-33.  Draws from the GLOBAL numpy RNG exactly like the
    reference so that the same seed yields the same weights."""
    if len(dims) == 2:
        fanin, fanout = dims
        return np.random.randn(fanin, fanout) * np.sqrt(2.0 / float(fanin + fanout))
    nf, fc, fy, fx = dims
    return np.random.randn(nf, fc, fy, fx) * np.sqrt(2.0 / float(nf * fy * fx + fc * fy * fx))


# --------------------------------------------------------------------------
# layer glue + learner                layers/*.py, learner/backpropagation.py
# --------------------------------------------------------------------------
class Stack:
    """Spec-driven restatement of LayerStack + Backpropagation for the layers on
    the hot path.  ``spec`` entries:

      ("dense", neurons, activation)            layers/core.py:9-42
      ("act", activation)                       layers/core.py:45-52
      ("flatten",)                              layers/core.py:70-100
      ("conv", nfilters, fx, fy, activation)    layers/tensor.py:43-92
      ("pool", fdim)                            layers/tensor.py:7-40
      ("gap",)                                  layers/tensor.py:95-114 (GlobalAveragePooling)
      ("dropout", dropchance)                   layers/fancy.py:60-90 (set `stack.learning = True` while training)
      ("lstm", neurons, activation, return_seq) layers/recurrent.py:12-47,80-113
      ("rnn", neurons, activation, return_seq)  layers/recurrent.py:12-47,50-77 (RLayer)
      ("gru", neurons, activation, return_seq)  layers/recurrent.py:12-47,116-190
      ("clockwork", neurons, activation, return_seq[, blocksizes, ticktimes])  layers/recurrent.py:193-283

    Weights are drawn with ``white`` in stack order at construction, i.e. with
    the same RNG consumption as the reference's ``connect`` calls.
    """

    def __init__(self, input_shape, spec, cost="mse", optimizer="sgd", **opt_hp):
        if isinstance(input_shape, int):
            input_shape = (input_shape,)
        self.input_shape = tuple(input_shape)
        self.spec = [tuple(s) for s in spec]
        self.cost = cost
        self.L = []
        shape = self.input_shape
        for s in self.spec:
            kind = s[0]
            lay = dict(kind=kind, inshape=shape)
            if kind == "dense":
                if len(shape) != 1:
                    raise RuntimeError("Dense only accepts input shapes with 1 dimension!")
                n = int(s[1])
                lay.update(act=s[2] if len(s) > 2 else "linear",
                           W=white(shape[0], n), b=np.zeros(n))
                shape = (n,)
            elif kind == "act":
                lay.update(act=s[1])
            elif kind == "flatten":
                shape = (int(np.prod(shape)),)
            elif kind == "conv":
                nf, fx, fy = int(s[1]), int(s[2]), int(s[3])
                depth, iy, ix = shape[-3:]
                if iy < fy or ix < fx:
                    raise RuntimeError("Incompatible shapes")
                # layers/tensor.py:70: white(nfilters, depth, fx, fy); the op reads dims 2,3 as (fy,fx)
                lay.update(act=s[4] if len(s) > 4 else "linear",
                           W=white(nf, depth, fx, fy), b=np.zeros((1, nf, 1, 1)))
                shape = conv_outshape(shape, lay["W"].shape, "valid")
            elif kind == "pool":
                fdim = int(s[1])
                ic, iy, ix = shape[-3:]
                if iy % fdim or ix % fdim:
                    raise RuntimeError("Incompatible shapes: {} % {}".format((ix, iy), fdim))
                lay.update(fdim=fdim)
                shape = (ic, iy // fdim, ix // fdim)
            elif kind == "gap":                    # layers/tensor.py:95-114
                shape = (shape[0],)
            elif kind == "dropout":                # layers/fancy.py:60-90
                lay.update(keep=1.0 - float(s[1]), mask=None)
            elif kind == "lstm":
                n = int(s[1])
                zdim = shape[-1] + n
                lay.update(act=s[2] if len(s) > 2 else "tanh", n=n,
                           return_seq=bool(s[3]) if len(s) > 3 else False,
                           W=white(zdim, 4 * n), b=np.zeros(4 * n) + 7.0)  # layers/recurrent.py:82,98
                shape = (shape[0], n) if lay["return_seq"] else (n,)
            elif kind == "clockwork":
                n = int(s[1])
                W, tick = clockwork_init(shape[-1], n, s[4] if len(s) > 4 else None, s[5] if len(s) > 5 else None)
                lay.update(act=s[2] if len(s) > 2 else "tanh", n=n, return_seq=bool(s[3]) if len(s) > 3 else False,
                           W=W, b=np.zeros(n), tick=tick)
                shape = (shape[0], n) if lay["return_seq"] else (n,)
            elif kind == "gru":
                n = int(s[1])
                lay.update(act=s[2] if len(s) > 2 else "tanh", n=n,
                           return_seq=bool(s[3]) if len(s) > 3 else False,
                           W=white(shape[-1] + n, 3 * n), b=np.zeros(3 * n))     # layers/recurrent.py:118-121
                shape = (shape[0], n) if lay["return_seq"] else (n,)
            elif kind == "rnn":
                n = int(s[1])
                zdim = shape[-1] + n
                lay.update(act=s[2] if len(s) > 2 else "tanh", n=n,
                           return_seq=bool(s[3]) if len(s) > 3 else False,
                           W=white(zdim, n), b=np.zeros(n))                # layers/recurrent.py:61-64
                shape = (shape[0], n) if lay["return_seq"] else (n,)
            else:
                raise ValueError(kind)
            lay["outshape"] = shape
            if "W" in lay:
                lay["gW"] = np.zeros_like(lay["W"])
                lay["gb"] = np.zeros_like(lay["b"])
            self.L.append(lay)
        self.outshape = shape
        self.opt = Optimizer(optimizer, self.num_params, **opt_hp)
        # {layer index: array}: discrete decisions imposed on the backward pass -- the 0/1 mask of a pooling layer, the
        # 0/1 gate of a ReLU layer.  A tensor-core (TF32) forward pass differs from this float64 one by ~1e-5, enough to
        # flip the arg-max of a near-tied pooling window or the sign of a near-zero pre-activation in a few places per
        # batch; each flip moves a gradient discontinuously.  Tests that hold the device's ARITHMETIC to 1e-4 impose
        # the device's decisions here and check separately that every differing decision is a near-tie.
        self.forced = {}

    # ---- flat vectors: model/layerstack.py:46-61, layers/abstract_layer.py:40-55
    @property
    def trainable(self):
        return [l for l in self.L if "W" in l]

    @property
    def num_params(self):
        return sum(l["W"].size + l["b"].size for l in self.trainable)

    def get_weights(self):
        return np.concatenate([np.concatenate((l["W"].ravel(), l["b"].ravel())) for l in self.trainable])

    def set_weights(self, w):
        w = np.asarray(w, dtype=F64)
        s = 0
        for l in self.trainable:
            nW, nb = l["W"].size, l["b"].size
            l["W"] = w[s:s + nW].reshape(l["W"].shape).copy()
            l["b"] = w[s + nW:s + nW + nb].reshape(l["b"].shape).copy()
            s += nW + nb

    def get_gradients(self):
        return np.concatenate([np.concatenate((l["gW"].ravel(), l["gb"].ravel())) for l in self.trainable])

    def zero_gradients(self):                      # learner/backpropagation.py:68-73
        for l in self.trainable:
            l["gW"] = np.zeros_like(l["W"])
            l["gb"] = np.zeros_like(l["b"])

    # ---- forward: model/layerstack.py:37-40
    def feedforward(self, X):
        X = np.asarray(X, dtype=F64)
        for li, l in enumerate(self.L):
            _TF32["layer"] = li
            k = l["kind"]
            if k == "dense":                       # layers/core.py:27-32
                l["inputs"] = X
                X = act_forward(l["act"], dense_forward(X, l["W"], l["b"]))
            elif k == "act":                       # layers/core.py:47-49
                X = act_forward(l["act"], X)
            elif k == "flatten":                   # layers/core.py:82-83
                X = X.reshape(X.shape[0], -1)
            elif k == "conv":                      # layers/tensor.py:75-79 (bias AFTER the activation)
                l["inputs"] = X
                X = act_forward(l["act"], conv_forward(X, l["W"], "valid")) + l["b"]
            elif k == "pool":                      # layers/tensor.py:28-30
                X, l["filter"] = maxpool_forward(X, l["fdim"])
            elif k == "gap":                       # layers/tensor.py:102-104
                l["repeats"] = int(np.prod(X.shape[2:]))
                X = X.mean(axis=(2, 3))
            elif k == "dropout":                   # layers/fancy.py:73-77: one mask of the SAMPLE shape, global NumPy RNG
                l["mask"] = np.random.uniform(0, 1, l["inshape"]) < l["keep"]
                X = X * (l["mask"] if getattr(self, "learning", False) else l["keep"])
            elif k == "lstm":                      # layers/recurrent.py:26,101-106
                l["inputs"] = X.transpose(1, 0, 2)
                O, l["Z"], l["cache"] = lstm_forward(l["inputs"], l["W"], l["b"], l["act"])
                l["O"] = O
                X = O.transpose(1, 0, 2) if l["return_seq"] else O[-1]
            elif k == "clockwork":                 # layers/recurrent.py:239-257
                l["inputs"] = X.transpose(1, 0, 2)
                l["O"], l["Z"] = clockwork_forward(l["inputs"], l["W"], l["b"], l["tick"], l["act"])
                X = l["O"].transpose(1, 0, 2) if l["return_seq"] else l["O"][-1]
            elif k == "gru":                       # layers/recurrent.py:123-153
                l["inputs"] = X.transpose(1, 0, 2)
                hs, l["Z"], l["Zk"], l["gates"] = gru_forward(l["inputs"], l["W"], l["b"], l["act"])
                X = hs.transpose(1, 0, 2) if l["return_seq"] else hs[-1]
            elif k == "rnn":                       # layers/recurrent.py:26,66-69
                l["inputs"] = X.transpose(1, 0, 2)
                O, l["Z"] = rnn_forward(l["inputs"], l["W"], l["b"], l["act"])
                l["O"] = O
                X = O.transpose(1, 0, 2) if l["return_seq"] else O[-1]
            l["output"] = X
        return X

    # ---- backward: learner/backpropagation.py:32-35
    def backpropagate(self, error):
        error = np.array(error, dtype=F64, copy=True)
        for li in range(len(self.L) - 1, -1, -1):
            l = self.L[li]
            _TF32["layer"] = li
            l["delta_in"] = error          # kept for layer-by-layer comparisons in the tests
            k = l["kind"]
            if k == "dense":                       # layers/core.py:34-39 (ACCUMULATES)
                error = error * act_backward(l["act"], l["output"])
                dW, db, error = dense_backward(l["inputs"], error, l["W"])
                l["gW"] = l["gW"] + dW
                l["gb"] = l["gb"] + db
            elif k == "act":                       # layers/core.py:51-52
                gate = self.forced.get(li)         # tests: the DEVICE's gate decisions (see `forced`)
                error = error * (gate if gate is not None else act_backward(l["act"], l["output"]))
            elif k == "flatten":                   # layers/core.py:85-86
                error = error.reshape((-1,) + tuple(l["inshape"]))
            elif k == "conv":                      # layers/tensor.py:81-84 (ASSIGNS)
                error = error * act_backward(l["act"], l["output"])
                l["gW"], l["gb"], error = conv_backward(l["inputs"], error, l["W"])
            elif k == "pool":                      # layers/tensor.py:32-33
                error = maxpool_backward(error, self.forced.get(li, l["filter"]))
            elif k == "dropout":                   # layers/fancy.py:79-82 (the mask is reset to the keep probability)
                error = error * l["mask"]
                l["mask"] = np.ones_like(l["mask"], dtype=F64) * l["keep"]
            elif k == "gap":                       # layers/tensor.py:106-110
                m = len(error)
                error = np.repeat(error / float(l["repeats"]), l["repeats"]).reshape((m,) + tuple(l["inshape"]))
            elif k == "lstm":                      # layers/recurrent.py:31-40,108-113 (ASSIGNS)
                if l["return_seq"]:
                    E = error.transpose(1, 0, 2)
                else:
                    E = np.zeros((l["inputs"].shape[0], len(error), l["n"]))
                    E[-1] = error
                dX, l["gW"], l["gb"] = lstm_backward(l["Z"], l["O"], E, l["W"], l["cache"], l["act"])
                error = dX.transpose(1, 0, 2)
            elif k == "clockwork":                 # layers/recurrent.py:31-40,259-277
                if l["return_seq"]:
                    E = error.transpose(1, 0, 2)
                else:
                    E = np.zeros((l["inputs"].shape[0], len(error), l["n"]))
                    E[-1] = error
                dX, l["gW"], l["gb"] = clockwork_backward(l["Z"], l["O"], E, l["W"], l["tick"], l["act"])
                error = dX.transpose(1, 0, 2)
            elif k == "gru":                       # layers/recurrent.py:31-40,155-190
                if l["return_seq"]:
                    E = error.transpose(1, 0, 2)
                else:
                    E = np.zeros((l["inputs"].shape[0], len(error), l["n"]))
                    E[-1] = error
                dX, l["gW"], l["gb"] = gru_backward(l["Z"], l["Zk"], l["gates"], E, l["W"], l["act"])
                error = dX.transpose(1, 0, 2)
            elif k == "rnn":                       # layers/recurrent.py:31-40,71-77 (ASSIGNS)
                if l["return_seq"]:
                    E = error.transpose(1, 0, 2)
                else:
                    E = np.zeros((l["inputs"].shape[0], len(error), l["n"]))
                    E[-1] = error
                dX, l["gW"], l["gb"] = rnn_backward(l["Z"], l["O"], E, l["W"], l["act"])
                error = dX.transpose(1, 0, 2)
        return error

    # ---- learner/backpropagation.py:16-30,37-42
    def learn_batch(self, X, Y, w=None, update=True, metrics=()):
        m = len(X)
        preds = self.feedforward(X)
        delta = cost_derivative(preds, np.asarray(Y, dtype=F64), self.cost)
        if w is not None:
            delta = delta * np.asarray(w)[:, None]
        self.backpropagate(delta)
        if update:
            self.update(m)
        out = {"cost": cost_value(self.cost, preds, Y) / m}
        if "acc" in metrics or "accuracy" in metrics:
            out["accuracy"] = accuracy(preds, Y) / m
        return out

    def update(self, m):
        W = self.get_weights()
        gW = self.get_gradients()
        self.set_weights(self.opt.optimize(W, gW, m))
        self.zero_gradients()


# --------------------------------------------------------------------------
# neuro-evolution fitness    learner/neuroevolution.py:20-22,34-37 (+ SURVEY 7.2#10)
# --------------------------------------------------------------------------
def as_weights(genome):
    """learner/neuroevolution.py:20-22."""
    return (genome - 0.5) * 20.0


def fitness(stack, genome, X, Y):
    """Parity pin for the config-5 path: set_weights(as_weights(g)) ->
    feedforward(X) -> cost(out, Y)/m (the reference's default ``fitness`` raises
    TypeError as shipped, neuroevolution.py:36; this is what it intends)."""
    stack.set_weights(as_weights(np.asarray(genome, dtype=F64)))
    out = stack.feedforward(X)
    return cost_value(stack.cost, out, Y) / len(X)


def fitness_mlp_logsoftmax(genomes, X, Y, nin, nh, nout, chunk=64):
    """Vectorised, underflow-free float64 fitness of the Dense(nh,sigmoid) ->
    Dense(nout,softmax) MLP under cxent for a batch of genomes (P, loci):
    -sum(Y*log_softmax(logits))/m.  Equal to ``fitness`` wherever the literal
    float64 formula does not hit log(0).  The first layers of `chunk` genomes
    are one BLAS product X @ [W1_0 | W1_1 | ...] (site "fit.l1" under TF32
    emulation, where as_weights is evaluated in FP32 like the device does)."""
    genomes = np.asarray(genomes)
    X = np.asarray(X, dtype=F64)
    Y = np.asarray(Y, dtype=F64)
    o1 = nin * nh
    out = np.zeros(len(genomes))
    emul = _TF32["policy"] is not None
    for p0 in range(0, len(genomes), chunk):
        g = genomes[p0:p0 + chunk]
        if emul:
            G = ((g.astype(np.float32) - np.float32(0.5)) * np.float32(20.0)).astype(F64)
        else:
            G = as_weights(np.asarray(g, dtype=F64))
        P = len(G)
        W1 = G[:, :o1].reshape(P, nin, nh).transpose(1, 0, 2).reshape(nin, P * nh)
        b1 = G[:, o1:o1 + nh]
        W2 = G[:, o1 + nh:o1 + nh + nh * nout].reshape(P, nh, nout)
        b2 = G[:, o1 + nh + nh * nout:]
        Xa, W1a = _operands("fit.l1", X, W1)
        Hh = 1.0 / (1.0 + np.exp(-((Xa @ W1a).reshape(len(X), P, nh).transpose(1, 0, 2) + b1[:, None, :])))
        L = np.matmul(Hh, W2) + b2[:, None, :]
        L = L - L.max(axis=2, keepdims=True)
        lsm = L - np.log(np.exp(L).sum(axis=2, keepdims=True))
        out[p0:p0 + chunk] = -(Y[None] * lsm).sum(axis=(1, 2)) / len(X)
    return out
