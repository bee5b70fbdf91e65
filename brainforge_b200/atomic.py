"""`b200atomic`: the op layer, same names and signatures as the reference's `brainforge.atomic` /
`brainforge.llatomic` (atomic/__init__.py:1-6), executing on the B200 through libbf_b200.so.

This is the plug-in seam the reference already has (layers pick `atomic` or `llatomic` per layer,
layers/core.py:20-25, layers/tensor.py:11-17,56-59): this module is the third back-end behind it.
Every op accepts host NumPy arrays (copied in, result copied back as float64 -- what the reference's own
tests do) or `DeviceArray`s (stay in HBM, what the layers use).
"""
import ctypes

import numpy as np

from . import _lib
from .device import DeviceArray, as_device, call, precision, stream_ptr

L = _lib.lib
_NULL = ctypes.c_void_p(0)


def _ret(like, *outs):
    """Return results in the caller's currency: host float64 if the input was a host array."""
    if isinstance(like, DeviceArray):
        return outs[0] if len(outs) == 1 else outs
    outs = tuple(o.numpy() if isinstance(o, DeviceArray) else o for o in outs)
    return outs[0] if len(outs) == 1 else outs


# --------------------------------------------------------------------------- activations
class ActivationFunction:
    """atomic/activation_op.py:12-23.  `backward(A)` returns act'(.) computed from the OUTPUT A; the
    layers use the fused `backward_mul(A, delta)` = delta * act'(A) instead (one pass, no temporary)."""
    type = ""

    def forward(self, Z):
        z = as_device(Z)
        out = DeviceArray.empty(*z.shape)
        self._forward_into(z, out)
        return _ret(Z, out)

    def _forward_into(self, z, out):
        call(L.bf_act_forward, _lib.ACT[self.type], z.ptr, out.ptr, z.size, stream_ptr())

    def backward(self, A):
        a = as_device(A)
        ones = DeviceArray.empty(*a.shape)
        call(L.bf_affine, a.ptr, ones.ptr, a.size, 0.0, 1.0, stream_ptr())
        call(L.bf_act_backward, _lib.ACT[self.type], a.ptr, ones.ptr, ones.ptr, a.size, stream_ptr())
        return _ret(A, ones)

    def backward_mul(self, A, delta, out=None):
        """delta * act'(A) (layers/core.py:35,52); in place when out is None."""
        a, d = as_device(A), as_device(delta)
        out = d if out is None else out
        call(L.bf_act_backward, _lib.ACT[self.type], a.ptr, d.ptr, out.ptr, d.size, stream_ptr())
        return out

    def __str__(self):
        return self.type


class Sigmoid(ActivationFunction):
    type = "sigmoid"       # activation_op.py:26-34


class Tanh(ActivationFunction):
    type = "tanh"          # activation_op.py:50-58


class ReLU(ActivationFunction):
    type = "relu"          # activation_op.py:83-91 (relu'(0) = 0: the NumPy behaviour)


class Linear(ActivationFunction):
    type = "linear"        # activation_op.py:72-80

    def forward(self, Z):
        return Z

    def backward(self, A):
        return 1.0

    def backward_mul(self, A, delta, out=None):
        return delta if out is None else super().backward_mul(A, delta, out)


class SoftMax(ActivationFunction):
    type = "softmax"       # activation_op.py:115-135 (temperature 1)

    def _forward_into(self, z, out):
        if z.ndim != 2:
            raise ValueError("softmax expects a 2-D (m, n) array")
        call(L.bf_softmax_forward, z.ptr, out.ptr, z.shape[0], z.shape[1], stream_ptr())

    def backward(self, A):
        return 1.0         # activation_op.py:134-135: the Jacobian is folded into `outputs - targets`

    def backward_mul(self, A, delta, out=None):
        return delta if out is None else super().backward_mul(A, delta, out)


activations = {"sigmoid": Sigmoid, "tanh": Tanh, "linear": Linear, "relu": ReLU, "softmax": SoftMax}


# --------------------------------------------------------------------------- dense / reshape
class DenseOp:
    """atomic/core_op.py:9-20."""

    @staticmethod
    def forward(X, W, b=None, activation="linear"):
        x, w = as_device(X), as_device(W)
        m, k = x.shape
        k2, n = w.shape
        if k != k2:
            raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (m, k, k2, n))
        bd = as_device(b) if b is not None else None
        out = DeviceArray.empty(m, n)
        call(L.bf_dense_forward, x.ptr, w.ptr, bd.ptr if bd is not None else _NULL, out.ptr, m, k, n,
             _lib.ACT[activation], precision(), stream_ptr())
        return _ret(X, out)

    @staticmethod
    def backward(X, E, W, dW=None, db=None, accumulate=False, need_dX=True):
        x, e, w = as_device(X), as_device(E), as_device(W)
        m, k = x.shape
        n = w.shape[1]
        dW = DeviceArray.empty(k, n) if dW is None else dW
        db = DeviceArray.empty(n) if db is None else db
        dX = DeviceArray.empty(m, k) if need_dX else None
        call(L.bf_dense_backward, x.ptr, e.ptr, w.ptr, dW.ptr, db.ptr, dX.ptr if dX is not None else _NULL, m, k, n,
             1 if accumulate else 0, precision(), stream_ptr())
        return _ret(X, dW, db, dX)


class ReshapeOp:
    """atomic/core_op.py:23-33."""
    type = "Reshape"

    @staticmethod
    def forward(X, outshape):
        return X.reshape(-1, *outshape)

    @staticmethod
    def backward(E, inshape):
        return ReshapeOp.forward(E, inshape)


# --------------------------------------------------------------------------- convolution
class ConvolutionOp:
    """atomic/tensor_op.py:6-75 / llatomic/lltensor_op.py:75-121.  Filter layout (nf, fc, fy, fx)."""

    @staticmethod
    def forward(A, F, mode="valid", bias=None, activation="linear"):
        a, f = as_device(A), as_device(F)
        im, ic, iy, ix = a.shape
        nf, fc, fy, fx = f.shape
        if mode not in _lib.CONV_MODE:
            raise RuntimeError("Unsupported mode: %s" % (mode,))
        _, oy, ox = ConvolutionOp.outshape(a.shape, f.shape, mode)
        out = DeviceArray.empty(im, nf, max(oy, 0), max(ox, 0))
        bd = as_device(bias).reshape(-1) if bias is not None else None
        call(L.bf_conv_forward, a.ptr, f.ptr, bd.ptr if bd is not None else _NULL, out.ptr, im, ic, iy, ix, nf, fc, fy,
             fx, _lib.CONV_MODE[mode], _lib.ACT[activation], precision(), stream_ptr())
        return _ret(A, out)

    @staticmethod
    def valid(A, F):
        return ConvolutionOp.forward(A, F, "valid")

    @staticmethod
    def full(A, F):
        return ConvolutionOp.forward(A, F, "full")

    @staticmethod
    def backward(X, E, F, dF=None, db=None, need_dX=True):
        """-> (dF, db, dX); keyword call like layers/tensor.py:83 (the two reference back-ends disagree on
        the positional order).  db has the NumPy back-end's shape (1,nf,1,1)."""
        x, e, f = as_device(X), as_device(E), as_device(F)
        im, ic, iy, ix = x.shape
        nf, fc, fy, fx = f.shape
        dF = DeviceArray.empty(nf, fc, fy, fx) if dF is None else dF
        db = DeviceArray.empty(1, nf, 1, 1) if db is None else db
        dX = DeviceArray.empty(im, ic, iy, ix) if need_dX else None
        # two calls (filter-gradient half, data-gradient half) so that each kernel can be timed on its own
        call(L.bf_conv_backward, x.ptr, e.ptr, f.ptr, dF.ptr, db.ptr, _NULL, im, ic, iy, ix, nf, fy, fx, precision(),
             stream_ptr(), tag="dF")
        if dX is not None:
            call(L.bf_conv_backward, x.ptr, e.ptr, f.ptr, _NULL, _NULL, dX.ptr, im, ic, iy, ix, nf, fy, fx,
                 precision(), stream_ptr(), tag="dX")
        return _ret(X, dF, db, dX)

    @staticmethod
    def outshape(inshape, fshape, mode="valid"):
        ic, iy, ix = inshape[-3:]
        nf, fc, fy, fx = fshape
        if mode == "valid":
            return nf, iy - fy + 1, ix - fx + 1
        elif mode == "full":
            return nf, iy + fy - 1, ix + fx - 1
        raise RuntimeError("Unsupported mode: %s" % (mode,))

    def __str__(self):
        return "Convolution"


# --------------------------------------------------------------------------- max pooling
class PoolMask:
    """The reference's `filt` (A-shaped 0/1 float array, multi-hot on ties) stored compactly: one word
    of fdim*fdim bits per window.  `np.asarray(mask)` expands it to the reference's array."""

    def __init__(self, words, ashape, fdim):
        self.words, self.ashape, self.fdim = words, tuple(ashape), fdim

    @property
    def shape(self):
        return self.ashape

    def expand(self):
        im, ic, iy, ix = self.ashape
        out = DeviceArray.empty(*self.ashape)
        call(L.bf_pool_mask_expand, ctypes.c_void_p(self.words.data_ptr()), out.ptr, im, ic, iy, ix, self.fdim,
             stream_ptr())
        return out

    def __array__(self, dtype=None, copy=None):
        return self.expand().numpy(np.float64 if dtype is None else dtype)


class MaxPoolOp:
    """atomic/tensor_op.py:78-128 / llatomic/lltensor_op.py:124-152."""

    def __str__(self):
        return "MaxPool"

    @staticmethod
    def forward(A, fdim):
        import torch
        a = as_device(A)
        im, ic, iy, ix = a.shape
        nbytes = L.bf_pool_mask_bytes(im, ic, iy, ix, fdim)
        out = DeviceArray.empty(im, ic, iy // fdim, ix // fdim)
        words = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda")
        call(L.bf_pool_forward, a.ptr, out.ptr, ctypes.c_void_p(words.data_ptr()), im, ic, iy, ix, fdim, stream_ptr())
        mask = PoolMask(words, a.shape, fdim)
        if isinstance(A, DeviceArray):
            return out, mask
        return out.numpy(), np.asarray(mask)

    @staticmethod
    def backward(E, filt):
        """E (im,ic,oy,ox), filt = the mask from forward (PoolMask, or a host 0/1 array as the reference
        returns it).  Unlike the reference (tensor_op.py:118-119) the cached mask is not consumed."""
        import torch
        e = as_device(E)
        if isinstance(filt, PoolMask):
            mask = filt
        else:  # host mask array: re-encode it as bits
            f = np.asarray(filt)
            im, ic, iy, ix = f.shape
            fdim = iy // e.shape[2]
            w = f.reshape(im, ic, iy // fdim, fdim, ix // fdim, fdim).transpose(0, 1, 2, 4, 3, 5).reshape(im, ic, iy // fdim, ix // fdim, fdim * fdim)
            bits = (w != 0).astype(np.uint64) << np.arange(fdim * fdim, dtype=np.uint64)
            words = bits.sum(axis=-1).astype(np.uint8 if fdim * fdim <= 8 else np.uint64)
            wt = torch.from_numpy(words.view(np.uint8).copy()).cuda()
            mask = PoolMask(wt, f.shape, fdim)
        im, ic, iy, ix = mask.ashape
        dX = DeviceArray.empty(im, ic, iy, ix)
        call(L.bf_pool_backward, e.ptr, ctypes.c_void_p(mask.words.data_ptr()), dX.ptr, im, ic, iy, ix, mask.fdim,
             stream_ptr())
        return _ret(E, dX)

    @staticmethod
    def outshape(inshape, fdim):
        if len(inshape) == 3:
            m, iy, ix = inshape
            return m, iy // fdim, ix // fdim
        elif len(inshape) == 2:
            iy, ix = inshape
            return iy // fdim, ix // fdim


# --------------------------------------------------------------------------- LSTM
class LSTMOp:
    """atomic/recurrent_op.py:46-112.  X is time-major (T,B,D); gate order along 4H is [cand|f|i|o]."""

    def __init__(self, activation):
        self.activation = activation
        self.actfn = activations[activation]()

    def forward(self, X, W, b):
        x, w, bd = as_device(X), as_device(W), as_device(b)
        T, B, D = x.shape
        H = w.shape[-1] // 4
        if w.shape[0] != D + H:
            raise ValueError("LSTM weight shape %s does not match D+H=%d" % (w.shape, D + H))
        O = DeviceArray.empty(T, B, H)
        Z = DeviceArray.empty(T, B, D + H)
        cache = DeviceArray.empty(6, T, B, H)
        cache.bw = DeviceArray.empty(5, T, B, H)   # act'(.) of Ca,cand,f,i,o from the pre-activations
        call(L.bf_lstm_forward, x.ptr, w.ptr, bd.ptr, O.ptr, Z.ptr, cache.ptr, cache.bw.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(X, O, Z, cache)

    def backward(self, Z, O, E, W, cache, nablaW=None, nablab=None):
        z, o, e, w, c = as_device(Z), as_device(O), as_device(E), as_device(W), as_device(cache)
        bw = getattr(cache, "bw", None)
        if bw is None and not isinstance(cache, DeviceArray):
            # host float64 cache: form the reference's bwCa / bwgates (recurrent_op.py:85-88) before narrowing
            ch = np.asarray(cache, dtype=np.float64)
            def d(kind, a):
                return {"sigmoid": a * (1. - a), "tanh": 1. - a ** 2, "relu": (a > 0.).astype(np.float64),
                        "linear": np.ones_like(a)}[kind]
            bw = DeviceArray.from_host(np.stack([d(self.activation, ch[1]), d(self.activation, ch[2]),
                                                 d("sigmoid", ch[3]), d("sigmoid", ch[4]), d("sigmoid", ch[5])]))
        T, B, ZD = z.shape
        H = w.shape[-1] // 4
        D = ZD - H
        dX = DeviceArray.empty(T, B, D)
        nablaW = DeviceArray.empty(ZD, 4 * H) if nablaW is None else nablaW
        nablab = DeviceArray.empty(4 * H) if nablab is None else nablab
        call(L.bf_lstm_backward, z.ptr, o.ptr, e.ptr, w.ptr, c.ptr, bw.ptr if bw is not None else _NULL, dX.ptr,
             nablaW.ptr, nablab.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(Z, dX, nablaW, nablab)
