"""`b200atomic`: the op layer, same names and signatures as the reference's `brainforge.atomic` /
`brainforge.llatomic` (atomic/__init__.py:1-6), executing on the B200 through libbf_b200.so.

This is the plug-in seam the reference already has (layers pick `atomic` or `llatomic` per layer,
layers/core.py:20-25, layers/tensor.py:11-17,56-59): this module is the third back-end behind it.
Every op accepts host NumPy arrays (copied in, result copied back as float64 -- what the reference's own
tests do) or `DeviceArray`s (stay in HBM, what the layers use).
"""
import ctypes
import os

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
        out = DeviceArray.empty_like(z)      # elementwise: the storage layout rides along
        self._forward_into(z, out)
        return _ret(Z, out)

    def _forward_into(self, z, out):
        call(L.bf_act_forward, _lib.ACT[self.type], z.pptr, out.pptr, z.size, stream_ptr())

    def backward(self, A):
        a = as_device(A)
        ones = DeviceArray.empty_like(a).fill(1.0)
        call(L.bf_act_backward, _lib.ACT[self.type], a.pptr, ones.pptr, ones.pptr, a.size, stream_ptr())
        return _ret(A, ones)

    def backward_mul(self, A, delta, out=None):
        """delta * act'(A) (layers/core.py:35,52); in place when out is None."""
        a, d = as_device(A), as_device(delta)
        d = d.like_layout(a)
        out = d if out is None else out
        call(L.bf_act_backward, _lib.ACT[self.type], a.pptr, d.pptr, out.pptr, d.size, stream_ptr())
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
        if x.cl_flat is not None:
            if w.shape[1] <= 32 and x.shape[1] == w.shape[0] and L.bf_dense_cl_supported(w.shape[0], w.shape[1]):
                c, y, xx = x.cl_flat
                bd = as_device(b) if b is not None else None
                out = DeviceArray.empty(x.shape[0], w.shape[1])
                call(L.bf_dense_forward_cl, x.pptr, w.ptr, bd.ptr if bd is not None else _NULL, out.ptr, x.shape[0], c, y * xx,
                     w.shape[1], _lib.ACT[activation], stream_ptr())
                return out
            x = x.to_nchw()
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
        if x.cl_flat is not None:
            if w.shape[1] <= 32 and len(x) > 0 and L.bf_dense_cl_supported(w.shape[0], w.shape[1]):
                c, y, xx = x.cl_flat
                m, n = x.shape[0], w.shape[1]
                dW = DeviceArray.empty(*w.shape) if dW is None else dW
                db = DeviceArray.empty(n) if db is None else db
                dX = DeviceArray(DeviceArray.empty(m, c * y * xx).t, cl_flat=x.cl_flat) if need_dX else None
                call(L.bf_dense_backward_cl, x.pptr, e.ptr, w.ptr, dW.ptr, db.ptr, dX.pptr if dX is not None else _NULL, m, c,
                     y * xx, n, 1 if accumulate else 0, stream_ptr())
                return dW, db, dX
            x = x.to_nchw()
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
class FilterPack:
    """The two packed forms of a filter the channels-last kernels contract against (forward [tap][f][c]; data gradient,
    flipped and transposed, [tap][c][f]), written by ONE launch in the forward pass and reused by the backward pass of
    the same step.  `current(F)` is true only while no host-API call has written parameter memory since (device.mutations)."""

    def __init__(self):
        self.fwd = self.dx = None
        self.stamp = self.src = None

    def repack(self, f):
        from .device import mutations
        nf, ic, fy, fx = f.shape
        if self.fwd is None or self.fwd.size != f.size:
            self.fwd, self.dx = DeviceArray.empty(f.size), DeviceArray.empty(f.size)
        call(L.bf_conv_nhwc_repack, f.ptr, self.fwd.ptr, self.dx.ptr, nf, ic, fy, fx, stream_ptr())
        self.stamp, self.src = mutations(), f.t.data_ptr()

    def current(self, f):
        from .device import mutations
        return self.dx is not None and self.stamp == mutations() and self.src == f.t.data_ptr() and self.dx.size == f.size


class ConvolutionOp:
    """atomic/tensor_op.py:6-75 / llatomic/lltensor_op.py:75-121.  Filter layout (nf, fc, fy, fx).

    Kernel selection (BF_PREC_TF32 only; BF_PREC_FP32 always takes the SIMT implicit GEMM on NCHW):
      * input channels-last (or NCHW with ic % 32 == 0, converted once) and nf in {32,64,128}: the TMA + tcgen05
        channels-last kernels (`bf_conv_nhwc_*`), output channels-last;
      * NCHW input with ic <= 4 (the first layer): `bf_conv_c3_*`, NCHW in, channels-last out;
      * anything else: the gather-staged implicit GEMM on NCHW (`bf_conv_forward/backward`)."""

    channels_last = True      # class-wide switch: False keeps every convolution on the NCHW gather kernels
    db_from_pool = True       # take db from the channel sums the pooling backward pass leaves on E (no all-ones MMA)
    # channels-last convolutions absorb a following [ReLU ->] PoolLayer(2) in their epilogue (BF_FUSE_POOL_NHWC=0: two kernels)
    fuse_pool_nhwc = os.environ.get("BF_FUSE_POOL_NHWC", "1") != "0"

    @staticmethod
    def _route(a_nhwc, ic, iy, ix, nf, fy, fx, mode, activation):
        if not ConvolutionOp.channels_last:
            return "generic"
        if precision() != _lib.PREC["tf32"] or mode != "valid" or activation not in ("linear", "relu"):
            return "generic"
        if iy < fy or ix < fx:
            return "generic"
        if L.bf_conv_nhwc_supported(ic, nf, fy, fx) and ix <= 256:
            return "nhwc"
        if not a_nhwc and L.bf_conv_c3_supported(ic, iy, ix, nf, fy, fx):
            return "c3"
        return "generic"

    @staticmethod
    def forward(A, F, mode="valid", bias=None, activation="linear", pack=None):
        """`pack` (a FilterPack owned by the calling layer): the channels-last route re-lays F into both packed forms
        with one launch and leaves them there for the data gradient of the same step."""
        a, f = as_device(A), as_device(F)
        im, ic, iy, ix = a.shape
        nf, fc, fy, fx = f.shape
        if mode not in _lib.CONV_MODE:
            raise RuntimeError("Unsupported mode: %s" % (mode,))
        _, oy, ox = ConvolutionOp.outshape(a.shape, f.shape, mode)
        bd = as_device(bias).reshape(-1) if bias is not None else None
        bptr = bd.ptr if bd is not None else _NULL
        route = ConvolutionOp._route(a.nhwc, ic, iy, ix, nf, fy, fx, mode, activation) if fc == ic else "generic"
        if route != "generic":
            out = DeviceArray.empty_nhwc(im, nf, oy, ox)
            try:
                if route == "nhwc":
                    an = a.to_nhwc()      # (a named local: a temporary's block could be reused before the launch)
                    if pack is not None:
                        pack.repack(f)
                        call(L.bf_conv_nhwc_forward_packed, an.pptr, pack.fwd.ptr, bptr, out.pptr, im, ic, iy, ix, nf, fy, fx,
                             _lib.ACT[activation], stream_ptr())
                    else:
                        call(L.bf_conv_nhwc_forward, an.pptr, f.ptr, bptr, out.pptr, im, ic, iy, ix, nf, fy, fx,
                             _lib.ACT[activation], stream_ptr())
                else:
                    call(L.bf_conv_c3_forward, a.ptr, f.ptr, bptr, out.pptr, im, ic, iy, ix, nf, fy, fx,
                         _lib.ACT[activation], stream_ptr())
                return _ret(A, out)
            except _lib.Unsupported:
                pass                             # no tiling for this plane: the generic kernel takes it
        out = DeviceArray.empty(im, nf, max(oy, 0), max(ox, 0))
        ac = a.to_nchw()
        call(L.bf_conv_forward, ac.ptr, f.ptr, bptr, out.ptr, im, ic, iy, ix, nf, fc, fy,
             fx, _lib.CONV_MODE[mode], _lib.ACT[activation], precision(), stream_ptr())
        return _ret(A, out)

    @staticmethod
    def forward_pool(A, F, bias=None, activation="linear", relu_in=False, pack=None):
        """valid convolution (+ bias) -> [ReLU] -> 2x2 max pooling in ONE kernel: returns (pooled, mask) exactly as
        `MaxPoolOp.forward([relu](ConvolutionOp.forward(A, F, bias=..)), 2)` would, without the full-resolution activation
        ever reaching HBM; None when the shape cannot take a fused kernel.  First-layer route (ic <= 4, NCHW input):
        bf_conv_c3_forward_pool; channels-last route (needs the layer's FilterPack): bf_conv_nhwc_forward_pool_packed."""
        a, f = as_device(A), as_device(F)
        im, ic, iy, ix = a.shape
        nf, fc, fy, fx = f.shape
        if fc != ic or activation != "linear":
            return None
        route = ConvolutionOp._route(a.nhwc, ic, iy, ix, nf, fy, fx, "valid", activation)
        oy, ox = iy - fy + 1, ix - fx + 1
        if route == "generic" or oy <= 0 or ox <= 0 or oy % 2 or ox % 2:
            return None
        import torch
        bd = as_device(bias).reshape(-1) if bias is not None else None
        bptr = bd.ptr if bd is not None else _NULL
        if route == "c3":
            if a.nhwc or (128 // ix) < 2:
                return None
            pooled = DeviceArray.empty_nhwc(im, nf, oy // 2, ox // 2)
            words = torch.empty(max(pooled.size, 1), dtype=torch.uint8, device="cuda")
            try:
                call(L.bf_conv_c3_forward_pool, a.ptr, f.ptr, bptr, pooled.pptr, ctypes.c_void_p(words.data_ptr()), im, ic, iy, ix,
                     nf, fy, fx, _lib.ACT[activation], 1 if relu_in else 0, stream_ptr())
            except _lib.Unsupported:
                return None
        else:
            if pack is None or not ConvolutionOp.fuse_pool_nhwc:
                return None
            an = a.to_nhwc()
            pooled = DeviceArray.empty_nhwc(im, nf, oy // 2, ox // 2)
            words = torch.empty(max(pooled.size, 1), dtype=torch.uint8, device="cuda")
            pack.repack(f)
            try:
                call(L.bf_conv_nhwc_forward_pool_packed, an.pptr, pack.fwd.ptr, bptr, pooled.pptr, ctypes.c_void_p(words.data_ptr()),
                     im, ic, iy, ix, nf, fy, fx, 1 if relu_in else 0, stream_ptr())
            except _lib.Unsupported:
                return None
        return pooled, PoolMask(words, (im, nf, oy, ox), 2, nhwc=True)

    @staticmethod
    def backward_filter_unpool(X, dP, mask, gate, fshape, dF, db):
        """Filter gradient of conv -> [relu] -> 2x2 pool taken directly from the pooling layer's delta dP (the
        un-pooling and the ReLU gate are folded into the kernel's operand construction).  Writes dF / db and returns
        True, or returns False when the shape cannot take the fused kernel (caller materialises E instead)."""
        x = as_device(X)
        im, ic, iy, ix = x.shape
        nf, fc, fy, fx = fshape
        if x.nhwc or not mask.nhwc or mask.fdim != 2 or \
                ConvolutionOp._route(False, ic, iy, ix, nf, fy, fx, "valid", "linear") != "c3":
            return False
        dp = as_device(dP).to_nhwc()
        try:
            call(L.bf_conv_c3_backward_filter_unpool, x.ptr, dp.pptr, mask.wptr, gate.pptr if gate is not None else _NULL,
                 dF.ptr, db.ptr, im, ic, iy, ix, nf, fy, fx, stream_ptr(), tag="dF")
        except _lib.Unsupported:
            return False
        return True

    @staticmethod
    def valid(A, F):
        return ConvolutionOp.forward(A, F, "valid")

    @staticmethod
    def full(A, F):
        return ConvolutionOp.forward(A, F, "full")

    @staticmethod
    def backward(X, E, F, dF=None, db=None, need_dX=True, pack=None):
        """-> (dF, db, dX); keyword call like layers/tensor.py:83 (the two reference back-ends disagree on
        the positional order).  db has the NumPy back-end's shape (1,nf,1,1).  dX comes back in
        the storage layout of X where a channels-last kernel exists, NCHW otherwise."""
        x, e, f = as_device(X), as_device(E), as_device(F)
        im, ic, iy, ix = x.shape
        nf, fc, fy, fx = f.shape
        dF = DeviceArray.empty(nf, fc, fy, fx) if dF is None else dF
        db = DeviceArray.empty(1, nf, 1, 1) if db is None else db
        route = ConvolutionOp._route(x.nhwc, ic, iy, ix, nf, fy, fx, "valid", "linear")
        # two calls (filter-gradient half, data-gradient half) so that each kernel can be timed on its own.
        # Layout conversions are bound to locals: a temporary's block could be recycled before the launch.
        en = e.to_nhwc() if route != "generic" else None
        done = False
        if route != "generic":
            try:
                if route == "nhwc":
                    xn = x.to_nhwc()
                    dbp = getattr(e, "db_partial", None) if ConvolutionOp.db_from_pool else None
                    if dbp is not None and en is e:
                        call(L.bf_conv_nhwc_backward_filter_dbp, xn.pptr, en.pptr, dF.ptr, db.ptr, ctypes.c_void_p(dbp[0].data_ptr()),
                             dbp[1], im, ic, iy, ix, nf, fy, fx, stream_ptr(), tag="dF")
                    else:
                        call(L.bf_conv_nhwc_backward_filter, xn.pptr, en.pptr, dF.ptr, db.ptr, im, ic, iy, ix,
                             nf, fy, fx, stream_ptr(), tag="dF")
                else:
                    call(L.bf_conv_c3_backward_filter, x.ptr, en.pptr, dF.ptr, db.ptr, im, ic, iy, ix, nf, fy, fx,
                         stream_ptr(), tag="dF")
                done = True
            except _lib.Unsupported:
                pass
        xc = ec = None
        if not done:
            xc, ec = x.to_nchw(), e.to_nchw()
            call(L.bf_conv_backward, xc.ptr, ec.ptr, f.ptr, dF.ptr, db.ptr, _NULL, im, ic, iy, ix, nf, fy,
                 fx, precision(), stream_ptr(), tag="dF")
        dX = None
        if need_dX:
            if route == "nhwc":
                try:
                    dX = DeviceArray.empty_nhwc(im, ic, iy, ix)
                    if pack is not None and pack.current(f):
                        call(L.bf_conv_nhwc_backward_data_packed, en.pptr, pack.dx.ptr, dX.pptr, im, ic, iy, ix, nf, fy, fx,
                             stream_ptr(), tag="dX")
                    else:
                        call(L.bf_conv_nhwc_backward_data, en.pptr, f.ptr, dX.pptr, im, ic, iy, ix, nf, fy, fx,
                             stream_ptr(), tag="dX")
                except _lib.Unsupported:
                    dX = None
            if dX is None:
                dX = DeviceArray.empty(im, ic, iy, ix)
                xc = x.to_nchw() if xc is None else xc
                ec = e.to_nchw() if ec is None else ec
                call(L.bf_conv_backward, xc.ptr, ec.ptr, f.ptr, _NULL, _NULL, dX.ptr, im, ic, iy, ix, nf,
                     fy, fx, precision(), stream_ptr(), tag="dX")
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
    of fdim*fdim bits per window (window order follows the activation's storage layout).
    `np.asarray(mask)` expands it to the reference's NCHW array."""

    def __init__(self, words, ashape, fdim, nhwc=False):
        self.words, self.ashape, self.fdim, self.nhwc = words, tuple(ashape), fdim, nhwc

    @property
    def shape(self):
        return self.ashape

    @property
    def wptr(self):
        return ctypes.c_void_p(self.words.data_ptr())

    def expand(self):
        im, ic, iy, ix = self.ashape
        if self.nhwc:   # scatter ones through the mask, then the usual layout change
            ones = DeviceArray.empty_nhwc(im, ic, iy // self.fdim, ix // self.fdim).fill(1.0)
            out = DeviceArray.empty_nhwc(im, ic, iy, ix)
            call(L.bf_pool_nhwc_backward, ones.pptr, self.wptr, _NULL, out.pptr, im, ic, iy, ix, self.fdim, stream_ptr())
            return out.to_nchw()
        out = DeviceArray.empty(*self.ashape)
        call(L.bf_pool_mask_expand, self.wptr, out.ptr, im, ic, iy, ix, self.fdim, stream_ptr())
        return out

    def __array__(self, dtype=None, copy=None):
        return self.expand().numpy(np.float64 if dtype is None else dtype)


class MaxPoolOp:
    """atomic/tensor_op.py:78-128 / llatomic/lltensor_op.py:124-152."""

    def __str__(self):
        return "MaxPool"

    @staticmethod
    def forward(A, fdim, relu_in=False):
        """-> (out, filt).  Channels-last inputs stay channels-last.  relu_in (channels-last only) applies
        max(0, .) to A while reading it -- the fused form of `Activation("relu") -> PoolLayer`."""
        import torch
        a = as_device(A)
        im, ic, iy, ix = a.shape
        nbytes = L.bf_pool_mask_bytes(im, ic, iy, ix, fdim)
        words = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda")
        wptr = ctypes.c_void_p(words.data_ptr())
        if a.nhwc:
            out = DeviceArray.empty_nhwc(im, ic, iy // fdim, ix // fdim)
            call(L.bf_pool_nhwc_forward, a.pptr, out.pptr, wptr, im, ic, iy, ix, fdim, 1 if relu_in else 0, stream_ptr())
        else:
            assert not relu_in
            out = DeviceArray.empty(im, ic, iy // fdim, ix // fdim)
            call(L.bf_pool_forward, a.ptr, out.ptr, wptr, im, ic, iy, ix, fdim, stream_ptr())
        mask = PoolMask(words, a.shape, fdim, nhwc=a.nhwc)
        if isinstance(A, DeviceArray):
            return out, mask
        return out.numpy(), np.asarray(mask)

    @staticmethod
    def backward(E, filt, gate=None, want_db=False):
        """E (im,ic,oy,ox), filt = the mask from forward (PoolMask, or a host 0/1 array as the reference
        returns it).  Unlike the reference (tensor_op.py:118-119) the cached mask is not consumed.
        gate (channels-last only): the pooled forward output; multiplies E by (gate > 0).
        want_db (channels-last, device arrays): also leave per-CTA channel sums of the result on it (`.db_partial`) for the
        filter-gradient kernel of the convolution below."""
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
        if mask.nhwc:
            dX = DeviceArray.empty_nhwc(im, ic, iy, ix)
            en = e.to_nhwc()
            rows = L.bf_pool_nhwc_db_rows(im, ic, iy, ix, mask.fdim) if (want_db and isinstance(E, DeviceArray)) else 0
            if rows > 0:
                # the convolution below wants db = dX.sum((0,2,3)): per-CTA channel sums fall out of this pass for free
                part = torch.empty((rows, ic), dtype=torch.float32, device="cuda")
                call(L.bf_pool_nhwc_backward_db, en.pptr, mask.wptr, gate.pptr if gate is not None else _NULL, dX.pptr,
                     ctypes.c_void_p(part.data_ptr()), im, ic, iy, ix, mask.fdim, stream_ptr())
                dX.db_partial = (part, rows)
            else:
                call(L.bf_pool_nhwc_backward, en.pptr, mask.wptr, gate.pptr if gate is not None else _NULL, dX.pptr,
                     im, ic, iy, ix, mask.fdim, stream_ptr())
        else:
            assert gate is None
            dX = DeviceArray.empty(im, ic, iy, ix)
            ec = e.to_nchw()
            call(L.bf_pool_backward, ec.ptr, mask.wptr, dX.ptr, im, ic, iy, ix, mask.fdim, stream_ptr())
        return _ret(E, dX)

    @staticmethod
    def outshape(inshape, fdim):
        if len(inshape) == 3:
            m, iy, ix = inshape
            return m, iy // fdim, ix // fdim
        elif len(inshape) == 2:
            iy, ix = inshape
            return iy // fdim, ix // fdim


# --------------------------------------------------------------------------- plain recurrence
class RecurrentOp:
    """atomic/recurrent_op.py:9-43.  X is time-major (T,B,D), W (D+H,H).  `backward` returns the reference's
    results literally, including `dX = E[:, :, :indim]` (:40): a slice of the scaled error tensor."""

    def __init__(self, activation):
        self.activation = activation
        self.actfn = activations[activation]()

    def forward(self, X, W, b):
        x, w, bd = as_device(X), as_device(W), as_device(b)
        T, B, D = x.shape
        H = w.shape[-1]
        if w.shape[0] != D + H:
            raise ValueError("recurrent weight shape %s does not match D+H=%d" % (w.shape, D + H))
        O = DeviceArray.empty(T, B, H)
        Z = DeviceArray.empty(T, B, D + H)
        call(L.bf_rnn_forward, x.ptr, w.ptr, bd.ptr, O.ptr, Z.ptr, T, B, D, H, _lib.ACT[self.activation], precision(),
             stream_ptr())
        return _ret(X, O, Z)

    def backward(self, Z, O, E, W, nablaW=None, nablab=None):
        z, o, e, w = as_device(Z), as_device(O), as_device(E), as_device(W)
        T, B, ZD = z.shape
        H = w.shape[-1]
        D = ZD - H
        Eout = DeviceArray.empty(T, B, H)
        nablaW = DeviceArray.empty(ZD, H) if nablaW is None else nablaW
        nablab = DeviceArray.empty(H) if nablab is None else nablab
        call(L.bf_rnn_backward, z.ptr, o.ptr, e.ptr, w.ptr, Eout.ptr, nablaW.ptr, nablab.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        dX = Eout if D >= H else DeviceArray(Eout.t[:, :, :D].contiguous())      # recurrent_op.py:40
        return _ret(Z, dX, nablaW, nablab)


# --------------------------------------------------------------------------- Clockwork
class ClockworkOp:
    """The time loops of the reference's ClockworkLayer (layers/recurrent.py:239-277; no atomic op in the reference).
    X is time-major (T,B,D), W (D+H,H), tick (H) the layer's tick_array."""

    def __init__(self, activation):
        self.activation = activation
        self.actfn = activations[activation]()

    def forward(self, X, W, b, tick):
        x, w, bd, tk = as_device(X), as_device(W), as_device(b), as_device(tick)
        T, B, D = x.shape
        H = w.shape[-1]
        if w.shape[0] != D + H:
            raise ValueError("clockwork weight shape %s does not match D+H=%d" % (w.shape, D + H))
        O = DeviceArray.empty(T, B, H)
        Z = DeviceArray.empty(T, B, D + H)
        call(L.bf_clockwork_forward, x.ptr, w.ptr, bd.ptr, tk.ptr, O.ptr, Z.ptr, T, B, D, H, _lib.ACT[self.activation],
             precision(), stream_ptr())
        return _ret(X, O, Z)

    def backward(self, Z, O, E, W, tick, nablaW=None, nablab=None):
        z, o, e, w, tk = as_device(Z), as_device(O), as_device(E), as_device(W), as_device(tick)
        T, B, ZD = z.shape
        H = w.shape[-1]
        D = ZD - H
        dX = DeviceArray.empty(T, B, D)
        nablaW = DeviceArray.empty(ZD, H) if nablaW is None else nablaW
        nablab = DeviceArray.empty(H) if nablab is None else nablab
        call(L.bf_clockwork_backward, z.ptr, o.ptr, e.ptr, w.ptr, tk.ptr, dX.ptr, nablaW.ptr, nablab.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(Z, dX, nablaW, nablab)


# --------------------------------------------------------------------------- GRU
class GRUOp:
    """The time loops of the reference's GRU layer (layers/recurrent.py:123-190; it has no atomic op of its own).
    X is time-major (T,B,D), W (D+H,3H) with columns [U|R|O]."""

    def __init__(self, activation):
        self.activation = activation
        self.actfn = activations[activation]()

    def forward(self, X, W, b):
        x, w, bd = as_device(X), as_device(W), as_device(b)
        T, B, D = x.shape
        H = w.shape[-1] // 3
        if w.shape[0] != D + H:
            raise ValueError("GRU weight shape %s does not match D+H=%d" % (w.shape, D + H))
        hs = DeviceArray.empty(T, B, H)
        Z = DeviceArray.empty(T, B, D + H)
        Zk = DeviceArray.empty(T, B, D + H)
        gates = DeviceArray.empty(3, T, B, H)
        call(L.bf_gru_forward, x.ptr, w.ptr, bd.ptr, hs.ptr, Z.ptr, Zk.ptr, gates.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(X, hs, Z, Zk, gates)

    def backward(self, Z, Zk, gates, E, W, nablaW=None, nablab=None):
        z, zk, g, e, w = as_device(Z), as_device(Zk), as_device(gates), as_device(E), as_device(W)
        T, B, ZD = z.shape
        H = w.shape[-1] // 3
        D = ZD - H
        dX = DeviceArray.empty(T, B, D)
        nablaW = DeviceArray.empty(ZD, 3 * H) if nablaW is None else nablaW
        nablab = DeviceArray.empty(3 * H) if nablab is None else nablab
        call(L.bf_gru_backward, z.ptr, zk.ptr, g.ptr, e.ptr, w.ptr, dX.ptr, nablaW.ptr, nablab.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(Z, dX, nablaW, nablab)


# --------------------------------------------------------------------------- LSTM
class LSTMOp:
    """atomic/recurrent_op.py:46-112.  X is time-major (T,B,D); gate order along 4H is [cand|f|i|o]."""

    def __init__(self, activation):
        self.activation = activation
        self.actfn = activations[activation]()

    def forward(self, X, W, b, batch_major=False):
        """`batch_major=True`: X is (B,T,D) as the layer receives it (the transpose is folded into the kernel)."""
        x, w, bd = as_device(X), as_device(W), as_device(b)
        if batch_major:
            B, T, D = x.shape
        else:
            T, B, D = x.shape
        H = w.shape[-1] // 4
        if w.shape[0] != D + H:
            raise ValueError("LSTM weight shape %s does not match D+H=%d" % (w.shape, D + H))
        O = DeviceArray.empty(T, B, H)
        Z = DeviceArray.empty(T, B, D + H)
        cache = DeviceArray.empty(6, T, B, H)
        cache.bw = DeviceArray.empty(5, T, B, H)   # coefficients of the backward recurrence, from the pre-activations
        call(L.bf_lstm_forward_bm if batch_major else L.bf_lstm_forward, x.ptr, w.ptr, bd.ptr, O.ptr, Z.ptr, cache.ptr,
             cache.bw.ptr, T, B, D, H, _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(X, O, Z, cache)

    def backward(self, Z, O, E, W, cache, nablaW=None, nablab=None, need_dX=True):
        z, o, e, w, c = as_device(Z), as_device(O), as_device(E), as_device(W), as_device(cache)
        bw = getattr(cache, "bw", None)
        if bw is None and not isinstance(cache, DeviceArray):
            # host float64 cache: form the five coefficients of the backward recurrence from the reference's bwCa /
            # bwgates (recurrent_op.py:85-88) before narrowing: o*act'(Ca), i*act'(cand), C[t-1]*f', cand*i', Ca*o'
            ch = np.asarray(cache, dtype=np.float64)
            def d(kind, a):
                return {"sigmoid": a * (1. - a), "tanh": 1. - a ** 2, "relu": (a > 0.).astype(np.float64),
                        "linear": np.ones_like(a)}[kind]
            Cm, Ca, cand, fg, ig, og = ch
            Cprev = np.concatenate([np.zeros_like(Cm[:1]), Cm[:-1]])
            bw = DeviceArray.from_host(np.stack([og * d(self.activation, Ca), ig * d(self.activation, cand),
                                                 Cprev * d("sigmoid", fg), cand * d("sigmoid", ig), Ca * d("sigmoid", og)]))
        T, B, ZD = z.shape
        H = w.shape[-1] // 4
        D = ZD - H
        dX = DeviceArray.empty(T, B, D) if need_dX else None
        nablaW = DeviceArray.empty(ZD, 4 * H) if nablaW is None else nablaW
        nablab = DeviceArray.empty(4 * H) if nablab is None else nablab
        call(L.bf_lstm_backward, z.ptr, o.ptr, e.ptr, w.ptr, c.ptr, bw.ptr if bw is not None else _NULL,
             dX.ptr if dX is not None else _NULL,
             nablaW.ptr, nablab.ptr, T, B, D, H,
             _lib.ACT[self.activation], precision(), stream_ptr())
        return _ret(Z, dX, nablaW, nablab)


# --------------------------------------------------------------------------- operand handling on the TF32 path
_TMA, _RNA = ("tma", "tma"), ("rna", "rna")


def tf32_operand_modes(kind, learn=True, **s):
    """How the kernel family that `kind` + shape selects on the TF32 path hands its two operands to
    `tcgen05.mma.kind::tf32`: {site: (mode_A, mode_B) | None}.  "tma": staged by a tensor map of type TFLOAT32 (the TMA
    unit rounds FP32 to TF32, to nearest, while copying); "rna": rounded to nearest, ties away, by the SIMT operand
    builder (cvt.rna.tf32.f32); None: the site runs FP32 FMAs (no tensor core).  No operand reaches an MMA unrounded.  This is the documentation of record for the stated TF32 bounds
    (DESIGN.md 4); tests/test_gpu_tf32_emulated.py feeds it to the emulating oracle.  `learn`: inside learn_batch
    (first-layer filter gradient un-pools in-kernel) rather than op by op."""
    if precision() != _lib.PREC["tf32"]:
        return {}
    if kind == "dense":
        if s["n"] <= 32:
            return {"dense.fwd": None, "dense.dW": None, "dense.dX": None}
        return {site: (_TMA if L.bf_dense_tma_supported(s["m"], s["k"], s["n"], i) else _RNA)
                for i, site in enumerate(("dense.fwd", "dense.dW", "dense.dX"))}
    if kind == "conv":
        route = ConvolutionOp._route(s.get("nhwc", False), s["ic"], s["iy"], s["ix"], s["nf"], s["fy"], s["fx"], "valid", "linear")
        if route == "nhwc":
            # db: through an all-ones operand of the dF MMA (sees E as the TMA unit rounds it) unless the pooling layer
            # above summed its own output in FP32 (`pooled`: learn_batch / backpropagate through a PoolLayer(2))
            return {"conv.fwd": _TMA, "conv.dF": _TMA, "conv.dX": _TMA,
                    "conv.db": None if (s.get("pooled", False) and ConvolutionOp.db_from_pool) else "tma"}
        if route == "c3":
            # forward: pixel stream and filter rounded while they are re-laid; filter gradient: patch rows rounded by the
            # builder warps, E either rebuilt (rounded) by them from the pooling layer's delta (learn_batch) or a TMA
            # box; the data gradient of a first layer runs on the NCHW gather kernel
            return {"conv.fwd": _RNA, "conv.dF": _RNA if (learn and s.get("unpool", False)) else ("rna", "tma"), "conv.dX": _RNA}
        return {"conv.fwd": _RNA, "conv.dF": _RNA, "conv.dX": _RNA}
    if kind == "lstm":
        zd = s["D"] + s["H"]
        tma = zd % 4 == 0
        return {"lstm.fwd": _TMA if tma else _RNA, "lstm.dZ": _TMA, "lstm.dW": _TMA if tma else _RNA}
    raise ValueError(kind)


def tf32_policy_for(net, learn=True):
    """The per-layer form of `tf32_operand_modes` for a connected network: a callable (site, layer_index) ->
    (mode_A, mode_B) | None with layer_index counting the layers behind the input layer from 0."""
    from .layers.tensor import ConvLayer
    table = {}
    layers = net.layers.layers[1:]
    for i, layer in enumerate(layers):
        name = type(layer).__name__
        if name == "Dense":
            table[i] = tf32_operand_modes("dense", m=1 << 20, k=layer.weights.shape[0], n=layer.neurons)
        elif isinstance(layer, ConvLayer):
            ic, iy, ix = layer.inshape[-3:]
            prev_cl = i > 0 and any(isinstance(p, ConvLayer) for p in layers[:i])
            pool, _ = layer._pool_after()
            oy, ox = iy - layer.fy + 1, ix - layer.fx + 1
            pooled = pool is not None and bool(L.bf_pool_nhwc_db_rows(1, layer.nfilters, oy, ox, 2))
            table[i] = tf32_operand_modes("conv", learn=learn, ic=ic, iy=iy, ix=ix, nf=layer.nfilters, fy=layer.fy, fx=layer.fx,
                                          nhwc=prev_cl, unpool=pool is not None and i == 0 and layer.fuse_pool, pooled=pooled)
        elif name == "LSTM":
            table[i] = tf32_operand_modes("lstm", D=layer.inshape[-1], H=layer.neurons)
    def policy(site, li):
        return table.get(li, {}).get(site)
    policy.declares = lambda site, li: site in table.get(li, {})
    return policy
