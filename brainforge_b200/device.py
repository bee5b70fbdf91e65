"""Device plumbing: FP32 arrays resident in HBM behind a NumPy-like facade.

PyTorch is used here for device memory (caching allocator), the current CUDA stream, CUDA graphs and
`torch.distributed` -- plumbing only.  Every arithmetic kernel is in libbf_b200.so and is reached through
`call()`; there is no torch op and no CPU path behind any layer.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_state = {"workspace": None, "precision": _lib.PREC["fp32"], "ws_generation": 0, "ws_retired": [], "mutations": 0}


def note_mutation():
    """Called by everything that writes parameter memory from the host API (uploads, fills, optimizer steps): lets a layer
    tell whether operands it derived from its weights during the forward pass (packed filters) are still current."""
    _state["mutations"] += 1


def mutations():
    return _state["mutations"]
DEFAULT_WORKSPACE = 64 << 20


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("brainforge_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback. "
                           "The float64 CPU oracle lives under oracle/ and is test infrastructure only.")


UPLOAD_POOL_MIN = 1 << 16      # elements: smaller pageable arrays take the plain copy (+ the device-side narrowing)


try:                                    # raw handle of the current stream without building a torch.cuda.Stream object
    _raw_stream, _cur_device = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice      # (14 us -> <1 us per call:
except AttributeError:                  #  three calls per training step were a quarter of a tiny-batch step's host time)
    _raw_stream = _cur_device = None


def stream_ptr():
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(_cur_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_precision(name):
    """'fp32' (SIMT FMA, <=1e-5 vs the float64 reference) or 'tf32' (tcgen05 tensor cores)."""
    _state["precision"] = _lib.PREC[name]


def precision():
    return _state["precision"]


def precision_name():
    return {v: k for k, v in _lib.PREC.items()}[_state["precision"]]


def _ensure_workspace(nbytes):
    ws = _state["workspace"]
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("workspace must be sized before CUDA-graph capture (run the step once eagerly)")
        nbytes = max(int(nbytes), DEFAULT_WORKSPACE)
        torch.cuda.synchronize()
        if ws is not None:
            # captured CUDA graphs have the old block's address baked into their kernels: it must stay allocated (and
            # unused by anybody else) until those graphs are gone.  `workspace_generation()` tells graph owners to
            # re-capture; they call `release_retired_workspaces()` once none of their graphs is older than that.
            _state["ws_retired"].append(ws)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        _state["workspace"] = ws
        _state["ws_generation"] += 1
        _lib.check(_lib.lib.bf_set_workspace(ctypes.c_void_p(ws.data_ptr()), ws.numel()))
    return ws


def workspace_generation():
    """Bumped every time the scratch workspace is re-allocated: a CUDA graph captured under an older generation
    references a retired block and has to be captured again before its next replay."""
    return _state["ws_generation"]


def call(fn, *args, tag=None):
    """Invoke a libbf_b200 entry point; grows the scratch workspace and retries once if asked to.
    With timing enabled (bench.py's roofline pass) the call is bracketed by CUDA events on the launching
    stream and filed under `name(int args)[tag]`."""
    if _state["workspace"] is None:
        _ensure_workspace(DEFAULT_WORKSPACE)
    timing = _state.get("timing")
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    try:
        _lib.check(fn(*args))
    except _lib.WorkspaceTooSmall as e:
        _ensure_workspace(int(e.need * 1.25) + 4096)
        _lib.check(fn(*args))
    if timing is not None:
        e1.record()
        label = "%s(%s)%s" % (fn.__name__, ",".join(str(a) for a in args if isinstance(a, int)),
                              "[%s]" % tag if tag else "")
        timing.setdefault(label, []).append((e0, e1))


def start_timing():
    _state["timing"] = {}


def stop_timing():
    """-> {label: [ms per call]} for every library call made since start_timing()."""
    torch.cuda.synchronize()
    timing, _state["timing"] = _state.get("timing") or {}, None
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in timing.items()}


def launch_count():
    return int(_lib.lib.bf_launch_count())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class DeviceArray:
    """A C-contiguous FP32 tensor in HBM with the handful of ndarray attributes the reference's callers
    touch on layer state (`.shape`, `.size`, `.ndim`, `*= 0`, `np.asarray(..)`): SURVEY 7.2#7.

    `nhwc=True` marks a 4-D activation tensor whose LOGICAL shape is the reference's (im, c, y, x) but whose
    storage is channels-last (im, y, x, c) -- what the TMA + tcgen05 convolutions read and write.  `.shape`,
    `.numpy()`, `reshape`, indexing always speak the logical NCHW order (converting on the device when they
    must); only layout-aware code (ConvolutionOp, MaxPoolOp, elementwise ops) touches the physical buffer
    through `.pptr`.  `.ptr` refuses channels-last arrays so that a layout-unaware op cannot misread one."""

    __array_priority__ = 100

    def __init__(self, tensor, nhwc=False, cl_flat=None):
        assert tensor.dtype == torch.float32 and tensor.is_cuda and tensor.is_contiguous(), \
            (tensor.dtype, tensor.device, tensor.is_contiguous())
        assert not nhwc or tensor.dim() == 4
        self.t = tensor
        self.nhwc = nhwc
        # (c, y, x) when this 2-D array is the FLATTENED view of a channels-last activation: features are stored in
        # (y, x, c) order, the logical (reference) order is (c, y, x).  Only the skinny Dense kernels read it as is.
        self.cl_flat = cl_flat
        assert cl_flat is None or (tensor.dim() == 2 and not nhwc)

    # ---- construction
    @staticmethod
    def empty(*shape):
        require_cuda()
        return DeviceArray(torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device="cuda"))

    @staticmethod
    def empty_nhwc(im, c, y, x):
        """Channels-last storage for the logical shape (im, c, y, x)."""
        require_cuda()
        return DeviceArray(torch.empty((int(im), int(y), int(x), int(c)), dtype=torch.float32, device="cuda"), nhwc=True)

    @staticmethod
    def empty_like(other):
        return DeviceArray(torch.empty_like(other.t), nhwc=other.nhwc, cl_flat=other.cl_flat)

    @staticmethod
    def zeros(*shape):
        require_cuda()
        return DeviceArray(torch.zeros(tuple(int(s) for s in shape), dtype=torch.float32, device="cuda"))

    @staticmethod
    def from_host(a):
        """H2D.  float32 C-contiguous arrays (ideally pinned) are copied as they are; anything else is
        shipped as float64 and narrowed on the device (bf_cast_f64_to_f32)."""
        out = DeviceArray.empty(*np.shape(a))
        out.copy_from_host(a)
        return out

    def copy_from_host(self, a):
        assert not self.nhwc
        note_mutation()
        a = np.asarray(a)
        if tuple(a.shape) != self.shape:
            a = a.reshape(self.shape)
        if a.size == 0:
            return self
        if a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]:
            pinned = _is_pinned(a)
            if pinned or a.size < UPLOAD_POOL_MIN or torch.cuda.is_current_stream_capturing():
                call(_lib.lib.bf_memcpy_h2d, self.ptr, ctypes.c_void_p(a.ctypes.data), a.nbytes, stream_ptr())
                if not pinned:
                    call(_lib.lib.bf_stream_sync, stream_ptr())  # pageable source may be reused by the caller
            else:
                # pageable: host threads copy it into pinned staging chunks (the driver's own staging runs on one thread)
                call(_lib.lib.bf_upload_host, self.ptr, ctypes.c_void_p(a.ctypes.data), a.size, 0, stream_ptr())
        else:
            a64 = np.ascontiguousarray(a, dtype=np.float64)
            if a64.size >= UPLOAD_POOL_MIN and not torch.cuda.is_current_stream_capturing():
                # the reference API's native arrays: narrowed on the host by a thread pool into pinned staging chunks, half
                # the bytes over the host link (the values equal the device cast: round to nearest even)
                call(_lib.lib.bf_upload_host, self.ptr, ctypes.c_void_p(a64.ctypes.data), a64.size, 1, stream_ptr())
            else:
                stage = torch.empty(a64.size, dtype=torch.float64, device="cuda")
                call(_lib.lib.bf_memcpy_h2d, _ptr(stage), ctypes.c_void_p(a64.ctypes.data), a64.nbytes, stream_ptr())
                call(_lib.lib.bf_cast_f64_to_f32, _ptr(stage), self.ptr, a64.size, stream_ptr())
                call(_lib.lib.bf_stream_sync, stream_ptr())
        return self

    # ---- layout
    def to_nchw(self):
        """The reference's index order; a device transpose when the storage is channels-last."""
        if self.cl_flat is not None:
            c, y, x = self.cl_flat
            m = self.t.shape[0]
            return DeviceArray(DeviceArray(self.t.view(m, y, x, c), nhwc=True).to_nchw().t.view(m, c * y * x))
        if not self.nhwc:
            return self
        im, y, x, c = self.t.shape
        out = DeviceArray.empty(im, c, y, x)
        if self.size:
            call(_lib.lib.bf_layout_nhwc_to_nchw, self.pptr, out.pptr, im, c, y * x, stream_ptr())
        return out

    def to_nhwc(self):
        if self.nhwc:
            return self
        assert self.cl_flat is None
        im, c, y, x = self.t.shape
        out = DeviceArray.empty_nhwc(im, c, y, x)
        if self.size:
            call(_lib.lib.bf_layout_nchw_to_nhwc, self.pptr, out.pptr, im, c, y * x, stream_ptr())
        return out

    def like_layout(self, other):
        """This array in the storage layout of `other`."""
        return self.to_nhwc() if other.nhwc else self.to_nchw()

    # ---- ndarray-like surface
    @property
    def pptr(self):
        """Pointer to the physical buffer, whatever its layout (layout-aware callers only)."""
        return ctypes.c_void_p(self.t.data_ptr())

    @property
    def ptr(self):
        if self.nhwc or self.cl_flat is not None:
            raise RuntimeError("layout-unaware access to a channels-last DeviceArray (use .to_nchw() or .pptr)")
        return ctypes.c_void_p(self.t.data_ptr())

    @property
    def shape(self):
        if self.nhwc:
            im, y, x, c = self.t.shape
            return (im, c, y, x)
        return tuple(self.t.shape)

    @property
    def size(self):
        return self.t.numel()

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def __len__(self):
        return self.t.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return DeviceArray(self.to_nchw().t.view(*[int(s) for s in shape]))

    def ravel(self):
        return self.reshape(-1)

    def __getitem__(self, idx):
        v = self.to_nchw().t[idx]
        if not v.is_contiguous():
            raise IndexError("DeviceArray only supports slices that stay contiguous")
        return DeviceArray(v)

    def numpy(self, dtype=np.float64):
        """D2H copy (synchronises the stream); float64 by default like the reference's arrays."""
        src = self.to_nchw()
        host = np.empty(src.shape, dtype=np.float32)
        if host.size:
            call(_lib.lib.bf_memcpy_d2h, ctypes.c_void_p(host.ctypes.data), src.ptr, host.nbytes, stream_ptr())
            call(_lib.lib.bf_stream_sync, stream_ptr())
        return host.astype(dtype, copy=False)

    def __array__(self, dtype=None, copy=None):
        return self.numpy(np.float64 if dtype is None else dtype)

    def copy(self):
        out = DeviceArray.empty_like(self)
        if self.size:
            call(_lib.lib.bf_act_forward, _lib.ACT["linear"], self.pptr, out.pptr, self.size, stream_ptr())
        return out

    def fill(self, value):
        """Every element = value, without reading the old contents (a NaN / Inf already there does not survive)."""
        note_mutation()
        if self.size:
            call(_lib.lib.bf_fill, self.pptr, self.size, float(value), stream_ptr())
        return self

    def fill_zero(self):
        return self.fill(0.0)

    def __imul__(self, scalar):
        # `layer.nabla_w *= 0` (learner/backpropagation.py:72-73) and friends
        s = float(scalar)
        if s == 0.0:
            return self.fill_zero()      # the reference's idiom for zero_gradients: must also clear non-finite values
        note_mutation()
        if self.size:
            call(_lib.lib.bf_affine, self.pptr, self.pptr, self.size, s, 0.0, stream_ptr())
        return self

    def __repr__(self):
        return "DeviceArray(shape=%s, fp32 @ cuda%s)" % (self.shape, ", channels-last" if self.nhwc else "")


_pinned_ranges = []


def pinned_empty(shape, dtype=np.float32):
    """A NumPy array backed by page-locked host memory (fast, truly asynchronous H2D source)."""
    tdtype = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
    t = torch.empty(tuple(shape), dtype=tdtype, pin_memory=True)
    a = t.numpy()
    _pinned_ranges.append((a.ctypes.data, a.ctypes.data + a.nbytes, t))
    return a


def _is_pinned(a):
    p = a.ctypes.data
    return any(lo <= p < hi for lo, hi, _ in _pinned_ranges)


def as_device(x):
    """numpy / DeviceArray -> DeviceArray (no copy when already on the device)."""
    if isinstance(x, DeviceArray):
        return x
    return DeviceArray.from_host(x)


def synchronize():
    call(_lib.lib.bf_stream_sync, stream_ptr())
