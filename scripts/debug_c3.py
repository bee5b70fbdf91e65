"""GPU debug driver for the first-layer (few-channel) conv kernels."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import brainforge_b200 as bf
from brainforge_b200 import _lib
from brainforge_b200.device import call, stream_ptr
from oracle import bf_oracle as O
L = _lib.lib
def dev(a): return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
def P(t): return ctypes.c_void_p(t.data_ptr())
def err(a, b): return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-30))
shapes = [(2, 3, 30, 30, 32, 3, 3), (5, 1, 28, 28, 32, 3, 3), (3, 4, 17, 40, 64, 2, 3), (37, 3, 30, 30, 32, 3, 3), (2, 2, 12, 50, 128, 3, 2)]
for (im, ic, iy, ix, nf, fy, fx) in shapes:
    rs = np.random.RandomState(1)
    A = rs.randn(im, ic, iy, ix); F = rs.randn(nf, ic, fy, fx) * 0.2; b = rs.randn(nf)
    Y = O.conv_forward(A, F, "valid"); E = rs.randn(*Y.shape)
    rF, rb, rX = O.conv_backward(A, E, F)
    oy, ox = Y.shape[2:]
    tag = "%s" % ((im, ic, iy, ix, nf, fy, fx),)
    if not L.bf_conv_c3_supported(ic, iy, ix, nf, fy, fx):
        print(tag, "unsupported"); continue
    Ad, Fd, bd, Ed = dev(A), dev(F), dev(b), dev(E.transpose(0, 2, 3, 1))
    try:
        yd = torch.full((im, oy, ox, nf), float("nan"), device="cuda")
        call(L.bf_conv_c3_forward, P(Ad), P(Fd), P(bd), P(yd), im, ic, iy, ix, nf, fy, fx, 0, stream_ptr())
        torch.cuda.synchronize()
        print(tag, "fwd", err(yd.cpu().numpy().transpose(0, 3, 1, 2), Y + b[None, :, None, None]), flush=True)
        fd = torch.full((nf, ic, fy, fx), float("nan"), device="cuda"); dbd = torch.full((nf,), float("nan"), device="cuda")
        call(L.bf_conv_c3_backward_filter, P(Ad), P(Ed), P(fd), P(dbd), im, ic, iy, ix, nf, fy, fx, stream_ptr())
        torch.cuda.synchronize()
        print(tag, "dF ", err(fd.cpu().numpy(), rF), "db", err(dbd.cpu().numpy(), rb.ravel()), flush=True)
    except Exception as ex:
        print(tag, "FAILED:", repr(ex)[:300], flush=True)
        break
# timing at the cfg3 first layer
B = 4096
X = torch.randn(B, 3, 30, 30, device="cuda"); F = torch.randn(32, 3, 3, 3, device="cuda") * 0.1; bias = torch.zeros(32, device="cuda")
Y = torch.empty(B, 28, 28, 32, device="cuda"); E = torch.randn(B, 28, 28, 32, device="cuda")
dF = torch.empty_like(F); db = torch.empty(32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("fwd", lambda: call(L.bf_conv_c3_forward, P(X), P(F), P(bias), P(Y), B, 3, 30, 30, 32, 3, 3, 0, stream_ptr())),
                 ("dF ", lambda: call(L.bf_conv_c3_backward_filter, P(X), P(E), P(dF), P(db), B, 3, 30, 30, 32, 3, 3, stream_ptr()))):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    med = float(np.median(ts))
    print("c3 30x30x3->32 %s: %.1f us  (%.0f GB/s of the 455 MB algorithmic)" % (name, med * 1e3, 455.0 / med), flush=True)
