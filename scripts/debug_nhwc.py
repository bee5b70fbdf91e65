"""GPU debug driver for the NHWC conv entry points: compares with the float64 oracle on small and cfg3 shapes."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import brainforge_b200 as bf
from brainforge_b200 import _lib
from brainforge_b200.device import call, stream_ptr
from oracle import bf_oracle as O
L = _lib.lib
NULL = ctypes.c_void_p(0)
def dev(a): return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
def P(t): return ctypes.c_void_p(t.data_ptr())
def err(a, b): return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-30))
shapes = [(2, 32, 14, 14, 64, 3, 3), (9, 64, 6, 6, 128, 3, 3), (3, 32, 8, 8, 32, 3, 3), (2, 32, 40, 40, 64, 3, 3),
          (5, 128, 7, 9, 64, 2, 3), (300, 32, 14, 14, 64, 3, 3)]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if len(sys.argv) > 2: shapes = [tuple(int(v) for v in sys.argv[2].split(","))]
for (im, ic, iy, ix, nf, fy, fx) in shapes:
    rs = np.random.RandomState(0)
    A = rs.randn(im, ic, iy, ix); F = rs.randn(nf, ic, fy, fx) * 0.2; b = rs.randn(nf)
    Y = O.conv_forward(A, F, "valid"); E = rs.randn(*Y.shape)
    rF, rb, rX = O.conv_backward(A, E, F)
    oy, ox = Y.shape[2:]
    Ad, Fd, bd, Ed = dev(A.transpose(0, 2, 3, 1)), dev(F), dev(b), dev(E.transpose(0, 2, 3, 1))
    tag = "%s" % ((im, ic, iy, ix, nf, fy, fx),)
    try:
        if which in ("all", "fwd"):
            yd = torch.full((im, oy, ox, nf), float("nan"), device="cuda")
            call(L.bf_conv_nhwc_forward, P(Ad), P(Fd), P(bd), P(yd), im, ic, iy, ix, nf, fy, fx, 0, stream_ptr())
            torch.cuda.synchronize()
            print(tag, "fwd", err(yd.cpu().numpy().transpose(0, 3, 1, 2), Y + b[None, :, None, None]), flush=True)
        if which in ("all", "dX"):
            xd = torch.full((im, iy, ix, ic), float("nan"), device="cuda")
            call(L.bf_conv_nhwc_backward_data, P(Ed), P(Fd), P(xd), im, ic, iy, ix, nf, fy, fx, stream_ptr())
            torch.cuda.synchronize()
            print(tag, "dX ", err(xd.cpu().numpy().transpose(0, 3, 1, 2), rX), flush=True)
        if which in ("all", "dF"):
            fd = torch.full((nf, ic, fy, fx), float("nan"), device="cuda"); dbd = torch.full((nf,), float("nan"), device="cuda")
            call(L.bf_conv_nhwc_backward_filter, P(Ad), P(Ed), P(fd), P(dbd), im, ic, iy, ix, nf, fy, fx, stream_ptr())
            torch.cuda.synchronize()
            print(tag, "dF ", err(fd.cpu().numpy(), rF), "db", err(dbd.cpu().numpy(), rb.ravel()), flush=True)
    except Exception as ex:
        print(tag, "FAILED:", repr(ex)[:300], flush=True)
        break
