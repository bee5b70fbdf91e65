import sys, time, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import brainforge_b200 as bf
from brainforge_b200 import _lib
from brainforge_b200.device import DeviceArray, call, stream_ptr
from oracle import bf_oracle as O
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
shape = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (2, 8, 8, 8, 8, 3, 3)
im, ic, iy, ix, nf, fy, fx = shape
rs = np.random.RandomState(0)
A = rs.randn(im, ic, iy, ix); F = rs.randn(nf, ic, fy, fx) * 0.3
yv = O.conv_forward(A, F, "valid"); E = rs.randn(*yv.shape)
rF, rb, rX = O.conv_backward(A, E, F)
e = lambda a, b: float(np.linalg.norm(np.asarray(a) - b) / np.linalg.norm(b))
Ad, Fd, Ed = DeviceArray.from_host(A), DeviceArray.from_host(F), DeviceArray.from_host(E)
NULL = ctypes.c_void_p(0)
bf.set_precision("fp32")
bf.atomic.ConvolutionOp().forward(Ad, Fd, "valid"); torch.cuda.synchronize()     # warm-up: context, workspace
np.set_printoptions(precision=3, linewidth=220, suppress=True)
t = time.time()
try:
    if which == "fwd":
        y = DeviceArray.empty(*yv.shape)
        call(_lib.lib.bf_conv_forward, Ad.ptr, Fd.ptr, NULL, y.ptr, im, ic, iy, ix, nf, ic, fy, fx, 0, 0, 1, stream_ptr())
        torch.cuda.synchronize()
        print("path", _lib.lib.bf_last_path(), "fwd relerr", e(y, yv), "t=%.3f" % (time.time() - t))
        if "-v" in sys.argv: print(np.asarray(y)[0, 0]); print(yv[0, 0])
    elif which == "dF":
        dF, db = DeviceArray.empty(*F.shape), DeviceArray.empty(nf)
        call(_lib.lib.bf_conv_backward, Ad.ptr, Ed.ptr, Fd.ptr, dF.ptr, db.ptr, NULL, im, ic, iy, ix, nf, fy, fx, 1, stream_ptr())
        torch.cuda.synchronize()
        print("path", _lib.lib.bf_last_path(), "dF", e(dF, rF), "db", e(db, rb.ravel()), "t=%.3f" % (time.time() - t))
        if "-v" in sys.argv: print(np.asarray(dF)[0, 0], rF[0, 0]); print(np.asarray(db), rb.ravel())
    else:
        dX = DeviceArray.empty(*A.shape)
        call(_lib.lib.bf_conv_backward, Ad.ptr, Ed.ptr, Fd.ptr, NULL, NULL, dX.ptr, im, ic, iy, ix, nf, fy, fx, 1, stream_ptr())
        torch.cuda.synchronize()
        print("path", _lib.lib.bf_last_path(), "dX", e(dX, rX), "t=%.3f" % (time.time() - t))
        if "-v" in sys.argv: print(np.asarray(dX)[0, 0]); print(rX[0, 0])
except Exception as ex:
    print(which, "FAILED after %.3f s:" % (time.time() - t), repr(ex)[:200])
