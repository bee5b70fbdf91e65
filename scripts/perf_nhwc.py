"""Times the NHWC conv entry points at the cfg3 layer shapes (batch 4096) with CUDA events."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import brainforge_b200 as bf
from brainforge_b200 import _lib
from brainforge_b200.device import call, stream_ptr
L = _lib.lib
def P(t): return ctypes.c_void_p(t.data_ptr())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NIT = int(sys.argv[2]) if len(sys.argv) > 2 else 10
def timeit(fn, n=NIT):
    for _ in range(1 if NIT < 3 else 3): fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))
for (ic, iy, ix, nf) in [(32, 14, 14, 64), (64, 6, 6, 128)]:
    oy, ox = iy - 2, ix - 2
    X = torch.randn(B, iy, ix, ic, device="cuda"); F = torch.randn(nf, ic, 3, 3, device="cuda") * 0.1
    E = torch.randn(B, oy, ox, nf, device="cuda"); Y = torch.empty(B, oy, ox, nf, device="cuda")
    dX = torch.empty_like(X); dF = torch.empty_like(F); db = torch.empty(nf, device="cuda"); bias = torch.zeros(nf, device="cuda")
    flops = 2.0 * B * oy * ox * nf * ic * 9
    byt = 4.0 * (X.numel() + Y.numel())
    for name, fn in (("fwd", lambda: call(L.bf_conv_nhwc_forward, P(X), P(F), P(bias), P(Y), B, ic, iy, ix, nf, 3, 3, 0, stream_ptr())),
                     ("dX ", lambda: call(L.bf_conv_nhwc_backward_data, P(E), P(F), P(dX), B, ic, iy, ix, nf, 3, 3, stream_ptr())),
                     ("dF ", lambda: call(L.bf_conv_nhwc_backward_filter, P(X), P(E), P(dF), P(db), B, ic, iy, ix, nf, 3, 3, stream_ptr()))):
        med, mn = timeit(fn)
        print("conv %dx%dx%d->%d %s: median %.1f us (min %.1f)  %.1f TFLOP/s  %.0f GB/s algorithmic" %
              (iy, ix, ic, nf, name, med * 1e3, mn * 1e3, flops / med / 1e9, byt / med / 1e6), flush=True)
