"""GPU timing driver: first-layer filter gradient with in-kernel un-pooling at the cfg3 shape (batch 4096)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import brainforge_b200 as bf
bf.set_precision("tf32")
op = bf.atomic.ConvolutionOp()
B = 4096
rs = np.random.RandomState(0)
a = bf.DeviceArray(torch.randn(B, 3, 30, 30, device="cuda")); f = bf.DeviceArray(torch.randn(32, 3, 3, 3, device="cuda") * 0.1)
b = bf.DeviceArray(torch.zeros(32, device="cuda"))
pooled, mask = op.forward_pool(a, f, bias=b, relu_in=True)
dP = bf.DeviceArray(torch.randn(B, 14, 14, 32, device="cuda"), nhwc=True)
dF, db = bf.DeviceArray.empty(32, 3, 3, 3), bf.DeviceArray.empty(1, 32, 1, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op.backward_filter_unpool(a, dP, mask, pooled, (32, 3, 3, 3), dF, db); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("c3 dF unpool: median %.1f us" % (np.median(ts[2:]) * 1e3))
