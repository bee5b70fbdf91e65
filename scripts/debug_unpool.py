"""GPU debug driver: fused un-pool filter gradient vs the materialised route at growing batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import brainforge_b200 as bf
bf.set_precision("tf32")
op, pool = bf.atomic.ConvolutionOp(), bf.atomic.MaxPoolOp()
def rel(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
for im in (37, 38, 74, 75, 130, 300):
    rs = np.random.RandomState(im)
    A, F, bias = rs.randn(im, 3, 30, 30), rs.randn(32, 3, 3, 3) * 0.3, rs.randn(32) * 0.1
    a, f, b = bf.DeviceArray.from_host(A), bf.DeviceArray.from_host(F), bf.DeviceArray.from_host(bias)
    pooled, mask = op.forward_pool(a, f, bias=b, relu_in=True)
    dP = bf.DeviceArray.from_host(rs.randn(*pooled.shape)).to_nhwc()
    dF, db = bf.DeviceArray.empty(*F.shape), bf.DeviceArray.empty(1, 32, 1, 1)
    ok = op.backward_filter_unpool(a, dP, mask, pooled, F.shape, dF, db)
    E = pool.backward(dP, mask, gate=pooled)
    dF2, db2, _ = op.backward(X=a, E=E, F=f, need_dX=False)
    # per-image contribution check: which images are wrong?  compare against single-image sums on the materialised route
    print(im, ok, "dF", rel(dF, dF2), "db", rel(db, db2), flush=True)
