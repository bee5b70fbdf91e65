"""GPU: per-role wait cycles (BF_PROF_IGEMM=1) of the channels-last igemm kernels at the cfg3 layer shapes, batch 4096."""
import os, sys
os.environ["BF_PROF_IGEMM"] = "1"
os.environ["BF_DEBUG_PLAN"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import brainforge_b200 as bf
from brainforge_b200.device import DeviceArray
bf.set_precision("tf32")
op = bf.atomic.ConvolutionOp()
B = 4096
rs = np.random.RandomState(0)
for (ic, iy, nf) in ((32, 14, 64), (64, 6, 128)):
    x = DeviceArray.from_host(rs.randn(B, ic, iy, iy).astype(np.float32)).to_nhwc()
    f = DeviceArray.from_host((rs.randn(nf, ic, 3, 3) * 0.1).astype(np.float32))
    e = DeviceArray.from_host(rs.randn(B, nf, iy - 2, iy - 2).astype(np.float32)).to_nhwc()
    for rep in range(2):
        print("== forward %d->%d" % (ic, nf), flush=True)
        y = op.forward(x, f)
        torch.cuda.synchronize()
        print("== forward + relu + pool %d->%d" % (ic, nf), flush=True)
        pk = bf.atomic.FilterPack()
        fp = op.forward_pool(x, f, bias=None, relu_in=True, pack=pk)
        torch.cuda.synchronize()
        print("== backward %d->%d" % (ic, nf), flush=True)
        dF, db, dX = op.backward(X=x, E=e, F=f)
        torch.cuda.synchronize()
