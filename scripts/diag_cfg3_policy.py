"""Diagnostic (GPU): per-layer gradient error of the cfg3 learn_batch path against the emulating oracle, for several
assumptions about how the first layer's fused un-pooling filter-gradient kernel rounds its operands."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brainforge_b200 as bf
from oracle import bf_oracle as O
from helpers import CFG, build_b200, build_oracle, onehot
from conftest import relerr

bf.set_precision("tf32")
ishape, spec, cost, opt = CFG["cfg3"]
net = build_b200(ishape, spec, cost, opt, seed=1234)
rs = np.random.RandomState(99)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
X, Y = rs.randn(B, *ishape).astype(np.float32), onehot(rs, B, 10)
a = net.learn_batch(X, Y, update=False)
g = net.get_gradients()
base = bf.atomic.tf32_policy_for(net, learn=True)
for ma, mb in itertools.product(["rna", "rne", "rz", None], repeat=2):
    def pol(site, li, ma=ma, mb=mb):
        if li == 0 and site == "conv.dF":
            return (ma, mb)
        return base(site, li)
    emu = build_oracle(ishape, spec, cost, opt, seed=1234)
    with O.tf32_emulation(pol):
        b = emu.learn_batch(X, Y, update=False)
    ge = emu.get_gradients()
    s, errs = 0, []
    for l in emu.trainable:
        nW, nb = l["W"].size, l["b"].size
        errs.append("%.1e/%.1e" % (relerr(g[s:s + nW], ge[s:s + nW]), relerr(g[s + nW:s + nW + nb], ge[s + nW:s + nW + nb])))
        s += nW + nb
    print("dF0 modes (X=%s, E=%s): total %.2e  per layer W/b: %s  cost %.2e" % (ma, mb, relerr(g, ge), " ".join(errs), abs(a["cost"] - b["cost"]) / b["cost"]), flush=True)
