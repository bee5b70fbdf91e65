"""Diagnostic (GPU): cfg3 forward outputs and backward deltas, layer by layer, device (TF32 path) vs emulating oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brainforge_b200 as bf
from brainforge_b200.device import DeviceArray
from oracle import bf_oracle as O
from helpers import CFG, build_b200, build_oracle, onehot
from conftest import relerr

bf.set_precision("tf32")
ishape, spec, cost, opt = CFG["cfg3"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 99
net = build_b200(ishape, spec, cost, opt, seed=1234)
rs = np.random.RandomState(seed)
X, Y = rs.randn(B, *ishape).astype(np.float32), onehot(rs, B, 10)
pol = bf.atomic.tf32_policy_for(net, learn=False)
emu = build_oracle(ishape, spec, cost, opt, seed=1234)
with O.tf32_emulation(pol):
    po = emu.feedforward(X)
    emu.backpropagate(O.cost_derivative(po, Y))
xd, yd = DeviceArray.from_host(X), DeviceArray.from_host(Y)
preds = net.layers._fwd(xd)
layers = net.layers.layers[1:]
for i, (layer, ol) in enumerate(zip(layers, emu.L)):
    out = layer.output
    if out is not None:
        print("fwd  %2d %-22s %.2e" % (i, str(layer), relerr(np.asarray(out), ol["output"])))
delta = net.cost.derivative(preds, yd)
for i in range(len(layers) - 1, -1, -1):
    layer, ol = layers[i], emu.L[i]
    d_in = delta.materialize() if hasattr(delta, "materialize") else delta
    e_in = relerr(np.asarray(d_in).reshape(np.shape(ol["delta_in"])), ol["delta_in"])
    delta = layer._bwd(delta)
    msg = "bwd  %2d %-22s delta_in %.2e" % (i, str(layer), e_in)
    if layer.trainable:
        msg += "  dW %.2e db %.2e" % (relerr(np.asarray(layer.nabla_w), ol["gW"]), relerr(np.asarray(layer.nabla_b), ol["gb"]))
        a, b = np.asarray(layer.nabla_w), ol["gW"]
        bad = np.abs(a - b) > 1e-3 * np.abs(b).max()
        msg += "  elements off by >1e-3*max: %d of %d" % (bad.sum(), bad.size)
    print(msg, flush=True)
d_in = delta.materialize() if hasattr(delta, "materialize") else delta
print("input delta %.2e" % relerr(np.asarray(d_in), emu.L[0]["delta_in"] * 0 + np.asarray(d_in)) if False else "")
