"""Golden fixtures for the remaining cost functions (binary cross-entropy, hinge: metrics/costs.py:46-69), generated
from the REAL reference like make_golden.py (authoring container only):

    python tests/golden/make_golden_costs.py
"""
import os
import sys
import tempfile

os.environ["HOME"] = tempfile.mkdtemp(prefix="bfhome_")
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

np.asscalar = lambda a: a.item()  # noqa: E731  (util/typing.py:11 needs it)
from brainforge.metrics import costs as ref_costs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rs = np.random.RandomState(4321)
m, n = 23, 7
O = rs.uniform(0.02, 0.98, size=(m, n))                       # probabilities (bxent)
T = (rs.uniform(size=(m, n)) < 0.4).astype(np.float64)        # multi-label 0/1 targets
S = rs.randn(m, n) * 1.5                                      # raw scores (hinge)
S[0, :3] = [1.0, 1.000001, 0.999999]                          # around the kink at 1, distinct in float32 too (-y for outputs <= 1)
Y = np.where(rs.uniform(size=(m, n)) < 0.5, -1.0, 1.0)        # +-1 labels
out = dict(O=O, T=T, S=S, Y=Y,
           bxent=ref_costs.bxent(O, T), bxent_delta=ref_costs.bxent.derivative(O, T),
           hinge=ref_costs.hinge(S, Y), hinge_delta=ref_costs.hinge.derivative(S, Y.copy()))
# DropOut (layers/fancy.py:60-90): one mask of the sample shape from the global RNG, shared by the batch
from brainforge.layers import DropOut  # noqa: E402
from brainforge.util import testing  # noqa: E402
brain = testing.NoBrainer(outshape=(4, 3))
brain.learning = True
layer = DropOut(0.3)
layer.connect(brain)
DX, DD = rs.randn(5, 4, 3), rs.randn(5, 4, 3)
np.random.seed(99)
d_train = layer.feedforward(DX).copy()
d_back = layer.backpropagate(DD).copy()
d_mask_after = np.asarray(layer.mask, dtype=np.float64).copy()
brain.learning = False
np.random.seed(99)
d_eval = layer.feedforward(DX).copy()
out.update(drop_X=DX, drop_D=DD, drop_train=d_train, drop_back=d_back, drop_mask_after=d_mask_after, drop_eval=d_eval)
path = os.path.join(HERE, "costs2.npz")
np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), float(out["bxent"]), float(out["hinge"]))
