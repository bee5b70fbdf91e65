"""Host-side cost of one training step at a tiny batch (cfg1: 784-60-10, batch 20): cProfile over Backpropagation.epoch."""
import cProfile, pstats, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brainforge_b200 as bf
from brainforge_b200.layers import Dense
bf.set_precision("tf32")
np.random.seed(1)
net = bf.Backpropagation(input_shape=(784,), layerstack=[Dense(60, activation="sigmoid"), Dense(10, activation="softmax")],
                         cost="cxent", optimizer="sgd", use_graph=True)
rs = np.random.RandomState(0)
X = [bf.pinned_empty((20, 784)) for _ in range(4)]
Y = [bf.pinned_empty((20, 10)) for _ in range(4)]
for x, y in zip(X, Y):
    x[...] = rs.randn(20, 784); y[...] = np.eye(10)[rs.randint(0, 10, 20)]
def gen():
    i = 0
    while True:
        yield X[i % 4], Y[i % 4]
        i += 1
net.epoch(gen(), updates_per_epoch=50, verbose=0)
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); net.epoch(gen(), updates_per_epoch=2000, verbose=0); torch.cuda.synchronize(); t1 = time.perf_counter()
print("epoch: %.1f us per step" % ((t1 - t0) / 2000 * 1e6))
pr = cProfile.Profile(); pr.enable(); net.epoch(gen(), updates_per_epoch=2000, verbose=0); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
