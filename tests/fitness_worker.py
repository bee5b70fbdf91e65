"""Sharded neuro-evolution fitness worker (torchrun, >= 2 ranks): see tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brainforge_b200 as bf  # noqa: E402
from brainforge_b200 import dist  # noqa: E402
from brainforge_b200.layers import Dense  # noqa: E402


def build(P, resident):
    np.random.seed(9)
    return bf.NeuroEvolution(input_shape=(784,), layerstack=[Dense(60, activation="sigmoid"), Dense(10, activation="softmax")],
                             cost="cxent", population_size=P, device_resident=resident)


def main():
    dist.init_from_env()
    r, ws = dist.rank(), dist.world_size()
    bf.set_precision("tf32")
    rs = np.random.RandomState(1)
    P, N = 301, 256                       # odd: uneven shards of individuals are fine here (no collective but the gather)
    X = rs.randn(N, 784).astype(np.float32)
    Y = np.eye(10, dtype=np.float32)[rs.randint(0, 10, N)]
    for resident in (False, True):
        with dist.local_mode():
            alone = build(P, resident)
            alone.population.update(None, 0, X=X, Y=Y)
        ne = build(P, resident)
        assert ne.population.device_resident == resident
        ne.population.update(None, 0, X=X, Y=Y)
        assert np.array_equal(ne.population.grades, alone.population.grades), "sharded grades differ from one rank's"
        sub = np.array([7, 300, 12, 5, 150])
        ne.population.grades[:] = 0
        ne.population.update(sub, 0, X=X, Y=Y)
        assert np.array_equal(ne.population.grades[sub], alone.population.grades[sub])
        g = torch.as_tensor(ne.population.grades, device="cuda")
        ref = g.clone()
        torch.distributed.broadcast(ref, 0)
        assert torch.equal(g, ref)
    torch.distributed.barrier()
    if r == 0:
        print("sharded fitness OK on %d GPUs" % ws, flush=True)


if __name__ == "__main__":
    main()
