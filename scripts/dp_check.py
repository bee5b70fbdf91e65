"""Data-parallel check on N GPUs of one box (run under torchrun): the fused all-reduce + optimizer kernel
(bf_opt_step_allreduce over symmetric memory) against the NCCL all-reduce + bf_opt_step path, eager and graph replay,
and both against one rank training on the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/dp_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brainforge_b200 as bf  # noqa: E402
from brainforge_b200 import dist  # noqa: E402
from brainforge_b200.layers import ConvLayer, PoolLayer, Activation, Flatten, Dense  # noqa: E402


def build(fused, use_graph, optimizer):
    os.environ["BF_FUSED_ALLREDUCE"] = "1" if fused else "0"
    np.random.seed(77)
    net = bf.Backpropagation(input_shape=(3, 12, 12), layerstack=[
        ConvLayer(8, 3, 3), Activation("relu"), PoolLayer(2), Flatten(), Dense(32, activation="tanh"),
        Dense(10, activation="softmax")], cost="cxent", optimizer=optimizer, use_graph=use_graph)
    return net


def main():
    dist.init_from_env()
    r, ws = dist.rank(), dist.world_size()
    rs = np.random.RandomState(5)
    gb = 16 * ws
    batches = [(rs.randn(gb, 3, 12, 12), np.eye(10)[rs.randint(0, 10, gb)]) for _ in range(6)]
    lo, hi = dist.shard_bounds(gb)
    results = {}
    for optimizer in ("adam", "momentum"):
        for fused in (False, True):
            for use_graph in (False, True):
                net = build(fused, use_graph, optimizer)
                assert (net.layers.symm is not None) == fused, "symmetric memory path not active"
                costs = [net.learn_batch(X[lo:hi], Y[lo:hi], metrics=["acc"]) for X, Y in batches]
                torch.cuda.synchronize()
                results[(optimizer, fused, use_graph)] = (net.get_weights(), [c["cost"] for c in costs], [c[[k for k in c if k != "cost"][0]] for c in costs])
        base = results[(optimizer, False, False)]
        for key, val in results.items():
            if key[0] != optimizer:
                continue
            dw = np.abs(val[0] - base[0]).max()
            dc = np.abs(np.array(val[1]) - np.array(base[1])).max()
            da = np.abs(np.array(val[2]) - np.array(base[2])).max()
            if r == 0:
                print(key, "max|dW| %.3e max|dcost| %.3e max|dacc| %.3e" % (dw, dc, da), flush=True)
            tol = 0.0 if ws == 2 else 1e-5
            assert dw <= tol + 1e-12 and dc <= 1e-6 and da == 0, (key, dw, dc, da)
    # every rank holds the same weights
    w = torch.as_tensor(results[("adam", True, True)][0], device="cuda")
    ref = w.clone()
    torch.distributed.broadcast(ref, 0)
    assert torch.equal(w, ref), "ranks diverged"
    if r == 0:
        print("dp_check OK on %d GPUs" % ws, flush=True)


if __name__ == "__main__":
    main()
