"""Data-parallel parity worker (run under torchrun by tests/test_gpu_multi.py, or by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dp_worker.py

Each rank trains on its shard of the same global batches (batch rows split over ranks, ONE exchange of
the flat gradient vector per learn_batch, identical fused update with the global m).  Checked, for Adam and Momentum,
eager and CUDA-graph replay:
  1. the fused push all-reduce + optimizer kernel (bf_opt_step_allreduce over symmetric memory) against the NCCL
     all-reduce + bf_opt_step path: weights, costs, accuracies;
  2. both against ONE rank training on the WHOLE batch (dist.local_mode) -- the data-parallel job must be the same
     optimisation, not merely self-consistent;
  3. against the float64 oracle on the whole batch (FP32 path, <= 2e-5 after 6 steps);
  4. every rank ends with bit-identical weights; an uneven split is detected (ValueError) instead of training with a
     wrong 1/m.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import brainforge_b200 as bf  # noqa: E402
from brainforge_b200 import dist  # noqa: E402
from helpers import build_b200, build_oracle  # noqa: E402

ISHAPE = (3, 12, 12)
SPEC = [("conv", 8, 3, 3, "linear"), ("act", "relu"), ("pool", 2), ("flatten",), ("dense", 32, "tanh"), ("dense", 10, "softmax")]


def relerr(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(a), np.linalg.norm(b), 1e-300))


def train(net, batches, lo, hi):
    out = [net.learn_batch(X[lo:hi], Y[lo:hi], metrics=["acc"]) for X, Y in batches]
    torch.cuda.synchronize()
    return net.get_weights(), np.array([o["cost"] for o in out]), np.array([o["accuracy"] for o in out])


def main():
    dist.init_from_env()
    r, ws = dist.rank(), dist.world_size()
    assert ws > 1, "run under torchrun with at least 2 ranks"
    rs = np.random.RandomState(5)
    gb = 16 * ws
    batches = [(rs.randn(gb, *ISHAPE), np.eye(10)[rs.randint(0, 10, gb)]) for _ in range(6)]
    lo, hi = dist.shard_bounds(gb)
    for precision in ("fp32", "tf32"):
        bf.set_precision(precision)
        for optimizer in ("adam", "momentum"):
            with dist.local_mode():                              # one rank, whole batch
                single = build_b200(ISHAPE, SPEC, "cxent", optimizer, seed=77)
                assert single.layers.symm is None
                w1, c1, a1 = train(single, batches, 0, gb)
            results = {}
            for fused in (False, True):
                for use_graph in (False, True):
                    os.environ["BF_FUSED_ALLREDUCE"] = "1" if fused else "0"
                    net = build_b200(ISHAPE, SPEC, "cxent", optimizer, seed=77, use_graph=use_graph)
                    assert (net.layers.symm is not None) == fused, "symmetric-memory path not active"
                    results[(fused, use_graph)] = train(net, batches, lo, hi)
            base = results[(False, False)]
            for key, (w, c, a) in results.items():
                # summation order over ranks differs between NCCL's ring / tree and the rank-ordered sum of the fused kernel
                # (2 ranks: a + b either way -> identical bits)
                dw, dc = np.abs(w - base[0]).max(), np.abs(c - base[1]).max()
                tol = 0.0 if ws == 2 else 2e-6
                assert dw <= tol and dc <= 1e-6 and np.array_equal(a, base[2]), (precision, optimizer, key, dw, dc)
                # vs one rank on the whole batch: same sums in another order (per-shard partial sums, then across ranks)
                e1 = relerr(w, w1)
                assert e1 <= (2e-6 if precision == "fp32" else 2e-4), (precision, optimizer, key, e1)
                assert np.abs(c - c1).max() <= 1e-5 * np.abs(c1).max() * (1 if precision == "fp32" else 50)
                if r == 0:
                    print("[%s %-8s fused=%d graph=%d] max|dW| vs NCCL %.2e, rel.err vs single rank %.2e" %
                          (precision, optimizer, key[0], key[1], dw, e1), flush=True)
            if precision == "fp32":
                ora = build_oracle(ISHAPE, SPEC, "cxent", optimizer, seed=77)
                oc = [ora.learn_batch(X, Y)["cost"] for X, Y in batches]
                eo = relerr(results[(True, True)][0], ora.get_weights())
                assert eo <= 2e-5 and np.abs(results[(True, True)][1] - np.array(oc)).max() <= 2e-5, eo
                if r == 0:
                    print("[fp32 %-8s] fused+graph vs float64 oracle on the whole batch: %.2e" % (optimizer, eo), flush=True)
            w = torch.as_tensor(results[(True, True)][0], device="cuda")
            ref = w.clone()
            torch.distributed.broadcast(ref, 0)
            assert torch.equal(w, ref), "ranks diverged"
    # uneven shards are detected, not silently mis-scaled
    bf.set_precision("fp32")
    os.environ["BF_FUSED_ALLREDUCE"] = "1"
    net = build_b200(ISHAPE, SPEC, "cxent", "sgd", seed=77)
    X, Y = batches[0]
    n_mine = 8 if r == 0 else 6
    try:
        net.learn_batch(X[:n_mine], Y[:n_mine])
        raise AssertionError("uneven shards went unnoticed")
    except ValueError as e:
        assert "equal shards" in str(e)
    torch.distributed.barrier()
    if r == 0:
        print("dp parity OK on %d GPUs" % ws, flush=True)


if __name__ == "__main__":
    main()
