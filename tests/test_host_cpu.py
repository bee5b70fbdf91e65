"""CPU tests of the host logic: sharding, the host GA, and the N>1 plumbing over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything_once():
    from brainforge_b200 import dist
    for n in (0, 1, 7, 8, 4096, 4097):
        for ws in (1, 2, 3, 8):
            spans = [dist.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_population_minimises_with_host_fitness():
    """The GA operators stay host NumPy; same toy as xperiments/xp_evolution.py."""
    from brainforge_b200.evolution import Population
    np.random.seed(3)
    target = np.array([3., 3.])
    pop = Population(loci=2, fitness_function=lambda ind: (np.linalg.norm(target - ind * 10.),), limit=60)
    means, stds, bests = pop.run(15, verbosity=0, mutation_rate=0.01)
    assert bests[-1] <= bests[0] and bests[-1] < 1.0
    assert pop.best.shape == (2,)


def test_population_batch_fitness_matches_serial():
    from brainforge_b200.evolution import Population
    f = lambda g: (float(np.sum((g - 0.25) ** 2)),)
    fb = lambda G: np.sum((G - 0.25) ** 2, axis=1, keepdims=True)
    np.random.seed(5)
    a = Population(loci=6, fitness_function=f, limit=20)
    np.random.seed(5)
    b = Population(loci=6, fitness_function=f, limit=20, batch_fitness_function=fb)
    a.update()
    b.update()
    assert np.allclose(a.grades, b.grades)
    b.update(np.array([3, 7, 11]))
    assert np.allclose(a.grades, b.grades)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:  # surface the failure instead of a queue timeout
        q.put((rank, "error: %r" % (e,)))


def _worker_body(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as td
    td.init_process_group("gloo", rank=rank, world_size=world)
    from brainforge_b200 import dist
    from brainforge_b200.evolution import Population
    # (1) the one training collective: all-reduce(sum) of a flat gradient vector with the scalars at its tail
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    dist.allreduce_sum(g)
    ok1 = torch.allclose(g, torch.arange(10, dtype=torch.float32) * sum(range(1, world + 1)))
    # (2) population fitness sharded over ranks, only fitness values gathered
    np.random.seed(11)                       # same population on every rank (host GA is replicated)
    seen = []

    def fb(G):
        seen.append(len(G))
        return np.sum(G ** 2, axis=1, keepdims=True)
    pop = Population(loci=5, fitness_function=None, limit=9, batch_fitness_function=fb)
    pop.update()
    expect = np.sum(pop.individuals ** 2, axis=1)
    ok2 = np.allclose(pop.grades, expect) and seen[0] in (4, 5)
    pop.update(np.array([1, 2, 8]))
    ok3 = np.allclose(pop.grades, expect)
    q.put((rank, bool(ok1), bool(ok2), bool(ok3), dist.global_count(7)))
    td.destroy_process_group()


def test_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True, True, True, 14), (1, True, True, True, 14)], res


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys, timed on
    the unmodified reference when oracle/_ref is present (else the oracle port), no GPU touched."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    import os
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_bench_algorithmic_cost_formulas():
    """The per-call algorithmic bytes / FLOPs behind `roofline.achieved` (bench.py::algorithmic_cost) against closed forms
    of SURVEY 8d for the cfg3 layer shapes: every operand read once, every result written once, true-MAC FLOPs."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    B = 4096
    # 32 -> 64 on 14x14: X 32*196, E 64*144, F 64*32*9 floats
    X, E, F = B * 32 * 196, B * 64 * 144, 64 * 32 * 9
    flops = 2.0 * B * 144 * 64 * 32 * 9
    assert b.algorithmic_cost("bf_conv_nhwc_backward_filter_dbp(1184,4096,32,14,14,64,3,3)[dF]") == (4 * (X + E + F + 64), flops)
    assert b.algorithmic_cost("bf_conv_nhwc_backward_data_packed(4096,32,14,14,64,3,3)[dX]") == (4 * (E + F + X), flops)
    pooled = B * 64 * 36
    assert b.algorithmic_cost("bf_conv_nhwc_forward_pool_packed(4096,32,14,14,64,3,3,1)") == (4 * (X + F + 64 + pooled) + pooled, flops)
    # first layer, fused un-pooling filter gradient: X, the pooled-resolution delta and gate, one mask byte each
    X1, P1, F1 = B * 3 * 900, B * 32 * 196, 32 * 27
    by, fl = b.algorithmic_cost("bf_conv_c3_backward_filter_unpool(4096,3,30,30,32,3,3)[dF]")
    assert by == 4 * (X1 + 2 * P1 + F1 + 32) + P1 and fl == 2.0 * B * 784 * 32 * 27
    # per sample the forward FLOPs of the three layers add up to SURVEY's 9 022 464 (dense head excluded)
    per = sum(b.algorithmic_cost(l)[1] for l in ("bf_conv_c3_forward_pool(4096,3,30,30,32,3,3,0,1)",
                                                  "bf_conv_nhwc_forward_pool_packed(4096,32,14,14,64,3,3,1)",
                                                  "bf_conv_nhwc_forward_pool_packed(4096,64,6,6,128,3,3,1)")) / B
    assert per == 1354752 + 5308416 + 2359296
