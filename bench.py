#!/usr/bin/env python
"""bench.py -- train samples/s (fwd + bwd + update) of the brainforge hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 back-end
    python bench.py --impl reference --steps K --warmup W    # the reference itself (oracle/_ref) on the host cores
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, NCCL

Workload (default): BASELINE.json configs[2] ("cfg3") -- the 3-conv stack
[ConvLayer(n,3,3) -> ReLU -> PoolLayer(2)] for n = 32,64,128 -> Flatten -> Dense(10,softmax), cxent, the
reference's Adam, synthetic X (4096,3,30,30) ~ N(0,1), one-hot Y; GLOBAL batch 4096 sharded over the N
ranks (strong scaling), one all-reduce of the flat gradient vector per step.  It is the config the metric
("at 1/2/4/8 B200") is quoted on and it fits one GPU; `--workload cfg1|cfg2|cfg4|cfg5` select the others.

One JSON line on stdout (rank 0):
  value          samples/s with the inputs already resident in HBM (device-timed, max over ranks)
  e2e            the same metric through the public API `Backpropagation.learn_batch(X_host, Y_host)`:
                 H2D of the step's inputs from pinned host memory + D2H of the cost inside the timed region
  roofline       the dominant kernel of the step: algorithmic bytes (or FLOPs) / its CUDA-event duration,
                 against the measured peak in MEASURED_PEAKS.json
  cpu_baseline   csxeba/brainforge unmodified (oracle/_ref; the oracle port if that directory is absent) timed on
                 this box's host cores on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = {
    "cfg1": dict(ishape=(784,), spec=[("dense", 60, "sigmoid"), ("dense", 10, "softmax")], cost="cxent",
                 opt="sgd", batch=20, ncls=10, cpu_batch=20),
    "cfg2": dict(ishape=(1, 28, 28), spec=[("conv", 8, 3, 3, "linear"), ("pool", 2), ("act", "relu"), ("flatten",),
                                          ("dense", 10, "softmax")], cost="cxent", opt="sgd", batch=32, ncls=10,
                 cpu_batch=32),
    "cfg3": dict(ishape=(3, 30, 30), spec=[("conv", 32, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                                          ("conv", 64, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                                          ("conv", 128, 3, 3, "linear"), ("act", "relu"), ("pool", 2),
                                          ("flatten",), ("dense", 10, "softmax")], cost="cxent", opt="adam",
                 batch=4096, ncls=10, cpu_batch=256),
    "cfg4": dict(ishape=(64, 128), spec=[("lstm", 256, "tanh", False), ("dense", 10, "softmax")], cost="cxent",
                 opt="sgd", batch=1024, ncls=10, cpu_batch=32),
}
WORKLOAD_NAME = {
    "cfg1": "Dense MLP 784-60-10 sigmoid/softmax, SGD, batch 20 (BASELINE configs[0])",
    "cfg2": "ConvLayer(8,3x3)+PoolLayer(2)+ReLU+Dense10 on 1x28x28, SGD, batch 32 (BASELINE configs[1])",
    "cfg3": "3-conv stack 32/64/128 3x3 + ReLU + MaxPool2 + Dense10 on 3x30x30, Adam, global batch 4096 (BASELINE configs[2])",
    "cfg4": "LSTM(256,tanh) T=64 D=128 + Dense10, SGD, global batch 1024 (BASELINE configs[3])",
    "cfg5": "NeuroEvolution fitness of 4096 genomes of the 784-60-10 MLP on X (1024,784) (BASELINE configs[4])",
}


# --------------------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16_burst=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_data(cfg, n, seed, dtype=np.float32):
    rs = np.random.RandomState(seed)
    X = rs.randn(n, *cfg["ishape"]).astype(dtype)
    Y = np.eye(cfg["ncls"], dtype=dtype)[rs.randint(0, cfg["ncls"], n)]
    return X, Y


def build_oracle(cfg):
    from oracle import bf_oracle as O
    np.random.seed(1234)
    return O.Stack(cfg["ishape"], cfg["spec"], cost=cfg["cost"], optimizer=cfg["opt"])


def build_b200(cfg, use_graph):
    import brainforge_b200 as bf
    from brainforge_b200.layers import Dense, ConvLayer, PoolLayer, Flatten, Activation, LSTM
    np.random.seed(1234)     # same seed on every rank -> identical replicas (weights are never broadcast)
    layers = []
    for s in cfg["spec"]:
        k = s[0]
        layers.append({"dense": lambda: Dense(s[1], activation=s[2]), "act": lambda: Activation(s[1]),
                       "flatten": lambda: Flatten(), "conv": lambda: ConvLayer(s[1], s[2], s[3], activation=s[4]),
                       "pool": lambda: PoolLayer(s[1]),
                       "lstm": lambda: LSTM(s[1], activation=s[2], return_seq=s[3])}[k]())
    return bf.Backpropagation(input_shape=cfg["ishape"], layerstack=layers, cost=cfg["cost"], optimizer=cfg["opt"],
                              use_graph=use_graph)


# --------------------------------------------------------------------------------------- CPU arms
def time_oracle(cfg, steps, warmup, batch):
    net = build_oracle(cfg)
    X, Y = make_data(cfg, batch, 7, np.float64)
    for _ in range(warmup):
        net.learn_batch(X, Y)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.learn_batch(X, Y)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def time_oracle_fitness(P, N, steps, warmup):
    from oracle import bf_oracle as O
    rs = np.random.RandomState(3)
    X = rs.randn(N, 784)
    Y = np.eye(10)[rs.randint(0, 10, N)]
    G = rs.uniform(size=(P, 47710))
    for _ in range(warmup):
        O.fitness_mlp_logsoftmax(G[:4], X, Y, 784, 60, 10)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.fitness_mlp_logsoftmax(G, X, Y, 784, 60, 10)
    dt = time.perf_counter() - t0
    return P * steps / dt, dt / steps


def reference_available():
    try:
        from oracle import ref_runner
        return ref_runner.available()
    except Exception:
        return False


def time_reference(cfg, steps, warmup, batch):
    """The UNMODIFIED reference (oracle/_ref, installed by oracle/build_ref.sh) through its own Backpropagation API."""
    from oracle import ref_runner
    net = ref_runner.build(cfg["ishape"], cfg["spec"], cfg["cost"], cfg["opt"])
    X, Y = make_data(cfg, batch, 7, np.float64)
    return ref_runner.time_learn_batch(net, X, Y, steps, warmup)


def cpu_arm(cfg, steps, warmup, batch):
    """-> (samples/s, s/step, kind): the reference itself when oracle/_ref travelled with the repo, else the oracle port."""
    if reference_available():
        v, spu = time_reference(cfg, steps, warmup, batch)
        return v, spu, "reference"
    v, spu = time_oracle(cfg, steps, warmup, batch)
    return v, spu, "port"


def pick_cpu_batch(cfg, steps, warmup, budget_s):
    """Largest power-of-two fraction of the config's batch whose (steps + warmup) reference steps fit `budget_s`
    (the reference's time per sample is flat in the batch size: per-sample Numba loops / BLAS)."""
    probe = min(cfg["batch"], cfg["cpu_batch"])
    _, spu, _ = cpu_arm(cfg, 1, 1, probe)
    per_sample = spu / probe
    b = cfg["batch"]
    while b > probe and per_sample * b * (steps + warmup) > budget_s:
        b //= 2
    return max(b, probe)


def run_reference(args):
    """The reference's own CPU implementation of the path, on this box's host cores, all the threads NumPy's BLAS and
    Numba take by default: csxeba/brainforge unmodified from oracle/_ref (`kind: reference`), the oracle port only when
    that directory did not travel (`kind: port`).  Each step is a bounded sample of the workload (largest batch that
    keeps the run within a few minutes; the reference's cost per sample does not depend on the batch size)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count()
    K, W = args.steps, max(args.warmup, 1)
    if args.workload == "cfg5":
        P, N = 32, 1024
        rs = np.random.RandomState(3)
        X = rs.randn(N, 784)
        Y = np.eye(10)[rs.randint(0, 10, N)]
        if reference_available():
            from oracle import ref_runner
            val, spu = ref_runner.fitness_individuals_per_s(784, 60, 10, P, X, Y, K, 1)
            kind = "reference"
        else:
            val, spu = time_oracle_fitness(P, N, K, 1)
            kind = "port"
        sample, unit, metric = "%d genomes x N=%d per step (of the 4096-genome population)" % (P, N), "individuals/s", \
            "fitness evaluations/s"
        cfgd = {"workload": WORKLOAD_NAME["cfg5"], "population": 4096, "samples": N}
    else:
        cfg = CFG[args.workload]
        b = args.batch or pick_cpu_batch(cfg, K, W, 120.0)
        val, spu, kind = cpu_arm(cfg, K, W, b)
        sample, unit, metric = "batch %d per step (of the %d-sample step)" % (b, cfg["batch"]), "samples/s", \
            "train samples/s (fwd+bwd+update)"
        cfgd = {"workload": WORKLOAD_NAME[args.workload], "global_batch": cfg["batch"], "sample_batch": b}
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": spu * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfgd,
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": kind, "sample": sample,
                             "code": "csxeba/brainforge unmodified (oracle/_ref): Backpropagation.learn_batch, Dense/LSTM on "
                                     "NumPy+BLAS, Conv/Pool on Numba" if kind == "reference" else "oracle/bf_oracle.py (float64 NumPy port)"},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- roofline
def parse_label(label):
    name = label.split("(")[0]
    ints = [int(v) for v in label[label.index("(") + 1:label.index(")")].split(",") if v.strip().lstrip("-").isdigit()]
    tag = label[label.index("[") + 1:-1] if "[" in label else ""
    return name, ints, tag


def algorithmic_cost(label):
    """(bytes, flops) a launch must move / execute: every operand read once, every result written once."""
    name, a, tag = parse_label(label)
    f = 4
    if name == "bf_conv_forward":
        im, ic, iy, ix, nf, fc, fy, fx, mode = a[:9]
        oy, ox = (iy - fy + 1, ix - fx + 1) if mode == 0 else (iy + fy - 1, ix + fx - 1)
        return f * (im * ic * iy * ix + nf * ic * fy * fx + nf + im * nf * oy * ox), 2.0 * im * oy * ox * nf * ic * fy * fx
    if name == "bf_conv_backward":
        im, ic, iy, ix, nf, fy, fx = a[:7]
        oy, ox = iy - fy + 1, ix - fx + 1
        X, E, F = im * ic * iy * ix, im * nf * oy * ox, nf * ic * fy * fx
        flops = 2.0 * im * oy * ox * nf * ic * fy * fx
        return (f * (X + E + F + nf), flops) if tag == "dF" else (f * (E + F + X), flops)
    if name == "bf_conv_nhwc_backward_filter_dbp":   # leading int = rows of per-CTA channel sums left by the pooling pass
        a = a[1:]
        name = "bf_conv_nhwc_backward_filter"
    if name == "bf_conv_nhwc_backward_data_packed":
        name = "bf_conv_nhwc_backward_data"
    if name == "bf_conv_nhwc_forward_pool_packed":   # read X, packed F, bias; write the POOLED output + 1 mask byte each
        im, ic, iy, ix, nf, fy, fx = a[:7]
        oy, ox = iy - fy + 1, ix - fx + 1
        pooled = im * nf * (oy // 2) * (ox // 2)
        return f * (im * ic * iy * ix + nf * ic * fy * fx + nf + pooled) + pooled, 2.0 * im * oy * ox * nf * ic * fy * fx
    if name == "bf_pool_nhwc_backward_db":  # read E, mask, gate (pooled output); write dX + the per-CTA channel sums
        im, c, iy, ix, fd = a[:5]
        n = im * c * iy * ix
        return f * n + 2 * f * n // (fd * fd) + n // (fd * fd), 0.0
    if name == "bf_dense_head":             # X, W, b, Y read; out, dX written; dW, db read-modify-written
        m, chan, hw, n = a[:4]
        k = chan * hw
        return f * (2 * m * k + 3 * m * n + 3 * k * n + 3 * n), 6.0 * m * k * n
    if name in ("bf_conv_nhwc_forward", "bf_conv_c3_forward", "bf_conv_nhwc_forward_packed"):
        im, ic, iy, ix, nf, fy, fx = a[:7]
        oy, ox = iy - fy + 1, ix - fx + 1
        return f * (im * ic * iy * ix + nf * ic * fy * fx + nf + im * nf * oy * ox), 2.0 * im * oy * ox * nf * ic * fy * fx
    if name == "bf_conv_c3_forward_pool":   # read X, F, bias; write the POOLED output + 1 mask byte per pooled element
        im, ic, iy, ix, nf, fy, fx = a[:7]
        oy, ox = iy - fy + 1, ix - fx + 1
        pooled = im * nf * (oy // 2) * (ox // 2)
        return f * (im * ic * iy * ix + nf * ic * fy * fx + nf + pooled) + pooled, 2.0 * im * oy * ox * nf * ic * fy * fx
    if name == "bf_conv_c3_backward_filter_unpool":   # read X, the pooled-resolution delta + gate + 1 mask byte each; write dF, db
        im, ic, iy, ix, nf, fy, fx = a[:7]
        oy, ox = iy - fy + 1, ix - fx + 1
        pooled = im * nf * (oy // 2) * (ox // 2)
        return f * (im * ic * iy * ix + 2 * pooled + nf * ic * fy * fx + nf) + pooled, 2.0 * im * oy * ox * nf * ic * fy * fx
    if name in ("bf_conv_nhwc_backward_filter", "bf_conv_c3_backward_filter", "bf_conv_nhwc_backward_data"):
        im, ic, iy, ix, nf, fy, fx = a[:7]
        oy, ox = iy - fy + 1, ix - fx + 1
        X, E, F = im * ic * iy * ix, im * nf * oy * ox, nf * ic * fy * fx
        flops = 2.0 * im * oy * ox * nf * ic * fy * fx
        return (f * (E + F + X), flops) if name.endswith("data") else (f * (X + E + F + nf), flops)
    if name == "bf_pool_nhwc_forward":      # read A, write pooled out + 1 mask byte per window
        im, c, iy, ix, fd = a[:5]
        n = im * c * iy * ix
        return f * n + f * n // (fd * fd) + n // (fd * fd), 0.0
    if name == "bf_pool_nhwc_backward":     # read E, mask (+ gate = pooled output), write dX
        im, c, iy, ix, fd = a[:5]
        n = im * c * iy * ix
        return f * n + 2 * f * n // (fd * fd) + n // (fd * fd), 0.0
    if name in ("bf_layout_nchw_to_nhwc", "bf_layout_nhwc_to_nchw"):
        im, c, hw = a[:3]
        return 2 * f * im * c * hw, 0.0
    if name == "bf_pool_forward":
        im, ic, iy, ix, fd = a[:5]
        n = im * ic * iy * ix
        return f * n + f * n // (fd * fd) + n // (fd * fd), 0.0
    if name == "bf_pool_backward":
        im, ic, iy, ix, fd = a[:5]
        n = im * ic * iy * ix
        return f * n + f * n // (fd * fd) + n // (fd * fd), 0.0
    if name in ("bf_act_forward",):
        return 2 * f * a[1], 0.0
    if name == "bf_act_backward":
        return 3 * f * a[1], 0.0
    if name == "bf_dense_forward":
        m, k, n = a[:3]
        return f * (m * k + k * n + n + m * n), 2.0 * m * k * n
    if name == "bf_dense_backward":
        m, k, n = a[:3]
        return f * (2 * m * k + 2 * m * n + 2 * k * n + n), 4.0 * m * k * n
    if name in ("bf_lstm_forward", "bf_lstm_forward_bm"):
        T, B, D, H = a[:4]
        return f * T * B * (D + (D + H) + H + 11 * H) + f * (D + H) * 4 * H, 2.0 * T * B * (D + H) * 4 * H
    if name == "bf_lstm_backward":
        T, B, D, H = a[:4]
        return f * T * B * ((D + H) + 12 * H + D + 8 * H) + 2 * f * (D + H) * 4 * H, 4.0 * T * B * (D + H) * 4 * H
    if name == "bf_opt_step":
        n = a[1] if len(a) > 1 else a[0]
        return 7 * f * n, 0.0
    if name == "bf_population_fitness":     # genomes read in place: the K-major operand (64 padded hidden rows) + the tails
        ldg, Pstore, P, N, nin, nh, nout = a[:7]
        return f * (P * (nin * 64 + nh + nh * nout + nout) + N * nin + N * nout + P), 2.0 * P * N * (nin * nh + nh * nout)
    if name == "bf_fitness_mlp":
        P, N, nin, nh, nout = a[:5]
        return f * (P * (nin * nh + nh + nh * nout + nout) + N * nin + N * nout + P), 2.0 * P * N * (nin * nh + nh * nout)
    return 0, 0.0


def measured_traffic(label):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the call's dominant kernel, from the newest committed
    `ncu --set full` summary under profiles/ (None when that launch was not captured).  Entries are keyed by the call label
    `name(int args)`: one per layer shape (scripts/summarise_profiles.py maps each captured kernel instance to its call)."""
    pdir = os.path.join(ROOT, "profiles")
    if not os.path.isdir(pdir):
        return None
    files = sorted(f for f in os.listdir(pdir) if f.endswith("_traffic.json"))
    if not files:
        return None
    try:
        table = json.load(open(os.path.join(pdir, files[-1])))
    except Exception:
        return None
    key = label.split("[")[0]
    if key in table:
        return table[key].get("bytes")
    return None          # this launch (shape) was not captured: no figure rather than another instance's


def roofline_entry(times, peaks, tf32):
    """times: {label: [ms]} from the instrumented pass.  Picks the call with the largest total time."""
    if not times:
        return None, []
    tot = {k: float(np.sum(v)) for k, v in times.items()}
    step_total = sum(tot.values())
    ranked = sorted(tot, key=tot.get, reverse=True)
    top = ranked[0]
    avg_ms = float(np.mean(times[top]))
    nbytes, flops = algorithmic_cost(top)
    gbs = nbytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    tfs = flops / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    # ridge point of the machine balance: which roof bounds this launch
    # kind::tf32 runs at half the bf16 rate (nominal 1.1 vs 2.25 PF): the denominator is the larger of measured-bf16 / 2
    # and what this library's own TMA + tcgen05 TF32 GEMM sustains on a 4096^3 product (measured in this run)
    own = peaks.get("tf32_gemm_measured") or 0.0
    tensor_peak = max(peaks["bf16_sustained"] / 2.0, own)
    t_hbm = nbytes / (peaks["hbm"] * 1e9)
    t_tensor = flops / (tensor_peak * 1e12) if flops else 0.0
    if t_tensor > t_hbm:
        entry = {"bound": "tensor", "achieved": tfs, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tfs / tensor_peak,
                 "peak_note": "max(measured sustained bf16 (%s) / 2 = %.1f, own TF32 GEMM 4096^3 measured here = %.1f TFLOP/s)"
                              % (peaks["source"], peaks["bf16_sustained"] / 2.0, own)}
    else:
        entry = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                 "peak_note": "HBM copy bandwidth, %s" % peaks["source"]}
    entry.update({"kernel": top, "avg_ms": avg_ms, "share_of_step": tot[top] / step_total if step_total else None,
                  "algorithmic_bytes": nbytes, "algorithmic_flops": flops,
                  "traffic": measured_traffic(top),
                  "path": "tcgen05-tf32" if tf32 else "simt-fp32"})
    table = [{"kernel": k, "calls": len(times[k]), "avg_ms": float(np.mean(times[k])),
              "share": tot[k] / step_total if step_total else None} for k in ranked[:12]]
    return entry, table


# --------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import brainforge_b200 as bf
    from brainforge_b200 import dist, device
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback (use --impl reference)"
    dist.init_from_env()
    rank, ws = dist.rank(), dist.world_size()
    if ws == 1:
        torch.cuda.set_device(0)
    bf.set_precision(args.precision)
    peaks = load_peaks()
    K, W = args.steps, max(args.warmup, 3)

    if args.workload == "cfg5":
        return run_b200_fitness(args, bf, dist, device, torch, peaks, K, W)

    cfg = CFG[args.workload]
    gb = args.batch or cfg["batch"]
    lo, hi = dist.shard_bounds(gb)
    mb = hi - lo
    net = build_b200(cfg, use_graph=not args.no_graph)
    # distinct input batches, rotated every step: with cfg3 the rotation set (4 x 44 MB) plus the step's
    # activations (>1 GB) exceeds the 126 MB L2, so no step finds its inputs cached
    nrot = 4
    hostX, hostY, devX, devY = [], [], [], []
    for r in range(nrot):
        X, Y = make_data(cfg, gb, 100 + r)
        px, py = bf.pinned_empty((mb,) + cfg["ishape"]), bf.pinned_empty((mb, cfg["ncls"]))
        px[...] = X[lo:hi]; py[...] = Y[lo:hi]
        hostX.append(px); hostY.append(py)
        devX.append(bf.DeviceArray.from_host(px)); devY.append(bf.DeviceArray.from_host(py))
    h2d = hostX[0].nbytes + hostY[0].nbytes
    bf.synchronize()

    def barrier():
        if ws > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    graph_ok = not args.no_graph
    # ---- (A) end-to-end through the public API.  Two flavours, both with every step's H2D of its inputs (pinned
    # host memory) and D2H of its cost inside the timed region:
    #   e2e             Backpropagation.epoch(generator, K): the reference's training loop (learner/
    #                   abstract_learner.py:48-74) -> learn_batch per step; this back-end uploads batch i+1 on a copy
    #                   stream while step i computes
    #   e2e_learn_batch K bare learn_batch(X_host, Y_host) calls: copy, compute and read-back strictly in series
    def batches():
        i = 0
        while True:
            yield hostX[i % nrot], hostY[i % nrot]
            i += 1

    try:
        for i in range(W + 1):
            out = net.learn_batch(hostX[i % nrot], hostY[i % nrot])
    except Exception as e:            # CUDA-graph capture not possible in this configuration: run eagerly
        if args.no_graph:
            raise
        sys.stderr.write("[bench] graph capture failed (%r); falling back to eager launches\n" % (e,))
        graph_ok = False
        net = build_b200(cfg, use_graph=False)
        for i in range(W + 1):
            out = net.learn_batch(hostX[i % nrot], hostY[i % nrot])
    net.epoch(batches(), updates_per_epoch=W, verbose=0)
    barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    # three K-step epochs, the MEDIAN is reported (all three are in the line): a host-side hiccup -- the loop is bound by the
    # host link and one Python thread -- otherwise decides a 15 ms measurement
    e2e_runs = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hist = net.epoch(batches(), updates_per_epoch=K, verbose=0)
        e1.record()
        barrier()
        e2e_runs.append(e0.elapsed_time(e1))
        assert len(hist["cost"]) == K
    e2e_ms = sorted(e2e_runs)[1]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(K):
        out = net.learn_batch(hostX[i % nrot], hostY[i % nrot])
    s1.record()
    barrier()
    e2e_serial_ms = s0.elapsed_time(s1)

    # ---- (A2) the same through `fit(X_host, Y_host)`: the data set (S steps' worth of this rank's shard) is uploaded ONCE
    # inside the timed call, every mini-batch of the E passes is a shuffled on-device gather (ResidentDataset); and K bare
    # learn_batch calls on float64 PAGEABLE arrays -- the reference API's native currency (staged as f64, narrowed on the device)
    S, E = min(K, 8), 3
    Xf = np.concatenate([hostX[i % nrot] for i in range(S)])
    Yf = np.concatenate([hostY[i % nrot] for i in range(S)])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    fhist = net.fit(Xf, Yf, batch_size=mb, epochs=E, verbose=0, shuffle=True)
    f1.record()
    barrier()
    fit_ms = f0.elapsed_time(f1)
    assert len(fhist["cost"]) == S * E
    del Xf, Yf
    K64 = min(K, 5)
    X64, Y64 = np.asarray(hostX[0], dtype=np.float64), np.asarray(hostY[0], dtype=np.float64)
    net.learn_batch(X64, Y64)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(K64):
        net.learn_batch(X64, Y64)
    p1.record()
    barrier()
    f64_ms = p0.elapsed_time(p1)
    del X64, Y64

    # ---- (B) inputs resident in HBM: the device part of the step, K times back to back
    launches0 = device.launch_count()
    metrics = []

    def dev_step(i):
        net._step(devX[i % nrot], devY[i % nrot], None, metrics, True, mb, gb)

    graphs = []
    if graph_ok:
        try:
            for i in range(nrot):
                dev_step(i)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(nrot):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        dev_step(i)
                    graphs.append(g)
            torch.cuda.current_stream().wait_stream(side)
        except Exception as e:
            sys.stderr.write("[bench] device-step graph capture failed (%r); eager\n" % (e,))
            graphs = []
    per_step_launches = None
    l0 = device.launch_count()
    dev_step(0)
    per_step_launches = device.launch_count() - l0
    for i in range(W):
        graphs[i % nrot].replay() if graphs else dev_step(i)
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for i in range(K):
        graphs[i % nrot].replay() if graphs else dev_step(i)
    d1.record()
    barrier()
    dev_ms = d0.elapsed_time(d1)
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms, e2e_serial_ms, fit_ms, f64_ms], dtype=torch.float64, device="cuda")
    if ws > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_serial_ms, fit_ms, f64_ms = (float(v) for v in t)

    # ---- (C) per-kernel CUDA-event timing of the same step (eager, instrumented) for the roofline
    roof, table = None, []
    if rank == 0 or ws > 1:
        device.start_timing()
        for i in range(max(2, min(K, 5))):
            dev_step(i)
        times = device.stop_timing()
        if rank == 0:
            if args.precision == "tf32":
                peaks["tf32_gemm_measured"] = measure_tf32_gemm(bf, torch)
            roof, table = roofline_entry(times, peaks, args.precision == "tf32")
    barrier()

    if rank != 0:
        return
    # ---- CPU baseline: the oracle on a bounded sample of the same workload (N=1 only)
    cpu = None
    if ws == 1 and not args.no_cpu:
        b = cfg["cpu_batch"]
        val, spu, kind = cpu_arm(cfg, 3 if spu_guess(cfg) < 3 else 1, 1, b)
        cpu = {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": kind,
               "sample": "%s, float64, batch %d, %.2f s/step" % ("csxeba/brainforge unmodified (oracle/_ref)" if kind == "reference"
                                                                 else "NumPy oracle port", b, spu)}
    line = {
        "metric": "train samples/s (fwd+bwd+update)", "value": gb * K / (dev_ms * 1e-3), "unit": "samples/s",
        "n_gpus": ws, "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[args.workload], "global_batch": gb, "per_gpu_batch": mb,
                   "parallelism": "dp%d" % ws, "precision_path": args.precision, "cuda_graph": bool(graphs),
                   "allreduce": ("none" if ws == 1 else "fused with the optimizer step: gradient slices pushed into the peers' "
                                 "receive buffers over NVLink (bf_opt_step_allreduce)" if getattr(net.layers, "symm", None) is not None
                                 else "NCCL all-reduce + bf_opt_step"),
                   "input_delta": "first layer's dX is not computed inside learn_batch (the reference discards it; "
                                  "its FLOPs are excluded per SURVEY 8d)",
                   "l2": "4 rotating input batches + >1 GB of activations per step exceed the 126 MB L2",
                   "cost_last_step": out["cost"]},
        "e2e": {"value": gb * K / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms / K,
                "ms_per_step_runs": [r / K for r in e2e_runs], "statistic": "median of 3 epochs of K steps (this rank)",
                "h2d_bytes_per_step": h2d * ws, "d2h_bytes_per_step": 256 * ws,
                "api": "Backpropagation.epoch(generator of pinned f32 host batches, K): learn_batch per step, "
                       "next batch's H2D overlapped on a copy stream"},
        "e2e_learn_batch": {"value": gb * K / (e2e_serial_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_serial_ms / K,
                            "api": "Backpropagation.learn_batch(X_host_pinned_f32, Y_host_pinned_f32), copy and compute in series"},
        "e2e_fit": {"value": gb * S * E / (fit_ms * 1e-3), "unit": "samples/s", "ms_per_step": fit_ms / (S * E),
                    "h2d_bytes_per_step": h2d * ws / E, "steps": S * E,
                    "api": "Backpropagation.fit(X_host_f32, Y_host_f32, batch_size, epochs=%d, shuffle=True): one upload of the %d-step "
                           "data set inside the timed call, shuffled mini-batches gathered on the device" % (E, S)},
        "e2e_f64_pageable": {"value": gb * K64 / (f64_ms * 1e-3), "unit": "samples/s", "ms_per_step": f64_ms / K64,
                             "h2d_bytes_per_step": 2 * h2d * ws,
                             "api": "Backpropagation.learn_batch(X_f64_pageable, Y_f64_pageable): the reference API's native arrays"},
        "gpu_launches": int(per_step_launches * K),
        "clocks": clocks, "roofline": roof, "kernels": table, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def measure_tf32_gemm(bf, torch, n=4096, reps=10):
    """TFLOP/s of this library's own TMA-staged tcgen05 TF32 GEMM (bf_dense_forward, n > 32 route) on an n^3 product."""
    try:
        rs = np.random.RandomState(0)
        x = bf.DeviceArray.from_host(rs.randn(n, n).astype(np.float32))
        w = bf.DeviceArray.from_host((rs.randn(n, n) * 0.05).astype(np.float32))
        op = bf.atomic.DenseOp()
        for _ in range(3):
            op.forward(x, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            op.forward(x, w)
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    except Exception as exc:      # noqa: BLE001
        sys.stderr.write("[bench] TF32 GEMM measurement failed: %r\n" % (exc,))
        return None


def spu_guess(cfg):
    return {4096: 1.5, 1024: 1.0}.get(cfg["batch"], 0.01)


def run_b200_fitness(args, bf, dist, device, torch, peaks, K, W):
    """cfg5: one 'step' = `Population.update` of the whole population (P genomes) on the shared (X, Y).  The population is
    device resident: genomes and their K-major first-layer operand live in HBM between generations."""
    from brainforge_b200.layers import Dense
    rank, ws = dist.rank(), dist.world_size()
    P, N = args.batch or 4096, 1024
    lo, hi = dist.shard_bounds(P)
    np.random.seed(1234)
    ne = bf.NeuroEvolution(input_shape=(784,), layerstack=[Dense(60, activation="sigmoid"), Dense(10, activation="softmax")],
                           cost="cxent", population_size=P)
    pop = ne.population
    assert pop.device_resident, "cfg5 expects the device-resident population (TF32 path)"
    store = pop._store
    rs = np.random.RandomState(5)
    X = bf.pinned_empty((N, 784)); X[...] = rs.randn(N, 784)
    Y = bf.pinned_empty((N, 10)); Y[...] = np.eye(10, dtype=np.float32)[rs.randint(0, 10, N)]
    xd, yd = bf.DeviceArray.from_host(X), bf.DeviceArray.from_host(Y)
    mine = np.arange(lo, hi)
    import ctypes
    from brainforge_b200 import _lib
    idx = torch.as_tensor(mine.astype(np.int32), device="cuda")
    out = bf.DeviceArray.empty(hi - lo)

    def dev_step():       # what Population.update launches for this rank's shard, without the read-back
        device.call(_lib.lib.bf_population_fitness, store.g.ptr, store.ldg, store.P, store.bt.ptr, ctypes.c_void_p(idx.data_ptr()),
                    hi - lo, xd.ptr, yd.ptr, out.ptr, N, 784, 60, 10, device.precision(), device.stream_ptr())

    def barrier():
        if ws > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        dev_step()
        pop.update(None, 0, X=X, Y=Y)
    barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    l0 = device.launch_count()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(K):
        dev_step()
    d1.record()
    barrier()
    launches = device.launch_count() - l0
    # end to end through the public API: Population.update(X_host, Y_host) -- H2D of X and Y, the fan-out on every rank's
    # shard, D2H + gather of the fitness values, grades refreshed on the host
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        pop.update(None, 0, X=X, Y=Y)
    e1.record()
    barrier()
    # one generation step (selection + pairing on the host, mating / mutation / re-layout on the device), for the record
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host = time.perf_counter()
    g0.record()
    pop.run(2, survival_rate=0.5, mutation_rate=0.1, verbosity=0, X=xd, Y=yd)
    g1.record()
    barrier()
    gen_wall = (time.perf_counter() - t_host) / 2
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([d0.elapsed_time(d1), e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if ws > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    device.start_timing()
    dev_step()
    times = device.stop_timing()
    if rank != 0:
        return
    roof, table = roofline_entry(times, peaks, args.precision == "tf32")
    cpu = None
    if ws == 1 and not args.no_cpu:
        if reference_available():
            from oracle import ref_runner
            val, spu = ref_runner.fitness_individuals_per_s(784, 60, 10, 32, np.asarray(X, np.float64), np.asarray(Y, np.float64), 2, 1)
            kind = "reference"
        else:
            val, spu = time_oracle_fitness(32, N, 2, 1)
            kind = "port"
        cpu = {"value": val, "unit": "individuals/s", "cores": os.cpu_count(), "kind": kind,
               "sample": "32 genomes x N=%d through Population.update, %.2f s/step" % (N, spu)}
    line = {"metric": "fitness evaluations/s", "value": P * K / (dev_ms * 1e-3), "unit": "individuals/s", "n_gpus": ws,
            "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME["cfg5"], "population": P, "samples": N, "parallelism": "dp%d" % ws,
                       "store": "device resident: genomes (P x %d FP32) + K-major first-layer operand stay in HBM" % store.ldg,
                       "l2": "the K-major genome operand (%.0f MB per rank) exceeds the 126 MB L2" % ((hi - lo) * 64 * 784 * 4 / 1e6),
                       "mean_fitness": float(np.mean(pop.grades)),
                       "generation_step_ms": {"wall": gen_wall * 1e3, "note": "one epoch of Population.run incl. re-evaluation "
                                              "of survivors + mutants; selection/pairing on the host"}},
            "e2e": {"value": P * K / (e2e_ms * 1e-3), "unit": "individuals/s", "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": int(X.nbytes + Y.nbytes) * ws, "d2h_bytes_per_step": 4 * P,
                    "api": "Population.update(None, X=X_host_pinned_f32, Y=Y_host_pinned_f32) on the device-resident population"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": table, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--precision", default=os.environ.get("BF_PRECISION", "tf32"), choices=["fp32", "tf32"],
                    help="tf32: tcgen05 tensor-core contractions (default, north_star); fp32: SIMT FMA path (<=1e-5 parity)")
    ap.add_argument("--batch", type=int, default=0, help="override the global batch / population")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
