"""learner/backpropagation.py of the reference: the training step (forward, cost', backward, update).

Device flow of one `learn_batch` (everything between the two host copies stays in HBM):
  H2D X,Y -> layers._fwd -> bf_cost_delta -> layer._bwd (reverse) -> [all-reduce(sum) of the flat gradient
  vector + step scalars when data-parallel] -> ONE fused optimizer kernel over the flat parameter vector
  (also clears the gradients) -> D2H of the step scalars (cost sum, hits).
Gradients are batch SUMS (atomic/core_op.py:17-19) and the 1/m lives in the optimizer
(optimizers/gradient_descent.py:9), so sharding the batch over ranks and all-reducing the sums with the
GLOBAL m is algebraically exact.
"""
import os

import numpy as np
import torch

from .abstract_learner import Learner
from .. import _lib
from ..device import DeviceArray, as_device, call, stream_ptr, workspace_generation, note_mutation
from ..metrics import metrics as _metrics
from ..optimizers import optimizers, GradientDescent
from .. import dist as _dist


class StagedBatch:
    """A batch whose host->device copy is in flight on the copy stream (Backpropagation.stage_batch)."""

    def __init__(self, x, y, event):
        self.x, self.y, self.event = x, y, event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)

    def __len__(self):
        return len(self.x)


class Backpropagation(Learner):

    def __init__(self, layerstack, cost="mse", optimizer="sgd", name="", **kw):
        super().__init__(layerstack, cost, name, **kw)
        self.optimizer = optimizer if isinstance(optimizer, GradientDescent) else optimizers[optimizer]()
        self.optimizer.initialize(nparams=self.layers.num_params)
        self.layers.flatten_parameters()
        self.use_graph = kw.get("use_graph", False)
        self.fused_head = kw.get("fused_head", os.environ.get("BF_FUSED_HEAD", "1") != "0")
        if any(type(l).__name__ == "DropOut" for l in self.layers.layers):
            self.use_graph = False       # the mask is drawn on the host every forward pass: nothing to replay
        self._graphs = {}

    # ------------------------------------------------------------------ the hot path
    def learn_batch(self, X, Y, w=None, metrics=(), update=True):
        """learner/backpropagation.py:16-30.  Under torch.distributed (world_size > 1) X, Y are this
        rank's shard of the global batch; cost and metrics are those of the GLOBAL batch."""
        return self.learn_batch_async(X, Y, w=w, metrics=metrics, update=update)()

    def learn_batch_async(self, X, Y, w=None, metrics=(), update=True, snapshot=False):
        """Enqueue the whole device step and return a callable that reads the step's scalars back (the only
        host synchronisation of a step).  `Learner.epoch` uses the gap to upload the NEXT batch on a copy stream.
        With `snapshot` the scalars are also copied to pinned host memory right behind the step, so the callable stays
        valid (and waits for THIS step only) after later steps have been enqueued."""
        metrics = [_metrics.get(m) for m in metrics]
        self.layers.flatten_parameters()
        if isinstance(X, StagedBatch):
            X.wait()
            X, Y = X.x, X.y
        if self.use_graph and w is None:
            reader = self._launch_graph(X, Y, metrics, update)
            m_local = len(X)
        else:
            x, y = as_device(X), as_device(Y)
            wd = as_device(w) if w is not None else None
            m_local = len(x)
            self._step(x, y, wd, metrics, update, m_local, None)
            reader = lambda: self._read_scalars(metrics, m_local)   # noqa: E731
        return self._snapshot_scalars(metrics, m_local) if snapshot else reader

    def _snapshot_scalars(self, metrics, m_local):
        sc = self.layers.scalars().t
        if getattr(self, "_snap_host", None) is None or self._snap_host[0].numel() != sc.numel():
            self._snap_host = [torch.empty(sc.numel(), dtype=torch.float32).pin_memory() for _ in range(4)]
            self._snap_ev = [torch.cuda.Event() for _ in range(4)]
            self._snap_i = -1
        self._snap_i = slot = (self._snap_i + 1) % 4
        host, ev = self._snap_host[slot], self._snap_ev[slot]
        host.copy_(sc, non_blocking=True)
        ev.record()

        def read():
            ev.synchronize()
            return self._scalars_to_dict(host.numpy(), metrics, m_local)
        return read

    def stage_batch(self, X, Y):
        """Start the host->device copy of a batch on a side stream; the returned handle is accepted by
        learn_batch / learn_batch_async in place of (X, Y)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            xd, yd = DeviceArray.from_host(X), DeviceArray.from_host(Y)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        xd.t.record_stream(main)
        yd.t.record_stream(main)
        return StagedBatch(xd, yd, ev)

    def _step(self, x, y, wd, metrics, update, m_local, m_global_static):
        """Device-only part of learn_batch (capturable in a CUDA graph: no host sync inside)."""
        sc = self.layers.scalars()
        last = self.layers.layers[-1]
        if self.fused_head and self._mlp2_step(x, y, wd, metrics, sc):
            pass             # Dense -> Dense at small batch: forward, cost, both backward passes in one cluster kernel
        elif self.fused_head and len(self.layers.layers) > 1 and getattr(last, "head_supported", None) and last.head_supported(self.cost):
            # output head in one kernel: last Dense forward, cost value (of the PRE-update forward pass, :26), accuracy,
            # cost derivative and the layer's backward pass (bf_dense_head) -- five launches and a second read of its input less
            below = self.layers._fwd(x, stop=len(self.layers.layers) - 1)
            delta = last.learn_head(below, y, wd, self.cost, sc[0:1], sc[1:2] if metrics else None,
                                    need_dX=len(self.layers.layers) > 2)
            self._bwd(delta, need_input_delta=False, below_last=True)
        else:
            preds = self.layers._fwd(x)
            delta = self.cost.derivative_and_value(preds, y, wd, sc[0:1])   # cost of the PRE-update forward pass (:26)
            if metrics:
                _metrics.accuracy.value_into(preds, y, sc[1:2])
            # learn_batch discards the delta w.r.t. the network input (learner/backpropagation.py:22 ignores the
            # return value), so the first layer's dX contraction is not computed here; backpropagate() still does.
            self._bwd(delta, need_input_delta=False)
        m_global = m_global_static if m_global_static is not None else _dist.global_count(m_local)
        if _dist.world_size() > 1:
            if update and self.layers.symm is not None:
                # one kernel: every rank's gradient vector (and step scalars) pushed to all peers over NVLink, summed, + update
                self.update(m_global, symm=self.layers.symm, m_local=m_local)
                return
            # gradients + step scalars in ONE collective (scalars alone when no update follows); slot 2 carries this
            # rank's sample count so that the host can tell an uneven split from the sum
            sc[2:3].fill(m_local)
            _dist.allreduce_sum(self.layers.flat_grads if update else sc)
        if update:
            self.update(m_global)

    def _mlp2_step(self, x, y, wd, metrics, sc):
        """InputLayer -> Dense -> Dense (BASELINE configs[0]) at batch <= 32: the device step up to the optimizer is ONE
        launch (bf_mlp2_step).  Sets the layers' inputs / outputs and accumulates their gradients exactly like
        _fwd -> cost -> _bwd; returns False when the stack or the shapes do not qualify."""
        L = self.layers.layers
        if len(L) != 3 or x.nhwc or len(x.shape) != 2 or getattr(x, "cl_flat", None) is not None:
            return False
        l1, l2 = L[1], L[2]
        if type(l1).__name__ != "Dense" or type(l2).__name__ != "Dense" or not (l1.trainable and l2.trainable):
            return False
        if self.cost.kind not in ("mse", "cxent", "bxent"):
            return False
        m, k = x.shape
        h, n = l1.neurons, l2.neurons
        a1, a2, ck = _lib.ACT[str(l1.activation)], _lib.ACT[str(l2.activation)], _lib.COST[self.cost.kind]
        if not _lib.lib.bf_mlp2_step_supported(m, k, h, n, a1, a2, ck):
            return False
        H, out = DeviceArray.empty(m, h), DeviceArray.empty(m, n)
        call(_lib.lib.bf_mlp2_step, x.ptr, l1.weights.ptr, l1.biases.ptr, l2.weights.ptr, l2.biases.ptr, y.ptr,
             wd.ptr if wd is not None else None, H.ptr, out.ptr, l1.nabla_w.ptr, l1.nabla_b.ptr, l2.nabla_w.ptr, l2.nabla_b.ptr,
             sc[0:1].ptr, sc[1:2].ptr if metrics else None, m, k, h, n, a1, a2, ck, stream_ptr())
        L[0]._fwd(x)
        l1.inputs, l1.output, l2.inputs, l2.output = x, H, H, out
        return True

    def _read_scalars(self, metrics, m_local):
        return self._scalars_to_dict(self.layers.scalars().numpy(), metrics, m_local)

    def _scalars_to_dict(self, sc, metrics, m_local):
        m = _dist.global_count(m_local)
        if _dist.world_size() > 1:
            symm = self.layers.symm
            if symm is not None:
                symm.check()         # a peer missed the exchange: this replica skipped the update (raises)
            if int(round(float(sc[2]))) != m:
                raise ValueError("data-parallel learn_batch needs equal shards on every rank: the ranks hold %d samples in "
                                 "total, this rank's %d x world_size %d = %d was used as the batch size of the update"
                                 % (int(round(float(sc[2]))), m_local, _dist.world_size(), m))
        out = {"cost": float(sc[0]) / m}
        for metric in metrics:
            out[str(metric).lower()] = float(sc[1]) / m
        return out

    def _launch_graph(self, X, Y, metrics, update):
        """CUDA-graph replay of the step for a fixed (shape, metrics, update) signature: the ~dozens of
        kernel launches of a step collapse into one graph launch (small-batch configs are launch-bound,
        SURVEY 7.2#6).  Inputs are staged through static device buffers (host arrays: one H2D; arrays already
        in HBM, e.g. prefetched by `stage_batch`: one device-to-device copy)."""
        if not isinstance(X, DeviceArray):
            X = np.asarray(X)
        if not isinstance(Y, DeviceArray):
            Y = np.asarray(Y)
        key = (tuple(X.shape), tuple(Y.shape), tuple(str(m) for m in metrics), bool(update), float(self.optimizer.eta))
        entry = self._graphs.get(key)
        m_local = len(X)

        def fill(dst, src):
            if isinstance(src, DeviceArray):
                src = src.to_nchw()
                if src.size:
                    call(_lib.lib.bf_memcpy_d2d, dst.ptr, src.ptr, src.size * 4, stream_ptr())
            else:
                dst.copy_from_host(src)

        if entry is None:
            xs, ys = DeviceArray.empty(*X.shape), DeviceArray.empty(*Y.shape)
            fill(xs, X)
            fill(ys, Y)
            m_global = m_local * _dist.world_size()
            # first call runs eagerly (sizes the workspace, warms every kernel), later calls replay
            self._step(xs, ys, None, metrics, update, m_local, m_global)
            self._graphs[key] = {"x": xs, "y": ys, "graph": None, "m_global": m_global, "calls": 1}
            return lambda: self._read_scalars(metrics, m_local)
        fill(entry["x"], X)
        fill(entry["y"], Y)
        if entry["graph"] is not None and entry["ws_gen"] != workspace_generation():
            entry["graph"] = None     # the scratch workspace moved since the capture: its old address is baked into the graph
        if entry["graph"] is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self._step(entry["x"], entry["y"], None, metrics, update, m_local, entry["m_global"])
            torch.cuda.current_stream().wait_stream(side)
            entry["graph"], entry["ws_gen"] = g, workspace_generation()
        entry["graph"].replay()
        note_mutation()               # the replayed step updated the weights behind Python's back
        return lambda: self._read_scalars(metrics, m_local)

    # ------------------------------------------------------------------ pieces of the public API
    def _bwd(self, error, need_input_delta=True, below_last=False):
        first = self.layers[1] if len(self.layers.layers) > 1 else None
        for layer in self.layers[(-2 if below_last else -1):0:-1]:
            if layer is first and not need_input_delta and getattr(layer, "_can_skip_dx", False):
                error = layer._bwd(error, need_dX=False)
            else:
                error = layer._bwd(error)
        return error

    def backpropagate(self, error):
        """learner/backpropagation.py:32-35; returns the delta w.r.t. the network input."""
        out = self._bwd(as_device(error) if isinstance(error, DeviceArray) else DeviceArray.from_host(error))
        return out if isinstance(error, DeviceArray) else out.numpy()

    def update(self, m, symm=None, m_local=None):
        """learner/backpropagation.py:37-42 without the host round trip: one fused kernel reads W, g (and
        the optimizer state), writes W (and state) and clears g (zero_gradients, :68-73).  With `symm` (data-parallel
        jobs) the same kernel first sums the gradient vectors of all ranks through their peer-mapped buffers."""
        L = self.layers
        L.flatten_parameters()
        n = L.flat_size
        if symm is not None:
            self.optimizer.optimize_allreduce(L.flat_weights, L.flat_grads, n, m, symm,
                                              m if m_local is None else m_local)
            return
        self.optimizer.optimize(L.flat_weights[0:n], L.flat_grads[0:n], m, zero_grad=True)

    def get_weights(self, unfold=True):
        return self.layers.get_weights(unfold=unfold)

    def set_weigts(self, ws, fold=True):   # (sic) the reference's spelling, learner/backpropagation.py:47
        self.layers.set_weights(ws=ws, fold=fold)

    set_weights = set_weigts

    def get_gradients(self, unfold=True):
        return self.layers.get_gradients(unfold=unfold)

    def zero_gradients(self):
        self.layers.flatten_parameters()
        self.layers.flat_grads.fill_zero()
        for layer in self.layers:
            if layer.weights is not None and not layer.trainable:
                layer.nabla_w.fill_zero()
                layer.nabla_b.fill_zero()

    @property
    def num_params(self):
        return self.layers.num_params

    @property
    def output_shape(self):
        return self.layers.output_shape

    @property
    def input_shape(self):
        return self.layers.input_shape
