"""The training / evaluation loops behind the reference's `Learner` API (learner/abstract_learner.py: fit,
fit_generator, epoch, evaluate, evaluate_stream, evaluate_batch, predict -- same names, arguments and results),
organised around what the device needs instead of around host arrays:

* `fit(X, Y, ...)` uploads the data set ONCE (`ResidentDataset`) and draws every mini-batch on the device: the
  reference reshuffles the host arrays on every pass (util/_nnutil.py:4-15, a full copy of X and Y per pass); here the
  same permutation is drawn from the same NumPy stream, but only its index vector travels and `bf_gather_rows` assembles
  the batch in HBM.  Unshuffled batches are views.  Data sets that do not fit the memory budget stream from the host.
* `epoch(generator, n)` drives a `StepPipeline`: a learner that can enqueue a step without waiting for it
  (`learn_batch_async` + `stage_batch`, Backpropagation) has batch i+1's host->device copy in flight on a copy stream
  while step i computes, and its scalars are read one step late from a pinned snapshot -- the device never waits for
  Python.  The generator is still drawn exactly n times and the history holds the same numbers in the same order as n
  serial `learn_batch` calls.
"""
import time

import numpy as np

from ..model.layerstack import LayerStack
from ..metrics import costs as _costs, metrics as _metrics
from ..util import batch_stream, MetricLogs


class ResidentDataset:
    """Arrays with a common first dimension kept in HBM; `batches()` is `util.batch_stream` over them (same batch
    sequence, same use of the global NumPy RNG) yielding DeviceArrays."""

    BUDGET_FRACTION = 0.5      # of the currently free device memory

    def __init__(self, arrays):
        from ..device import DeviceArray
        self.n = len(arrays[0])
        self.shapes = [tuple(np.shape(a)[1:]) for a in arrays]
        self.dev = [a if isinstance(a, DeviceArray) else DeviceArray.from_host(a) for a in arrays]

    @classmethod
    def fits(cls, arrays):
        import torch
        need = sum(4 * int(np.prod(np.shape(a))) for a in arrays)
        free, _ = torch.cuda.mem_get_info()
        return need <= cls.BUDGET_FRACTION * free

    def _take(self, order_dev, start, stop):
        import ctypes
        from .. import _lib
        from ..device import DeviceArray, call, stream_ptr
        out = []
        for arr, tail in zip(self.dev, self.shapes):
            dst = DeviceArray.empty(stop - start, *tail)
            row = int(np.prod(tail)) if tail else 1
            call(_lib.lib.bf_gather_rows, arr.ptr, ctypes.c_void_p(order_dev.data_ptr() + 4 * start), dst.ptr, stop - start, row,
                 stream_ptr())
            out.append(dst)
        return tuple(out)

    def batches(self, m, shuffle=True, infinite=True):
        import torch
        while True:
            order = np.arange(self.n)
            if shuffle:
                np.random.shuffle(order)                  # util/_nnutil.py:8-9: one permutation per pass
                order_dev = torch.as_tensor(order.astype(np.int32), device="cuda")
            for start in range(0, self.n, m):
                stop = min(start + m, self.n)
                if shuffle:
                    yield self._take(order_dev, start, stop)
                else:
                    yield tuple(arr[start:stop] for arr in self.dev)
            if not infinite:
                return


class StepPipeline:
    """Feeds batches to a learner and collects the per-step results in order.  With an asynchronous learner the
    result of step i is collected after step i+1 has been enqueued and batch i+2's upload has been started."""

    def __init__(self, learner, metrics, sink, learn_kw):
        self.learner, self.metrics, self.sink, self.kw = learner, metrics, sink, learn_kw
        self.overlapped = hasattr(learner, "learn_batch_async") and "w" not in learn_kw
        self.batch_size = 0

    SMALL_BATCH_BYTES = 1 << 20     # below this a batch goes straight into the step's input buffers (see _stage)

    def _stage(self, batch):
        from ..device import DeviceArray
        x, y = batch[0], batch[1]
        if isinstance(x, DeviceArray) and isinstance(y, DeviceArray):
            return x, y                                   # already in HBM (ResidentDataset): nothing to copy
        if getattr(x, "nbytes", 1 << 62) + getattr(y, "nbytes", 1 << 62) <= self.SMALL_BATCH_BYTES:
            # a few KB: the copy takes microseconds, while a side-stream upload (two allocations, an event, two more
            # device copies into the graph's input buffers) costs more host time than the whole device step
            return x, y
        return self.learner.stage_batch(x, y), None

    def run(self, generator, steps):
        if not self.overlapped:
            for _ in range(steps):
                batch = next(generator)
                self.batch_size = len(batch[0])
                self.sink(self.learner.learn_batch(*batch, metrics=self.metrics, **self.kw))
            return
        in_flight = None                                  # reader of the previous step's scalars
        staged = self._stage(next(generator)) if steps else None
        for i in range(steps):
            self.batch_size = len(staged[0])
            reader = self.learner.learn_batch_async(staged[0], staged[1], metrics=self.metrics, snapshot=True, **self.kw)
            staged = self._stage(next(generator)) if i + 1 < steps else None
            if in_flight is not None:
                self.sink(in_flight())
            in_flight = reader
        if in_flight is not None:
            self.sink(in_flight())


class Learner:

    def __init__(self, layerstack, cost="mse", name="", **kw):
        if not isinstance(layerstack, LayerStack):
            if "input_shape" not in kw:
                raise RuntimeError("Please supply input_shape as a keyword argument!")
            layerstack = LayerStack(kw["input_shape"], layers=layerstack)
        self.layers = layerstack
        self.name = name
        self.age = 0
        self.cost = _costs.get(cost)

    # ------------------------------------------------------------------ training
    def fit(self, X, Y, batch_size=20, epochs=30, metrics=(), validation=(), validation_steps=None, verbose=1,
            shuffle=True, **kw):
        """learner/abstract_learner.py:41-54."""
        return self.fit_generator(self._batch_source(X, Y, batch_size, shuffle), len(X) // batch_size, epochs,
                                  [_metrics.get(m) for m in metrics], validation, validation_steps, verbose, **kw)

    def _batch_source(self, X, Y, batch_size, shuffle, infinite=True):
        resident = getattr(self, "resident_data", True) and hasattr(self, "learn_batch_async")
        if resident and ResidentDataset.fits((X, Y)):
            return ResidentDataset((X, Y)).batches(batch_size, shuffle=shuffle, infinite=infinite)
        return batch_stream(X, Y, m=batch_size, shuffle=shuffle, infinite=infinite)

    def fit_generator(self, generator, lessons_per_epoch, epochs=30, metrics=(), validation=(),
                      validation_steps=None, verbose=1, **kw):
        """learner/abstract_learner.py:21-39."""
        metrics = [_metrics.get(m) for m in metrics]
        history = MetricLogs.from_metric_list(lessons_per_epoch, ("cost",), metrics)
        width = len(str(epochs))
        for number in range(1, epochs + 1):
            if verbose:
                print("Epoch {:>{w}}/{}".format(number, epochs, w=width))
            history.update(self.epoch(generator, updates_per_epoch=lessons_per_epoch, metrics=metrics, validation=validation,
                                      validation_steps=validation_steps, verbose=verbose, **kw))
        return history

    def epoch(self, generator, updates_per_epoch, metrics=(), validation=None, validation_steps=None, verbose=1, **kw):
        """learner/abstract_learner.py:56-87: `updates_per_epoch` lessons drawn from `generator`, then the validation
        pass; returns the epoch's MetricLogs."""
        began = time.time()
        metrics = [_metrics.get(m) for m in metrics]
        history = MetricLogs.from_metric_list(updates_per_epoch, ["cost"], metrics)

        def sink(step_metrics):
            history.record(step_metrics)
            if verbose:
                history.log(prefix="\rTraining ", end="", add_progress=True)

        self.layers.learning = True
        pipe = StepPipeline(self, metrics, sink, kw)
        try:
            pipe.run(generator, updates_per_epoch)
        finally:
            self.layers.learning = False
        if verbose and validation:
            if isinstance(validation, (tuple, list)):
                report = self.evaluate(*validation, batch_size=pipe.batch_size, metrics=metrics, verbose=False)
            elif validation_steps is None:
                raise RuntimeError("If validating on a stream, validation_steps must be set to a positive integer.")
            else:
                report = self.evaluate_stream(validation, validation_steps, metrics, verbose=False)
            report.log(prefix=" Validation ", suffix="", add_progress=False)
        if verbose:
            print(" took {:.2f} minutes".format((time.time() - began) / 60))
        self.age += updates_per_epoch
        return history

    def learn_batch(self, X, Y, metrics=(), **kw) -> dict:
        raise NotImplementedError

    # ------------------------------------------------------------------ inference / evaluation
    def predict(self, X):
        return self.layers.feedforward(X)

    def evaluate_batch(self, x, y, metrics=()):
        """learner/abstract_learner.py:92-99: cost and metrics of one batch, per sample."""
        from ..device import as_device
        count = len(x)
        xd, yd = as_device(x), as_device(y)
        preds = self.layers._fwd(xd)
        report = {"cost": self.cost(preds, yd) / count}
        for metric in metrics or ():
            report[str(metric).lower()] = metric(preds, yd) / count
        return report

    def evaluate_stream(self, stream, steps, metrics=(), verbose=False):
        """learner/abstract_learner.py:101-114: mean of `steps` batch reports."""
        metrics = [_metrics.get(m) for m in metrics]
        history = MetricLogs.from_metric_list(steps, ["cost"], metrics)
        for seen, (x, y) in enumerate(stream, start=1):
            history.record(self.evaluate_batch(x, y, metrics))
            if verbose:
                history.log("\r", end="")
            if seen >= steps:
                break
        if verbose:
            print()
        history.reduce_mean()
        return history

    def evaluate(self, X, Y, batch_size=32, metrics=(), verbose=False):
        """learner/abstract_learner.py:116-123."""
        total = X.shape[0]
        batch_size = min(batch_size, total)
        source = self._batch_source(X, Y, batch_size, shuffle=False, infinite=False)
        return self.evaluate_stream(source, int(round(total / batch_size)), metrics, verbose)

    # ------------------------------------------------------------------ shortcuts
    @property
    def output(self):
        return self.layers[-1].output

    @property
    def outshape(self):
        return self.layers.outshape

    @property
    def num_params(self):
        return self.layers.num_params

    @property
    def trainable_layers(self):
        return self.layers.trainable_layers
