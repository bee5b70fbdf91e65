"""model/layerstack.py of the reference: the ordered layer container + flat weight vector fold/unfold.

B200 layout: the parameters and gradients of all trainable layers live in TWO flat FP32 buffers in HBM
(`flat_weights`, `flat_grads`); each layer's weights / biases / nabla_w / nabla_b are views into them
(the reference converges to the same thing after its first set_weights, layers/abstract_layer.py:45-49).
Segments are padded to 64 floats so every view is 256-byte aligned (128-bit loads, TMA descriptors); the
padding stays zero and is invisible through get_weights / set_weights.  The optimizer step and the
data-parallel all-reduce therefore touch one contiguous vector and never round-trip through the host.
"""
import numpy as np
import torch

from .. import dist as _dist
from ..device import DeviceArray, as_device, note_mutation
from ..layers.abstract_layer import LayerBase
from ..layers.core import InputLayer

SEG_ALIGN = 64


class LayerStack:

    def __init__(self, input_shape, layers=()):
        self.input_shape = input_shape
        self.floatX = "float64"
        self.compiled = True
        self.layers = []
        self.architecture = []
        self.learning = False
        self._iterme = None
        self.flat_weights = None
        self.flat_grads = None
        self._segments = []
        self._flat_dirty = True
        self.symm = None
        self._add_input_layer(input_shape)
        for layer in layers:
            self.add(layer)

    def _add_input_layer(self, input_shape):
        if isinstance(input_shape, (int, np.integer)):
            input_shape = (int(input_shape),)
        inl = InputLayer(tuple(input_shape))
        self.layers.append(inl)
        self.architecture.append(str(inl))

    def add(self, layer):
        layer.connect(self)
        self.layers[-1]._next_layer = layer
        self.layers.append(layer)
        self.architecture.append(str(layer))
        self._flat_dirty = True

    def pop(self):
        self.layers.pop()
        self.layers[-1]._next_layer = None
        self.architecture.pop()
        self._flat_dirty = True

    # ---- flat device buffers
    def flatten_parameters(self):
        """(Re)build the flat buffers and re-point every trainable layer's arrays at views into them."""
        if not self._flat_dirty and self.flat_weights is not None:
            return
        segs, off = [], 0
        for layer in self.trainable_layers:
            for wname, gname in (("weights", "nabla_w"), ("biases", "nabla_b")):
                size = getattr(layer, wname).size
                segs.append((layer, wname, gname, off, size))
                off += (size + SEG_ALIGN - 1) // SEG_ALIGN * SEG_ALIGN
        total = max(off, SEG_ALIGN)
        # two extra aligned slots at the tail carry the per-step scalars (cost sum, hit count, local m)
        # so that the data-parallel path needs exactly ONE all-reduce per learn_batch
        self._scalar_off = total
        total += SEG_ALIGN
        W = torch.zeros(total, dtype=torch.float32, device="cuda")
        # data-parallel jobs own a symmetric (peer-mapped) receive buffer: the step tail then is one kernel that pushes
        # the gradient vector to every rank over NVLink, sums the copies and applies the optimizer (bf_opt_step_allreduce)
        self.symm = _dist.symmetric_recv(total)
        G = torch.zeros(total, dtype=torch.float32, device="cuda")
        for layer, wname, gname, o, size in segs:
            w_old, g_old = getattr(layer, wname), getattr(layer, gname)
            W[o:o + size].copy_(w_old.t.reshape(-1))
            G[o:o + size].copy_(g_old.t.reshape(-1))
            setattr(layer, wname, DeviceArray(W[o:o + size].view(w_old.shape)))
            setattr(layer, gname, DeviceArray(G[o:o + size].view(g_old.shape)))
        self.flat_weights, self.flat_grads, self._segments = DeviceArray(W), DeviceArray(G), segs
        self._flat_dirty = False
        note_mutation()

    @property
    def flat_size(self):
        self.flatten_parameters()
        return self._scalar_off

    def scalars(self):
        """View of the 3 step scalars riding at the tail of the gradient buffer."""
        self.flatten_parameters()
        return self.flat_grads[self._scalar_off:self._scalar_off + SEG_ALIGN]

    # ---- compute
    def _fwd(self, x, stop=None):
        """Forward through every layer (or the layers before index `stop`)."""
        for layer in (self.layers if stop is None else self.layers[:stop]):
            x = layer._fwd(x)
        return x

    def feedforward(self, X):
        out = self._fwd(as_device(X))
        return out if isinstance(X, DeviceArray) else out.numpy()

    predict = feedforward

    # ---- weights (host float64 in/out; model/layerstack.py:42-61)
    def get_states(self, unfold=True):
        hs = [np.asarray(layer.output) for layer in self.layers]
        return np.concatenate(hs) if unfold else hs

    def get_weights(self, unfold=True):
        if unfold and self.flat_weights is not None and not self._flat_dirty:
            host = self.flat_weights.numpy()          # ONE D2H of the padded vector, unpadded on the host
            return np.concatenate([host[o:o + n] for _, _, _, o, n in self._segments]) if self._segments \
                else np.zeros(0)
        ws = [layer.get_weights(unfold=unfold) for layer in self.layers if layer.trainable]
        return np.concatenate(ws) if unfold else ws

    def set_weights(self, ws, fold=True):
        trl = (l for l in self.layers if l.trainable)
        if fold:
            ws = np.asarray(ws)
            if self.flat_weights is not None and not self._flat_dirty:
                host = np.zeros(self.flat_weights.size, dtype=np.float32)
                start = 0
                for _, _, _, o, n in self._segments:
                    host[o:o + n] = ws[start:start + n]
                    start += n
                self.flat_weights.copy_from_host(host)   # ONE H2D
                return
            start = 0
            for layer in trl:
                end = start + layer.num_params
                layer.set_weights(ws[start:end])
                start = end
        else:
            for w, layer in zip(ws, trl):
                layer.set_weights(w, fold=False)

    def get_gradients(self, unfold=True):
        if unfold and self.flat_grads is not None and not self._flat_dirty:
            host = self.flat_grads.numpy()
            return np.concatenate([host[o:o + n] for _, _, _, o, n in self._segments]) if self._segments \
                else np.zeros(0)
        grads = [l.get_gradients(unfold=unfold) for l in self.layers if l.trainable]
        return np.concatenate(grads) if unfold else grads

    def describe(self):
        return "Architecture: " + "->".join(self.architecture),

    def reset(self):
        for layer in (l for l in self.layers if l.trainable):
            layer.shuffle()

    @property
    def outshape(self):
        return self.layers[-1].outshape

    output_shape = outshape

    @property
    def num_params(self):
        return sum(layer.num_params for layer in self.layers if layer.trainable)

    @property
    def output(self):
        return self.layers[-1].output

    @property
    def trainable_layers(self):
        return [layer for layer in self.layers if layer.trainable]

    def __iter__(self):
        return self

    def __next__(self) -> LayerBase:
        if self._iterme is None:
            self._iterme = iter(self.layers)
        try:
            return next(self._iterme)
        except StopIteration:
            self._iterme = None
            raise

    def __getitem__(self, item):
        return self.layers.__getitem__(item)
