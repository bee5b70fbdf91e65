"""gradientcheck/raw_gradients.py of the reference, with the numerical sweep batched on the device.

Reference (:15-42): for each of the nparams weights, two `set_weights` (a full host->layer copy), two forward passes
and two host cost evaluations.  Here one sweep iteration -- save w_i, w_i+eps, forward, cost, w_i-eps, forward, cost,
restore, advance -- is a fixed sequence of device launches that reads its parameter index from a counter in device
memory (`bf_gradcheck_perturb`, `bf_gradcheck_cost`): it is captured once in a CUDA graph and replayed nparams
times.  Nothing crosses PCIe during the sweep; the 3 x nparams doubles (two costs and the step actually taken) come
back in one copy."""
import ctypes

import numpy as np


def analytical_gradients(network, X, Y):
    """:4-12.  Gradients are zeroed first (Dense accumulates, layers/core.py:37-38)."""
    network.zero_gradients()
    preds = network.predict(X)
    network.backpropagate(network.cost.derivative(preds, Y))
    return network.get_gradients(unfold=True)


def _parameter_offsets(stack):
    """Position of every parameter of `get_weights(unfold=True)` inside the padded flat device vector."""
    stack.flatten_parameters()
    return np.concatenate([np.arange(off, off + size) for _, _, _, off, size in stack._segments]).astype(np.int32)


def numerical_gradients(network, X, Y, epsilon, use_graph=True):
    import torch
    from .. import _lib
    from ..device import as_device, call, stream_ptr
    stack = network.layers
    offsets = _parameter_offsets(stack)
    nparams = offsets.size
    kind = _lib.COST[network.cost.kind]
    xd, yd = as_device(np.asarray(X)), as_device(np.asarray(Y))
    dev = xd.t.device
    offs_d = torch.as_tensor(offsets, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    saved = torch.zeros(1, dtype=torch.float32, device=dev)
    record = torch.zeros(3 * nparams, dtype=torch.float64, device=dev)
    W = stack.flat_weights
    p = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731

    def iteration():
        for phase in (0, 1):
            call(_lib.lib.bf_gradcheck_perturb, W.ptr, p(offs_d), p(counter), p(saved), p(record), float(epsilon), phase, stream_ptr())
            preds = stack._fwd(xd)
            m = preds.shape[0]
            call(_lib.lib.bf_gradcheck_cost, kind, preds.ptr, yd.ptr, m, preds.size // m, p(record), p(counter), phase, stream_ptr())
        call(_lib.lib.bf_gradcheck_perturb, W.ptr, p(offs_d), p(counter), p(saved), p(record), float(epsilon), 2, stream_ptr())

    was_learning, stack.learning = stack.learning, False
    try:
        done = 0
        if nparams:
            iteration()                                  # eager: sizes the workspace, warms every kernel (parameter 0)
            done = 1
        graph = None
        if use_graph and nparams > 1 and not any(type(l).__name__ == "DropOut" for l in stack.layers):
            try:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    with torch.cuda.graph(graph, stream=side):
                        iteration()
                torch.cuda.current_stream().wait_stream(side)
            except Exception:                            # a layer that cannot be captured: plain launches
                graph = None
                torch.cuda.synchronize()
                counter.fill_(done)
        for _ in range(done, nparams):
            graph.replay() if graph is not None else iteration()
        torch.cuda.synchronize()
    finally:
        stack.learning = was_learning
    rec = record.cpu().numpy().reshape(nparams, 3)
    return (rec[:, 0] - rec[:, 1]) / rec[:, 2]
