"""learner/neuroevolution.py of the reference: a Population of flat weight genomes around a LayerStack."""
import numpy as np

from .abstract_learner import Learner
from .. import _lib
from ..device import DeviceArray, as_device, call, precision, stream_ptr
from ..evolution import Population
from ..layers.core import Dense


class NeuroEvolution(Learner):
    """When the stack is the fused-fitness MLP (Dense(h<=64, sigmoid) -> Dense(o<=16, softmax), cxent, default
    fitness) the population is DEVICE RESIDENT on the TF32 path (`device_resident=False` opts out): genomes never
    cross PCIe again after construction, `Population.update` evaluates them in place."""

    def __init__(self, layerstack, cost="mse", population_size=100, name="", **kw):
        super().__init__(layerstack, cost, name, **kw)
        ff = kw.pop("fitness_function", self.fitness)
        fw = kw.pop("fitness_weights", [1.])
        oa = kw.pop("on_accuracy", False)
        resident = kw.pop("device_resident", None)
        kw.pop("input_shape", None)
        self.on_accuracy = oa
        self.layers.flatten_parameters()
        dims = self._fused_mlp_dims()
        fused = ff == self.fitness and not oa and dims is not None
        custom_ops = kw.get("mate_function") is not None or kw.get("mutate_function") is not None
        if resident is None:
            resident = fused and not custom_ops and precision() == _lib.PREC["tf32"] and dims[0] % 4 == 0
        elif resident and not (fused and not custom_ops):
            raise ValueError("device_resident=True needs the fused MLP fitness (Dense-sigmoid -> Dense-softmax under cxent) "
                             "and the default mate / mutate operators")
        self.population = Population(loci=self.layers.num_params, fitness_function=ff, fitness_weights=fw,
                                     limit=population_size, device_resident=bool(resident), store_dims=dims,
                                     batch_fitness_function=(self._store_fitness if resident else self.batch_fitness if fused else None),
                                     **kw)

    @staticmethod
    def as_weights(genome):
        """learner/neuroevolution.py:20-22."""
        return (genome - 0.5) * 20.

    def learn_batch(self, X, Y, **kw):
        """learner/neuroevolution.py:24-32.  X, Y are uploaded ONCE here and stay in HBM for every fitness evaluation of
        the run (the reference re-reads the host arrays per individual)."""
        xd, yd = as_device(X), as_device(Y)
        self.population.run(epochs=kw.get("epochs", 1), survival_rate=kw.get("survival_rate", 0.8),
                            mutation_rate=kw.get("mutation_rate", 0.1), verbosity=kw.get("verbosity", 0), X=xd, Y=yd)
        self.layers.set_weights(self.as_weights(self.population.best))
        return self.population.mean_grade()

    def fitness(self, genome, X, Y):
        """What learner/neuroevolution.py:34-37 intends (as shipped it raises TypeError):
        weights <- genome, forward on (X, Y), fitness = cost/m (or 1-accuracy with on_accuracy)."""
        self.layers.set_weights(self.as_weights(np.asarray(genome)))
        x, y = as_device(X), as_device(Y)
        preds = self.layers._fwd(x)
        if self.on_accuracy:
            from ..metrics import metrics as _metrics
            return 1. - _metrics.accuracy(preds, y) / len(x)
        return self.cost(preds, y) / len(x)

    # ------------------------------------------------------------------ fused fan-out for the config-5 MLP
    def _fused_mlp_dims(self):
        """(nin, nh, nout) when the stack is Dense(nh, sigmoid) -> Dense(nout, softmax) under cxent."""
        ls = self.layers.layers[1:]
        if len(ls) != 2 or not all(isinstance(l, Dense) and l.trainable for l in ls):
            return None
        if str(ls[0].activation) != "sigmoid" or str(ls[1].activation) != "softmax" or self.cost.kind != "cxent":
            return None
        if ls[1].neurons > 16 or ls[0].neurons > 64:
            return None
        return ls[0].weights.shape[0], ls[0].neurons, ls[1].neurons

    def batch_fitness(self, genomes, X, Y):
        """`genomes` (n, loci) host array evaluated on the shared (X, Y) in one device pass (bf_fitness_mlp).  X and Y are
        uploaded per call unless they already are DeviceArrays (learn_batch passes those): a host array may have been
        refilled in place since the last call, so nothing is cached by identity."""
        nin, nh, nout = self._fused_mlp_dims()
        g = np.ascontiguousarray(genomes, dtype=np.float32)
        x, y = as_device(X), as_device(Y)
        gd = DeviceArray.from_host(g)
        out = DeviceArray.empty(len(g))
        call(_lib.lib.bf_fitness_mlp, gd.ptr, x.ptr, y.ptr, out.ptr, len(g), x.shape[0], nin, nh, nout, precision(),
             stream_ptr())
        return out.numpy().reshape(-1, 1)

    def _store_fitness(self, store, inds, X, Y):
        """Fitness of the individuals `inds` of a device-resident store (bf_population_fitness): reads the genomes and
        their K-major first-layer operand in place; only len(inds) fitness values cross PCIe."""
        return store.fitness(inds, as_device(X), as_device(Y)).reshape(-1, 1)
