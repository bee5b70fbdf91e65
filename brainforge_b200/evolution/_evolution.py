"""evolution/_evolution.py of the reference: a generic, fitness-MINIMISING genetic algorithm -- with the population
resident in HBM.

Two stores behind one `Population` API (names, arguments and return values of the reference's class):

* host store (`device_resident=False`, or whenever the caller supplies its own mate / mutate functions or a plain
  per-individual fitness function): genomes are a host float64 array and the GA operators are NumPy, like the
  reference; only the fitness fan-out of `update` (the reference's serial loop, :83-97) goes to the device when a
  `batch_fitness_function` is given.
* device store (`device_resident=True`, what `NeuroEvolution` asks for when its fused MLP fitness applies): the genome
  matrix lives in HBM as FP32 (P, ldg) next to the K-major operand of the fitness kernel (`GenomeStore`).  `update`
  evaluates in place (`bf_population_fitness`), a generation step is one pass over the store
  (`bf_population_next_generation`: survivors copied, the rest mated from two parents, every locus mutated with
  probability rate) followed by a re-layout of the rows that changed.  What stays on the host is what works on P
  scalars: selection by grade (:136-141) and the accept/reject pairing of parents (:99-134), consuming the global
  NumPy RNG draw for draw like the reference; the per-locus draws (P x loci coins and mutation uniforms, the
  reference's bottleneck: SURVEY 8f row 1) come from Philox on the device.
  `population.individuals` stays readable and assignable (a host float64 copy is fetched / uploaded on demand).

Under torch.distributed the individuals to (re)evaluate are sharded over ranks and only fitness values are gathered
(SURVEY 8e row 2); with a device store every rank holds the whole population and steps it with the same seed.
"""
import ctypes
import warnings

import numpy as np

from .. import dist as _dist


class GenomeStore:
    """(P, loci) genomes in HBM: `g` FP32 (P, ldg) double-buffered across generations, `bt` the (P*64, nin) K-major
    TF32 first-layer operand of the fitness kernel (include/bf_b200.h: bf_population_*)."""

    def __init__(self, host_genomes, nin, nh, nout):
        import torch
        from ..device import DeviceArray
        host = np.ascontiguousarray(host_genomes, dtype=np.float32)
        self.P, self.loci = host.shape
        self.nin, self.nh, self.nout = nin, nh, nout
        self.ldg = (self.loci + 3) // 4 * 4
        self.g = DeviceArray.zeros(self.P, self.ldg)
        self.g_next = DeviceArray.zeros(self.P, self.ldg)
        self.bt = DeviceArray.zeros(self.P * 64, nin)
        self._torch = torch
        self.upload(host)

    def upload(self, host):
        host = np.asarray(host)
        assert host.shape == (self.P, self.loci), (host.shape, (self.P, self.loci))
        padded = np.zeros((self.P, self.ldg), dtype=np.float32)
        padded[:, :self.loci] = host
        self.g.copy_from_host(padded)
        self.relayout(None)

    def download(self):
        return self.g.numpy()[:, :self.loci]

    def row(self, index):
        """One genome as a host float64 vector (D2H of a single row)."""
        return self.g[int(index)].numpy()[:self.loci]

    def _ints(self, values):
        return self._torch.as_tensor(np.ascontiguousarray(values, dtype=np.int32), device="cuda")

    def relayout(self, inds):
        from .. import _lib
        from ..device import call, stream_ptr
        if inds is None:
            call(_lib.lib.bf_population_relayout, self.g.ptr, self.ldg, self.bt.ptr, None, self.P, self.nin, self.nh, stream_ptr())
        elif len(inds):
            idx = self._ints(inds)
            call(_lib.lib.bf_population_relayout, self.g.ptr, self.ldg, self.bt.ptr, ctypes.c_void_p(idx.data_ptr()), len(inds),
                 self.nin, self.nh, stream_ptr())

    def fitness(self, inds, x, y):
        """cxent/m of the listed individuals (all when None) on the device arrays (x, y) -> host float64 vector."""
        from .. import _lib
        from ..device import DeviceArray, call, precision, stream_ptr
        n = self.P if inds is None else len(inds)
        out = DeviceArray.empty(n)
        idx = None if inds is None else self._ints(inds)
        call(_lib.lib.bf_population_fitness, self.g.ptr, self.ldg, self.P, self.bt.ptr,
             None if idx is None else ctypes.c_void_p(idx.data_ptr()), n, x.ptr, y.ptr, out.ptr, x.shape[0], self.nin, self.nh,
             self.nout, precision(), stream_ptr())
        return out.numpy()

    def next_generation(self, survive, left, right, rate, seed, generation, injected=None):
        """One pass: survivors kept, the others mated from (left[i], right[i]), every locus mutated with probability
        `rate`.  Returns the indices of the mutated rows.  `injected` = (coin, mut_u, mut_val) host arrays (P, loci)
        replaces the device RNG (parity tests)."""
        from .. import _lib
        from ..device import DeviceArray, call, stream_ptr
        torch = self._torch
        sv = torch.as_tensor(np.ascontiguousarray(survive, dtype=np.uint8), device="cuda")
        lf, rt = self._ints(left), self._ints(right)
        mutant = torch.zeros(self.P, dtype=torch.uint8, device="cuda")
        extra = [None, None, None]
        if injected is not None:
            extra = [DeviceArray.from_host(np.ascontiguousarray(a, dtype=np.float32)) for a in injected]
        call(_lib.lib.bf_population_next_generation, self.g.ptr, self.g_next.ptr, self.ldg, self.loci, self.P,
             ctypes.c_void_p(sv.data_ptr()), ctypes.c_void_p(lf.data_ptr()), ctypes.c_void_p(rt.data_ptr()), float(rate),
             int(seed) & 0xFFFFFFFF, int(generation) & 0xFFFFFFFF, *[e.ptr if e is not None else None for e in extra],
             ctypes.c_void_p(mutant.data_ptr()), stream_ptr())
        self.g, self.g_next = self.g_next, self.g
        mutants = np.flatnonzero(mutant.cpu().numpy())
        changed = np.union1d(np.flatnonzero(np.asarray(survive) == 0), mutants)
        self.relayout(changed)
        return mutants


class Population:

    def __init__(self, loci, fitness_function, fitness_weights=None, limit=100, grade_function=None,
                 mate_function=None, mutate_function=None, batch_fitness_function=None, device_resident=False,
                 store_dims=None, seed=None, **kw):
        del kw
        self.fitness = fitness_function
        self.batch_fitness = batch_fitness_function
        self.limit = limit
        self.grades = np.zeros(limit)
        custom_grade = grade_function is not None
        if custom_grade and fitness_weights is not None:
            warnings.warn("grade_function supplied, fitness_weights ignored!")
        if fitness_weights is None:
            fitness_weights = np.array([1.])
        self.fitness_w = np.asarray(fitness_weights, dtype=np.float64)
        self.grade_function = grade_function if custom_grade else self._default_grade_function
        self._custom_operators = mate_function is not None or mutate_function is not None
        self.mate_function = mate_function or self._default_mate_function
        self.mutation = mutate_function or self._default_mutate_function
        self.fitnesses = np.zeros((limit, len(self.fitness_w)))
        first = np.random.uniform(size=(limit, loci))       # the reference's initial draw (:60), same RNG consumption
        self.age = 0
        self.initialized = False
        self._store = None
        self._host = first
        # device RNG key: taken from the bits of the first genome (no extra draw from the global NumPy stream, equal on
        # every rank that seeded NumPy identically)
        self._seed = int(first.ravel()[:1].view(np.uint64)[0] & 0x7FFFFFFF) if seed is None else int(seed)
        if device_resident and not self._custom_operators:
            if store_dims is None:
                raise ValueError("device_resident=True needs store_dims=(nin, nh, nout) of the fused MLP fitness")
            self._store = GenomeStore(first, *store_dims)
            self._host = None

    # ------------------------------------------------------------------ the genome matrix, wherever it lives
    @property
    def device_resident(self):
        return self._store is not None

    @property
    def individuals(self):
        """Host float64 (limit, loci) view of the population (a fresh copy when the store is on the device: write it
        back with `population.individuals = array`)."""
        if self._store is not None:
            return self._store.download()
        return self._host

    @individuals.setter
    def individuals(self, value):
        if self._store is not None:
            self._store.upload(value)
        else:
            self._host = value

    @property
    def best(self):
        winner = int(np.argmin(self.grades))
        return self._store.row(winner) if self._store is not None else self._host[winner]

    @classmethod
    def from_capsule(cls, capsule, fitness_fn, grade_fn=None, mate_fn=None, mutate_fn=None):
        limit, loci = capsule["individuals"].shape
        pop = cls(loci, fitness_fn, capsule["fitness_w"], limit, grade_fn, mate_fn, mutate_fn)
        pop.individuals, pop.grades, pop.age = capsule["individuals"], capsule["grades"], capsule["age"]
        return pop

    def capsule(self):
        return dict(age=self.age, individuals=self.individuals, grades=self.grades, fitness_w=self.fitness_w)

    # ------------------------------------------------------------------ fitness fan-out (the hot part)
    def update(self, inds=None, verbose=0, *fitness_args, **fitness_kw):
        """evolution/_evolution.py:83-97: (re)evaluate `inds` (everybody when None) and refresh their grades."""
        inds = np.arange(self.limit) if inds is None else np.asarray(inds).ravel().astype(np.int64)
        if inds.size == 0:
            return
        ws = _dist.world_size()
        lo, hi = _dist.shard_bounds(inds.size) if ws > 1 else (0, inds.size)
        mine = inds[lo:hi]
        nfit = self.fitnesses.shape[1]
        if self._store is not None:
            fit = np.asarray(self.batch_fitness(self._store, mine, *fitness_args, **fitness_kw), dtype=np.float64)
        elif self.batch_fitness is not None:
            fit = np.asarray(self.batch_fitness(self._host[mine], *fitness_args, **fitness_kw), dtype=np.float64)
        else:
            fit = np.array([np.atleast_1d(self.fitness(self._host[i], *fitness_args, **fitness_kw)) for i in mine],
                           dtype=np.float64).reshape(len(mine), nfit)
        fit = fit.reshape(len(mine), nfit)
        if ws > 1:
            counts = [(b - a) * nfit for a, b in (_dist.shard_bounds(inds.size, r, ws) for r in range(ws))]
            fit = _dist.gather_concat(fit.ravel(), counts).reshape(inds.size, nfit)
        self.fitnesses[inds] = fit
        if self.grade_function == self._default_grade_function:
            self.grades[inds] = fit @ self.fitness_w
        else:
            for i in inds:
                self.grades[i] = self.grade_function(self.fitnesses[i])
        if verbose:
            print("\rUpdating {0}/{0}".format(self.limit))

    # ------------------------------------------------------------------ selection / pairing (host: P scalars)
    def selection(self, rate):
        """evolution/_evolution.py:136-141: 0/1 mask of the best (lowest-grade) `rate` fraction."""
        mask = np.zeros_like(self.grades)
        keep = int(self.limit * rate) if rate else 0
        if keep:
            mask[np.argsort(self.grades)[:keep]] = 1.
        return mask

    def _resolve_survivors(self, survivors):
        if isinstance(survivors, tuple):
            survivors = survivors[0]
        if survivors is None or survivors.size == 0:
            survivors = np.where(self.selection(0.3))[0]
        return survivors

    def pair_parents(self, survivors=None):
        """The accept/reject loop of get_candidates (:99-134) without the matings: `limit` (left, right) parent pairs.
        Parents are drawn by shuffling the survivors twice per pass; a pair is accepted with probability = mean of
        the parents' rescaled grades.  RNG use is the reference's, draw for draw: two shuffles per pass, one scalar
        uniform per examined pair (taken a pass at a time, the surplus of the last pass handed back by rewinding the
        generator)."""
        survivors = self._resolve_survivors(survivors)
        prbs = rescale(self.grades)
        left, right = np.empty(self.limit, dtype=np.int64), np.empty(self.limit, dtype=np.int64)
        made, lmt = 0, len(survivors)
        while made < self.limit:
            a, b = np.copy(survivors), np.copy(survivors)
            np.random.shuffle(a)
            np.random.shuffle(b)
            state = np.random.get_state()
            rolls = np.random.uniform(size=lmt)
            ok = ~((prbs[a] + prbs[b]) / 2. < rolls)
            taken = np.flatnonzero(ok)
            need = self.limit - made
            if len(taken) >= need:
                last = taken[need - 1]                    # the reference stops examining pairs right here
                taken = taken[:need]
                np.random.set_state(state)
                np.random.uniform(size=int(last) + 1)
                if last == lmt - 1:                       # ... but its pair generator has already reshuffled for the
                    np.random.shuffle(np.copy(survivors))  # next pass when the limit falls on the last pair of one (:105-118)
                    np.random.shuffle(np.copy(survivors))
            left[made:made + len(taken)] = a[taken]
            right[made:made + len(taken)] = b[taken]
            made += len(taken)
        return left, right

    def get_candidates(self, survivors=None):
        """evolution/_evolution.py:99-134 on the host store: `limit` offspring of accepted parent pairs."""
        if self._store is not None:
            raise RuntimeError("get_candidates materialises a host matrix; a device-resident population mates inside "
                               "run() (bf_population_next_generation)")
        survivors = self._resolve_survivors(survivors)
        prbs = rescale(self.grades)
        candidates = np.zeros_like(self._host)
        made = 0
        while True:       # one pass over two fresh shuffles of the survivors; the limit is checked on the NEXT pair drawn,
            a, b = np.copy(survivors), np.copy(survivors)      # i.e. after a reshuffle if a pass just ended (:105-118,124-126)
            np.random.shuffle(a)
            np.random.shuffle(b)
            for l, r in zip(a, b):
                if made >= self.limit:
                    return candidates
                if np.mean([prbs[l], prbs[r]]) < np.random.uniform():
                    continue
                candidates[made] = self.mate_function(self._host[l], self._host[r])
                made += 1

    # ------------------------------------------------------------------ the generation loop
    def run(self, epochs, survival_rate=0.5, mutation_rate=0.1, force_update_at_every=0, verbosity=1,
            *fitness_args, **fitness_kwargs):
        """evolution/_evolution.py:143-208 -> (mean grades, grade stds, best grades), one entry per epoch."""
        if not self.initialized:
            self.update(None, verbosity, *fitness_args, **fitness_kwargs)
            if verbosity:
                print("EVOLUTION: initial mean grade :", self.grades.mean())
                print("EVOLUTION: initial std of mean:", self.grades.std())
                print("EVOLUTION: initial best grade :", self.grades.min())
        loci = self._store.loci if self._store is not None else self._host.shape[-1]
        per_locus_rate = mutation_rate / loci                                    # :176
        log = {"mean": [], "std": [], "best": []}
        for epoch in range(1, epochs + 1):
            if verbosity:
                print("-" * 50)
                print("Epoch {0:>{w}}/{1}".format(epoch, epochs, w=len(str(epochs))))
            survmask = self.selection(survival_rate)
            if self._store is not None:
                left, right = self.pair_parents(np.where(survmask)[0])
                mutants = self._store.next_generation(survmask, left, right, per_locus_rate, self._seed, self.age)
            else:
                candidates = self.get_candidates(survivors=np.where(survmask))
                newgen = np.where(survmask[:, None], self._host, candidates)
                self._host, mutants = self.mutation(per_locus_rate, newgen)
            if force_update_at_every and epoch % force_update_at_every == 0:
                inds = None
            else:
                inds = np.unique(np.append(np.where(survmask)[0], mutants))     # :185-189 (sic: the survivors)
            self.update(inds, verbosity, *fitness_args, **fitness_kwargs)
            if verbosity:
                self.describe(verbosity - 1)
            log["mean"].append(self.grades.mean())
            log["std"].append(self.grades.std())
            log["best"].append(self.grades.min())
            self.age += 1
        if verbosity:
            print()
        return np.array(log["mean"]), np.array(log["std"]), np.array(log["best"])

    def total_grade(self):
        return self.grades.sum()

    def mean_grade(self):
        return self.grades.mean()

    def describe(self, show=0):
        order = np.argsort(self.grades)
        report = ["-" * 50]
        report += ["TOP {}: F = {} G = {:.4f}".format(rank, self.fitnesses[i], self.grades[i])
                   for rank, i in enumerate(order[:show], start=1)]
        report.append("Best Grade : {:.4f} Fitnesses: {}".format(self.grades[order[0]], list(self.fitnesses[order[0]])))
        report.append("Mean Grade : {:.4f}, STD: {:.4f}".format(self.grades.mean(), self.grades.std()))
        print("\n".join(report))

    # ------------------------------------------------------------------ default operators of the host store
    @staticmethod
    def _default_mate_function(gen1, gen2):
        """:245-247: uniform crossover."""
        return np.where(np.random.uniform(size=gen1.shape) < 0.5, gen1, gen2)

    def _default_grade_function(self, fitnesses):
        return np.dot(fitnesses, self.fitness_w)

    @staticmethod
    def _default_mutate_function(rate, individuals):
        """:253-261: every locus redrawn with probability `rate`; returns (individuals, indices of mutated rows)."""
        if not rate:
            return individuals, np.array([], dtype=int)
        hit = np.random.uniform(size=individuals.shape) < rate
        individuals[np.nonzero(hit)] = np.random.uniform(size=(hit.sum(),))
        return individuals, np.where(hit.sum(axis=1))[0]


def to_phenotype(ind, ranges):
    if len(ranges) != ind.shape[0]:
        raise RuntimeError("Specified ranges are incompatible with the supplied individuals")
    lo = np.array([r[0] for r in ranges], dtype=np.float64)
    hi = np.array([r[1] for r in ranges], dtype=np.float64)
    shape = (-1,) + (1,) * (ind.ndim - 1)
    return ind * (hi - lo).reshape(shape) + lo.reshape(shape)


def rescale(vector):
    """:274-279: grades mapped onto [0.05, 1.0]."""
    shifted = vector - vector.min()
    return shifted / (shifted.max() + 1e-8) * 0.95 + 0.05
