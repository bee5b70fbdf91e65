"""evolution/_evolution.py of the reference: a generic, fitness-MINIMISING genetic algorithm.

Scope: the GA operators (selection, mating, mutation) stay host NumPy exactly as in the
reference; what moves to the B200 is the fitness-evaluation fan-out of `update` -- the reference's serial
per-individual loop (:90-95).  A `batch_fitness_function(genomes (n,loci), *args, **kw) -> (n, n_fitness)`
evaluates all selected individuals in one device pass; under torch.distributed the selected individuals
are sharded over ranks and only the fitness values are gathered.
"""
import warnings

import numpy as np

from .. import dist as _dist


class Population:

    def __init__(self, loci, fitness_function, fitness_weights=None, limit=100, grade_function=None,
                 mate_function=None, mutate_function=None, batch_fitness_function=None, **kw):
        del kw
        self.fitness = fitness_function
        self.batch_fitness = batch_fitness_function
        self.limit = limit
        self.grades = np.zeros(limit)
        if grade_function is None:
            if fitness_weights is None:
                fitness_weights = np.array([1.])
            self.grade_function = self._default_grade_function
        else:
            if fitness_weights is not None:
                warnings.warn("grade_function supplied, fitness_weights ignored!")
            self.grade_function = grade_function
            if fitness_weights is None:
                fitness_weights = np.array([1.])
        self.mate_function = self._default_mate_function if mate_function is None else mate_function
        self.mutation = self._default_mutate_function if mutate_function is None else mutate_function
        self.fitnesses = np.zeros((limit, len(fitness_weights)))
        self.fitness_w = np.asarray(fitness_weights, dtype=np.float64)
        self.individuals = np.random.uniform(size=(limit, loci))
        self.age = 0
        self.initialized = False

    @classmethod
    def from_capsule(cls, capsule, fitness_fn, grade_fn=None, mate_fn=None, mutate_fn=None):
        indiv = capsule["individuals"]
        limit, loci = indiv.shape
        pop = cls(loci, fitness_fn, capsule["fitness_w"], limit, grade_fn, mate_fn, mutate_fn)
        pop.individuals = indiv
        pop.grades = capsule["grades"]
        pop.age = capsule["age"]
        return pop

    def capsule(self):
        return dict(age=self.age, individuals=self.individuals, grades=self.grades, fitness_w=self.fitness_w)

    # ------------------------------------------------------------------ fitness fan-out (the hot part)
    def update(self, inds=None, verbose=0, *fitness_args, **fitness_kw):
        """evolution/_evolution.py:83-97.  Evaluates `inds` (all individuals when None)."""
        if inds is None:
            inds = np.arange(self.individuals.shape[0])
        else:
            inds = np.asarray(inds).ravel().astype(np.int64)
        if inds.size == 0:
            return
        ws = _dist.world_size()
        lo, hi = _dist.shard_bounds(inds.size) if ws > 1 else (0, inds.size)
        mine = inds[lo:hi]
        if self.batch_fitness is not None:
            fit = np.asarray(self.batch_fitness(self.individuals[mine], *fitness_args, **fitness_kw), dtype=np.float64)
            fit = fit.reshape(len(mine), -1)
        else:
            fit = np.zeros((len(mine), self.fitnesses.shape[1]))
            for row, ind in enumerate(mine):
                fit[row] = self.fitness(self.individuals[ind], *fitness_args, **fitness_kw)
        if ws > 1:
            nfit = self.fitnesses.shape[1]
            counts = [(_dist.shard_bounds(inds.size, r, ws)[1] - _dist.shard_bounds(inds.size, r, ws)[0]) * nfit
                      for r in range(ws)]
            fit = _dist.gather_concat(fit.ravel(), counts).reshape(inds.size, nfit)
        self.fitnesses[inds] = fit
        for row, ind in enumerate(inds):
            self.grades[ind] = self.grade_function(self.fitnesses[ind])
        if verbose:
            print("\rUpdating {0}/{1}".format(self.limit, self.limit))

    # ------------------------------------------------------------------ GA operators (host, as the reference)
    def get_candidates(self, survivors=None):
        """evolution/_evolution.py:99-134: offspring from randomly paired survivors, accepted with the
        mean of the parents' rescaled grades as probability."""
        if isinstance(survivors, tuple):
            survivors = survivors[0]
        if survivors is None or survivors.size == 0:
            survivors = np.where(self.selection(0.3))[0]

        def pairs():
            while 1:
                left, right = np.copy(survivors), np.copy(survivors)
                np.random.shuffle(left)
                np.random.shuffle(right)
                for pair in zip(left, right):
                    yield pair

        prbs = rescale(self.grades)
        candidates = np.zeros_like(self.individuals)
        made = 0
        for left, right in pairs():
            if made >= self.limit:
                break
            if np.mean([prbs[left], prbs[right]]) < np.random.uniform():
                continue
            candidates[made] = self.mate_function(self.individuals[left], self.individuals[right])
            made += 1
        return candidates

    def selection(self, rate):
        """evolution/_evolution.py:136-141: mask of the best (lowest-grade) `rate` fraction."""
        survmask = np.zeros_like(self.grades)
        if rate:
            survmask[np.argsort(self.grades)[:int(self.limit * rate)]] = 1.
        return survmask

    def run(self, epochs, survival_rate=0.5, mutation_rate=0.1, force_update_at_every=0, verbosity=1,
            *fitness_args, **fitness_kwargs):
        """evolution/_evolution.py:143-208 -> (mean grades, grade stds, best grades) per epoch."""
        if not self.initialized:
            self.update(None, verbosity, *fitness_args, **fitness_kwargs)
            if verbosity:
                print("EVOLUTION: initial mean grade :", self.grades.mean())
                print("EVOLUTION: initial std of mean:", self.grades.std())
                print("EVOLUTION: initial best grade :", self.grades.min())
        mean_grades, grades_std, bests = [], [], []
        mutation_rate /= self.individuals.shape[-1]
        for epoch in range(1, epochs + 1):
            if verbosity:
                print("-" * 50)
                print("Epoch {0:>{w}}/{1}".format(epoch, epochs, w=len(str(epochs))))
            survmask = self.selection(survival_rate)
            candidates = self.get_candidates(survivors=np.where(survmask))
            newgen = np.where(survmask[:, None], self.individuals, candidates)
            newgen, mutants = self.mutation(mutation_rate, newgen)
            self.individuals = newgen
            if force_update_at_every and epoch % force_update_at_every == 0:
                inds = None
            else:
                inds = np.unique(np.append(np.where(survmask)[0], mutants))
            self.update(inds, verbosity, *fitness_args, **fitness_kwargs)
            if verbosity:
                self.describe(verbosity - 1)
            mean_grades.append(self.grades.mean())
            grades_std.append(self.grades.std())
            bests.append(self.grades.min())
            self.age += 1
        if verbosity:
            print()
        return np.array(mean_grades), np.array(grades_std), np.array(bests)

    def total_grade(self):
        return self.grades.sum()

    def mean_grade(self):
        return self.grades.mean()

    def describe(self, show=0):
        best = self.grades.argmin()
        lines = ["-" * 50]
        for i, index in enumerate(np.argsort(self.grades)[:show], start=1):
            lines.append("TOP {}: F = {} G = {:.4f}".format(i, self.fitnesses[index], self.grades[index]))
        lines.append("Best Grade : {:.4f} Fitnesses: {}".format(self.grades[best], list(self.fitnesses[best])))
        lines.append("Mean Grade : {:.4f}, STD: {:.4f}".format(self.grades.mean(), self.grades.std()))
        print("\n".join(lines))

    @property
    def best(self):
        return self.individuals[np.argmin(self.grades)]

    @staticmethod
    def _default_mate_function(gen1, gen2):
        return np.where(np.random.uniform(size=gen1.shape) < 0.5, gen1, gen2)

    def _default_grade_function(self, fitnesses):
        return np.dot(fitnesses, self.fitness_w)

    @staticmethod
    def _default_mutate_function(rate, individuals):
        if rate:
            mut_mask = np.random.uniform(size=individuals.shape) < rate
            mutants = np.where(mut_mask.sum(axis=1))[0]
            individuals[np.nonzero(mut_mask)] = np.random.uniform(size=(mut_mask.sum(),))
        else:
            mutants = np.array([], dtype=int)
        return individuals, mutants


def to_phenotype(ind, ranges):
    if len(ranges) != ind.shape[0]:
        raise RuntimeError("Specified ranges are incompatible with the supplied individuals")
    phenotype = ind.copy()
    for locus, (mini, maxi) in enumerate(ranges):
        phenotype[locus] = phenotype[locus] * (maxi - mini) + mini
    return phenotype


def rescale(vector):
    output = vector - vector.min()
    output = output / (output.max() + 1e-8)
    return output * 0.95 + 0.05
