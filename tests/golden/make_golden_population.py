"""Golden fixture for the device-resident Population, generated from the REAL reference.

    python tests/golden/make_golden_population.py        (authoring container only: needs /root/reference)

The reference's `Population.run` (evolution/_evolution.py:143-208) is executed UNMODIFIED for three epochs on a small
neuro-evolution problem (4-5-3 MLP, 43 loci, 16 individuals).  Its per-locus random draws are made observable through
the hooks the class itself offers -- `mate_function=` and `mutate_function=` -- which here are the reference's own
default operators (_evolution.py:245-261) drawing from a SEPARATE, recorded RandomState, so that the global NumPy
stream carries only what stays on the host in the B200 back-end (the shuffles and accept/reject rolls of
get_candidates, :99-134).  Recorded per epoch: survivor mask, the (left, right) parents of every candidate, the coin /
mutation-mask / mutation-value draws as dense (P, loci) arrays, the resulting individuals, the individuals re-evaluated
and the grades.  The B200 tests replay the draws through bf_population_next_generation and compare bit for bit.
"""
import os
import sys
import tempfile

os.environ["HOME"] = tempfile.mkdtemp(prefix="bfhome_")
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

np.asscalar = lambda a: a.item()  # noqa: E731

from brainforge.evolution import Population  # noqa: E402
from brainforge.learner.neuroevolution import NeuroEvolution  # noqa: E402
from brainforge.layers import Dense  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NIN, NH, NOUT, P, N, EPOCHS = 4, 5, 3, 16, 32, 3
SURVIVAL, MUTATION = 0.5, 0.8
SEED = 77


def main():
    np.random.seed(SEED)
    ne = NeuroEvolution(input_shape=(NIN,), layerstack=[Dense(NH, activation="sigmoid"), Dense(NOUT, activation="softmax")],
                        cost="cxent", population_size=2)
    loci = ne.layers.num_params
    rs = np.random.RandomState(5)
    X = rs.randn(N, NIN)
    Y = np.eye(NOUT)[rs.randint(0, NOUT, N)]

    def fitness(genome, X, Y):                       # what learner/neuroevolution.py:34-37 intends
        ne.layers.set_weights(ne.as_weights(genome))
        return (ne.cost(ne.layers.feedforward(X), Y) / len(X),)

    side = np.random.RandomState(99)                 # the per-locus draws, observable
    rec = {"coin": [], "pairs": [], "mut_u": None, "mut_val": None}
    pop_ref = {}

    def mate(g1, g2):                                # _evolution.py:245-247 with the draws recorded
        pop = pop_ref["pop"]
        base = pop.individuals.__array_interface__["data"][0]
        row = pop.individuals.strides[0]
        li = (g1.__array_interface__["data"][0] - base) // row
        ri = (g2.__array_interface__["data"][0] - base) // row
        assert np.shares_memory(g1, pop.individuals) and np.array_equal(pop.individuals[li], g1)
        u = side.uniform(size=g1.shape)
        rec["coin"].append(u)
        rec["pairs"].append((li, ri))
        return np.where(u < 0.5, g1, g2)

    def mutate(rate, individuals):                   # _evolution.py:253-261 with the draws recorded
        u = side.uniform(size=individuals.shape)
        mask = u < rate
        assert np.abs(u - rate).min() > 1e-7         # no draw sits on the float32 / float64 boundary of the comparison
        vals = side.uniform(size=(mask.sum(),)).astype(np.float32).astype(np.float64)   # FP32-representable: the device stores FP32
        dense = np.zeros(individuals.shape)
        dense[np.nonzero(mask)] = vals
        rec["mut_u"], rec["mut_val"] = u, dense
        mutants = np.where(mask.sum(axis=1))[0]
        individuals[np.nonzero(mask)] = vals
        return individuals, mutants

    np.random.seed(SEED + 1)
    pop = Population(loci=loci, fitness_function=fitness, fitness_weights=np.array([1.]), limit=P, mate_function=mate,
                     mutate_function=mutate)
    pop.individuals = pop.individuals.astype(np.float32).astype(np.float64)      # FP32-representable start
    pop_ref["pop"] = pop
    out = dict(X=X, Y=Y, dims=np.array([NIN, NH, NOUT]), individuals0=pop.individuals.copy(), survival=SURVIVAL,
               mutation=MUTATION, seed=SEED + 1, epochs=EPOCHS)
    np.random.seed(SEED + 2)                         # the host-side stream the B200 back-end must consume identically
    pop.update(None, 0, X=X, Y=Y)
    out["grades0"] = pop.grades.copy()
    means, stds, bests = [], [], []
    for e in range(EPOCHS):
        rec["coin"], rec["pairs"] = [], []
        before = pop.individuals.copy()
        survmask = pop.selection(SURVIVAL)
        pop.initialized = True                       # run() would re-evaluate everybody first (same values)
        m, s, b = pop.run(1, survival_rate=SURVIVAL, mutation_rate=MUTATION, verbosity=0, X=X, Y=Y)
        assert len(rec["pairs"]) == P
        out["survmask%d" % e] = survmask
        out["left%d" % e] = np.array([p[0] for p in rec["pairs"]])
        out["right%d" % e] = np.array([p[1] for p in rec["pairs"]])
        out["coin%d" % e] = np.array(rec["coin"])
        out["mut_u%d" % e] = rec["mut_u"]
        out["mut_val%d" % e] = rec["mut_val"]
        out["individuals%d" % (e + 1)] = pop.individuals.copy()
        out["grades%d" % (e + 1)] = pop.grades.copy()
        out["rng_probe%d" % e] = np.array(np.random.get_state()[2])      # position in the global stream after the epoch
        means.append(m[0]); stds.append(s[0]); bests.append(b[0])
        assert np.diff(np.sort(pop.grades)).min() > 1e-5       # FP32 fitness cannot reorder the selection
        assert not np.array_equal(before, pop.individuals)
    out.update(mean_grades=np.array(means), grade_stds=np.array(stds), best_grades=np.array(bests))
    # reference quirk kept on purpose (_evolution.py:185-189): run() re-evaluates the SURVIVORS and the mutants, so an
    # offspring that did not mutate carries the stale grade of the individual it replaced until it survives a selection
    stale = sum(abs(fitness(pop.individuals[i], X, Y)[0] - pop.grades[i]) > 1e-12 for i in range(P))
    print("individuals with a stale grade after the last epoch:", stale)
    path = os.path.join(HERE, "population.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "best grades", bests)


if __name__ == "__main__":
    main()
