"""Device-resident Population against fixtures recorded from the reference's own `Population.run`
(tests/golden/make_golden_population.py: the unmodified class, its per-locus draws recorded through its mate_function /
mutate_function hooks).  Host-side pairing: CPU test.  Generation kernel, in-place fitness, whole runs: GPU tests."""
import numpy as np
import pytest

from conftest import load_golden, relerr


@pytest.fixture(scope="module")
def fx():
    return load_golden("population")


def test_pair_parents_consumes_the_numpy_stream_like_the_reference(fx):
    """selection + the accept/reject pairing of get_candidates (_evolution.py:99-141) stay on the host: same survivor
    masks, same parent pairs and the same position in the global NumPy stream as the reference, epoch after epoch."""
    from brainforge_b200.evolution import Population
    P, loci = fx["individuals0"].shape
    np.random.seed(int(fx["seed"]))
    pop = Population(loci=loci, fitness_function=None, limit=P)
    np.random.seed(int(fx["seed"]) + 1)
    for e in range(int(fx["epochs"])):
        pop.grades = fx["grades%d" % e].copy()
        survmask = pop.selection(float(fx["survival"]))
        assert np.array_equal(survmask, fx["survmask%d" % e])
        left, right = pop.pair_parents(np.where(survmask)[0])
        assert np.array_equal(left, fx["left%d" % e]) and np.array_equal(right, fx["right%d" % e])
        assert np.random.get_state()[2] == int(fx["rng_probe%d" % e])


def test_host_store_run_matches_the_reference(fx):
    """The host store with the reference's operators injected the same way the fixture was recorded: individuals and
    grades after every epoch equal the reference's (float64, exact)."""
    from brainforge_b200.evolution import Population
    from oracle import bf_oracle as O
    nin, nh, nout = (int(v) for v in fx["dims"])
    X, Y = fx["X"], fx["Y"]
    side = np.random.RandomState(99)

    def fitness(g, X, Y):
        return O.fitness_mlp_logsoftmax(g[None], X, Y, nin, nh, nout)

    def mate(g1, g2):
        return np.where(side.uniform(size=g1.shape) < 0.5, g1, g2)

    def mutate(rate, individuals):
        mask = side.uniform(size=individuals.shape) < rate
        individuals[np.nonzero(mask)] = side.uniform(size=(mask.sum(),)).astype(np.float32).astype(np.float64)
        return individuals, np.where(mask.sum(axis=1))[0]

    np.random.seed(int(fx["seed"]))
    pop = Population(loci=fx["individuals0"].shape[1], fitness_function=fitness, limit=len(fx["individuals0"]),
                     mate_function=mate, mutate_function=mutate)
    pop.individuals = fx["individuals0"].copy()
    np.random.seed(int(fx["seed"]) + 1)
    pop.update(None, 0, X=X, Y=Y)
    assert relerr(pop.grades, fx["grades0"]) < 1e-12
    pop.initialized = True          # as in the recording: run() must not re-evaluate everybody between the epochs
    for e in range(int(fx["epochs"])):
        m, s, b = pop.run(1, survival_rate=float(fx["survival"]), mutation_rate=float(fx["mutation"]), verbosity=0, X=X, Y=Y)
        assert np.array_equal(pop.individuals, fx["individuals%d" % (e + 1)])
        assert relerr(pop.grades, fx["grades%d" % (e + 1)]) < 1e-12
        assert abs(m[0] - fx["mean_grades"][e]) < 1e-10 and abs(b[0] - fx["best_grades"][e]) < 1e-10


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture()
def bf():
    import torch
    assert torch.cuda.is_available()
    import brainforge_b200 as bf
    bf.set_precision("tf32")
    yield bf
    bf.set_precision("fp32")


@pytest.mark.gpu
def test_next_generation_replays_the_reference_draws(bf, fx):
    """bf_population_next_generation with the reference's recorded coins / mutation draws: the store after every epoch is
    the reference's population bit for bit (FP32-representable values throughout), the mutated rows are the reference's,
    and the resident K-major operand follows (fitness of everybody == fitness from a fresh upload)."""
    from brainforge_b200.evolution._evolution import GenomeStore
    from brainforge_b200.device import DeviceArray
    nin, nh, nout = (int(v) for v in fx["dims"])
    store = GenomeStore(fx["individuals0"], nin, nh, nout)
    assert np.array_equal(store.download(), fx["individuals0"])
    x, y = DeviceArray.from_host(fx["X"]), DeviceArray.from_host(fx["Y"])
    P, loci = fx["individuals0"].shape
    rate = float(fx["mutation"]) / loci
    for e in range(int(fx["epochs"])):
        mutants = store.next_generation(fx["survmask%d" % e], fx["left%d" % e], fx["right%d" % e], rate, 0, e,
                                        injected=(fx["coin%d" % e], fx["mut_u%d" % e], fx["mut_val%d" % e]))
        want = fx["individuals%d" % (e + 1)]
        assert np.array_equal(store.download(), want), "generation %d differs from the reference" % (e + 1)
        assert np.array_equal(mutants, np.where((fx["mut_u%d" % e] < rate).sum(axis=1))[0])
        fresh = GenomeStore(want, nin, nh, nout)
        assert np.array_equal(store.fitness(None, x, y), fresh.fitness(None, x, y))
        assert np.array_equal(store.bt.numpy(np.float32), fresh.bt.numpy(np.float32))


@pytest.mark.gpu
def test_resident_fitness_matches_oracle_and_upload_path(bf, fx):
    """bf_population_fitness (genomes + K-major operand read in place, index list) == bf_fitness_mlp on uploaded genomes,
    bit for bit, and both equal the TF32-emulating oracle / the reference's grades."""
    from brainforge_b200.layers import Dense
    from brainforge_b200.device import DeviceArray
    from oracle import bf_oracle as O
    nin, nh, nout = (int(v) for v in fx["dims"])
    np.random.seed(3)
    ne = bf.NeuroEvolution(input_shape=(nin,), layerstack=[Dense(nh, activation="sigmoid"), Dense(nout, activation="softmax")],
                           cost="cxent", population_size=len(fx["individuals0"]))
    assert ne.population.device_resident
    ne.population.individuals = fx["individuals0"]
    X, Y = fx["X"], fx["Y"]
    ne.population.update(None, 0, X=X, Y=Y)
    assert relerr(ne.population.grades, fx["grades0"]) < 1e-2                      # TF32 operands vs the float64 reference
    with O.tf32_emulation({"fit.l1": ("tma", "rna")}):
        emu = O.fitness_mlp_logsoftmax(fx["individuals0"].astype(np.float32), X.astype(np.float32), Y, nin, nh, nout)
    assert relerr(ne.population.grades, emu) < 1e-4
    up = ne.batch_fitness(fx["individuals0"], X, Y).ravel()
    assert np.array_equal(up, ne.population.grades.astype(np.float32))
    # a subset / permutation through the index list, odd count
    inds = np.array([5, 0, 11, 3, 3, 15, 7])
    sub = ne.population._store.fitness(inds, DeviceArray.from_host(X), DeviceArray.from_host(Y))
    assert np.array_equal(sub.astype(np.float32), up[inds])


@pytest.mark.gpu
def test_resident_population_run_at_scale(bf):
    """A whole `learn_batch` on a device-resident population of 1024 genomes of the 784-60-10 MLP: the store never
    leaves HBM, Philox draws have the statistics the reference's NumPy draws have, grades stay consistent with a
    re-evaluation of the final genomes, and the best grade does not get worse."""
    from brainforge_b200.layers import Dense
    from oracle import bf_oracle as O
    rs = np.random.RandomState(4)
    np.random.seed(4)
    P, N = 1024, 256
    ne = bf.NeuroEvolution(input_shape=(784,), layerstack=[Dense(60, activation="sigmoid"), Dense(10, activation="softmax")],
                           cost="cxent", population_size=P)
    pop = ne.population
    assert pop.device_resident
    X = rs.randn(N, 784).astype(np.float32)
    Y = np.eye(10, dtype=np.float32)[rs.randint(0, 10, N)]
    g0 = pop.individuals
    pop.update(None, 0, X=X, Y=Y)
    first = pop.grades.copy()
    ref = O.fitness_mlp_logsoftmax(g0[:64], X, Y, 784, 60, 10)
    assert relerr(first[:64], ref) < 1e-2
    best0 = first.min()
    ne.learn_batch(X, Y, epochs=2, survival_rate=0.5, mutation_rate=20.0)
    g2 = pop.individuals
    assert g2.shape == g0.shape and g2.min() >= 0.0 and g2.max() <= 1.0      # (a float64 draw just below 1 narrows to 1.0f)
    assert pop.grades.min() <= best0 + 1e-6
    # children are loci-wise mixtures of two parents of the previous generation; after two epochs at most a few loci per
    # row are fresh draws (rate 20/47710 per locus and epoch): nearly every locus of every row exists in generation 0
    col_sets_hit = np.mean([np.isin(g2[i, :2000], g0[:, :2000]).mean() for i in range(0, P, 97)])
    assert col_sets_hit > 0.99, col_sets_hit
    # grades of the re-evaluated rows equal a fresh evaluation of the final genomes
    pop.update(None, 0, X=X, Y=Y)
    again = ne.batch_fitness(g2, X, Y).ravel()
    assert np.array_equal(pop.grades.astype(np.float32), again)


@pytest.mark.gpu
def test_philox_draw_statistics(bf):
    """Coins are fair and mutation hits are Binomial(P*loci, rate) within 5 sigma; mutated values are U(0,1)."""
    from brainforge_b200.evolution._evolution import GenomeStore
    P, loci = 512, 4000 + 6
    a = np.zeros((P, loci), np.float32)
    a[1::2] = 0.25                                   # even rows 0.0, odd rows 0.25: a child shows which parent gave each locus
    store = GenomeStore(a, 4, 5, 3) if False else None
    # the store asserts loci = nin*nh+nh+nh*nout+nout only in the fitness call: build one with matching dims
    nin, nh, nout = 60, 64, 1                        # 60*64 + 64 + 64 + 1 = 3969
    loci = nin * nh + nh + nh * nout + nout
    a = np.zeros((P, loci), np.float32)
    a[1::2] = 0.25
    store = GenomeStore(a, nin, nh, nout)
    survive = np.zeros(P)
    survive[:2] = 1
    left, right = np.zeros(P, np.int64), np.ones(P, np.int64)
    rate = 1e-3
    mutants = store.next_generation(survive, left, right, rate, seed=123, generation=0)
    g = store.download()
    # survivors keep their genome except for the (rare) mutated loci
    assert (g[0] == 0.0).mean() > 0.99 and (g[1] == 0.25).mean() > 0.99
    kids = g[2:]
    from_left = (kids == 0.0).sum()
    from_right = (kids == 0.25).sum()
    mutated = kids.size - from_left - from_right
    n = kids.size
    assert abs(from_left - from_right) < 5 * np.sqrt(n), (from_left, from_right)
    assert abs(mutated - n * rate) < 5 * np.sqrt(n * rate) + n * rate * 0.26, (mutated, n * rate)   # a hit may redraw 0.0/0.25 never
    vals = kids[(kids != 0.0) & (kids != 0.25)]
    assert 0.45 < vals.mean() < 0.55 and vals.min() >= 0 and vals.max() < 1
    rows = np.unique(np.where((g != 0.0) & (g != 0.25))[0])
    assert np.array_equal(np.sort(mutants), rows)
    # another generation index gives another stream; the same one reproduces
    s2 = GenomeStore(a, nin, nh, nout)
    s2.next_generation(survive, left, right, rate, seed=123, generation=0)
    assert np.array_equal(s2.download(), g)
    s3 = GenomeStore(a, nin, nh, nout)
    s3.next_generation(survive, left, right, rate, seed=123, generation=1)
    assert not np.array_equal(s3.download(), g)
