#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the LIVE reference (ianwhale/tblup).

Run in the build container only (the reference is mounted read-only at /root/reference and does not
exist on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference package, calls its own static methods / classes on small seeded
inputs and stores inputs + outputs as ``.npz`` files.  The tests never import the reference; they
read these files.  Everything is deterministic for the library versions printed into the fixture
(``versions`` field).
"""
import json
import os
import random
import sys
import tempfile

import numpy as np

REF = os.environ.get("TBLUP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, REF)

if not hasattr(np, "asscalar"):          # removed in numpy 1.23; the reference's monitor still calls it
    np.asscalar = lambda a: a.item()

import scipy  # noqa: E402
import sklearn  # noqa: E402
import tblup  # noqa: E402  (the live reference)

from oracle.gblup_oracle import synth_genotypes  # noqa: E402

VERSIONS = json.dumps({"numpy": np.__version__, "scipy": scipy.__version__, "sklearn": sklearn.__version__})


def write_dataset(tmp, x, y):
    g, p = os.path.join(tmp, "geno.npy"), os.path.join(tmp, "pheno.npy")
    np.save(g, x.astype(np.float64))
    np.save(p, y)
    return g, p


def ragged_pack(lists):
    flat = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists])
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(l) for l in lists])
    return flat, off


def fitness_cases(name, n, m, seed, h2, ks, offset=0.0, n_folds=4):
    """Static-method outputs of the reference evaluator on a seeded data set."""
    x, y = synth_genotypes(n, m, h2=h2, seed=seed, offset=offset)
    xf = x.astype(np.float64)
    with tempfile.TemporaryDirectory() as tmp:
        g, p = write_dataset(tmp, x, y)
        random.seed(seed)
        np.random.seed(seed)
        ev = tblup.InterGCVBlupParallelEvaluator(g, p, h2, n_procs=1, n_folds=n_folds,
                                                 snp_remover=tblup.SNPRemovalHandler(10, 0.0, h2, False))
    train, valid, test = list(ev.training_indices), list(ev.validation_indices), list(ev.testing_indices)
    folds = ev.fold_indices

    rng = np.random.default_rng(seed + 1000)
    genomes = []
    for k in ks:
        if k < 0:                                   # negative k: sample WITH replacement (duplicates)
            genomes.append(rng.integers(0, m, size=-k))
        else:
            genomes.append(rng.choice(m, size=k, replace=False))
    E = tblup.BlupParallelEvaluator
    out = {"gblup": [], "snp_blup": [], "blup": [], "blup_testing": [], "blup_folds": []}
    tv = np.concatenate((train, valid))
    for gen in genomes:
        gen = gen.astype(int)
        out["gblup"].append(E.gblup(gen, train, valid, xf, y, h2))
        out["snp_blup"].append(E.snp_blup(gen, train, valid, xf, y, h2))
        out["blup"].append(E.blup(gen, train, valid, xf, y, h2))
        out["blup_testing"].append(E.blup(gen, tv, test, xf, y, h2))
        out["blup_folds"].append([E.blup(gen, ft, fv, xf, y, h2) for ft, fv in folds])
    flat, off = ragged_pack(genomes)
    fold_flat_t, fold_off_t = ragged_pack([f[0] for f in folds])
    fold_flat_v, fold_off_v = ragged_pack([f[1] for f in folds])
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        x=x, y=y, h2=h2, train=np.array(train), valid=np.array(valid), test=np.array(test),
        genomes_flat=flat, genomes_off=off,
        fold_train_flat=fold_flat_t, fold_train_off=fold_off_t,
        fold_valid_flat=fold_flat_v, fold_valid_off=fold_off_v,
        seed=seed, n_folds=n_folds,
        ref_gblup=np.array(out["gblup"]), ref_snp_blup=np.array(out["snp_blup"]), ref_blup=np.array(out["blup"]),
        ref_blup_testing=np.array(out["blup_testing"]), ref_blup_folds=np.array(out["blup_folds"]),
        versions=VERSIONS)
    print(name, "cases", len(genomes), "gblup", out["gblup"][:3], "snp", out["snp_blup"][:3])


def trajectory_case(name, n, m, seed, h2, features, pop, gens, regressor="blup", cv_folds=3):
    """Drive the reference's own main loop (Population / evolver / selector / monitor) with a recording
    subclass of its evaluator and store every genome batch it evaluated plus what it selected."""
    x, y = synth_genotypes(n, m, h2=h2, seed=seed)
    record = {"genomes": [], "positions": [], "fitness": [], "generation": [], "splits": []}

    base = {"blup": tblup.BlupParallelEvaluator, "intracv_blup": tblup.IntraGCVBlupParallelEvaluator,
            "intercv_blup": tblup.InterGCVBlupParallelEvaluator}[regressor]

    class Recording(base):
        """In-process evaluation with the reference's static blup(); records inputs and outputs."""

        def __enter__(self):
            self.consumers = ["in-process"]
            self._data = np.load(self.data_path)
            self._labels = np.load(self.labels_path)

        def __exit__(self, *a):
            self.consumers = []

        def _evaluate(self, population, to_evaluate, indices, generation):
            if regressor == "intracv_blup":
                split_list = [self.train_validation_indices(f) for f in range(self.n_folds)]
            else:
                split_list = [self.train_validation_indices(generation)]
            fits = []
            for genome in to_evaluate:
                vals = [tblup.BlupParallelEvaluator.blup(genome, tr, va, self._data, self._labels, self.h2)
                        for tr, va in split_list]
                fits.append(sum(vals) / len(vals) if regressor == "intracv_blup" else vals[0])
            record["genomes"].append([np.asarray(g) for g in to_evaluate])
            record["positions"].append(list(indices))
            record["fitness"].append(list(fits))
            record["generation"].append(generation)
            record["splits"].append(generation % getattr(self, "n_folds", 1) if regressor == "intercv_blup" else 0)
            for index, fitness in zip(indices, fits):
                population[index].set_fitness(fitness)
                self.archive[population[index].uid] = population[index].fitness
            return population

    def factory(args):
        kw = dict(n_procs=1, splitter=None,
                  snp_remover=tblup.SNPRemovalHandler(args.features, args.h2_alpha, args.heritability, False))
        if regressor != "blup":
            kw["n_folds"] = args.cv_folds
        return Recording(args.geno, args.pheno, args.heritability, **kw)

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        g, p = write_dataset(tmp, x, y)
        os.chdir(tmp)                                  # Monitor writes ./results/<run>/
        try:
            tblup.get_evaluator = factory
            from tblup.config import parser
            args = parser.parse_args(["--geno", g, "--pheno", p, "--seed", str(seed), "--features", str(features),
                                      "--population_size", str(pop), "--generations", str(gens),
                                      "--heritability", str(h2), "--regressor", regressor,
                                      "--cv_folds", str(cv_folds), "-p", "1"])
            random.seed(args.seed)
            np.random.seed(args.seed)
            kwargs = tblup.build_kwargs(args)
            ev = kwargs["evaluator"]
            best_fit, best_genome, pop_fit = [], [], []
            with ev:
                population = tblup.Population(**kwargs)
                for _ in range(gens + 1):
                    b = max(population, key=lambda ind: ind.fitness)
                    best_fit.append(float(b.fitness))
                    best_genome.append(np.sort(np.asarray(b.genome)))
                    pop_fit.append([float(ind.fitness) for ind in population])
                    if population.generation > gens:
                        break
                    population.do_generation()
                testing = ev.evaluate_testing if False else None  # the pool-based testing path is not replayed
        finally:
            os.chdir(cwd)
    assert testing is None
    flat, off = ragged_pack([g_ for batch in record["genomes"] for g_ in batch])
    batch_sizes = np.array([len(b) for b in record["genomes"]])
    folds = getattr(ev, "fold_indices", None)
    extra = {}
    if folds is not None:
        ft, fto = ragged_pack([f[0] for f in folds])
        fv, fvo = ragged_pack([f[1] for f in folds])
        extra = dict(fold_train_flat=ft, fold_train_off=fto, fold_valid_flat=fv, fold_valid_off=fvo)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        x=x, y=y, h2=h2, seed=seed, features=features, pop=pop, gens=gens, regressor=regressor,
        train=np.array(ev.training_indices), valid=np.array(ev.validation_indices),
        test=np.array(ev.testing_indices),
        genomes_flat=flat, genomes_off=off, batch_sizes=batch_sizes,
        positions=np.concatenate([np.asarray(p_) for p_ in record["positions"]]),
        fitness=np.concatenate([np.asarray(f_) for f_ in record["fitness"]]),
        generation=np.array(record["generation"]), split_of_batch=np.array(record["splits"]),
        best_fitness=np.array(best_fit), best_genome=np.stack(best_genome), pop_fitness=np.array(pop_fit),
        versions=VERSIONS, **extra)
    print(name, "batches", len(batch_sizes), "best", best_fit)


if __name__ == "__main__":
    # k < n exercises snp_blup, k > n exercises gblup; -k = duplicates; 1 and n, n+1 are the dispatch boundary
    fitness_cases("fit_small", n=120, m=400, seed=3, h2=0.4,
                  ks=[1, 7, 40, 119, 120, 121, 150, 260, 400, -90, -200, -330])
    fitness_cases("fit_offset", n=96, m=300, seed=11, h2=0.25, offset=50.0,
                  ks=[16, 95, 97, 128, 200, 300, -128, -250], n_folds=5)
    fitness_cases("fit_mid", n=400, m=1500, seed=5, h2=0.6, ks=[64, 400, 401, 700, 1100, -900], n_folds=3)
    trajectory_case("traj_gblup", n=150, m=600, seed=0, h2=0.4, features=200, pop=12, gens=6)
    trajectory_case("traj_snpblup", n=150, m=600, seed=1, h2=0.4, features=60, pop=10, gens=5)
    trajectory_case("traj_intracv", n=150, m=500, seed=2, h2=0.4, features=180, pop=8, gens=4,
                    regressor="intracv_blup", cv_folds=3)
    trajectory_case("traj_intercv", n=150, m=500, seed=4, h2=0.4, features=180, pop=8, gens=5,
                    regressor="intercv_blup", cv_folds=3)
