"""Knockout local search of tblup/local.py on the GPU.

``KnockoutLocalSearch.search()`` keeps the reference's contract (local.py:50-76: walk the best individual's markers in
order, drop a marker when the fitness without it is higher, return ``(genome[mask], best_fitness)``) but hands the
whole greedy sequence to the device (``tb_knockout``: speculative batches through the ordinary pipeline, every decision
identical to the sequential loop) instead of calling ``evaluator.blup`` once per marker.  ``tblup_b200.install`` binds it
to ``tblup.local.get_local_search``; without that the reference's own class still works against our evaluator, one
single-genome evaluation at a time, through the static ``BlupParallelEvaluator.blup``.
"""
import abc
from copy import deepcopy

import numpy as np

from .engine import MODE_AUTO
from .evaluator import BlupParallelEvaluator


def get_local_search(args, population):
    """tblup/local.py:8-18."""
    if args.local_search == args.LOCAL_SEARCH_KNOCKOUT:
        return KnockoutLocalSearch(population)
    raise NotImplementedError("Local search method {} not implemented.".format(args.local_search))


class LocalSearch(abc.ABC):
    def __init__(self, population):
        self.population = population

    @abc.abstractmethod
    def search(self):
        raise NotImplementedError()


class KnockoutLocalSearch(LocalSearch):
    def __init__(self, population):
        super().__init__(population)
        assert issubclass(population.evaluator.__class__,
                          BlupParallelEvaluator), "Knockout only implemented for BLUP regressors."
        self.evaluations = 0        # evaluations the greedy sequence consumed (what the reference would have run)
        self.batches = 0            # batched pipeline passes it took here

    def search(self):
        evaluator = self.population.evaluator
        best = deepcopy(max(self.population, key=lambda individual: individual.fitness))
        genome = evaluator.snp_remover.combine_with_removed(best.genome)
        # main.py calls this after the evaluator's ``with`` block has closed its device contexts (main.py:42-45), like
        # the reference, which reloads the data here (local.py:57): use a live engine if there is one, else an ad-hoc one
        if evaluator.consumers:
            engine, slot = evaluator.consumers[0], 0
        else:
            from .genoio import load_genotypes
            data = load_genotypes(evaluator.data_path, evaluator.n_samples)
            labels = np.load(evaluator.labels_path)
            engine = BlupParallelEvaluator._adhoc_engine(data, labels, evaluator.training_indices,
                                                         evaluator.validation_indices)
            slot = 0
        keep, fitness, self.evaluations, self.batches = engine.knockout(genome, best.fitness, slot=slot, h2=evaluator.h2,
                                                                         mode=MODE_AUTO)
        return np.asarray(genome)[keep], fitness
