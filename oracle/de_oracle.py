"""CPU restatement of the reference's differential-evolution step on random-key individuals.

TEST INFRASTRUCTURE ONLY (see oracle/gblup_oracle.py for the rules).  Restates, with explicit random inputs so a
device implementation can be fed the very same draws:

* ``draw_like_reference``  -- consumes the global ``random`` / ``numpy.random`` streams exactly like one call of
  ``DERandOneEvolver.de_rand_one`` does: three ``exclusive_randrange`` parent picks (tblup/utils.py:21-36,
  tblup/evolver.py:118-121), one ``random.randrange`` forced crossover position (evolver.py:78) and one
  ``np.random.rand(dim)`` crossover mask (evolver.py:79).
* ``de_rand_one``          -- mutant = a + F (b - c); binary crossover with one forced position; optional clip to
  [0, dim - 1] (evolver.py:104-139, :63-83).
* ``decode``               -- the ``dimensionality``-long key vector selects the indices of its ``length`` largest
  keys, in ascending key order (tblup/individual.py:155-156).
* ``select``               -- child replaces parent iff strictly fitter; NaN never wins (tblup/selector.py:18-34).
* ``mutation_intensity``   -- F = 5 on every 5th generation (evolver.py:147-151).

Pinned against the live reference by tests/golden/make_golden_de.py -> tests/golden/de_*.npz.
"""
import random

import numpy as np


def exclusive_randrange(begin, end, exclude):
    r = random.randrange(begin, end)
    exclude = set(exclude)
    while r in exclude:
        r = random.randrange(begin, end)
    return r


def draw_like_reference(pop_len, dim, parent_idx):
    a = exclusive_randrange(0, pop_len, [parent_idx])
    b = exclusive_randrange(0, pop_len, [parent_idx, a])
    c = exclusive_randrange(0, pop_len, [parent_idx, a, b])
    fixed = random.randrange(0, dim)
    mask = np.random.rand(dim)
    return a, b, c, fixed, mask


def mutation_intensity(generation, configured):
    return 5 if generation % 5 == 0 else configured


def de_rand_one(keys, parent_idx, a, b, c, fixed, mask_uniform, F, CR, clip, dim):
    """Offspring key vector of parent ``parent_idx``; ``mask_uniform`` are the U(0,1) draws of the crossover."""
    mutant = keys[a] + F * (keys[b] - keys[c])
    cross = mask_uniform < CR
    cross[fixed] = True
    child = np.where(cross, mutant, keys[parent_idx])
    if clip:
        child = np.clip(child, 0, dim - 1)
    return child


def decode(key_vector, length):
    return np.argsort(key_vector)[-int(length):]


def select(parent_fitness, child_fitness):
    """Boolean vector: child replaces parent."""
    return np.asarray(child_fitness) > np.asarray(parent_fitness)


# ---- SNP removal (tblup/evaluator.py:569-633, SNPRemovalHandler) ------------------------------------------------

def removal_threshold(h2, alpha):
    """Fitness above which the best individual's markers are removed (evaluator.py:585)."""
    return np.sqrt(h2) * (1 + alpha)


def remove_best(removed, best_genome, r):
    """New removed set after the handler fires: ``best.genome[-n:]`` with n = len(best) if r < len(best) else r
    (evaluator.py:603-606) -- in both cases the whole genome -- united with what was removed before."""
    n = len(best_genome) if r < len(best_genome) else r
    return np.union1d(removed, np.asarray(best_genome)[-n:])


def filtered_genome(genome, removed):
    """What the fitness is computed on once markers have been removed (evaluator.py:617); empty -> fitness 0.0."""
    return np.setdiff1d(genome, removed)


def testing_genome(genome, removed):
    """What the testing accuracy is computed on (evaluator.py:627-633)."""
    return np.union1d(genome, removed).astype(int)
