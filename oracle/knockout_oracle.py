"""CPU restatement of the knockout local search (tblup/local.py:50-76) -- TEST INFRASTRUCTURE ONLY.

``ref_knockout`` follows the reference loop line by line: mask out marker i, score ``genome[mask]`` with ``blup`` on the
training / validation split, keep it out when the fitness is strictly higher than the best so far, else put it back.
Pinned against the live reference class by ``tests/golden/make_golden_ko.py`` -> ``tests/golden/ko_*.npz``
(``tests/test_oracle_golden.py``)."""
import numpy as np

from . import gblup_oracle as O


def ref_knockout(genome, start_fitness, train, valid, data, labels, h2, blup=O.exact_blup):
    """Returns (keep mask, best fitness, fitness of every step's candidate)."""
    genome = np.asarray(genome)
    best = start_fitness
    mask = np.ones(len(genome), dtype=bool)
    trace = []
    for i in range(len(genome)):
        mask[i] = False                                           # local.py:63
        f = blup(genome[mask], train, valid, data, labels, h2)    # local.py:65-66
        trace.append(f)
        if f > best:                                              # local.py:68 (NaN never improves)
            best = f
        else:
            mask[i] = True                                        # local.py:74
    return mask, best, np.array(trace)


def leave_one_out(genome, train, valid, data, labels, h2, blup=O.exact_blup):
    """Fitness of genome without its i-th entry, for every i."""
    genome = np.asarray(genome)
    return np.array([blup(np.delete(genome, i), train, valid, data, labels, h2) for i in range(len(genome))])
