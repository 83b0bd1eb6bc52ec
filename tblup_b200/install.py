"""Plug the B200 evaluators into an importable reference package (ianwhale/tblup) without editing it.

``tblup/utils.py:48`` does ``from tblup import get_evaluator`` at call time, so assigning
``tblup.get_evaluator`` before ``main()`` runs is the whole seam.  The installed classes
also inherit from the reference's own classes so ``isinstance`` / ``issubclass`` checks such as
tblup/local.py:47 keep holding; every method is ours (ours come first in the MRO and never call up)."""
import numpy as np

from . import evaluator as ours

_NAMES = ["BlupParallelEvaluator", "InterGCVBlupParallelEvaluator", "IntraGCVBlupParallelEvaluator",
          "MonteCarloCVBlupParallelEvaluator"]


def shim_numpy():
    """``np.asscalar`` was removed in numpy 1.23 but tblup/monitor.py:244-245 still calls it."""
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()


def install(tblup_module=None):
    """Returns the factory now bound to ``tblup.get_evaluator``."""
    shim_numpy()
    if tblup_module is None:
        import tblup as tblup_module
    hybrid = {}
    for name in _NAMES:
        hybrid[name] = type(name, (getattr(ours, name), getattr(tblup_module, name)), {"__module__": __name__})

    def get_evaluator(args):
        saved = {n: getattr(ours, n) for n in _NAMES}
        try:
            for n in _NAMES:
                setattr(ours, n, hybrid[n])
            return ours.get_evaluator(args)
        finally:
            for n in _NAMES:
                setattr(ours, n, saved[n])

    tblup_module.get_evaluator = get_evaluator
    return get_evaluator
