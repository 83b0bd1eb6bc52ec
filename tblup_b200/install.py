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
    # knockout local search (main.py:7,42-45 imports tblup.local.get_local_search when it runs): the batched device
    # search; the class also derives from the reference's so isinstance checks keep holding
    from . import local as ours_local
    ref_local = getattr(tblup_module, "local", None)
    if ref_local is not None and hasattr(ref_local, "KnockoutLocalSearch"):
        hybrid_ko = type("KnockoutLocalSearch", (ours_local.KnockoutLocalSearch, ref_local.KnockoutLocalSearch),
                         {"__module__": __name__})

        def get_local_search(args, population):
            if args.local_search == args.LOCAL_SEARCH_KNOCKOUT:
                return hybrid_ko(population)
            raise NotImplementedError("Local search method {} not implemented.".format(args.local_search))

        ref_local.get_local_search = get_local_search
        tblup_module.get_local_search = get_local_search
    # top-SNPs seeder (tblup/seeder.py:7-45 get_seeder builds TopSNPsSeedStrategy by its module-level name): the marker
    # scan runs on the GPU; the class keeps the reference class as a base
    from . import seeder as ours_seeder
    ref_seeder = getattr(tblup_module, "seeder", None)
    if (ref_seeder is not None and hasattr(ref_seeder, "TopSNPsSeedStrategy")
            and ref_seeder.TopSNPsSeedStrategy.__module__ != __name__):        # (install() may run more than once)
        ref_seeder.TopSNPsSeedStrategy = type("TopSNPsSeedStrategy",
                                              (ours_seeder.TopSNPsSeedStrategy, ref_seeder.TopSNPsSeedStrategy),
                                              {"__module__": __name__})
    return get_evaluator
