"""CPU stand-in for ``tblup_b200.engine.GblupEngine`` backed by the oracle -- TEST INFRASTRUCTURE ONLY.
Lets the host-side logic (evaluator classes, sharding, the drop-in seam into the reference's main loop) run in
the GPU-less build container.  Never importable from the product package."""
import numpy as np

from oracle import gblup_oracle as O


class OracleEngine:
    instances = []

    def __init__(self, geno, pheno, perm=None, device=0, storage="packed2"):
        self.storage = storage
        self.x = (geno.unpack() if hasattr(geno, "unpack") else np.asarray(geno)).astype(np.int8)
        self.y = np.asarray(pheno, dtype=np.float64).ravel()
        self.n, self.m = self.x.shape
        self.device = device
        self.rowsets = {}
        self.calls = []
        self.closed = False
        OracleEngine.instances.append(self)

    def set_rowset(self, slot, train, valid):
        self.rowsets[int(slot)] = (np.asarray(train), np.asarray(valid))

    def evaluate_packed(self, flat, off, slots=(0,), h2=0.4, mode=0, out=None):
        assert not self.closed
        P = off.size - 1
        res = np.empty((P, len(slots)))
        for i in range(P):
            g = flat[off[i]:off[i + 1]]
            for s, slot in enumerate(slots):
                tr, va = self.rowsets[int(slot)]
                if mode == 0:
                    res[i, s] = O.exact_blup(g, tr, va, self.x, self.y, h2)
                else:
                    res[i, s] = O.exact_fitness(g, tr, va, self.x, self.y, h2,
                                                O.MODE_GBLUP if mode == 1 else O.MODE_SNPBLUP)
        self.calls.append((P, tuple(slots)))
        return res

    def evaluate(self, genomes, slots=(0,), h2=0.4, mode=0):
        from tblup_b200.engine import pack_genomes
        flat, off = pack_genomes(genomes, self.m)
        return self.evaluate_packed(flat, off, slots, h2, mode)

    def close(self):
        self.closed = True
