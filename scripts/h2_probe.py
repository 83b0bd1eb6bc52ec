import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import load_golden, unpack
from oracle import gblup_oracle as O
from tblup_b200 import GblupEngine, engine as E
g = load_golden("fit_mid")
x, y = g["x"], g["y"]
tr, va, te = g["train"], g["valid"], g["test"]
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
eng.set_rowset(0, tr, va)
rng = np.random.default_rng(1)
m = x.shape[1]
genomes = [rng.choice(m, size=k, replace=False) for k in (40, 255, 300, 401, 1500)]
for h2 in (0.02, 0.2, 0.9, 0.99, 0.999, 0.9999):
    for mode, om in ((E.MODE_GBLUP, O.MODE_GBLUP), (E.MODE_SNPBLUP, O.MODE_SNPBLUP)):
        got = eng.evaluate(genomes, slots=[0], h2=h2, mode=mode)[:, 0]
        sw = [int(eng.debug_fetch(E.DBG_SWEEPS, j)[0]) for j in range(len(genomes))]
        want = np.array([O.exact_fitness(gen, tr, va, x, y, h2, om) for gen in genomes])
        print("h2 %.4f mode %d max|diff| %.2e sweeps %s got %s" % (h2, mode, np.nanmax(np.abs(got - want)), sw, np.round(got, 5)), "NaN" if np.isnan(got).any() else "")
