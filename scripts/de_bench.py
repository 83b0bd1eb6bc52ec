"""Device-driven DE generations (evolve -> decode -> evaluate -> select, all on the GPU) at the headline shape.
usage: python scripts/de_bench.py [pop] [generations]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tblup_b200 import GblupEngine, synth
from tblup_b200.de import DeviceDE, mutation_intensity

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n, m, k = 5000, 50000, 5001
x, y = synth.synth_dataset(n, m, seed=0)
tr, va, te = synth.split_indices(n, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
eng.set_rowset(0, tr, va)
de = DeviceDE(eng, P, k, seed=1)
t0 = time.perf_counter()
f0 = de.evaluate()
t_init = time.perf_counter() - t0
best = [float(f0.max())]
ts = []
for g in range(1, G + 1):
    t0 = time.perf_counter()
    take = de.step(mutation_intensity(g, 0.5), 0.8, seed=g)
    ts.append(time.perf_counter() - t0)
    best.append(float(de.fitness().max()))
ts = np.array(ts)
print("pop %d, %d x %d, k=%d: generation 0 evaluation %.1f ms; DE generation median %.1f ms (min %.1f) -> %.0f individuals/s"
      % (P, n, m, k, 1e3 * t_init, 1e3 * np.median(ts), 1e3 * ts.min(), P / np.median(ts)))
print("best fitness per generation:", [round(b, 4) for b in best])
