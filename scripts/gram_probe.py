"""Timing probe: the Gram stage alone (pipeline stopped after it) at the headline shape, for schedule / epilogue variants."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tblup_b200 import GblupEngine, MODE_AUTO, synth
n, m, k, P = 5000, 50000, int(os.environ.get("K", 5001)), int(os.environ.get("P", 1000))
x, y = synth.synth_dataset(n, m, h2=0.4, seed=0)
tr, va, te = synth.split_indices(n, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
eng.set_rowset(0, tr, va)
flat, off = synth.random_genomes(P, m, k, seed=1)
eng.stage(flat=flat, off=off)
fit = torch.empty(P, dtype=torch.float64, device="cuda")
eng.set_option("stop_after", 3)
eng.set_option("profile", 1)
ops = 2.0 * k * (3200 * 3201 / 2 + 800 * 3200) * P
for name, opts in (("cg2 6-stage", {}), ("cg2 no stores", {"gram_experiment": 1}), ("cg2 no epilogue", {"gram_experiment": 2}),
                   ("single CTA", {"gram_pair": 0}), ("single no epilogue", {"gram_pair": 0, "gram_experiment": 2}),
                   ("multicast pair", {"gram_pair": 1})):
    for kk, v in {"gram_pair": 2, "gram_experiment": 0, **opts}.items():
        eng.set_option(kk, v)
    for rep in range(3):
        eng.reset_counters()
        eng.evaluate_staged([0], h2=0.4, mode=MODE_AUTO, out_device_ptr=fit.data_ptr())
        torch.cuda.synchronize()
        st = eng.stage_times()
    print("%-20s gram %.2f ms  %.0f TOP/s  (gather %.2f centre %.2f)" % (name, st["gram"][0], ops / (st["gram"][0] * 1e-3) / 1e12, st["gather"][0], st["centre"][0]), flush=True)
eng.close()
