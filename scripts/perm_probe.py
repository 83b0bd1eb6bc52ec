"""Monte-Carlo style row set (random 80/20 split of training + validation, tblup/evaluator.py:555-561) at the headline
shape: stage times with the gather-time row permutation on and off.  usage: python scripts/perm_probe.py [pop]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tblup_b200 import GblupEngine, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500
x, y = synth.synth_dataset(5000, 50000, seed=0)
tr, va, te = synth.split_indices(5000, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
pool = np.random.default_rng(3).permutation(np.concatenate([tr, va]))
eng.set_rowset(1, pool[:3200], pool[3200:])
flat, off = synth.random_genomes(P, 50000, 5001, seed=1)
eng.stage(flat=flat, off=off)
fit = torch.empty(P, dtype=torch.float64, device="cuda")
eng.set_stream(torch.cuda.current_stream().cuda_stream)
res = {}
for on in (1, 0):
    eng.set_option("perm_rows", on)
    for _ in range(2):
        eng.evaluate_staged([1], out_device_ptr=fit.data_ptr())
    eng.set_option("profile", 1)
    agg = {}
    for _ in range(3):
        eng.reset_counters()
        eng.evaluate_staged([1], out_device_ptr=fit.data_ptr())
        for k, v in eng.stage_times().items():
            agg.setdefault(k, []).append(v[0])
    eng.set_option("profile", 0)
    res[on] = fit.cpu().numpy().copy()
    med = {k: round(float(np.median(v)), 2) for k, v in agg.items()}
    print("perm_rows=%d pop %d: stage ms %s sum %.1f -> %.0f evals/s" % (on, P, med, sum(med.values()), P / sum(med.values()) * 1e3))
print("max |fitness difference| between the two paths: %.2e" % np.abs(res[1] - res[0]).max())
