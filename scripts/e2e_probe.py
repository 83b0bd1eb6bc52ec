"""Where the end-to-end call spends host time: tb_stage_genomes vs tb_eval_staged (host output), wall clock per call,
against the device-only time of the same evaluation.  usage: python scripts/e2e_probe.py [pop]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tblup_b200 import GblupEngine, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x, y = synth.synth_dataset(5000, 50000, seed=0)
tr, va, te = synth.split_indices(5000, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
eng.set_rowset(0, tr, va)
flat, off = synth.random_genomes(P, 50000, 5001, seed=1)
pf = torch.from_numpy(flat).pin_memory().numpy()
out = np.empty((P, 1))
for _ in range(3):
    eng.evaluate_packed(pf, off, [0], 0.4, 0, out=out)
ts, te_, tt = [], [], []
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.stage(flat=pf, off=off)
    t1 = time.perf_counter()
    eng.evaluate_staged([0], out=out)
    t2 = time.perf_counter()
    ts.append(t1 - t0); te_.append(t2 - t1); tt.append(t2 - t0)
print("stage %.2f ms, eval_staged(host out) %.2f ms, total %.2f ms" % (1e3 * np.median(ts), 1e3 * np.median(te_), 1e3 * np.median(tt)))
eng.set_option("profile", 1)
eng.reset_counters()
eng.evaluate_staged([0], out=out)
st = eng.stage_times()
print("device stages sum %.2f ms" % sum(v[0] for v in st.values()), {k: round(v[0], 2) for k, v in st.items()})
