"""One generation through the drop-in evaluator class (the object the reference's Population calls): Python list of index
arrays in, Python floats out, on .npy files of the headline shape.  usage: python scripts/evaluator_e2e.py [pop]"""
import os, sys, time, tempfile, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tblup_b200 import synth
from tblup_b200 import evaluator as EV

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n, m, k = 5000, 50000, 5001
x, y = synth.synth_dataset(n, m, seed=0)
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
with tempfile.TemporaryDirectory(dir=base) as tmp:
    np.save(os.path.join(tmp, "geno.npy"), x)             # int8 .npy (the reference would hold float64: 2 GB)
    np.save(os.path.join(tmp, "pheno.npy"), y)
    random.seed(0)
    np.random.seed(0)
    ev = EV.BlupParallelEvaluator(os.path.join(tmp, "geno.npy"), os.path.join(tmp, "pheno.npy"), 0.4,
                                  snp_remover=EV.SNPRemovalHandler(k, 0.1, 0.4, False))
    t0 = time.perf_counter()
    with ev:
        t_up = time.perf_counter() - t0
        flat, off = synth.random_genomes(P, m, k, seed=1)
        genomes = [flat[off[i]:off[i + 1]].astype(np.int64) for i in range(P)]
        for _ in range(2):
            ev._fitness_matrix(genomes, [0])
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            f = ev._fitness_matrix(genomes, [0])
            ts.append(time.perf_counter() - t0)
        print("upload (load .npy, validate, transpose, pack to 2 bits) %.1f s; one generation of %d genomes through "
              "the evaluator: median %.1f ms (min %.1f) -> %.0f evals/s; best fitness %.4f"
              % (t_up, P, 1e3 * np.median(ts), 1e3 * min(ts), P / np.median(ts), float(np.nanmax(f))))
