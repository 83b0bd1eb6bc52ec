import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import ctypes as C
from tblup_b200 import _lib
# the round-1 library lacks the symbols added since: bind only what exists
lib = C.CDLL(_lib.LIB_PATH)
for name, (res, args) in _lib.SYMBOLS.items():
    if hasattr(lib, name):
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
_lib._lib = lib
from conftest import load_golden, unpack
from tblup_b200 import GblupEngine, engine as E
for name in ("fit_small", "fit_offset", "fit_mid"):
    g = load_golden(name)
    tr, va, te = g["train"], g["valid"], g["test"]
    eng = GblupEngine(g["x"], g["y"], perm=np.concatenate([tr, va, te]))
    eng.set_rowset(0, tr, va)
    genomes = unpack(g["genomes_flat"], g["genomes_off"])
    f = eng.evaluate(genomes, slots=[0], h2=float(g["h2"]), mode=E.MODE_GBLUP)[:, 0]
    print(os.path.basename(_lib.LIB_PATH), name, "n_t", len(tr), "fallbacks", eng.info("last_fallbacks"), "err", np.abs(f - g["ref_gblup"]).max())
    eng.close()
