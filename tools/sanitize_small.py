"""Developer tool: a small fit, LML + gradient, small-batch queries, a short rollout and an append -- every kernel added in round 2 at
sizes of one to seven tiles -- as a quick smoke run (or under a memory checker where one is available).
usage: python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
for N in (100, 300, 834):
    rng = np.random.default_rng(0)
    X = rng.random((N, 3)); Y = 0.05 * np.sin(6 * X)
    eng = L.Engine(0)
    eng.set_train(X, Y)
    ell = np.full(3, 0.1)
    print(N, "fit", eng.factorize(0.1, ell, 1e-4, 1e-10, want_lml=True))
    print(N, "lml", eng.lml(0.1, ell, 1e-4, 1e-10, want_grad=True)[:2])
    eng.factorize(0.1, ell, 1e-4, 1e-10, want_lml=False)
    eng.prepare_variance()
    for M, fl in ((1, L.MEAN | L.STD), (1, L.MEAN | L.STD | L.DVAR), (9, L.MEAN | L.STD), (100, L.MEAN | L.STD | L.JAC | L.JACVAR)):
        o = eng.query(rng.random((M, 3)), fl)
        print(N, M, "std", float(o["std"].sum()))
    tr = eng.rollout_min_variance(rng.random((1, 3)), 5)
    print(N, "rollout", tr.shape, float(tr.sum()))
    eng.append_point(rng.random(3), rng.random(3) * 0.05)
    eng.close()
print("done")
