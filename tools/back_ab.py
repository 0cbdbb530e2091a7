"""Developer tool: back-substitution variants (one flag-chained launch vs one launch per block): fit time and alpha agreement.
usage: python tools/back_ab.py [N ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
for N in [int(a) for a in sys.argv[1:]] or [100, 834, 4096, 16384]:
    rng = np.random.default_rng(0)
    X = rng.random((N, 3)); Y = 0.05 * np.sin(6 * X)
    eng = L.Engine(0)
    eng.set_train(X, Y)
    ell = np.full(3, 0.1)
    res = {}
    for variant in (0, 1, 0, 1):
        eng.set_debug_option("back_substitution_variant", variant)
        ts = []
        for _ in range(6):
            t0 = time.perf_counter(); info, _ = eng.factorize(0.1, ell, 1e-4, 1e-10, want_lml=False); ts.append(time.perf_counter() - t0)
        a = eng.export_alpha().copy()
        res.setdefault(variant, []).append((min(ts), a))
    a0, a1 = res[0][0][1], res[1][0][1]
    print("N", N, "per-block ms", [round(1e3 * t, 3) for t, _ in res[0]], "chained ms", [round(1e3 * t, 3) for t, _ in res[1]],
          "max rel diff of alpha %.2e" % (np.max(np.abs(a0 - a1)) / np.max(np.abs(a0))), "repeatable", np.array_equal(res[1][0][1], res[1][1][1]), flush=True)
    eng.close()
