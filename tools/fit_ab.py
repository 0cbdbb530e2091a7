"""Developer tool: factorisation schedules A/B -- (spine_variant, back_substitution_variant) in {0,1}^2: fit time, LML+gradient time
and agreement of alpha / LML with the round-1 schedule (0, 0).  usage: python tools/fit_ab.py [N ...]"""
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
    ref = None
    for spine, back in ((0, 0), (0, 1), (1, 1), (1, 1)):
        eng.set_debug_option("spine_variant", spine)
        eng.set_debug_option("back_substitution_variant", back)
        ts, tl = [], []
        for _ in range(6):
            t0 = time.perf_counter(); info, _ = eng.factorize(0.1, ell, 1e-4, 1e-10, want_lml=False); ts.append(time.perf_counter() - t0)
        a = eng.export_alpha().copy()
        for _ in range(4):
            t0 = time.perf_counter(); info2, lml, g = eng.lml(0.1, ell, 1e-4, 1e-10, want_grad=True); tl.append(time.perf_counter() - t0)
        if ref is None:
            ref = (a, lml, g)
        print("N %6d spine %d back %d: fit %.3f ms, lml+grad %.3f ms, info %d, alpha rel diff %.2e, lml rel diff %.2e, grad rel diff %.2e" % (
            N, spine, back, 1e3 * min(ts), 1e3 * min(tl), info, np.max(np.abs(a - ref[0])) / np.max(np.abs(ref[0])),
            abs(lml - ref[1]) / abs(ref[1]), np.max(np.abs(g - ref[2])) / np.max(np.abs(ref[2]))), flush=True)
    eng.close()
