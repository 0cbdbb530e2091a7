"""A/B of the Cholesky trailing-update schedules (developer tool)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
for N in [int(a) for a in sys.argv[1:]] or [4096, 16384]:
    S, T = synthetic_pairs(N, 3, seed=0)
    eng = L.Engine(0)
    eng.set_train(S, T - S)
    res = {"N": N}
    Ls = []
    for v in (0, 1):
        eng.set_trailing_variant(v)
        eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10, want_lml=False); best = min(best, time.perf_counter() - t0)
        eng.timing(True); eng.timing_reset()
        eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10, want_lml=False)
        ms, n = eng.kernel_time(2)
        eng.timing(False); eng.timing_reset()
        res[f"v{v}_fit_ms"] = best * 1e3; res[f"v{v}_tflops"] = N ** 3 / 3 / best * 1e-12; res[f"v{v}_trailing_ms"] = ms
        if N <= 4096:
            Ls.append(eng.export_L())
    if Ls:
        res["L_identical"] = bool(np.array_equal(Ls[0], Ls[1]))
    print(json.dumps(res), flush=True)
    eng.close()
