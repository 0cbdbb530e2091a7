"""Developer tool: N-point fit (gptb_factorize, no LML) and one LML+gradient evaluation, timed; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list.  usage: python tools/fit_launches.py [N] [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(0)
X = rng.random((N, 3)); Y = 0.05 * np.sin(6 * X)
eng = L.Engine(0)
eng.set_train(X, Y)
ell = np.full(3, 0.1)
ts, tl = [], []
for _ in range(reps):
    t0 = time.perf_counter(); info, _ = eng.factorize(0.1, ell, 1e-4, 1e-10, want_lml=False); ts.append(time.perf_counter() - t0)
for _ in range(reps):
    t0 = time.perf_counter(); eng.lml(0.1, ell, 1e-4, 1e-10, want_grad=True); tl.append(time.perf_counter() - t0)
print("N", N, "info", info, "fit ms", [round(1e3 * t, 3) for t in ts], "lml+grad ms", [round(1e3 * t, 3) for t in tl])
