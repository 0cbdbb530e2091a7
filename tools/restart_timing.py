"""Optimised fit at N = 834 with 5 restarts, sequential and concurrent, five calls each (the first creates the restart engines).
usage: python tools/restart_timing.py"""
import contextlib, io, time, warnings
import numpy as np
import gaussian_process_transportation_b200 as g
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
warnings.filterwarnings("ignore")
rng = np.random.default_rng(5)
X = rng.random((834, 3)) * np.array([0.35, 0.47, 0.01])
Y = 0.05 * np.sin(8 * X) + 0.003 * rng.standard_normal((834, 3))
for par in (False, True, False, True):
    t0 = time.perf_counter()
    gp = g.GaussianProcess(kernel=C(0.1) * RBF(length_scale=[0.1]) + WhiteKernel(1e-4), parallel_restarts=par)
    make = time.perf_counter() - t0
    ts = []
    for _ in range(5):
        np.random.seed(0)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            gp.fit(X, Y)
        ts.append(round(time.perf_counter() - t0, 4))
    print("concurrent" if par else "sequential", "construct %.4f" % make, "fits", ts, "lml %.6f" % gp.gp.log_marginal_likelihood_value_, flush=True)
