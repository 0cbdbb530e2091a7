"""Developer tool: latency of small query batches through the FP64 variance path, split-k product on / off, and of the min-variance
rollouts that are made of such batches.  usage: python tools/small_batch_time.py [N ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
for N in [int(a) for a in sys.argv[1:]] or [834, 4096, 16384]:
    rng = np.random.default_rng(0)
    X = rng.random((N, 3)); Y = 0.05 * np.sin(6 * X)
    eng = L.Engine(0)
    eng.set_train(X, Y)
    eng.factorize(0.1, np.full(3, 0.1), 1e-4, 1e-10, want_lml=False)
    eng.prepare_variance()
    for M, flags, tag in ((1, L.MEAN | L.STD, "mean+std"), (1, L.MEAN | L.STD | L.DVAR, "mean+std+dvar"), (2, L.MEAN | L.STD | L.JAC | L.JACVAR, "mean+std+jac+jacvar"),
                          (8, L.MEAN | L.STD, "mean+std"), (9, L.MEAN | L.STD, "mean+std"), (102, L.MEAN | L.STD | L.JAC, "mean+std+jac"), (102, L.MEAN | L.STD | L.JAC | L.JACVAR, "mean+std+jac+jacvar"),
                          (400, L.MEAN | L.STD | L.JAC, "mean+std+jac"), (2048, L.MEAN | L.STD | L.JAC, "mean+std+jac")):
        xq = rng.random((M, 3))
        res = {}
        for on in (0, 1):
            eng.set_debug_option("variance_splitk", on)
            for _ in range(3):
                o = eng.query(xq, flags)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter(); o = eng.query(xq, flags); ts.append(time.perf_counter() - t0)
            res[on] = (np.median(ts), o)
        d = max(np.max(np.abs(res[0][1][k] - res[1][1][k])) for k in res[0][1])
        print("N %6d M %5d %-20s one CTA per tile %8.1f us   split-k %8.1f us   max abs diff %.2e" % (N, M, tag, 1e6 * res[0][0], 1e6 * res[1][0], d), flush=True)
    if N <= 4096:
        for K in (1, 256):
            st = rng.random((K, 3))
            for on in (0, 1):
                eng.set_debug_option("variance_splitk", on)
                eng.rollout_min_variance(st, 20)
                t0 = time.perf_counter(); eng.rollout_min_variance(st, 500); dt = time.perf_counter() - t0
                print("N %6d rollout K %3d, 500 steps, split-k %d: %.1f us per step" % (N, K, on, 1e6 * dt / 500), flush=True)
    eng.close()
