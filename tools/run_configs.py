"""BASELINE configs 4 and 5 at full N on one B200 (developer/record tool): config 4 = N=16384 with the LML optimisation
loop (L-BFGS-B on host, LML+gradient on the GPU) then a slice of the 64M-query stream; config 5 = N=32768 fit + a
slice of the dense grid.  Parity at these sizes is checked through size-independent properties (K alpha = y)."""
import json, os, sys, time, warnings, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import gaussian_process_transportation_b200 as g
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def config4(n_restarts):
    N = 16384
    S, T = synthetic_pairs(N, 3, seed=0)
    kern = C(0.1) * RBF([0.3, 0.3, 0.3]) + WhiteKernel(1e-3)
    t = g.GaussianProcessTransportation(kernel_transport=kern)
    t.method = g.PolicyTransportation(g.GaussianProcess(kernel=kern, n_restarts_optimizer=n_restarts))
    t.source_distribution, t.target_distribution = S, T
    np.random.seed(0)
    eng = t.method.delta_map._engine
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        t.fit_transportation()
    fit_s = time.perf_counter() - t0
    gp = t.method.delta_map
    M = 1 << 18
    rng = np.random.default_rng(1)
    t.training_traj = -0.1 + 1.2 * rng.random((M, 3))
    v = rng.standard_normal((M, 3)); t.training_delta = v / np.linalg.norm(v, axis=1, keepdims=True)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        t.apply_transportation()
    apply_s = time.perf_counter() - t0
    alpha = gp.gp.alpha_
    idx = np.arange(0, N, 64)
    m = gp.predict(gp.X[idx])
    prop = rel(m, gp.Y[idx] - gp.noise_var_ * alpha[idx])
    return {"config": "c4", "N": N, "n_restarts": n_restarts, "fit_transportation_s": fit_s, "kernel": str(gp.kernel),
            "lml": float(gp.gp.log_marginal_likelihood_value_), "launches_during_fit": eng.launch_count() - l0,
            "apply_transportation_modeB_queries": M, "apply_s": apply_s, "apply_qps": M / apply_s,
            "property_mean_at_train_rel_err": prop, "std_min_max": [float(t.std.min()), float(t.std.max())]}

def config5():
    N = 32768
    S, T = synthetic_pairs(N, 3, seed=0)
    eng = L.Engine(0)
    eng.set_train(S, T - S)
    c, ell, s2, jit = 0.1, [0.1] * 3, 1e-4, 1e-10
    t0 = time.perf_counter(); info, lml = eng.factorize(c, ell, s2, jit); t1 = time.perf_counter()
    t2 = time.perf_counter(); info, lml = eng.factorize(c, ell, s2, jit, want_lml=False); fit_s = time.perf_counter() - t2
    t3 = time.perf_counter(); eng.prepare_variance(); prep_s = time.perf_counter() - t3
    M = 1 << 15
    xq = -0.1 + 1.2 * np.random.default_rng(2).random((M, 3))
    t4 = time.perf_counter(); o = eng.query(xq, L.MEAN | L.STD | L.JAC); q_s = time.perf_counter() - t4
    alpha = eng.export_alpha()
    idx = np.arange(0, N, 128)
    oo = eng.query(S[idx], L.MEAN | L.STD)
    prop = rel(oo["mean"], (T - S)[idx] - (s2 + jit) * alpha[idx])
    return {"config": "c5", "N": N, "info": info, "fit_ms": fit_s * 1e3, "potrf_tflops": N ** 3 / 3 / fit_s * 1e-12,
            "prepare_variance_ms": prep_s * 1e3, "modeA_queries": M, "modeA_qps": M / q_s,
            "property_mean_at_train_rel_err": prop, "std_at_train_max": float(oo["std"].max()), "std_grid_minmax": [float(o["std"].min()), float(o["std"].max())]}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("c5", "all"):
        print(json.dumps(config5()), flush=True)
    if which in ("c4", "all"):
        print(json.dumps(config4(int(sys.argv[2]) if len(sys.argv) > 2 else 1)), flush=True)
