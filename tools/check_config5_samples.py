"""CPU oracle check of the lattice samples a config-5 run returned (tools/run_config5_full.py writes them to gpurun_out/): mean and
std at the sampled lattice points from a CPU Cholesky solve (+ one refinement step) at N = 32768.  Runs anywhere (no GPU); ~10 min and
~20 GB on 8 cores.    python tools/check_config5_samples.py gpurun_out/config5_samples_log2p29_N32768.npz"""
import json, os, sys, time
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_pairs, kabsch, KERNEL
from oracle.gp_oracle import rbf_cross
z = np.load(sys.argv[1])
idx, smp, dims, origin, step = z["idx"], z["sample"], z["dims"], z["origin"], z["step"]
N = int(os.path.basename(sys.argv[1]).split("_N")[1].split(".")[0])
t0 = time.perf_counter()
c, ell, s2 = KERNEL["c"], np.array(KERNEL["ell"]), KERNEL["s2"]
S, T = synthetic_pairs(N, 3, seed=0)
R, Sc, Tc = kabsch(S, T)
Sr = (R @ (S - Sc).T).T + Tc
D = T - Sr
xs = origin + step * np.stack(np.unravel_index(idx, tuple(int(v) for v in dims)), axis=1)
xa = (R @ (xs - Sc).T).T + Tc
K = rbf_cross(Sr, Sr, c, ell); K[np.diag_indices(N)] += s2 + KERNEL["jitter"]
ks = rbf_cross(xa, Sr, c, ell)
cf = sla.cho_factor(K.copy(), lower=True, overwrite_a=True, check_finite=False)
alpha = sla.cho_solve(cf, D, check_finite=False)
zz = sla.cho_solve(cf, ks.T, check_finite=False)
zz += sla.cho_solve(cf, ks.T - K @ zz, check_finite=False)
var = c + s2 - np.einsum("mn,nm->m", ks, zz)
std = np.sqrt(np.maximum(var, 0.0)) - np.sqrt(s2)
mean = ks @ alpha
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
print(json.dumps({"file": sys.argv[1], "N": N, "samples": int(len(idx)), "cpu_s": time.perf_counter() - t0,
                  "mean_rel": rel(smp[:, :3], mean), "std_abs_over_sqrt_prior": float(np.max(np.abs(smp[:, 3] - std)) / np.sqrt(c + s2)),
                  "tolerance": {"mean_rel": 1e-9, "std_abs_over_sqrt_prior": 1e-7}}))
