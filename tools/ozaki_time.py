"""Kernel-class timing of the INT8-sliced variance path (developer tool)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
S, T = synthetic_pairs(N, 3, seed=0)
eng = L.Engine(0)
eng.set_train(S, T - S)
eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
xq = -0.1 + 1.2 * np.random.default_rng(0).random((M, 3))
xd = torch.from_numpy(xq).cuda()
mean = torch.empty(M, 3, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda")
kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
fl = L.MEAN | L.STD | L.JAC
for mode, sl in [(0, 6), (1, 6), (2, 5)]:
    eng.set_variance_mode(mode, sl)
    eng.query_dev(xd.data_ptr(), M, fl, **kw)
    eng.timing(True); eng.timing_reset()
    eng.query_dev(xd.data_ptr(), M, fl, **kw)
    t0, n0 = eng.kernel_time(0); t1, n1 = eng.kernel_time(1)
    eng.timing(False); eng.timing_reset()
    print(json.dumps({"N": N, "M": M, "mode": mode, "slices": sl, "variance_ms": t0, "variance_launches": n0, "generator_ms": t1,
                      "int8_tops_equiv": (2.0 * (sl * (sl + 1) / 2) * M * (N + 64.0) * N / 2 / (t0 * 1e-3) * 1e-12) if mode else None}))
