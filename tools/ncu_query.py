"""Small single-GPU driver for ncu captures: one fit + a few mode-A query launches (developer tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
S, T = synthetic_pairs(N, 3, seed=0)
eng = L.Engine(0)
eng.set_train(S, T - S)
info, lml = eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
eng.prepare_variance()
xq = np.random.default_rng(0).random((M, 3))
for _ in range(reps):
    o = eng.query(xq, L.MEAN | L.STD | L.JAC)
print("ok", info, lml, float(o["std"].mean()), "launches", eng.launch_count())
