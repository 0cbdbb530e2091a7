"""Driver for ncu captures of the Gram build and the mean/Jacobian generator (A0 mode) (developer tool)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
S, T = synthetic_pairs(N, 3, seed=0)
eng = L.Engine(0)
eng.set_train(S, T - S)
print(eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10))
xq = np.random.default_rng(0).random((M, 3))
for _ in range(2):
    o = eng.query(xq, L.MEAN | L.JAC)
print(float(o["mean"].mean()))
