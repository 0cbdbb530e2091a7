"""One fit (+ optional LML gradient) at size N for ncu launch lists (developer tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
grad = len(sys.argv) > 2 and sys.argv[2] == "grad"
S, T = synthetic_pairs(N, 3, seed=0)
eng = L.Engine(0)
eng.set_train(S, T - S)
if grad:
    print(eng.lml(0.1, [0.1] * 3, 1e-4, 1e-10, True))
else:
    print(eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10))
