"""Developer tool: phase clocks of potrf_diag_kernel (128 x 128 diagonal tile: Cholesky + explicit inverse), the serial spine of the
blocked factorisation."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
eng = L.Engine(0)
prof = (ctypes.c_int64 * 64)()
eng.lib.gptb_debug_read_profile(eng.h, prof)
rng = np.random.default_rng(0)
M = rng.standard_normal((128, 128)); A = M @ M.T + 128 * np.eye(128)
Lt = np.zeros((128, 128)); Li = np.zeros((128, 128)); info = ctypes.c_int(0)
for _ in range(3):
    eng.lib.gptb_test_potrf_tile(eng.h, L.ptr(A), L.ptr(Lt), L.ptr(Li), ctypes.byref(info))
eng.lib.gptb_debug_read_profile(eng.h, prof)
p = list(prof)[:13]
names = ["load", "blk0", "upd0", "blk1", "upd1", "blk2", "upd2", "blk3", "upd3", "write L", "inverse off-diag", "write inv + fwd"]
print({n: p[i + 1] - p[i] for i, n in enumerate(names)}, "total cycles", p[12] - p[0])
ns = prof[21] - prof[20]
print("wall time of the same span: %.1f us -> SM clock %.0f MHz" % (ns / 1e3, (p[12] - p[0]) / ns * 1e3))
import time
t0 = time.perf_counter()
for _ in range(200):
    eng.lib.gptb_test_potrf_tile(eng.h, L.ptr(A), L.ptr(Lt), L.ptr(Li), ctypes.byref(info))
print("host time per tile call (H2D + kernel + D2H + sync): %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
