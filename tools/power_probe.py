"""Clock / power trace of the INT8-sliced query path phase by phase (developer tool): nvidia-smi is sampled every 100 ms
while (1) the product kernel + generator run serialised, (2) overlapped, (3) the FP64 DMMA path runs.  Prints per-phase
median SM clock and power."""
import json, os, subprocess, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
S, T = synthetic_pairs(N, 3, seed=0)
eng = L.Engine(0)
eng.set_train(S, T - S)
eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
xq = -0.1 + 1.2 * np.random.default_rng(0).random((M, 3))
xd = torch.from_numpy(xq).cuda()
mean = torch.zeros(M, 3, dtype=torch.float64, device="cuda"); std = torch.zeros_like(mean)
jac = torch.zeros(M, 3, 3, dtype=torch.float64, device="cuda")
kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
samples = []
stop = False
def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"],
                         stdout=subprocess.PIPE, text=True)
    while not stop:
        line = p.stdout.readline()
        if not line:
            break
        try:
            c, w = [float(v) for v in line.split(",")]
            samples.append((time.perf_counter(), c, w))
        except Exception:
            pass
    p.kill()
th = threading.Thread(target=sampler, daemon=True); th.start()
time.sleep(1.0)
def phase(name, setup, flags, secs=4.0):
    setup()
    eng.query_dev(xd.data_ptr(), M, flags, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < secs:
        eng.query_dev(xd.data_ptr(), M, flags, **kw); torch.cuda.synchronize(); n += 1
    t1 = time.perf_counter()
    sel = [(c, w) for (t, c, w) in samples if t0 + 0.5 < t < t1]
    print(json.dumps({"phase": name, "N": N, "ms_per_step": (t1 - t0) / n * 1e3, "sm_mhz_median": float(np.median([c for c, _ in sel])) if sel else None,
                      "sm_mhz_min": min([c for c, _ in sel]) if sel else None, "power_w_median": float(np.median([w for _, w in sel])) if sel else None,
                      "power_w_max": max([w for _, w in sel]) if sel else None, "samples": len(sel)}), flush=True)
A = L.MEAN | L.STD | L.JAC
phase("int8w5 serialised", lambda: (eng.set_variance_mode("int8w5"), eng.set_query_pipeline(0)), A)
phase("int8w5 overlapped", lambda: (eng.set_variance_mode("int8w5"), eng.set_query_pipeline(1)), A)
phase("int8x6 serialised", lambda: (eng.set_variance_mode("int8x6"), eng.set_query_pipeline(0)), A)
phase("fp64 dmma", lambda: eng.set_variance_mode("fp64"), A)
phase("generator only (A0)", lambda: eng.set_variance_mode("fp64"), L.MEAN | L.JAC, secs=3.0)
a = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device="cuda"); b = torch.randint(-64, 64, (8192, 8192), dtype=torch.int8, device="cuda")
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(20): torch._int_mm(a, b)
    torch.cuda.synchronize(); n += 20
t1 = time.perf_counter()
sel = [(c, w) for (t, c, w) in samples if t0 + 0.5 < t < t1]
print(json.dumps({"phase": "cuBLASLt int8 8192^3 loop", "tops": 2 * 8192 ** 3 * n / (t1 - t0) * 1e-12, "sm_mhz_median": float(np.median([c for c, _ in sel])),
                  "power_w_median": float(np.median([w for _, w in sel])), "samples": len(sel)}))
stop = True
