"""Developer tool: device-resident query throughput (mean + std + Jacobian, int8w5 + spatial) against the batch size cap.
usage: python tools/batch_cap_ab.py [N] [M]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
rng = np.random.default_rng(0)
X = rng.random((N, 3)); Y = 0.05 * np.sin(6 * X)
xq = -0.1 + 1.2 * rng.random((M, 3))
xd = torch.from_numpy(xq).cuda()
mean = torch.empty(M, 3, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda")
kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
fl = L.MEAN | L.STD | L.JAC
for cap in (65536, 131072, 262144, 524288, 131072):
    eng = L.Engine(0)
    eng.set_variance_mode(*L.parse_variance_mode("int8w5"))
    eng.set_spatial(1)
    eng.set_debug_option("batch_cap", cap)
    eng.set_train(X, Y)
    eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
    eng.prepare_variance()
    for _ in range(2):
        eng.query_dev(xd.data_ptr(), M, fl, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); eng.query_dev(xd.data_ptr(), M, fl, **kw); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(json.dumps({"N": N, "M": M, "batch_cap": cap, "ms": round(1e3 * min(ts), 2), "Mq_per_s": round(M / min(ts) / 1e6, 3), "std_checksum": float(std.sum())}), flush=True)
    eng.close()
