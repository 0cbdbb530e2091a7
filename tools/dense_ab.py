"""Developer A/B: the DENSE case (natural order, no block masks) through the original dense loops vs through the skipping kernel's loops
(three issuing warps, plane-slot ring) with null masks -- gptb_set_debug_option("oz_force_skip_variant")."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
import torch
for N, M in ((4096, 262144), (16384, 131072)):
    rng = np.random.default_rng(0)
    X = rng.random((N, 3)); Y = 0.05 * np.sin(4 * X) + 0.01 * rng.standard_normal((N, 3))
    eng = L.Engine(0)
    eng.set_variance_mode("int8w5")
    eng.set_train(X, Y)
    eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
    xd = torch.from_numpy(-0.1 + 1.2 * rng.random((M, 3))).cuda()
    mean = torch.empty(M, 3, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
    jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda")
    kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
    fl = L.MEAN | L.STD | L.JAC
    res, outs = {}, {}
    for force in (0, 1, 0, 1):
        eng.set_debug_option("oz_force_skip_variant", force)
        eng.query_dev(xd.data_ptr(), M, fl, **kw)
        eng.timing(True); eng.timing_reset()
        eng.query_dev(xd.data_ptr(), M, fl, **kw)
        t0, n0 = eng.kernel_time(0)
        eng.timing(False); eng.timing_reset()
        torch.cuda.synchronize()
        res.setdefault(force, []).append(t0 / n0 * 65536 / (M / n0))
        outs[force] = std.clone()
    print(json.dumps({"N": N, "dense_loops_ms_per_65536": res[0], "skip_loops_null_masks_ms_per_65536": res[1],
                      "bit_identical": bool(torch.equal(outs[0], outs[1]))}), flush=True)
    eng.close()
