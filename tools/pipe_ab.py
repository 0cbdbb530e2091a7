"""A/B of the generator/product overlap on the INT8-sliced path: step time with the pipeline off and on, outputs compared
bit for bit (developer tool).  usage: pipe_ab.py N M [mode]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
mode = sys.argv[3] if len(sys.argv) > 3 else "int8w5"
S, T = synthetic_pairs(N, 3, seed=0)
eng = L.Engine(0)
eng.set_train(S, T - S)
eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
eng.set_variance_mode(mode)
xq = -0.1 + 1.2 * np.random.default_rng(0).random((M, 3))
xd = torch.from_numpy(xq).cuda()
st = torch.cuda.ExternalStream(eng.stream())
fl = L.MEAN | L.STD | L.JAC
outs = {}
for pipe in (0, 1, 0, 1):
    eng.set_query_pipeline(pipe)
    mean = torch.zeros(M, 3, dtype=torch.float64, device="cuda"); std = torch.zeros_like(mean)
    jac = torch.zeros(M, 3, 3, dtype=torch.float64, device="cuda")
    kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
    eng.query_dev(xd.data_ptr(), M, fl, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.timing(True); eng.timing_reset()
    e0.record(st)
    for _ in range(3):
        eng.query_dev(xd.data_ptr(), M, fl, **kw)
    e1.record(st); e1.synchronize()
    t0, n0 = eng.kernel_time(0); t1, n1 = eng.kernel_time(1)
    eng.timing(False); eng.timing_reset()
    ms = e0.elapsed_time(e1) / 3
    cur = (mean.cpu().numpy(), std.cpu().numpy(), jac.cpu().numpy())
    same = None
    if pipe in outs:
        pass
    if (1 - pipe) in outs:
        same = all(np.array_equal(a, b) for a, b in zip(cur, outs[1 - pipe]))
    outs[pipe] = cur
    print(json.dumps({"N": N, "M": M, "mode": mode, "pipeline": pipe, "ms_per_step": ms, "qps": M / ms * 1e3, "products_ms": t0 / 3, "launches": n0 // 3,
                      "generator_ms_elapsed": t1 / 3, "bit_identical_to_other": same}), flush=True)
