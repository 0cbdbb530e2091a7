"""Accuracy and speed of the spatial mode (Morton order + zero-plane skipping) against the FP64 DMMA path in natural order
(developer tool).  usage: spatial_check.py N [N ...]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
import torch

def run(N, mode, spatial, xs, xq):
    S, T = synthetic_pairs(N, 3, seed=0)
    eng = L.Engine(0)
    eng.set_variance_mode(mode)
    eng.set_spatial(spatial)
    eng.set_train(S, T - S)
    eng.factorize(0.1, [0.1] * 3, 1e-4, 1e-10)
    o = eng.query(xs, L.MEAN | L.STD | L.JAC | L.JACVAR)
    M = xq.shape[0]
    xd = torch.from_numpy(xq).cuda()
    mean = torch.empty(M, 3, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
    jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda")
    st = torch.cuda.ExternalStream(eng.stream())
    kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
    fl = L.MEAN | L.STD | L.JAC
    eng.query_dev(xd.data_ptr(), M, fl, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); eng.query_dev(xd.data_ptr(), M, fl, **kw); e1.record(st); e1.synchronize()
    qps = M / (e0.elapsed_time(e1) * 1e-3)
    eng.close()
    return o, qps

for N in [int(a) for a in sys.argv[1:]] or [4096]:
    S, _ = synthetic_pairs(N, 3, seed=0)
    rng = np.random.default_rng(0)
    xs = np.vstack([-0.1 + 1.2 * rng.random((4096, 3)), S[:2048] + 1e-3, S[-2048:] + 1e-4])
    M = 1 << 18 if N <= 4096 else (1 << 17 if N <= 16384 else 1 << 16)
    xq = -0.1 + 1.2 * rng.random((M, 3))
    ref, qref = run(N, "fp64", 0, xs, xq)
    res = {"N": N, "fp64_natural_qps": qref}
    sc = np.sqrt(0.1 + 1e-4)
    for mode, sp in (("int8w5", 0), ("int8w5", 1), ("int8w6", 1), ("int8x6", 0)):
        o, qps = run(N, mode, sp, xs, xq)
        tag = f"{mode}_sp{sp}"
        res[tag] = {"std_err": float(np.max(np.abs(o["std"] - ref["std"])) / sc),
                    "jacvar_rel": float(np.linalg.norm(o["jacvar"] - ref["jacvar"]) / np.linalg.norm(ref["jacvar"])),
                    "mean_rel": float(np.linalg.norm(o["mean"] - ref["mean"]) / np.linalg.norm(ref["mean"])), "qps": qps}
    print(json.dumps(res), flush=True)
