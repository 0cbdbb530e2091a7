"""INT8-sliced variance path vs the FP64 DMMA path: accuracy of std / Jacobian variance and throughput (developer tool)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs
import torch

def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1000, 4096]
    for N in sizes:
        S, T = synthetic_pairs(N, 3, seed=0)
        eng = L.Engine(0)
        eng.set_train(S, T - S)
        c, ell, s2 = 0.1, [0.1] * 3, 1e-4
        eng.factorize(c, ell, s2, 1e-10)
        rng = np.random.default_rng(0)
        xs = np.vstack([-0.1 + 1.2 * rng.random((4096, 3)), S[:1024] + 1e-3])
        eng.set_variance_mode(0)
        ref = eng.query(xs, L.MEAN | L.STD | L.JAC | L.JACVAR)
        M = 1 << 18 if N <= 4096 else 1 << 16
        xq = -0.1 + 1.2 * rng.random((M, 3))
        xd = torch.from_numpy(xq).cuda()
        mean = torch.empty(M, 3, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
        jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda")
        st = torch.cuda.ExternalStream(eng.stream())
        def qps():
            kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
            fl = L.MEAN | L.STD | L.JAC
            eng.query_dev(xd.data_ptr(), M, fl, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); eng.query_dev(xd.data_ptr(), M, fl, **kw); e1.record(st); e1.synchronize()
            return M / (e0.elapsed_time(e1) * 1e-3)
        res = {"N": N, "fp64_qps": qps()}
        for md, sl, tag in ((1, 5, "oz5"), (1, 6, "oz6"), (1, 7, "oz7"), (2, 4, "ow4"), (2, 5, "ow5"), (2, 6, "ow6")):
            eng.set_variance_mode(md, sl)
            o = eng.query(xs, L.MEAN | L.STD | L.JAC | L.JACVAR)
            res[f"{tag}_std_err"] = float(np.max(np.abs(o["std"] - ref["std"])) / np.sqrt(c + s2))
            res[f"{tag}_jacvar_rel"] = float(np.linalg.norm(o["jacvar"] - ref["jacvar"]) / np.linalg.norm(ref["jacvar"]))
            res[f"{tag}_mean_same"] = bool(np.array_equal(o["mean"], ref["mean"]))
            res[f"{tag}_qps"] = qps()
        print(json.dumps(res), flush=True)
        eng.close()

if __name__ == "__main__":
    main()
