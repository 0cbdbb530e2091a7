"""Phase timings of the fit path and query modes on a B200 (developer tool; wall clock around synchronous ABI calls)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import synthetic_pairs

def t(f, reps=3):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best * 1e3

def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384]
    for N in sizes:
        S, T = synthetic_pairs(N, 3, seed=0)
        Y = T - S
        eng = L.Engine(0)
        eng.set_train(S, Y)
        c, ell, s2, j = 0.1, [0.1] * 3, 1e-4, 1e-10
        eng.factorize(c, ell, s2, j)
        res = {"N": N}
        res["factorize_ms"] = t(lambda: eng.factorize(c, ell, s2, j, want_lml=False))
        res["potrf_tflops"] = (N ** 3 / 3) / (res["factorize_ms"] * 1e-3) * 1e-12
        eng.timing(True); eng.timing_reset()
        eng.factorize(c, ell, s2, j, want_lml=False)
        ms, n = eng.kernel_time(2)
        eng.timing(False); eng.timing_reset()
        res["trailing_ms"] = ms; res["trailing_launches"] = n
        def prep():
            eng.factorize(c, ell, s2, j, want_lml=False); eng.prepare_variance()
        res["factorize_plus_trtri_ms"] = t(prep, 2)
        res["lml_grad_ms"] = t(lambda: eng.lml(c, ell, s2, j, True), 2)
        eng.factorize(c, ell, s2, j, want_lml=False); eng.prepare_variance()
        M = 1 << 20 if N <= 4096 else 1 << 17
        xq = np.random.default_rng(0).random((M, 3))
        import torch
        xd = torch.from_numpy(xq).cuda()
        mean = torch.empty(M, 3, dtype=torch.float64, device="cuda"); std = torch.empty_like(mean)
        jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda"); jv = torch.empty_like(jac)
        torch.cuda.synchronize()
        st = torch.cuda.ExternalStream(eng.stream())
        def timed(flags, **kw):
            eng.query_dev(xd.data_ptr(), M, flags, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); eng.query_dev(xd.data_ptr(), M, flags, **kw); e1.record(st); e1.synchronize()
            return M / (e0.elapsed_time(e1) * 1e-3)
        res["qps_A0_mean_jac"] = timed(L.MEAN | L.JAC, mean=mean.data_ptr(), jac=jac.data_ptr())
        res["qps_A_mean_std_jac"] = timed(L.MEAN | L.STD | L.JAC, mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr())
        Mb = M // 4
        def timedB():
            fl = L.MEAN | L.STD | L.JAC | L.JACVAR
            kw = dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr(), jacvar=jv.data_ptr())
            eng.query_dev(xd.data_ptr(), Mb, fl, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); eng.query_dev(xd.data_ptr(), Mb, fl, **kw); e1.record(st); e1.synchronize()
            return Mb / (e0.elapsed_time(e1) * 1e-3)
        res["qps_B_with_jacvar"] = timedB()
        print(json.dumps(res))
        eng.close()

if __name__ == "__main__":
    main()
