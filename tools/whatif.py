"""Developer tool: what bounds the skipping INT8-sliced product kernel, and A/B of the generator/product overlap.

    GPTB_LIB_PATH=.../libgptb200_whatif.so python tools/whatif.py whatif [N ...]   # needs `make whatif`
    python tools/whatif.py pipeline                                              # product library

whatif bits (compiled in only with -DGPTB_OZ_WHATIF): 1 = the producer issues no TMA loads, 2 = the MMA warp issues no MMAs
(barrier traffic unchanged), 4 = the epilogue does no TMEM loads / arithmetic.  Results are wrong by construction; only the
launch time of the product kernel is read."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
import torch


def synthetic(N, seed=0):
    rng = np.random.default_rng(seed)
    S = rng.random((N, 3))
    return S, 0.05 * np.sin(4 * S) + 0.01 * rng.standard_normal((N, 3))


def engine(N, spatial=1, ell=0.1, mode="int8w5"):
    X, Y = synthetic(N)
    eng = L.Engine(0)
    eng.set_variance_mode(mode)
    eng.set_spatial(spatial)
    eng.set_train(X, Y)
    info, _ = eng.factorize(0.1, [ell] * 3, 1e-4, 1e-10)
    assert info == 0
    eng.prepare_variance()
    return eng


def buffers(M, seed=0):
    xq = -0.1 + 1.2 * np.random.default_rng(seed).random((M, 3))
    xd = torch.from_numpy(xq).cuda()
    mean = torch.empty(M, 3, dtype=torch.float64, device="cuda")
    std = torch.empty_like(mean)
    jac = torch.empty(M, 3, 3, dtype=torch.float64, device="cuda")
    return xd, dict(mean=mean.data_ptr(), std=std.data_ptr(), jac=jac.data_ptr()), (mean, std, jac)


FL = L.MEAN | L.STD | L.JAC


def whatif(Ns):
    for N in Ns:
        for ell in (0.1, 0.68):
            eng = engine(N, 1, ell)
            M = 65536
            xd, kw, keep = buffers(M)
            eng.query_dev(xd.data_ptr(), M, FL, **kw)
            import ctypes
            prof = (ctypes.c_int64 * 64)()
            eng.lib.gptb_debug_read_profile(eng.h, prof)       # allocates the counter buffer
            for wi in (0, 1, 2, 4, 3, 5, 6, 7):
                eng.set_debug_option("oz_whatif", wi)
                eng.executed_products(reset=True)
                eng.timing(True); eng.timing_reset()
                for _ in range(2):
                    eng.query_dev(xd.data_ptr(), M, FL, **kw)
                t0, n0 = eng.kernel_time(0)
                t1, n1 = eng.kernel_time(1)
                eng.timing(False)
                ex = eng.executed_products(reset=True) / 2
                Npad = (N + 127) // 128 * 128
                T64 = Npad // 64
                dense = (M // 128) * (T64 * (T64 + 1) // 2) * 15
                mma_cycles_per_sm = ex * 64 / 148          # 64 cycles per plane pair per chunk (2 MMAs of 128x64x32)
                eng.lib.gptb_debug_read_profile(eng.h, prof)
                pv = list(prof)
                if wi in (0, 1, 7):
                    k = lambda v: round(v / 1e3, 1)             # kilo-cycles, CTA 0, last launch
                    print(json.dumps({"N": N, "ell": ell, "whatif": wi, "profile_kcycles_cta0": {
                        "producer": {"total": k(pv[0]), "tile_claim": k(pv[1]), "tile_decode_masks": k(pv[2]), "wait_empty": k(pv[3]), "chunks": pv[4]},
                        **{f"mma_warp{w}": {"total": k(pv[8 * w]), "wait_tile": k(pv[8 * w + 1]), "wait_acc_empty": k(pv[8 * w + 2]), "masks": k(pv[8 * w + 3]),
                                            "wait_full": k(pv[8 * w + 4]), "wait_tile_go": k(pv[8 * w + 5]), "issue": k(pv[8 * w + 6]), "chunks": pv[8 * w + 7]} for w in (1, 2, 3)},
                        "epilogue_warp4": {"total": k(pv[32]), "wait_tile": k(pv[33]), "wait_acc_full": k(pv[34])}}}), flush=True)
                print(json.dumps({"N": N, "ell": ell, "whatif": wi, "products_ms": t0 / n0, "generator_ms": t1 / n1, "executed_pairs": ex,
                                  "executed_frac": ex / dense, "mma_floor_ms_at_1965MHz": mma_cycles_per_sm / 1.965e6}), flush=True)
            eng.close()


def pipeline():
    for N, M, ell in ((4096, 1 << 20, 0.1), (16384, 1 << 18, 0.1), (16384, 1 << 18, 0.68), (4096, 1 << 20, 0.68)):
        eng = engine(N, 1, ell)
        if os.environ.get("BATCH_CAP"):
            eng.set_debug_option("batch_cap", int(os.environ["BATCH_CAP"]))
        xd, kw, keep = buffers(M)
        st = torch.cuda.ExternalStream(eng.stream())
        res = {}
        for pipe in (0, 1, 0, 1):
            eng.lib.gptb_set_query_pipeline(eng.h, pipe)
            eng.query_dev(xd.data_ptr(), M, FL, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(3):
                eng.query_dev(xd.data_ptr(), M, FL, **kw)
            e1.record(st); e1.synchronize(); torch.cuda.synchronize()
            res.setdefault(pipe, []).append(e0.elapsed_time(e1) / 3)
            if pipe == 0:
                ref = [t.clone() for t in keep]
            else:
                same = all(torch.equal(a, b) for a, b in zip(ref, keep))
        print(json.dumps({"N": N, "M": M, "ell": ell, "serial_ms": res[0], "overlap_ms": res[1], "bit_identical": bool(same),
                          "qps_serial": M / min(res[0]) * 1e3, "qps_overlap": M / min(res[1]) * 1e3}), flush=True)
        eng.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "whatif"
    if what == "whatif":
        whatif([int(a) for a in sys.argv[2:]] or [4096, 16384])
    else:
        pipeline()
