"""Developer self-check on a B200: every kernel family against numpy/the oracle, printing errors (does not stop at the
first failure).  Not part of the product; tests/ holds the asserted versions."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussian_process_transportation_b200 import _lib as L
from oracle.gp_oracle import ChoGP, synthetic_pairs, helix_queries


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def main():
    eng = L.Engine(0)
    lib = eng.lib
    rng = np.random.default_rng(0)
    # ---- tile engine
    for (mt, nt, K, ma, mb) in [(1, 1, 128, 0, 0), (2, 3, 384, 0, 0), (2, 2, 256, 1, 2), (3, 2, 384, 2, 1)]:
        A = rng.standard_normal((mt * 128, K)); B = rng.standard_normal((nt * 128, K))
        Cc = np.zeros((mt * 128, nt * 128))
        rc = lib.gptb_test_gemm_nt(eng.h, L.ptr(A), L.ptr(B), L.ptr(Cc), mt, nt, K, ma, mb)
        Am, Bm = A.copy(), B.copy()
        for (Mx, mask, tiles) in [(Am, ma, mt), (Bm, mb, nt)]:
            if mask:
                for t in range(tiles):
                    if t * 128 < K:
                        blk = Mx[t * 128:(t + 1) * 128, t * 128:(t + 1) * 128]
                        blk[:] = np.tril(blk) if mask == 1 else np.triu(blk)
        print(f"gemm mt={mt} nt={nt} K={K} masks=({ma},{mb}) rc={rc} rel={rel(Cc, Am @ Bm.T):.2e}")
    # ---- diagonal tile potrf + inverse
    Mx = rng.standard_normal((128, 128)); Apd = Mx @ Mx.T + 128 * np.eye(128)
    Lt = np.zeros((128, 128)); Li = np.zeros((128, 128)); info = C.c_int(0)
    rc = lib.gptb_test_potrf_tile(eng.h, L.ptr(Apd), L.ptr(Lt), L.ptr(Li), C.byref(info))
    Lref = np.linalg.cholesky(Apd)
    print(f"potrf tile rc={rc} info={info.value} L rel={rel(np.tril(Lt), Lref):.2e} mirror={rel(np.triu(Lt), Lref.T):.2e} "
          f"inv rel={rel(Li, np.linalg.inv(Lref)):.2e}")
    bad = Apd.copy(); bad[40, 40] = -1.0
    rc = lib.gptb_test_potrf_tile(eng.h, L.ptr(bad), L.ptr(Lt), L.ptr(Li), C.byref(info))
    print(f"potrf non-PD info={info.value} (expect 41)")
    # ---- full fit / lml / grad / query
    for (n, d, ell, c, s2, m) in [(20, 2, [4.0, 4.0], 10.0, 0.01, 50), (300, 3, [0.1, 0.15, 0.2], 0.1, 1e-4, 200),
                                  (834, 3, [0.0403] * 3, 0.0219 ** 2, 1e-5, 102), (1000, 3, [0.1] * 3, 0.1, 1e-4, 300),
                                  (2500, 3, [0.1, 0.12, 0.09], 0.1, 1e-4, 1000), (700, 1, [0.1], 0.5, 1e-3, 77)]:
        S, T = synthetic_pairs(n, d, seed=2)
        Y = T - S
        if d == 1:
            Y = Y[:, :1]
        ora = ChoGP(c, ell, s2).fit(S, Y)
        t0 = time.time()
        eng.set_train(S, Y)
        info, lml = eng.factorize(c, ell, s2, 1e-10)
        t1 = time.time()
        Lg = eng.export_L(); ag = eng.export_alpha()
        print(f"N={n} d={d}: info={info} fit {1e3*(t1-t0):.1f} ms  L rel={rel(Lg, ora.L):.2e} alpha rel={rel(ag, ora.alpha):.2e} "
              f"lml gpu={lml:.10g} ref={ora.lml():.10g}")
        info, lml2, g = eng.lml(c, ell, s2, 1e-10, True)
        v, gc, gl, gs = ora.lml(eval_gradient=True)
        gref = np.concatenate([[gc], gl, [gs]])
        print(f"      lml+grad info={info} lml rel={abs(lml2 - v)/abs(v):.2e} grad rel={rel(g, gref):.2e}")
        Kinv = eng.export_Kinv()
        from scipy.linalg import cho_solve
        Kref = cho_solve((ora.L, True), np.eye(n))
        print(f"      Kinv rel={rel(Kinv, Kref):.2e}")
        xq, vq = helix_queries(m, d)
        xq = xq * 0.8 + 0.1
        fl = L.MEAN | L.STD | L.JAC | L.JACVAR | L.DVAR
        out = eng.query(xq, fl)
        mean, std = ora.predict(xq, return_std=True)
        J, Jv = ora.derivative(xq, return_var=True)
        dv = ora.derivative_of_variance(xq)
        sc = np.sqrt(c + s2)
        print(f"      query mean={rel(out['mean'], mean):.2e} std={np.max(np.abs(out['std']-std))/sc:.2e} J={rel(out['jac'], J):.2e} "
              f"Jvar={rel(out['jacvar'], Jv):.2e} dvar={rel(out['dvar'], dv):.2e}")
        out2 = eng.query(xq, L.MEAN | L.JAC)
        print(f"      A0 query mean={rel(out2['mean'], mean):.2e} J={rel(out2['jac'], J):.2e}")
    print("launches", eng.launch_count())


if __name__ == "__main__":
    main()
