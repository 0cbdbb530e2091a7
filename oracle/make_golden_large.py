"""Golden vectors at the HEADLINE size from the UNMODIFIED reference (build container only: needs /root/reference).
TEST INFRASTRUCTURE.

    python oracle/make_golden_large.py          # ~2 min of CPU; writes tests/golden/n4096_*.npz

N = 4096 training pairs of the benchmark workload (BASELINE config 3: synthetic_pairs(4096, 3, seed 0) through the
reference's AffineTransform, residual GP on the aligned points) at six hyper-parameter settings -- the benchmark's own and
five adversarial ones for the INT8-sliced variance path (larger signal-to-noise ratios, a long length-scale, the
length-scale the N = 16384 optimisation converges to, an ARD kernel with a 10x spread).  1024 queries per setting, half of
them next to training points (where the predictive variance nearly cancels and the std is most sensitive).
Reference calls: GaussianProcess.fit / predict(return_std=True) / derivative(return_var=True)
(policy_transportation/models/gaussian_process.py:25-102)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import import_reference, OUT  # noqa: E402

CASES = {
    # tag: (c, ell, s2)
    "c3": (0.1, [0.1, 0.1, 0.1], 1e-4),                        # the benchmark's hyper-parameters
    "fitted": (3.92e-3, [0.674, 0.691, 0.69], 9.98e-5),        # what L-BFGS-B converges to at N = 16384 (profiles/r01_config4_*.json)
    "snr1e4": (1.0, [0.3, 0.3, 0.3], 1e-4),
    "snr1e5": (10.0, [0.3, 0.3, 0.3], 1e-4),
    "long": (20.0, [1.5, 1.5, 1.5], 1.19e-3),                  # C1-demo-like amplitude / noise, kernel wider than the domain
    "ard10": (0.1, [0.05, 0.5, 0.15], 1e-4),
}


def main():
    warnings.filterwarnings("ignore")
    pt, _ = import_reference()
    from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
    from oracle.gp_oracle import synthetic_pairs
    import contextlib, io
    N = 4096
    S, T = synthetic_pairs(N, 3, seed=0)
    aff = pt.AffineTransform()
    with contextlib.redirect_stdout(io.StringIO()):
        aff.fit(S, T)
    X = aff.predict(S)
    Y = T - X
    rng = np.random.default_rng(11)
    lo, hi = X.min(axis=0), X.max(axis=0)
    far = lo - 0.1 * (hi - lo) + 1.2 * (hi - lo) * rng.random((512, 3))
    near = X[rng.choice(N, 512, replace=False)] + 1e-3 * rng.standard_normal((512, 3))
    xq = np.ascontiguousarray(np.vstack([far, near]))
    np.savez_compressed(os.path.join(OUT, "n4096_train.npz"), X=X, Y=Y, xq=xq)
    for tag, (c, ell, s2) in CASES.items():
        gp = pt.GaussianProcess(kernel=C(c) * RBF(ell) + WhiteKernel(s2), optimizer=None)
        with contextlib.redirect_stdout(io.StringIO()):
            gp.fit(X, Y)
        mean, std = gp.predict(xq, return_std=True)
        J, Jv = [], []
        for i in range(0, len(xq), 256):                      # the reference materialises (d, M, N) temporaries
            a, b = gp.derivative(xq[i:i + 256], return_var=True)
            J.append(a); Jv.append(b)
        J, Jv = np.concatenate(J), np.concatenate(Jv)
        assert np.array_equal(std[:, 0], std[:, 1]) and np.array_equal(Jv[:, 0], Jv[:, 1])
        np.savez_compressed(os.path.join(OUT, f"n4096_{tag}.npz"), c=c, ell=np.asarray(ell, float), s2=s2, mean=mean, std0=std[:, 0].copy(),
                            J=J, Jvar0=Jv[:, 0].copy(), lml=gp.gp.log_marginal_likelihood_value_)
        print(tag, "done: std range", float(std.min()), float(std.max()), "lml", gp.gp.log_marginal_likelihood_value_, flush=True)


if __name__ == "__main__":
    main()
