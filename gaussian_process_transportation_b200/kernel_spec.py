"""Reading the reference's configuration object -- an sklearn kernel tree -- into the engine's hyper-parameters.

The reference passes `ConstantKernel * RBF + WhiteKernel` objects around as its only configuration
(policy_transportation/transportation/gaussian_process_transportation.py:12; SURVEY.md section 5 "Config").
sklearn's kernel classes are used here purely as that container (theta / bounds / "fixed" bookkeeping and
`get_params()` names the reference reads back at models/gaussian_process.py:38-41); no sklearn arithmetic runs.
"""
from __future__ import annotations

import numpy as np
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, Product, Sum, WhiteKernel


class UnsupportedKernel(NotImplementedError):
    pass


def check_supported(kernel):
    # exact types: sklearn's Matern subclasses RBF but has a different radial profile
    ok = (type(kernel) is Sum and type(kernel.k1) is Product and type(kernel.k1.k1) is ConstantKernel
          and type(kernel.k2) is WhiteKernel
          and (type(kernel.k1.k2) is RBF or (type(kernel.k1.k2) is Matern and kernel.k1.k2.nu in (0.5, 1.5, 2.5))))
    if not ok:
        raise UnsupportedKernel(
            "gaussian_process_transportation_b200 implements ConstantKernel * {RBF | Matern(nu=0.5|1.5|2.5)} + WhiteKernel only "
            f"(got {kernel!r}); there is no CPU fallback for other kernels")


def kernel_kind(kernel):
    """Engine code of the stationary factor's radial profile: 0 RBF, 1 Matern-1.5, 2 Matern-2.5, 3 Matern-0.5."""
    k = kernel.k1.k2
    if type(k) is RBF:
        return 0
    return {0.5: 3, 1.5: 1, 2.5: 2}[k.nu]


def read_params(kernel, d):
    """(c, ell[d], s2) of a supported kernel; a scalar / 1-element length-scale is isotropic (sklearn semantics)."""
    prm = kernel.get_params()
    c = float(prm["k1__k1__constant_value"])
    ell = np.atleast_1d(np.asarray(prm["k1__k2__length_scale"], dtype=np.float64)).ravel()
    if ell.size == 1:
        ell = np.repeat(ell, d)
    if ell.size != d:
        raise ValueError(f"Anisotropic kernel must have the same number of dimensions as data ({ell.size}!={d})")
    return c, np.ascontiguousarray(ell), float(prm["k2__noise_level"])


def map_gradient(kernel, grad_full, d):
    """Full gradient [dlog c, dlog ell_0..ell_{d-1}, dlog s2] -> the kernel's active theta layout
    (sklearn:kernels.py:739-790: fixed hyper-parameters are dropped, isotropic length-scale is one entry)."""
    out = []
    for hp in kernel.hyperparameters:
        if hp.fixed:
            continue
        if hp.name == "k1__k1__constant_value":
            out.append(grad_full[0])
        elif hp.name == "k1__k2__length_scale":
            if hp.n_elements > 1:
                out.extend(grad_full[1:1 + d])
            else:
                out.append(np.sum(grad_full[1:1 + d]))
        elif hp.name == "k2__noise_level":
            out.append(grad_full[1 + d])
        else:  # pragma: no cover
            raise UnsupportedKernel(f"unexpected hyper-parameter {hp.name}")
    return np.asarray(out, dtype=np.float64)
