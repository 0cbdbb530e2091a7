"""CPU oracle: a restatement of the reference's exact-GP transport path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Two independent restatements live here:

* ``SkGaussianProcess`` -- the reference's thin wrapper logic
  (``policy_transportation/models/gaussian_process.py:16-126``) restated on top of the *real* third-party dependency that
  does the arithmetic there: ``sklearn.gaussian_process.GaussianProcessRegressor`` (reference pins scikit-learn==1.3.1 in
  ``setup.py:16``; this image has 1.9.0 -- the installed version is the oracle of record).
* ``ChoGP`` -- a Cholesky-only numpy/scipy restatement of the published algorithm (Rasmussen & Williams Alg. 2.1 as
  implemented in sklearn ``_gpr.py:347-368, 444-500, 541-656`` and ``kernels.py:1273-1296, 1403-1419, 1558-1587``), which
  shares no code with sklearn and is what the size-independent property tests use.

Parity pin: ``oracle/make_golden.py`` imports the unmodified reference from ``/root/reference`` (in the build container
only) and stores its outputs under ``tests/golden``; ``tests/test_oracle.py`` checks both restatements against those.
The orientation step depends on ``numpy-quaternion`` which is absent from the reference tree and from this image:
that sub-step is restated from the published algorithm (Bar-Itzhack) and is *parity unpinned*.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular, eigh


# --------------------------------------------------------------------------------------------------------------
# Kernel algebra: c * RBF(ell) + White(s2)                       (sklearn kernels.py:1273-1296,1403-1419,1558-1587)
# --------------------------------------------------------------------------------------------------------------
def _ell_vec(ell, d):
    ell = np.atleast_1d(np.asarray(ell, dtype=np.float64)).ravel()
    if ell.size == 1:
        ell = np.repeat(ell, d)
    assert ell.size == d
    return ell


def rbf_cross(xq, X, c, ell):
    """k(x*, X) = c * exp(-0.5 * sum_a ((x*_a/ell_a) - (X_a/ell_a))^2); no White term off the training diagonal."""
    d = X.shape[1]
    ell = _ell_vec(ell, d)
    A = xq / ell
    B = X / ell
    r2 = np.zeros((A.shape[0], B.shape[0]))
    for a in range(d):
        diff = A[:, a][:, None] - B[:, a][None, :]
        r2 += diff * diff
    return c * np.exp(-0.5 * r2)


def train_gram(X, c, ell, s2, jitter):
    K = rbf_cross(X, X, c, ell)
    K[np.diag_indices_from(K)] += s2 + jitter
    return K


class ChoGP:
    """Cholesky-only exact GP with kernel c*RBF(ell)+White(s2) and sklearn's `alpha` jitter."""

    def __init__(self, c, ell, s2, jitter=1e-10):
        self.c, self.s2, self.jitter = float(c), float(s2), float(jitter)
        self.ell_in = ell

    def fit(self, X, Y):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        if Y.ndim == 1:
            Y = Y[:, None]
        self.X, self.Y = X, Y
        self.ell = _ell_vec(self.ell_in, X.shape[1])
        K = train_gram(X, self.c, self.ell, self.s2, self.jitter)
        self.L = cholesky(K, lower=True, check_finite=False)              # sklearn _gpr.py:350
        self.alpha = cho_solve((self.L, True), Y, check_finite=False)     # sklearn _gpr.py:364
        return self

    # sklearn _gpr.py:541-656
    def lml(self, eval_gradient=False):
        X, Y, L, al = self.X, self.Y, self.L, self.alpha
        n, p = Y.shape
        val = -0.5 * np.einsum("ik,ik->", Y, al) - p * np.log(np.diag(L)).sum() - p * n / 2.0 * math.log(2 * math.pi)
        if not eval_gradient:
            return val
        d = X.shape[1]
        Kinv = cho_solve((L, True), np.eye(n), check_finite=False)
        W = al @ al.T - p * Kinv                                            # sum over outputs of (a a^T - K^-1)
        R = rbf_cross(X, X, self.c, self.ell)                               # c * R
        g_c = 0.5 * np.sum(W * R)
        g_ell = np.zeros(d)
        Xs = X / self.ell
        for a in range(d):
            diff = Xs[:, a][:, None] - Xs[:, a][None, :]
            g_ell[a] = 0.5 * np.sum(W * R * diff * diff)
        g_s2 = 0.5 * self.s2 * np.trace(W)
        return val, g_c, g_ell, g_s2

    # sklearn _gpr.py:444-500 + reference gaussian_process.py:46-49 (the "- sqrt(noise)" shift, quirk Q1)
    def predict(self, xq, return_std=False):
        Ks = rbf_cross(xq, self.X, self.c, self.ell)
        mean = Ks @ self.alpha
        if not return_std:
            return mean
        V = solve_triangular(self.L, Ks.T, lower=True, check_finite=False)
        var = (self.c + self.s2) - np.einsum("ij,ij->j", V, V)
        var = np.where(var < 0, 0.0, var)
        std = np.sqrt(var) - math.sqrt(self.s2)
        return mean, np.repeat(std[:, None], self.Y.shape[1], axis=1)

    # reference gaussian_process.py:63-102, Cholesky form of K_inv (Appendix A.5 of SURVEY.md)
    def derivative(self, xq, return_var=False):
        d = self.X.shape[1]
        p = self.Y.shape[1]
        Ks = rbf_cross(xq, self.X, self.c, self.ell)
        J = np.zeros((xq.shape[0], p, d))
        Var = np.zeros((xq.shape[0], p, d))
        for a in range(d):
            G = Ks * (self.X[:, a][None, :] - xq[:, a][:, None]) / self.ell[a] ** 2
            J[:, :, a] = G @ self.alpha
            if return_var:
                Wt = solve_triangular(self.L, G.T, lower=True, check_finite=False)
                Var[:, :, a] = (self.c / self.ell[a] ** 2 - np.einsum("ij,ij->j", Wt, Wt))[:, None]
        return (J, Var) if return_var else J

    # reference gaussian_process.py:104-126
    def derivative_of_variance(self, xq):
        d = self.X.shape[1]
        Ks = rbf_cross(xq, self.X, self.c, self.ell)
        Vk = cho_solve((self.L, True), Ks.T, check_finite=False)          # K^-1 k*
        out = np.zeros((d, xq.shape[0]))
        for a in range(d):
            G = Ks * (self.X[:, a][None, :] - xq[:, a][:, None]) / self.ell[a] ** 2
            out[a] = -2.0 * np.einsum("ij,ji->i", G, Vk)
        return out


# --------------------------------------------------------------------------------------------------------------
# The reference wrapper restated over the real sklearn regressor
# --------------------------------------------------------------------------------------------------------------
class SkGaussianProcess:
    """Restates ``policy_transportation/models/gaussian_process.py:16-126`` (same attribute names, same quirks)."""

    def __init__(self, kernel, alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=5, n_targets=None):
        from sklearn.gaussian_process import GaussianProcessRegressor
        kw = dict(kernel=kernel, alpha=alpha, optimizer=optimizer, n_targets=n_targets)
        if optimizer is not None:                                            # gaussian_process.py:18-21
            kw["n_restarts_optimizer"] = n_restarts_optimizer
        self.gp = GaussianProcessRegressor(**kw)
        self.kernel = kernel
        self.alpha = alpha

    def fit(self, X, Y):                                                      # gaussian_process.py:25-44
        self.n_features = X.shape[1]
        self.n_samples = X.shape[0]
        self.n_outputs = Y.shape[1]
        keep = ~np.isnan(Y).any(axis=1)
        self.X, self.Y = X[keep], Y[keep]
        self.gp.fit(self.X, self.Y)
        self.kernel = self.gp.kernel_
        prm = self.kernel.get_params()
        self.kernel_params_ = [prm["k1__k2__length_scale"], prm["k1"]]
        self.noise_var_ = self.gp.alpha + prm["k2__noise_level"]
        self.prior_var = prm["k1__k1__constant_value"]
        self.K_inv = np.linalg.inv(self.kernel(self.X, self.X) + self.noise_var_ * np.eye(len(self.X)))

    def predict(self, x, return_std=False, return_cov=False):                 # gaussian_process.py:46-55
        if return_std:
            y, std = self.gp.predict(x, return_std=True)
            return np.array(y), np.array(std - np.sqrt(self.kernel.get_params()["k2__noise_level"]))
        if return_cov:
            y, cov = self.gp.predict(x, return_cov=True)
            return np.array(y), np.array(cov)
        return np.array(self.gp.predict(x))

    def samples(self, x):                                                     # gaussian_process.py:57-60
        return np.transpose(self.gp.sample_y(x, n_samples=10), (2, 0, 1))

    def _dk(self, x):
        ls = np.asarray(self.kernel_params_[0]).reshape(-1, 1)                # (d,1) or (1,1)
        ks = self.kernel(x, self.X)                                           # (M,N)
        coef = (self.X.T[:, None, :] - x.T[:, :, None]) / (ls[:, :, None] ** 2)
        return ls, ks, coef * ks                                              # (d,M,N)

    def derivative(self, x, return_var=False):                                # gaussian_process.py:63-102
        ls, _, dk = self._dk(x)
        alfa = self.K_inv @ self.Y
        J = (dk.transpose(1, 0, 2) @ alfa).transpose(0, 2, 1)                 # (M,p,d)
        if not return_var:
            return J
        quad = np.sum((dk @ self.K_inv) * dk, axis=2)                         # (d,M)
        var = self.prior_var / (ls ** 2) - quad
        S = np.repeat(var[None, :, :], self.n_outputs, axis=0).transpose(2, 0, 1)
        return J, S

    def derivative_of_variance(self, x):                                      # gaussian_process.py:104-126
        _, ks, dk = self._dk(x)
        return -2.0 * np.sum((dk @ self.K_inv) * ks, axis=2)


# --------------------------------------------------------------------------------------------------------------
# Affine pre-alignment and the transport flow
# --------------------------------------------------------------------------------------------------------------
class OracleAffine:
    """Restates ``policy_transportation/models/affine_trasformation.py:8-57`` (Kabsch + optional LS scale)."""

    def __init__(self, do_scale=False, do_rotation=True):
        self.do_scale, self.do_rotation, self.scale = do_scale, do_rotation, 1

    def fit(self, S, T):
        assert len(S) == len(T)
        n, d = S.shape
        self.S_centroid, self.T_centroid = S.mean(axis=0), T.mean(axis=0)
        Sc, Tc = S - self.S_centroid, T - self.T_centroid
        if (not self.do_rotation) or (d == 2 and n < 2) or (d == 3 and n < 3):
            self.rotation_matrix = np.eye(d)
        else:
            U, _, Vt = np.linalg.svd(Sc.T @ Tc)
            V = Vt.T
            self.rotation_matrix = V @ U.T
            if np.linalg.det(self.rotation_matrix) < 0:
                V[:, -1] *= -1
                self.rotation_matrix = V @ U.T
        if self.do_scale:
            Sr = (self.rotation_matrix @ Sc.T).T
            self.scale = np.sum(Sr * Tc) / np.sum(Sr ** 2)
        self.translation = self.T_centroid - self.S_centroid

    def predict(self, x):
        return self.scale * (self.rotation_matrix @ (x - self.S_centroid).T).T + self.T_centroid

    def derivative(self, x):                                                  # scale is ignored (quirk Q6)
        return np.repeat(self.rotation_matrix[None], x.shape[0], axis=0)


def quat_from_matrix_nonorthogonal(M):
    """Bar-Itzhack quaternion of a (possibly non-orthogonal) 3x3 matrix -- numpy-quaternion's
    ``from_rotation_matrix(rot, nonorthogonal=True)`` restated from the published algorithm (PARITY UNPINNED:
    the dependency is absent).  Returns (..., 4) as (w, x, y, z); sign is arbitrary."""
    M = np.asarray(M, dtype=np.float64)
    out = np.empty(M.shape[:-2] + (4,))
    for idx in np.ndindex(*M.shape[:-2]):
        R = M[idx]
        K3 = np.empty((4, 4))
        K3[0, 0] = (R[0, 0] - R[1, 1] - R[2, 2]) / 3.0
        K3[0, 1] = (R[1, 0] + R[0, 1]) / 3.0
        K3[0, 2] = (R[2, 0] + R[0, 2]) / 3.0
        K3[0, 3] = (R[1, 2] - R[2, 1]) / 3.0
        K3[1, 1] = (R[1, 1] - R[0, 0] - R[2, 2]) / 3.0
        K3[1, 2] = (R[2, 1] + R[1, 2]) / 3.0
        K3[1, 3] = (R[2, 0] - R[0, 2]) / 3.0
        K3[2, 2] = (R[2, 2] - R[0, 0] - R[1, 1]) / 3.0
        K3[2, 3] = (R[0, 1] - R[1, 0]) / 3.0
        K3[3, 3] = (R[0, 0] + R[1, 1] + R[2, 2]) / 3.0
        for i in range(4):
            for j in range(i):
                K3[i, j] = K3[j, i]
        _, vec = eigh(K3, subset_by_index=(3, 3))
        e = vec[:, 0]
        out[idx] = (e[3], -e[0], -e[1], -e[2])
    return out


def quat_mul(a, b):
    """Hamilton product, scalar first."""
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)


class OraclePolicyTransportation:
    """Restates ``policy_transportation/transportation/policy_transportation.py:11-84``."""

    def __init__(self, method):
        self.delta_map = method

    def fit(self, S, T, do_scale=False, do_rotation=True):
        self.affine_transform = OracleAffine(do_scale=do_scale, do_rotation=do_rotation)
        self.affine_transform.fit(S, T)
        S2 = self.affine_transform.predict(S)
        self.delta_distribution = T - S2
        self.delta_map.fit(S2, self.delta_distribution)

    def transport(self, pos, return_std=True):
        pr = self.affine_transform.predict(pos)
        mean, std = self.delta_map.predict(pr, return_std=True)
        return pr + mean, std

    def transport_velocity(self, pos, vel, return_var=True):
        pr = self.affine_transform.predict(pos)
        Jg = self.affine_transform.derivative(pos)
        Jp, Jpv = self.delta_map.derivative(pr, return_var=True)
        Jphi = Jg + Jp @ Jg
        v = vel[:, :, None]
        vr = Jg @ v
        return (Jphi @ v)[:, :, 0], (Jpv @ vr ** 2)[:, :, 0]

    def transport_orientation(self, pos, ori):
        Jp = self.delta_map.derivative(pos)                                   # un-rotated pos (quirk Q7)
        Jg = self.affine_transform.derivative(pos)
        Jphi = Jg + Jp @ Jg
        return quat_mul(quat_from_matrix_nonorthogonal(Jphi), ori)


class OracleGPT:
    """Restates ``policy_transportation/transportation/gaussian_process_transportation.py:11-30``."""

    def __init__(self, kernel_transport=None, **gp_kw):
        if kernel_transport is None:
            from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C
            kernel_transport = C(0.1) * RBF(length_scale=[0.1]) + WhiteKernel(0.0001)
        self.method = OraclePolicyTransportation(SkGaussianProcess(kernel=kernel_transport, **gp_kw))

    def fit_transportation(self, do_scale=False, do_rotation=True):
        self.method.fit(self.source_distribution, self.target_distribution, do_scale=do_scale, do_rotation=do_rotation)

    def apply_transportation(self):
        self.training_traj_old = self.training_traj
        self.training_traj, self.std = self.method.transport(self.training_traj_old)
        if hasattr(self, "training_delta"):
            self.training_delta, self.var_vel_transported = self.method.transport_velocity(
                self.training_traj_old, self.training_delta)
        if hasattr(self, "training_ori"):
            self.training_ori = self.method.transport_orientation(self.training_traj_old, self.training_ori)


class OracleDiffeo:
    """Restates ``policy_transportation/transportation/gaussian_process_transportation_diffeomorphic.py:15-141`` (the
    delta-map GP owned directly, Jacobian at the ROTATED positions, inverse-map consistency error)."""

    def __init__(self, kernel_transport):
        self.kernel_transport = kernel_transport

    def fit_transportation(self, optimize=True, do_scale=False, do_rotation=True):
        self.affine_transform = OracleAffine(do_scale=do_scale, do_rotation=do_rotation)
        self.affine_transform.fit(self.source_distribution, self.target_distribution)
        S2 = self.affine_transform.predict(self.source_distribution)
        self.delta_distribution = self.target_distribution - S2
        kw = dict(n_restarts_optimizer=5) if optimize else dict(optimizer=None)
        self.gp_delta_map = SkGaussianProcess(kernel=self.kernel_transport, **kw)
        self.gp_delta_map.fit(S2, self.delta_distribution)
        self.kernel_transport = self.gp_delta_map.kernel

    def apply_transportation(self):
        self.training_traj_old = self.training_traj
        self.traj_rotated = self.affine_transform.predict(self.training_traj)
        self.delta_map_mean, self.std = self.gp_delta_map.predict(self.traj_rotated, return_std=True)
        self.training_traj = self.traj_rotated + self.delta_map_mean
        if hasattr(self, "training_delta") or hasattr(self, "training_ori"):
            J, Jv = self.gp_delta_map.derivative(self.traj_rotated, return_var=True)
            rot_gp = np.eye(J[0].shape[0]) + J
            Jg = self.affine_transform.derivative(self.traj_rotated)
        if hasattr(self, "training_delta"):
            v = Jg @ self.training_delta[:, :, None]
            self.var_vel_transported = (Jv @ v ** 2)[:, :, 0]
            self.training_delta = (rot_gp @ v)[:, :, 0]
        if hasattr(self, "training_ori"):
            q_aff = quat_from_matrix_nonorthogonal(self.affine_transform.rotation_matrix)
            self.training_ori = quat_mul(quat_from_matrix_nonorthogonal(rot_gp), quat_mul(q_aff, self.training_ori))

    def check_invertibility(self):
        self.training_traj_old = self.training_traj
        self.traj_rotated = self.affine_transform.predict(self.training_traj)
        self.delta_map_mean, self.std = self.gp_delta_map.predict(self.traj_rotated, return_std=True)
        self.training_traj = self.traj_rotated + self.delta_map_mean
        self.gp_delta_inv = SkGaussianProcess(kernel=self.kernel_transport, optimizer=None)
        self.gp_delta_inv.fit(self.target_distribution, -self.delta_distribution)
        self.delta_map_inv_mean = self.gp_delta_inv.predict(self.training_traj)[0]        # first ROW (file:121)
        self.traj_rotated_inv = self.training_traj + self.delta_map_inv_mean
        return np.sum(np.linalg.norm(self.delta_map_mean + self.delta_map_inv_mean, axis=1))


class OracleGPAL:
    """Restates ``policy_transportation/models/gaussian_process_al.py:15-107`` (greedy max-std subset + exact GP)."""

    def __init__(self, kernel, alpha=1e-10, n_restarts_optimizer=5, n_samples_max=20000):
        self.kernel0, self.alpha, self.nro, self.n_samples_max = kernel, alpha, n_restarts_optimizer, n_samples_max

    def fit(self, X, Y):
        X, Y = np.asarray(X, float), np.asarray(Y, float)
        n = X.shape[0]
        if n > self.n_samples_max:
            n_initial = int(0.1 * self.n_samples_max)
            idx = np.random.choice(range(n), size=n_initial, replace=False)
            Xs, Ys = X[idx], Y[idx]
            Xt, Yt = np.delete(X, idx, axis=0), np.delete(Y, idx, axis=0)
            from sklearn.gaussian_process import GaussianProcessRegressor
            act = GaussianProcessRegressor(kernel=self.kernel0, alpha=self.alpha)
            act.fit(Xs, Ys)
            for _ in range(self.n_samples_max - n_initial):
                _, std = act.predict(Xt, return_std=True)
                q = np.argmax(std.reshape(-1, Yt.shape[1])[:, 0])
                Xs, Ys = np.vstack([Xs, Xt[q]]), np.vstack([Ys, Yt[q]])
                Xt, Yt = np.delete(Xt, q, axis=0), np.delete(Yt, q, axis=0)
                act.fit(Xs, Ys)
            X, Y = Xs, Ys
        self.X, self.Y = X, Y
        self.model = SkGaussianProcess(kernel=self.kernel0, alpha=self.alpha, n_restarts_optimizer=self.nro)
        self.model.fit(X, Y)
        self.kernel = self.model.kernel
        return self

    def predict(self, x):
        return self.model.predict(x, return_std=True)

    def derivative(self, x):
        J = self.model.derivative(x)
        return np.transpose(J, (0, 2, 1)), np.transpose(self.model.derivative_of_variance(x))[:, :, None]


# --------------------------------------------------------------------------------------------------------------
# Synthetic workloads; shared by tests and bench so GPU and CPU arms see identical inputs
# --------------------------------------------------------------------------------------------------------------
def synthetic_pairs(n, d=3, seed=0):
    rng = np.random.default_rng(seed)
    S = rng.random((n, d))
    th = math.radians(20.0)
    R0 = np.eye(d)
    if d >= 2:
        R0[0, 0], R0[0, 1], R0[1, 0], R0[1, 1] = math.cos(th), -math.sin(th), math.sin(th), math.cos(th)
    t0 = np.linspace(0.3, -0.2, d)
    T = S @ R0.T + t0 + 0.05 * np.sin(4.0 * S) + 0.01 * rng.standard_normal((n, d))
    return S, T


def helix_queries(m, d=3):
    t = np.linspace(0.0, 1.0, m)
    cols = [0.5 + 0.4 * np.cos(6 * np.pi * t), 0.5 + 0.4 * np.sin(6 * np.pi * t), 0.05 + 0.9 * t]
    x = np.stack(cols[:d], axis=1)
    v = np.gradient(x, axis=0)
    v /= np.linalg.norm(v, axis=1, keepdims=True) + 1e-300
    return np.ascontiguousarray(x), np.ascontiguousarray(v)
