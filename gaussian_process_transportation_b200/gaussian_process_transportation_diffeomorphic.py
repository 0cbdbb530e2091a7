"""Drop-in for policy_transportation/transportation/gaussian_process_transportation_diffeomorphic.py:15-167
(`GaussianProcessTransportationDiffeo`): the attribute-driven transport flow with the delta-map GP owned directly,
plus the inverse-map consistency check used to tune the length-scale bound.

Every posterior quantity comes from the B200 engine: `apply_transportation` is ONE fused query (affine prologue,
mean, std, Jacobian, Jacobian variance, x_hat, v_hat, Var v_hat), the inverse-map check is a second engine fit + mean
query.  What stays on the host is what the reference keeps on the host: the d x d Kabsch SVD, the optuna study driver
(`optimize_diffeomorphism`, only if optuna is importable).
Differences to `GaussianProcessTransportation` that are reproduced on purpose (they are the reference's behaviour):
the Jacobian of the orientation branch is evaluated at the ROTATED positions and composed as
quat(I + J) * (quat(R) * q) (file:97-101) instead of quat(R + J R) * q at the un-rotated ones.
"""
import numpy as np
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

from . import _lib
from .affine_transform import AffineTransform
from .gaussian_process import GaussianProcess


class GaussianProcessTransportationDiffeo():
    def __init__(self, kernel_transport=C(0.1) * RBF(length_scale=[0.1]) + WhiteKernel(0.0001)):
        super(GaussianProcessTransportationDiffeo, self).__init__()
        self.kernel_transport = kernel_transport

    # -- fit (file:50-66) ----------------------------------------------------------------------------------------
    def fit_transportation(self, optimize=True, do_scale=False, do_rotation=True):
        self.affine_transform = AffineTransform(do_scale=do_scale, do_rotation=do_rotation)
        self.affine_transform.fit(self.source_distribution, self.target_distribution)
        source_distribution = self.affine_transform.predict(self.source_distribution)
        self.delta_distribution = self.target_distribution - source_distribution
        print("Kernel:", self.kernel_transport)
        if optimize == True:  # noqa: E712
            self.gp_delta_map = GaussianProcess(kernel=self.kernel_transport, n_restarts_optimizer=5)
        else:
            self.gp_delta_map = GaussianProcess(kernel=self.kernel_transport, optimizer=None)
        self.gp_delta_map.fit(source_distribution, self.delta_distribution)
        self.kernel_transport = self.gp_delta_map.kernel
        a = self.affine_transform
        self.gp_delta_map._engine.set_affine(a.rotation_matrix, float(a.scale), a.S_centroid, a.T_centroid)

    # -- apply (file:69-101) -------------------------------------------------------------------------------------
    def _forward(self, traj, vel=None):
        flags = _lib.MEAN | _lib.STD | _lib.AFFINE_IN | _lib.TRANSPORT
        if vel is not None:
            flags |= _lib.JAC | _lib.JACVAR | _lib.VELOCITY
        return self.gp_delta_map._query(traj, flags, vel=vel)

    def apply_transportation(self):
        self.training_traj_old = self.training_traj
        has_delta, has_ori = hasattr(self, 'training_delta'), hasattr(self, 'training_ori')
        o = self._forward(self.training_traj, self.training_delta if has_delta else None)
        self.traj_rotated = self.affine_transform.predict(self.training_traj)
        self.delta_map_mean, self.std = o["mean"], o["std"]
        self.training_traj = o["xhat"]                      # traj_rotated + delta_map_mean
        if has_delta:
            # v_hat = (I + J)(R v), Var v_hat = Var[J] (R v)^2   (file:86-93)
            self.training_delta = o["vhat"]
            self.var_vel_transported = o["vvar"]
        if has_ori:
            # quat(I + J(gamma(x))) * (quat(R) * q)   (file:94-101): Jacobian, both 4x4 eigen-problems and the products on the GPU
            self.gp_delta_map._ensure_fitted_factor()
            self.training_ori = self.gp_delta_map._engine.transport_orientation_diffeo(self.training_traj_old, self.training_ori)

    def sample_transportation(self):
        delta_map_samples = self.gp_delta_map.samples(self.traj_rotated)
        return self.traj_rotated + delta_map_samples

    # -- inverse-map consistency (file:109-141) --------------------------------------------------------------------
    def _inverse_error(self, forward_points):
        """Fit the inverse delta map target -> -delta with the forward kernel (no optimisation) and measure how far
        delta_inv(phi(x)) is from -delta(x) along the trajectory."""
        delta_inv = -self.delta_distribution
        self.gp_delta_inv = GaussianProcess(kernel=self.kernel_transport, optimizer=None)
        self.gp_delta_inv.fit(self.target_distribution, delta_inv)
        # the reference indexes `predict(...)[0]` -- the FIRST ROW of the mean -- and broadcasts it (file:121,138)
        self.delta_map_inv_mean = self.gp_delta_inv.predict(forward_points)[0]
        self.traj_rotated_inv = forward_points + self.delta_map_inv_mean
        return np.sum(np.linalg.norm(self.delta_map_mean + self.delta_map_inv_mean, axis=1))

    def check_invertibility(self):
        self.training_traj_old = self.training_traj
        o = self._forward(self.training_traj)
        self.traj_rotated = self.affine_transform.predict(self.training_traj)
        self.delta_map_mean, self.std = o["mean"], o["std"]
        self.training_traj = o["xhat"]
        return self._inverse_error(self.training_traj)

    def diffeomorphism_error(self, trial):
        max_lengthscale = trial.suggest_float("max_lengthscale", 2, 20, log=True)
        self.kernel_transport = C(0.1) * RBF(length_scale=[2, 2], length_scale_bounds=[0.1, max_lengthscale]) + WhiteKernel(0.0001)
        self.fit_transportation()
        o = self._forward(self.training_traj)
        self.traj_rotated = self.affine_transform.predict(self.training_traj)
        self.delta_map_mean, self.std = o["mean"], o["std"]
        self.training_traj_target = o["xhat"]
        return self._inverse_error(self.training_traj_target)

    def optimize_diffeomorphism(self, n_trials=100):
        try:
            import optuna
        except ImportError as exc:  # the study driver is the reference's third-party dependency; there is no substitute search here
            raise ImportError("optimize_diffeomorphism needs optuna (the reference's own dependency); "
                              "diffeomorphism_error(trial) is available without it") from exc
        study = optuna.create_study(direction="minimize")
        study.optimize(self.diffeomorphism_error, n_trials=n_trials)
        trial = study.best_trial
        self.kernel_transport = C(0.1) * RBF(length_scale=np.ones(self.training_traj.shape[1]),
                                             length_scale_bounds=[1, trial.params['max_lengthscale']]) + WhiteKernel(0.0001)
        self.fit_transportation()
        o = self._forward(self.training_traj)
        self.traj_rotated = self.affine_transform.predict(self.training_traj)
        self.delta_map_mean, self.std = o["mean"], o["std"]
        self.training_traj_target = o["xhat"]
        self._inverse_error(self.training_traj_target)
        return study
