"""The transport flow phi(x) = gamma(x) + psi(gamma(x)) -- drop-in for
policy_transportation/transportation/policy_transportation.py:11-84.

`method` is the duck-typed delta-map plugin (fit / predict / derivative / samples).  When it is the B200
GaussianProcess the affine prologue, the posterior and the Jacobian algebra of each call run as ONE fused GPU query
(include/gptb200.h flags); any other plugin goes through the same array algebra as the reference.
"""
import numpy as np

from . import _lib
from .affine_transform import AffineTransform
from .gaussian_process import GaussianProcess
from .quaternion import from_rotation_matrix_nonorthogonal, multiply as quat_multiply


class PolicyTransportation():
    def __init__(self, method):
        super(PolicyTransportation, self).__init__()
        self.delta_map = method

    @property
    def _fused(self):
        return isinstance(self.delta_map, GaussianProcess)

    def fit(self, source_distribution, target_distribution, do_scale=False, do_rotation=True):
        self.affine_transform = AffineTransform(do_scale=do_scale, do_rotation=do_rotation)
        self.affine_transform.fit(source_distribution, target_distribution)
        source_distribution = self.affine_transform.predict(source_distribution)
        self.delta_distribution = target_distribution - source_distribution
        self.delta_map.fit(source_distribution, self.delta_distribution)
        if self._fused:
            a = self.affine_transform
            self.delta_map._engine.set_affine(a.rotation_matrix, float(a.scale), a.S_centroid, a.T_centroid)

    # -- position ---------------------------------------------------------------------------------------------------
    def transport(self, pos, return_std=True):
        if self._fused:
            if not return_std:
                # the reference dereferences an unbound name here (quirk Q8); keep the failure mode
                raise UnboundLocalError("cannot access local variable 'delta_map_std' where it is not associated with a value")
            o = self.delta_map._query(pos, _lib.MEAN | _lib.STD | _lib.AFFINE_IN | _lib.TRANSPORT)
            return o["xhat"], o["std"]
        pos_rotated = self.affine_transform.predict(pos)
        if return_std == True:  # noqa: E712
            delta_map_mean, delta_map_std = self.delta_map.predict(pos_rotated, return_std=return_std)
        else:
            delta_map_mean = self.delta_map.predict(pos_rotated, return_std=return_std)
        return pos_rotated + delta_map_mean, delta_map_std

    # -- velocity ---------------------------------------------------------------------------------------------------
    def transport_velocity(self, pos, vel, return_var=True):
        if self._fused:
            flags = _lib.JAC | _lib.AFFINE_IN | _lib.VELOCITY | _lib.JPHI
            if return_var:
                flags |= _lib.JACVAR
            o = self.delta_map._query(pos, flags, vel=vel)
            print("Is the map locally diffeomorphic?", np.all(np.abs(np.linalg.det(o["jphi"])) > 0))
            if not return_var:
                raise UnboundLocalError("cannot access local variable 'J_psi_var' where it is not associated with a value")
            return o["vhat"], o["vvar"]
        pos_rotated = self.affine_transform.predict(pos)
        J_gamma = self.affine_transform.derivative(pos)
        if return_var == True:  # noqa: E712
            J_psi, J_psi_var = self.delta_map.derivative(pos_rotated, return_var=return_var)
        else:
            J_psi = self.delta_map.derivative(pos_rotated, return_var=return_var)
        J_phi = J_gamma + J_psi @ J_gamma
        print("Is the map locally diffeomorphic?", np.all(np.abs(np.linalg.det(J_phi)) > 0))
        vel = vel[:, :, np.newaxis]
        vel_rotated = J_gamma @ vel
        var_vel_transported = J_psi_var @ vel_rotated ** 2
        vel_transported = J_phi @ vel
        return vel_transported[:, :, 0], var_vel_transported[:, :, 0]

    # -- orientation ------------------------------------------------------------------------------------------------
    def _jphi_unrotated(self, pos):
        if self._fused:
            return self.delta_map._query(pos, _lib.JAC | _lib.JPHI)["jphi"]          # un-rotated pos: quirk Q7
        J_phi = self.delta_map.derivative(pos)
        J_gamma = self.affine_transform.derivative(pos)
        return J_gamma + J_phi @ J_gamma

    def transport_orientation(self, pos, ori):
        if self._fused and np.shape(pos)[1] == 3 and self.delta_map.n_outputs == 3:
            # Jacobian, Jphi = R + Jpsi R, the 4x4 eigen-problem and the Hamilton product all run on the GPU
            self.delta_map._ensure_fitted_factor()
            ori_out, J_phi = self.delta_map._engine.transport_orientation(pos, ori)
            print("Is the map locally diffeomorphic?", np.all(np.linalg.det(J_phi) > 0))
            return ori_out
        J_phi = self._jphi_unrotated(pos)
        print("Is the map locally diffeomorphic?", np.all(np.linalg.det(J_phi) > 0))
        if J_phi[0].shape[0] == 3:
            quat_J_phi = from_rotation_matrix_nonorthogonal(J_phi)
            return quat_multiply(quat_J_phi, np.asarray(ori, dtype=np.float64))
        print("The Jacobain of the map as shape ", J_phi[0].shape, " but it should be (3x3)")
        print("Robot orientation is not transported")

    # -- stiffness -------------------------------------------------------------------------------------------------
    def transport_stiffness(self, pos, stiffness):
        """K_hat = Jphi K Jphi^T with Jphi = R + Jpsi(gamma(pos)) R, the linearisation the velocity transport uses.
        Not part of the reference code (README.md:6 announces it) -- provided because the transport
        of "position, velocity, orientation and stiffness" is the stated scope.  stiffness: (M, d, d)."""
        stiffness = np.asarray(stiffness, dtype=np.float64)
        if self._fused:
            self.delta_map._ensure_fitted_factor()
            out, _ = self.delta_map._engine.transport_stiffness(pos, stiffness)
            return out
        pos_rotated = self.affine_transform.predict(pos)
        J_gamma = self.affine_transform.derivative(pos)
        J_phi = J_gamma + self.delta_map.derivative(pos_rotated) @ J_gamma
        return J_phi @ stiffness @ np.transpose(J_phi, (0, 2, 1))

    # -- everything the façade needs, one generator pass --------------------------------------------------------------
    def transport_all(self, pos, vel=None):
        """Position (+std) and, when `vel` is given, velocity (+variance) in a single fused query (B200 path only)."""
        if not self._fused:
            raise NotImplementedError("transport_all needs the B200 GaussianProcess delta map")
        flags = _lib.MEAN | _lib.STD | _lib.AFFINE_IN | _lib.TRANSPORT
        if vel is not None:
            flags |= _lib.JAC | _lib.JACVAR | _lib.VELOCITY | _lib.JPHI
        o = self.delta_map._query(pos, flags, vel=vel)
        if vel is not None:
            print("Is the map locally diffeomorphic?", np.all(np.abs(np.linalg.det(o["jphi"])) > 0))
            return o["xhat"], o["std"], o["vhat"], o["vvar"]
        return o["xhat"], o["std"], None, None

    def sample_transportation(self, pos):
        pos_rotated = self.affine_transform.predict(pos)
        delta_map_samples = self.delta_map.samples(pos_rotated)
        return pos_rotated + delta_map_samples
