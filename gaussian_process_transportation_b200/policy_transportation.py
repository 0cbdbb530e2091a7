"""The transport flow phi(x) = gamma(x) + psi(gamma(x)) -- drop-in for
policy_transportation/transportation/policy_transportation.py:11-84.

`method` is the delta-map plugin slot of the reference (fit / predict / derivative / samples); here it must be the B200
`GaussianProcess` (or a subclass): the affine prologue, the posterior and the Jacobian algebra of each call run as ONE
fused GPU query (include/gptb200.h flags).  Any other plugin is refused at construction -- this package has no CPU path.
"""
import numpy as np

from . import _lib
from .affine_transform import AffineTransform
from .gaussian_process import GaussianProcess


class PolicyTransportation():
    def __init__(self, method):
        super(PolicyTransportation, self).__init__()
        if not isinstance(method, GaussianProcess):
            raise NotImplementedError(
                "gaussian_process_transportation_b200.PolicyTransportation drives the B200 GaussianProcess engine only "
                f"(got {type(method).__name__}); there is no CPU fallback for other delta maps")
        self.delta_map = method

    def fit(self, source_distribution, target_distribution, do_scale=False, do_rotation=True):
        self.affine_transform = AffineTransform(do_scale=do_scale, do_rotation=do_rotation)
        self.affine_transform.fit(source_distribution, target_distribution)
        source_distribution = self.affine_transform.predict(source_distribution)
        self.delta_distribution = target_distribution - source_distribution
        self.delta_map.fit(source_distribution, self.delta_distribution)
        a = self.affine_transform
        self.delta_map._engine.set_affine(a.rotation_matrix, float(a.scale), a.S_centroid, a.T_centroid)

    # -- position (file:26-35) ----------------------------------------------------------------------------------------
    def transport(self, pos, return_std=True):
        if not return_std:
            # the reference dereferences an unbound name here (quirk Q8); keep the failure mode
            raise UnboundLocalError("cannot access local variable 'delta_map_std' where it is not associated with a value")
        o = self.delta_map._query(pos, _lib.MEAN | _lib.STD | _lib.AFFINE_IN | _lib.TRANSPORT)
        return o["xhat"], o["std"]

    # -- velocity (file:37-59) ----------------------------------------------------------------------------------------
    def transport_velocity(self, pos, vel, return_var=True):
        flags = _lib.JAC | _lib.AFFINE_IN | _lib.VELOCITY | _lib.JPHI
        if return_var:
            flags |= _lib.JACVAR
        o = self.delta_map._query(pos, flags, vel=vel)
        print("Is the map locally diffeomorphic?", np.all(np.abs(np.linalg.det(o["jphi"])) > 0))
        if not return_var:
            raise UnboundLocalError("cannot access local variable 'J_psi_var' where it is not associated with a value")
        return o["vhat"], o["vvar"]

    # -- orientation (file:61-77) -------------------------------------------------------------------------------------
    def transport_orientation(self, pos, ori):
        if np.shape(pos)[1] == 3 and self.delta_map.n_outputs == 3:
            # Jacobian at the UN-rotated positions (quirk Q7), Jphi = R + Jpsi R, the 4x4 eigen-problem and the Hamilton
            # product all run on the GPU
            self.delta_map._ensure_fitted_factor()
            ori_out, J_phi = self.delta_map._engine.transport_orientation(pos, ori)
            print("Is the map locally diffeomorphic?", np.all(np.linalg.det(J_phi) > 0))
            return ori_out
        J_phi = self.delta_map._query(pos, _lib.JAC | _lib.JPHI)["jphi"]
        print("Is the map locally diffeomorphic?", np.all(np.linalg.det(J_phi) > 0))
        print("The Jacobain of the map as shape ", J_phi[0].shape, " but it should be (3x3)")
        print("Robot orientation is not transported")

    # -- stiffness ----------------------------------------------------------------------------------------------------
    def transport_stiffness(self, pos, stiffness):
        """K_hat = Jphi K Jphi^T with Jphi = R + Jpsi(gamma(pos)) R, the linearisation the velocity transport uses.
        Not part of the reference code (README.md:6 announces it) -- provided because the transport
        of "position, velocity, orientation and stiffness" is the stated scope.  stiffness: (M, d, d)."""
        self.delta_map._ensure_fitted_factor()
        out, _ = self.delta_map._engine.transport_stiffness(pos, np.asarray(stiffness, dtype=np.float64))
        return out

    # -- everything the facade needs, one generator pass --------------------------------------------------------------
    def transport_all(self, pos, vel=None):
        """Position (+std) and, when `vel` is given, velocity (+variance) in a single fused query."""
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
