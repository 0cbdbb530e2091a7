"""Attribute-driven façade used by every script of the reference -- drop-in for
policy_transportation/transportation/gaussian_process_transportation.py:11-30.  Mixin-friendly: no required constructor
arguments, calls super().__init__() (composed as `class GPT_surface(Transport, Surface_PointCloud_Detector, SIMPLe)`,
modules.py:13-17)."""
from sklearn.gaussian_process.kernels import RBF, WhiteKernel, ConstantKernel as C

from .gaussian_process import GaussianProcess
from .policy_transportation import PolicyTransportation


class GaussianProcessTransportation():
    def __init__(self, kernel_transport=C(0.1) * RBF(length_scale=[0.1]) + WhiteKernel(0.0001)):
        super(GaussianProcessTransportation, self).__init__()
        self.method = PolicyTransportation(GaussianProcess(kernel=kernel_transport))

    def fit_transportation(self, do_scale=False, do_rotation=True):
        self.method.fit(self.source_distribution, self.target_distribution, do_scale=do_scale, do_rotation=do_rotation)

    def apply_transportation(self):
        self.training_traj_old = self.training_traj
        if hasattr(self, 'training_delta'):
            # one fused pass: position, std, velocity and its variance share the regenerated k(x*, X)
            self.training_traj, self.std, self.training_delta, self.var_vel_transported = \
                self.method.transport_all(self.training_traj_old, self.training_delta)
        else:
            self.training_traj, self.std = self.method.transport(self.training_traj_old)
        if hasattr(self, 'training_ori'):
            self.training_ori = self.method.transport_orientation(self.training_traj_old, self.training_ori)
        if hasattr(self, 'training_stiff'):
            # not in the reference façade (its README announces stiffness transport, the code has none): K_hat = Jphi K Jphi^T
            self.training_stiff = self.method.transport_stiffness(self.training_traj_old, self.training_stiff)

    def sample_transportation(self):
        return self.method.sample_transportation(self.training_traj_old)
