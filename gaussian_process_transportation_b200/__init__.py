"""gaussian_process_transportation_b200 -- B200-native (sm_100a) exact-GP transport posterior engine behind the
reference's Python API (policy_transportation's GaussianProcess / AffineTransform / PolicyTransportation /
GaussianProcessTransportation).  All posterior numerics run in hand-written CUDA kernels behind a C ABI
(include/gptb200.h); there is no CPU fallback."""
from .affine_transform import AffineTransform
from .gaussian_process import GaussianProcess
from .policy_transportation import PolicyTransportation
from .gaussian_process_transportation import GaussianProcessTransportation
from .gaussian_process_transportation_diffeomorphic import GaussianProcessTransportationDiffeo
from .gaussian_process_al import GaussianProcess as GaussianProcessAL

__all__ = ['AffineTransform', 'GaussianProcessTransportation', 'GaussianProcess', 'PolicyTransportation',
           'GaussianProcessTransportationDiffeo', 'GaussianProcessAL']
__version__ = "0.1.0"
