"""diffdrr_b200 -- a Blackwell-native (sm_100a) differentiable DRR projector with DiffDRR's `DRR` module surface.

Hot path (Siddon + trilinear renderers, forward and backward) = hand-written CUDA kernels behind the C ABI of
include/b200drr.h; host code (pose, detector geometry, module plumbing) = PyTorch.  No CPU fallback.
"""
__version__ = "0.1.0"

from .detector import Detector  # noqa: F401
from .drr import DRR  # noqa: F401
from .pose import RigidTransform, convert  # noqa: F401
from .renderers import Siddon, Trilinear  # noqa: F401
from .metrics import NormalizedCrossCorrelation2d  # noqa: F401,E402
from .registration import Registration  # noqa: F401,E402
