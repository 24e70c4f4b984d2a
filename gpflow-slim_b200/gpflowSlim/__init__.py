"""gpflowSlim -- B200-native drop-in for the GP inference hot path of ssydasheng/GPflow-Slim.

Same import surface as the reference package (gpflowSlim/__init__.py): `settings`, `kernels`,
`models`, `conditionals`, `features`, `kullback_leiblers`, `densities`, `likelihoods`,
`mean_functions`, `transforms`, `neural_kernel_network`, `Param`.  Tensors are torch CUDA
float64; all O(n^2)/O(n^3) arithmetic runs in libgpslim_b200.so (hand-written sm_100a CUDA,
C ABI in include/gpslim_b200.h).  No TensorFlow, no CPU fallback.
"""
from ._settings import SETTINGS as settings

from . import misc
from . import transforms
from . import densities
from . import quadrature
from . import likelihoods
from . import kernels
from . import kernel_kitchen_sink
from . import conditionals
from . import features
from . import kullback_leiblers
from . import mean_functions
from . import priors
from . import models
from . import neural_kernel_network
from . import training
from . import LBFGS as _LBFGS_module  # noqa: F401  (importable as gpflowSlim.LBFGS)
from . import parallel

from .params import Parameter as Param
from .params import Parameter
from ._backend.lib import CholeskyError

__version__ = '0.1.0'
