"""batman_b200: B200-native drop-in for batman's Kriging-over-Saltelli -> Sobol' hot path.

    from batman_b200 import Kriging, SurrogateModel, UQ

mirror ``batman.surrogate.Kriging`` / ``SurrogateModel(kind='kriging')`` and
``batman.uq.UQ.sobol()``.  All arithmetic runs in hand-written sm_100a CUDA
kernels behind the C ABI of ``include/batman_b200.h`` (``libbatman_b200.so``,
loaded with ctypes).  There is no CPU fallback.
"""
from ._lib import BatmanB200Error, load as load_library  # noqa: F401
from .kriging import Kriging  # noqa: F401
from .mixture import Mixture  # noqa: F401
from .multifidelity import Evofusion  # noqa: F401
from .refine import Refiner  # noqa: F401
from .surrogate_model import Pod, SurrogateModel  # noqa: F401
from .uq import UQ  # noqa: F401

__version__ = '0.1.0'
