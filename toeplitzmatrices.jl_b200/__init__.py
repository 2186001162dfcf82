"""toeplitzmatrices.jl_b200 -- B200-native FFT hot path of ToeplitzMatrices.jl.

The directory name carries a dot, so it is imported through the repo-root shim
``toeplitz_b200`` (``import toeplitz_b200 as tb``).  Only what the hot path needs lives
here: ``csrc/`` (CUDA kernels + the C ABI), ``build.py``, ``_lib.py`` (ctypes binding) and
``api.py`` (host-side mirror of the reference's operator interface) and ``containers.py`` (O(n)
container algebra on the device-resident coefficient vectors).
"""
from ._lib import (ArgumentError, DimensionMismatch, ToeplitzB200Error, EXPORTS, LIB_PATH)  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (Toeplitz, SymmetricToeplitz, Circulant, LowerTriangularToeplitz, UpperTriangularToeplitz,  # noqa: F401
                  Hankel, ToeplitzFactorization, factorize, mul_, mul, ldiv_, ldiv, circ_ldiv_, eigvals, det, inv,
                  pinv, sqrt, circ_mul, adjoint, transpose, strang, chan, durbin, durbin_, levinson, levinson_, trench,
                  cgs, cg, cgs_batched, cg_batched, colmajor, to_colmajor, mul_host_, factorize_host, spectrum_tensor,
                  factorize_empty_like, spectrum_ready, launch_count, set_option, pass_times)
from . import containers  # noqa: F401,E402  (also installs + - * == on the matrix types)
from .containers import (tril, triu, tril_, triu_, reverse, copyto_, fill_, lmul_, rmul_, isdiag, iszero, getindex,  # noqa: F401,E402
                         dense)
