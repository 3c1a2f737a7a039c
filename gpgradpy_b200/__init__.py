"""gpgradpy_b200: B200-native (sm_100a CUDA) gradient-enhanced Gaussian-process hot path behind the
GpGradPy `GaussianProcess` API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"


def __getattr__(name):   # `from gpgradpy_b200 import GaussianProcess` without importing torch at package import
    if name == "GaussianProcess":
        from .gp import GaussianProcess
        return GaussianProcess
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
