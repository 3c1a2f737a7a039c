"""gpgradpy_b200: B200-native (sm_100a CUDA) gradient-enhanced Gaussian-process hot path behind the
GpGradPy `GaussianProcess` API.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
