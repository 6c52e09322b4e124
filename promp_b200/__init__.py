"""promp_b200 - B200-native (sm_100a) implementation of the ProMP data-parallel hot path.

The Python classes mirror the reference's operator interface for this path (same names, argument
meaning and error behaviour) and call hand-written CUDA kernels through the C ABI
of include/promp_b200.h.  PyTorch tensors are device buffers only.
"""
__version__ = "0.1.0"
