"""B200-native state-vector engine behind pennylane-lightning's hot-path interfaces.

Layout:
  csrc/      CUDA kernels (sm_100a) + the C ABI (include/plb200.h) -> lib/libplb200.so
  host/      C++ mirrors of StateVectorBase / Measurements / Observables / AdjointJacobian and the
             pybind11 module ``lightning_b200_ops`` exporting lightning_gpu_ops' Python surface
  _capi.py   ctypes view of the C ABI (tests, bench, distributed driver)
  dist.py    one-process-per-GPU sharded state vector over torch.distributed (NCCL / NVLink)

There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from ._capi import (  # noqa: F401
    LIB_PATH,
    B200Error,
    Observable,
    OpsBlob,
    StateVector,
    build,
    hermitian_eigh,
    jit_available,
    jit_enabled,
    jit_mode,
    jit_set_mode,
    jit_stats,
    jit_wait,
    lib,
)

__version__ = "0.1.0"
