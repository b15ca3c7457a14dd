import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import pennylane_lightning_b200 as plb
        import ctypes

        n = ctypes.c_int(0)
        if not os.path.exists(plb.LIB_PATH):
            return False
        rc = plb.lib().plb200_device_count(ctypes.byref(n))
        return rc == 0 and n.value > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference lightning.qubit core (oracle/_ref/liblq_ref.so)."""
    from oracle import lq_ref

    if not lq_ref.available():
        pytest.skip("oracle/_ref/liblq_ref.so not built")
    return lq_ref


@pytest.fixture(scope="session")
def plb():
    import pennylane_lightning_b200 as p

    return p


def random_state(n, dtype, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


TOL = {np.dtype(np.complex128): 1e-12, np.dtype(np.complex64): 1e-5}

# name -> (num wires, num params); -1 wires = variable   (core/gates/Constant.hpp)
GATES = {
    "Identity": (1, 0), "PauliX": (1, 0), "PauliY": (1, 0), "PauliZ": (1, 0), "Hadamard": (1, 0),
    "S": (1, 0), "SX": (1, 0), "T": (1, 0), "PhaseShift": (1, 1), "RX": (1, 1), "RY": (1, 1), "RZ": (1, 1),
    "Rot": (1, 3), "CNOT": (2, 0), "CY": (2, 0), "CZ": (2, 0), "SWAP": (2, 0), "IsingXX": (2, 1),
    "IsingXY": (2, 1), "IsingYY": (2, 1), "IsingZZ": (2, 1), "ControlledPhaseShift": (2, 1), "CRX": (2, 1),
    "CRY": (2, 1), "CRZ": (2, 1), "CRot": (2, 3), "SingleExcitation": (2, 1), "SingleExcitationMinus": (2, 1),
    "SingleExcitationPlus": (2, 1), "PSWAP": (2, 1), "Toffoli": (3, 0), "CSWAP": (3, 0),
    "DoubleExcitation": (4, 1), "DoubleExcitationMinus": (4, 1), "DoubleExcitationPlus": (4, 1),
    "MultiRZ": (-1, 1), "GlobalPhase": (-1, 1), "PCPhase": (-1, 2),
}
CONTROLLED_GATES = [
    "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T", "PhaseShift", "RX", "RY", "RZ", "Rot", "SWAP",
    "IsingXX", "IsingXY", "IsingYY", "IsingZZ", "SingleExcitation", "SingleExcitationMinus",
    "SingleExcitationPlus", "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus", "PSWAP",
    "MultiRZ", "GlobalPhase", "PCPhase",
]
GENERATORS = {
    "PhaseShift": 1, "RX": 1, "RY": 1, "RZ": 1, "IsingXX": 2, "IsingXY": 2, "IsingYY": 2, "IsingZZ": 2,
    "CRX": 2, "CRY": 2, "CRZ": 2, "ControlledPhaseShift": 2, "SingleExcitation": 2,
    "SingleExcitationMinus": 2, "SingleExcitationPlus": 2, "DoubleExcitation": 4, "DoubleExcitationMinus": 4,
    "DoubleExcitationPlus": 4, "PSWAP": 2, "MultiRZ": -1, "GlobalPhase": -1,
}
CONTROLLED_GENERATORS = [g for g in GENERATORS if g not in ("CRX", "CRY", "CRZ", "ControlledPhaseShift")]
