"""CPU test: the pybind11 module exports the Python-visible surface that lightning_gpu/*.py and
lightning_base/*.py consume (reference: tests/bindings/test_bindings_nb.py:51-140)."""
import numpy as np
import pytest

ops = pytest.importorskip("pennylane_lightning_b200.lightning_b200_ops")

GATES = ["Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T", "PhaseShift", "RX", "RY", "RZ", "Rot",
         "CNOT", "CY", "CZ", "SWAP", "IsingXX", "IsingXY", "IsingYY", "IsingZZ", "ControlledPhaseShift", "CRX", "CRY",
         "CRZ", "CRot", "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus", "PSWAP", "Toffoli",
         "CSWAP", "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus", "MultiRZ", "GlobalPhase",
         "PCPhase"]


def test_module_info_dicts():
    assert ops.backend_info()["NAME"] == "lightning.b200"
    assert isinstance(ops.compile_info(), dict) and isinstance(ops.runtime_info(), dict)
    assert ops.compile_info()["cuda.arch"] == "sm_100a"
    assert callable(ops.is_gpu_supported) and callable(ops.get_gpu_arch)
    assert hasattr(ops, "DevPool") and hasattr(ops, "DevTag")
    assert ops.DevTag(0).getDeviceID() == 0


@pytest.mark.parametrize("precision", ["64", "128"])
def test_classes_and_methods_exist(precision):
    sv = getattr(ops, f"StateVectorC{precision}")
    for g in GATES:
        assert hasattr(sv, g), g
    for m in ("applyMatrix", "applyControlledMatrix", "apply", "applyPauliRot", "resetStateVector", "setBasisState",
              "setStateVector", "updateData", "collapse", "DeviceToHost", "HostToDevice", "DeviceToDevice",
              "getState", "size", "__len__", "numQubits", "dataLength", "GetNumGPUs", "getCurrentGPU"):
        assert hasattr(sv, m), m
    meas = getattr(ops, f"MeasurementsC{precision}")
    for m in ("set_random_seed", "probs", "generate_samples", "expval", "var"):
        assert hasattr(meas, m), m
    for c in ("NamedObs", "HermitianObs", "TensorProdObs", "Hamiltonian", "Observable"):
        assert hasattr(ops.observables, f"{c}C{precision}")
    for c in ("AdjointJacobian", "OpsStruct", "create_ops_list"):
        assert hasattr(ops.algorithms, f"{c}C{precision}")


@pytest.mark.parametrize("precision", ["64", "128"])
def test_observable_objects_host_side(precision):
    o = ops.observables
    N, H, T, Ham = (getattr(o, f"{c}C{precision}") for c in ("NamedObs", "HermitianObs", "TensorProdObs",
                                                              "Hamiltonian"))
    z0, x1 = N("PauliZ", [0]), N("PauliX", [1])
    assert z0.get_wires() == [0] and "PauliZ" in repr(z0)
    assert z0 == N("PauliZ", [0]) and not (z0 == x1)
    t = T([z0, x1])
    assert t.get_wires() == [0, 1] and len(t.get_ops()) == 2
    h = Ham(np.array([0.3, -0.5]), [z0, t])
    np.testing.assert_allclose(h.get_coeffs(), [0.3, -0.5], rtol=1e-6)
    assert h.get_wires() == [0, 1]
    dt = np.complex64 if precision == "64" else np.complex128
    hm = H(np.array([[1, 0], [0, -1]], dtype=dt), [0])
    assert hm.get_wires() == [0]
    with pytest.raises(RuntimeError, match="size of matrix"):
        H(np.eye(2, dtype=dt), [0, 1])
    with pytest.raises(RuntimeError, match="disjoint"):
        T([z0, N("PauliY", [0])])


@pytest.mark.parametrize("precision", ["64", "128"])
def test_create_ops_list_signature(precision):
    """tests/bindings/test_adjoint_jacobian_nb.py:331-368"""
    dt = np.complex64 if precision == "64" else np.complex128
    create = getattr(ops.algorithms, f"create_ops_listC{precision}")
    s = create(["RX"], [[0.5]], [[0]], [False], [np.array([[1, 0], [0, -1]], dtype=dt)], [[]], [[]])
    assert s.__class__.__name__ == f"OpsStructC{precision}"
