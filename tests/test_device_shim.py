"""the ``lightning.b200`` Python device layer (pennylane-lightning_b200/device/).

PennyLane is not installed in this image, so the shim is exercised against STUBS of the few PennyLane /
lightning_base names it touches: what is checked is the shim's own logic — which binding each kind of operation is
routed to, in which order, with which arguments — and that the numbers coming back equal the reference core's.
CPU: the capability TOML parses and lists only gates the engine knows.  GPU: _apply_lightning -> Measurements ->
AdjointJacobian through the stubs."""
import ctypes as C
import importlib
import os
import sys
import tomllib
import types

import numpy as np
import pytest

from conftest import CONTROLLED_GATES, GATES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOML = os.path.join(ROOT, "pennylane-lightning_b200", "device", "lightning_b200.toml")


def test_capability_file_lists_engine_gates(plb):
    caps = tomllib.load(open(TOML, "rb"))
    gates = caps["operators"]["gates"]
    assert caps["schema"] == 3 and {"ExpectationMP", "SampleMP"} <= set(caps["measurement_processes"])
    lib = plb.lib()
    for name, props in gates.items():
        if name in ("PauliRot", "QubitUnitary"):
            continue
        nw, npar = GATES[name]
        nw = 3 if nw < 0 else nw
        w = (C.c_int64 * nw)(*range(nw))
        p = (C.c_double * max(npar, 1))(*([0.3] * max(npar, 1)))
        if name == "PCPhase":
            p[1] = 2.0
        assert lib.plb200_validate_op(C.c_int64(8), name.encode(), None, None, C.c_int64(0), w, C.c_int64(nw), 0, p,
                                      C.c_int64(npar)) == 0, (name, lib.plb200_last_error())
        assert ("controllable" in props.get("properties", [])) == (name in CONTROLLED_GATES), name
    assert "SparseHamiltonian" in caps["operators"]["observables"]


# ------------------------------------------------------------------------------------------- stubs
class _Op:
    def __init__(self, name, wires, params=(), **hyper):
        self.name, self.wires, self.parameters = name, list(wires), list(params)
        self.hyperparameters = self._hyperparameters = hyper
        self.hash = 1234


class _Adjoint(_Op):
    def __init__(self, base):
        super().__init__(f"Adjoint({base.name})", base.wires, base.parameters)
        self.base = base


class _Controlled(_Op):
    def __init__(self, base, control_wires, control_values):
        super().__init__(f"C({base.name})", list(control_wires) + list(base.wires), base.parameters)
        self.base, self.control_wires, self.control_values, self.target_wires = base, control_wires, control_values, base.wires


def _install_stubs():
    qp = types.ModuleType("pennylane")

    class Identity(_Op):
        pass

    class PauliRot(_Op):
        pass

    qp.Identity, qp.PauliRot = Identity, PauliRot
    qp.matrix = lambda op: op.matrix_value
    qp.ops = types.ModuleType("pennylane.ops")
    qp.ops.Controlled = _Controlled
    qp.ops.Conditional = type("Conditional", (_Op,), {})
    qp.ops.op_math = types.ModuleType("pennylane.ops.op_math")
    qp.ops.op_math.Adjoint = _Adjoint
    qp.exceptions = types.ModuleType("pennylane.exceptions")
    qp.exceptions.DeviceError = type("DeviceError", (Exception,), {})
    qp.measurements = types.ModuleType("pennylane.measurements")
    qp.measurements.MidMeasureMP = type("MidMeasureMP", (_Op,), {})
    qp.wires = types.ModuleType("pennylane.wires")
    qp.wires.Wires = lambda w: list(w)
    qp.pauli = types.ModuleType("pennylane.pauli")
    qp.pauli.pauli_word_to_string = lambda p: p.word
    mods = {"pennylane": qp, "pennylane.ops": qp.ops, "pennylane.ops.op_math": qp.ops.op_math,
            "pennylane.exceptions": qp.exceptions, "pennylane.measurements": qp.measurements, "pennylane.wires": qp.wires,
            "pennylane.pauli": qp.pauli}

    base_sv = types.ModuleType("pennylane_lightning.lightning_base._state_vector")

    class LightningBaseStateVector:
        def __init__(self, num_wires, dtype, rng=None):
            self._num_wires, self._dtype, self._rng = num_wires, dtype, rng

        dtype = property(lambda self: self._dtype)
        num_wires = property(lambda self: self._num_wires)
        state_vector = property(lambda self: self._qubit_state)

        def apply_operations(self, operations):
            self._apply_lightning(operations)

    base_sv.LightningBaseStateVector = LightningBaseStateVector
    base_m = types.ModuleType("pennylane_lightning.lightning_base._measurements")

    class LightningBaseMeasurements:
        def __init__(self, qubit_state):
            self._qubit_state, self._dtype = qubit_state, qubit_state.dtype

        dtype = property(lambda self: self._dtype)

    base_m.LightningBaseMeasurements = LightningBaseMeasurements
    base_a = types.ModuleType("pennylane_lightning.lightning_base._adjoint_jacobian")

    class LightningBaseAdjointJacobian:
        def __init__(self, qubit_state, batch_obs=False):
            self._qubit_state, self._batch_obs, self._dtype = qubit_state, batch_obs, qubit_state.dtype
            self._jacobian_lightning, self._create_ops_list_lightning = self._adjoint_jacobian_dtype()

        dtype = property(lambda self: self._dtype)

        def _handle_raises(self, tape, is_jacobian):
            return False

        def _process_jacobian_tape(self, tape, split_obs, use_mpi):
            names, params, wires, inv = zip(*[(o.name, o.parameters, o.wires, False) for o in tape["ops"]])
            dt = self.dtype
            ops = self._create_ops_list_lightning(list(names), [list(p) for p in params], [list(w) for w in wires],
                                                  list(inv), [np.zeros(0, dtype=dt)] * len(names), [[]] * len(names),
                                                  [[]] * len(names))
            return dict(state_vector=self._qubit_state.state_vector, obs_serialized=tape["obs"], ops_serialized=ops,
                        tp_shift=tape["tp"], record_tp_rows=list(range(len(tape["tp"]))), all_params=len(tape["tp"]),
                        obs_indices=list(range(len(tape["obs"]))))

        @staticmethod
        def _adjoint_jacobian_processing(jac):
            return np.squeeze(jac)

    base_a.LightningBaseAdjointJacobian = LightningBaseAdjointJacobian
    pl = types.ModuleType("pennylane_lightning")
    lb = types.ModuleType("pennylane_lightning.lightning_base")
    mods.update({"pennylane_lightning": pl, "pennylane_lightning.lightning_base": lb,
                 "pennylane_lightning.lightning_base._state_vector": base_sv,
                 "pennylane_lightning.lightning_base._measurements": base_m,
                 "pennylane_lightning.lightning_base._adjoint_jacobian": base_a})
    return mods


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_shim_call_order_with_stub_pennylane(plb, ref, dtype, monkeypatch):
    pytest.importorskip("pennylane_lightning_b200.lightning_b200_ops")
    for k, v in _install_stubs().items():
        monkeypatch.setitem(sys.modules, k, v)
    for m in ("_measurements", "_state_vector", "_adjoint_jacobian"):
        sys.modules.pop(f"pennylane_lightning_b200.device.{m}", None)
    svm = importlib.import_module("pennylane_lightning_b200.device._state_vector")
    mm = importlib.import_module("pennylane_lightning_b200.device._measurements")
    am = importlib.import_module("pennylane_lightning_b200.device._adjoint_jacobian")
    qp = sys.modules["pennylane"]
    n = 5
    sv = svm.LightningB200StateVector(n, dtype, rng=np.random.default_rng(3))
    u = np.array([[0, 1j], [1j, 0]], dtype=dtype)
    unitary = _Op("QubitUnitary", [4])
    unitary.matrix_value = u
    ops = [_Op("RX", [0], [0.3]), _Op("CNOT", [0, 1]), _Adjoint(_Op("RY", [2], [0.7])), qp.Identity("Identity", [3]),
           qp.PauliRot("PauliRot", [0, 1, 3], [0.5], pauli_word="XIZ"),
           _Controlled(_Op("RZ", [3], [0.9]), [1, 2], [True, False]), unitary, _Op("IsingXX", [2, 4], [0.2])]
    sv.apply_operations(ops)
    assert sv.state_vector.pendingOps() > 0 if hasattr(sv.state_vector, "pendingOps") else True
    # the same circuit on the reference core
    from pennylane_lightning_b200 import circuits
    r = ref.StateVector(n, dtype)
    r.apply("RX", [0], False, [0.3]); r.apply("CNOT", [0, 1]); r.apply("RY", [2], True, [0.7])
    r.apply_pauli_rot([0, 3], False, 0.5, "XZ")
    r.apply("RZ", [3], False, [0.9], [1, 2], [True, False]); r.apply_matrix(u, [4]); r.apply("IsingXX", [2, 4], False, [0.2])
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    np.testing.assert_allclose(sv.state, r.get_state(), rtol=0, atol=tol)
    # measurements: fused Pauli sentence
    meas = mm.LightningB200Measurements(sv)
    class word:  # stands for a qml.pauli.PauliWord: hashable, has .wires
        def __init__(self, w, ws):
            self.word, self.wires = w, np.array(ws)

    mp = types.SimpleNamespace(obs=types.SimpleNamespace(pauli_rep={word("XZ", [0, 2]): 0.5, word("Y", [1]): -1.5}))
    want = 0.5 * r.expval(ref.Observable.tensor([ref.Observable.named("PauliX", [0], dtype=dtype),
                                                 ref.Observable.named("PauliZ", [2], dtype=dtype)])) - 1.5 * r.expval(
        ref.Observable.named("PauliY", [1], dtype=dtype))
    assert abs(meas._expval_pauli_sentence(mp) - want) < 10 * tol
    # adjoint Jacobian through the serialised-tape path
    ops_mod = sys.modules["pennylane_lightning_b200.lightning_b200_ops"]
    bits = "128" if dtype == np.complex128 else "64"
    tape_ops = [_Op("RX", [0], [0.3]), _Op("RY", [1], [0.4]), _Op("CNOT", [0, 1]), _Op("RZ", [1], [0.5])]
    sv2 = svm.LightningB200StateVector(2, dtype)
    sv2.apply_operations(tape_ops)
    obs = [getattr(ops_mod.observables, f"NamedObsC{bits}")("PauliZ", [1]), getattr(ops_mod.observables, f"NamedObsC{bits}")("PauliX", [0])]
    jac = am.LightningB200AdjointJacobian(sv2).calculate_jacobian(dict(ops=tape_ops, obs=obs, tp=[0, 1, 2]))
    plain = [circuits.op(o.name, o.wires, o.parameters) for o in tape_ops]
    r2 = ref.StateVector(2, dtype)
    jr = r2.adjoint_jacobian([ref.Observable.named("PauliZ", [1], dtype=dtype), ref.Observable.named("PauliX", [0], dtype=dtype)],
                             plain, [0, 1, 2], apply_ops=True)
    np.testing.assert_allclose(jac, jr, rtol=0, atol=100 * tol)
