"""GPU tests of the pybind11 module, following the reference's raw-binding tests
(tests/bindings/test_statevector_nb.py, test_measurements_nb.py, test_adjoint_jacobian_nb.py,
test_backend_nb.py) with the module swapped for lightning_b200_ops, plus parity against the
reference lightning.qubit core through the same Python-visible calls."""
import numpy as np
import pytest

from pennylane_lightning_b200 import circuits

pytestmark = pytest.mark.gpu
ops = pytest.importorskip("pennylane_lightning_b200.lightning_b200_ops")
PREC = ["64", "128"]


def classes(p):
    dt = np.complex64 if p == "64" else np.complex128
    return (getattr(ops, f"StateVectorC{p}"), getattr(ops, f"MeasurementsC{p}"), dt)


def state(sv, dt):
    out = np.zeros(len(sv), dtype=dt)
    sv.getState(out)
    return out


@pytest.mark.parametrize("p", PREC)
def test_statevector_basics(p):
    SV, _, dt = classes(p)
    sv = SV(3)
    assert sv.size() == 8 and len(sv) == 8 and sv.numQubits() == 3
    exp = np.zeros(8, dtype=dt)
    exp[0] = 1
    np.testing.assert_allclose(state(sv, dt), exp)
    sv.PauliX([0], False, [])
    sv.Hadamard([0], True, [])
    sv.Hadamard([1], False, [])
    sv.CNOT([0, 1], True, [])
    sv.RX([0], False, [np.pi / 2])
    sv.RY([0], False, [np.pi])
    sv.RZ([0], False, [np.pi / 2])
    sv.resetStateVector()
    np.testing.assert_allclose(state(sv, dt), exp)
    sv.applyMatrix(np.array([[0, 1], [1, 0]], dtype=dt), [0], False)
    assert np.isclose(state(sv, dt)[4], 1.0)
    sv.setBasisState([1, 0, 1], [0, 1, 2])
    assert np.isclose(state(sv, dt)[5], 1.0)
    sv.resetStateVector()
    sv.setBasisState([1, 1, 0], [2, 1, 0])
    assert np.isclose(state(sv, dt)[3], 1.0)
    sv2 = SV(2)
    sup = np.array([1, 1], dtype=dt) / np.sqrt(2)
    sv2.setStateVector(sup, [0])
    np.testing.assert_allclose(state(sv2, dt), [sup[0], 0, sup[0], 0], atol=1e-6)
    data = (np.array([1, 2, 3, 4], dtype=dt) / np.sqrt(30)).astype(dt)
    sv2.updateData(data)
    np.testing.assert_allclose(state(sv2, dt), data, atol=1e-7)
    sv3 = SV(data)
    out = np.zeros(4, dtype=dt)
    sv3.DeviceToHost(out, False)
    np.testing.assert_allclose(out, data, atol=1e-7)
    sv3.HostToDevice(data[::-1].copy(), False)
    sv2.DeviceToDevice(sv3, False)
    np.testing.assert_allclose(state(sv2, dt), data[::-1], atol=1e-7)


@pytest.mark.parametrize("p", PREC)
def test_measurements(p):
    SV, M, dt = classes(p)
    N = getattr(ops.observables, f"NamedObsC{p}")
    H = getattr(ops.observables, f"HermitianObsC{p}")
    sv = SV(2)
    m = M(sv)
    np.testing.assert_allclose(m.probs([0, 1]), [1, 0, 0, 0], atol=1e-6)
    sv.Hadamard([0], False, [])
    sv.CNOT([0, 1], False, [])
    np.testing.assert_allclose(M(sv).probs([0, 1]), [0.5, 0, 0, 0.5], atol=1e-6)
    np.testing.assert_allclose(M(sv).probs(), [0.5, 0, 0, 0.5], atol=1e-6)
    s = m.generate_samples(2, 1000)
    assert s.shape == (1000, 2)
    c00, c11 = np.sum(np.all(s == [0, 0], axis=1)), np.sum(np.all(s == [1, 1], axis=1))
    assert c00 + c11 == 1000 and 400 <= c00 <= 600
    one = SV(1)
    one.Hadamard([0], False, [])
    mm = M(one)
    assert np.isclose(mm.expval(N("PauliX", [0])), 1.0, atol=1e-6)
    assert np.isclose(mm.expval(N("PauliZ", [0])), 0.0, atol=1e-6)
    assert np.isclose(mm.var(N("PauliZ", [0])), 1.0, atol=1e-6)
    assert np.isclose(mm.expval("PauliX", [0]), 1.0, atol=1e-6)
    assert np.isclose(mm.var("PauliX", [0]), 0.0, atol=1e-6)
    z = np.array([[1, 0], [0, -1]], dtype=dt)
    assert np.isclose(mm.expval(H(z, [0])), 0.0, atol=1e-6)
    assert np.isclose(mm.var(H(z, [0])), 1.0, atol=1e-6)
    assert np.isclose(mm.expval(z, [0]), 0.0, atol=1e-6)
    assert np.isclose(mm.expval(["X", "Z"], [[0], [0]], np.array([2.0, 3.0])), 2.0, atol=1e-6)
    # controlled-gate overloads (test_measurements_nb.py:336-454)
    g = SV(3)
    g.Hadamard([0], False, [])
    g.PauliX([0], [True], [1], False, [])
    g.PauliX([0, 1], [True, True], [2], False, [])
    pr = M(g).probs([0, 1, 2])
    assert np.isclose(pr[0], 0.5, atol=1e-6) and np.isclose(pr[7], 0.5, atol=1e-6)
    g.collapse(0, True)
    assert np.isclose(M(g).probs([0, 1, 2])[7], 1.0, atol=1e-6)


@pytest.mark.parametrize("p", PREC)
def test_adjoint_jacobian_binding(p):
    """tests/bindings/test_adjoint_jacobian_nb.py:99-330"""
    SV, _, dt = classes(p)
    N = getattr(ops.observables, f"NamedObsC{p}")
    H = getattr(ops.observables, f"HermitianObsC{p}")
    Ops = getattr(ops.algorithms, f"OpsStructC{p}")
    Adj = getattr(ops.algorithms, f"AdjointJacobianC{p}")
    th = 0.5
    for gate, expected in (("RX", (0.0, -np.cos(th), -np.sin(th))), ("RY", (np.cos(th), 0.0, -np.sin(th))),
                           ("RZ", (0.0, 0.0, 0.0))):
        sv = SV(2)
        getattr(sv, gate)([0], False, [th])
        o = Ops([gate], [[th]], [[0]], [False], [np.array([], dtype=dt)], [[]], [[]])
        res = Adj()(sv, [N("PauliX", [0]), N("PauliY", [0]), N("PauliZ", [0])], o, [0])
        assert isinstance(res, np.ndarray) and res.shape == (3,)
        np.testing.assert_allclose(res, expected, atol=1e-6)
    sv = SV(2)
    sv.RX([0], False, [0.3])
    sv.RY([1], False, [0.7])
    o = Ops(["RX", "RY"], [[0.3], [0.7]], [[0], [1]], [False, False], [np.array([], dtype=dt)] * 2, [[], []], [[], []])
    res = Adj()(sv, [N("PauliZ", [0]), N("PauliZ", [1])], o, [0, 1])
    np.testing.assert_allclose(res, [-np.sin(0.3), 0, 0, -np.sin(0.7)], atol=1e-6)
    z = np.array([[1, 0], [0, -1]], dtype=dt)
    res = Adj().batched(sv, [H(z, [0])], o, [0])
    np.testing.assert_allclose(res, [-np.sin(0.3)], atol=1e-6)
    # observable batching: with several GPUs every chunk of observables runs on its own device (one thread each);
    # the rows must come back in order and equal the single-device sweep
    many = [N("PauliZ", [0]), N("PauliZ", [1]), N("PauliX", [0]), N("PauliY", [1]), H(z, [1])]
    np.testing.assert_allclose(Adj().batched(sv, many, o, [0, 1]), Adj()(sv, many, o, [0, 1]), atol=1e-6)
    assert ops.DevPool.getTotalDevices() >= 1


def test_full_flow_matches_reference(ref):
    """The Python-visible call sequence of lightning_base (_apply_lightning -> Measurements -> adjoint)
    against lightning.qubit on config-1's circuit family."""
    n = 10
    tape, tp = circuits.strongly_entangling_layers(n, 2, 42)
    SV, M, dt = classes("128")
    sv = SV(n)
    for o in tape:
        getattr(sv, o["name"])(o["wires"], o["inverse"], o["params"])
    sv.applyPauliRot([3, 1], False, [0.41], "XY")
    sv.RY([7], [True], [2], False, [0.77])
    sv.apply("QubitUnitary", [4], False, [], np.array([[0, 1j], [1j, 0]]).ravel())
    r = ref.StateVector(n)
    r.apply_ops(tape)
    r.apply_pauli_rot([3, 1], False, 0.41, "XY")
    r.apply("RY", [2], False, [0.77], [7], [True])
    r.apply_matrix(np.array([[0, 1j], [1j, 0]]), [4])
    np.testing.assert_allclose(state(sv, dt), r.get_state(), rtol=0, atol=1e-12)
    N = ops.observables.NamedObsC128
    T = ops.observables.TensorProdObsC128
    Ham = ops.observables.HamiltonianC128
    ob = Ham(np.array([0.4, -1.1]), [N("PauliZ", [0]), T([N("PauliX", [1]), N("PauliY", [5])])])
    rob = ref.Observable.hamiltonian([0.4, -1.1], [ref.Observable.named("PauliZ", [0]), ref.Observable.tensor(
        [ref.Observable.named("PauliX", [1]), ref.Observable.named("PauliY", [5])])])
    m = M(sv)
    assert abs(m.expval(ob) - r.expval(rob)) < 1e-12
    assert abs(m.var(ob) - r.var(rob)) < 1e-11
    m.set_random_seed(42)
    np.testing.assert_array_equal(m.generate_samples(n, 100), r.generate_samples(100, seed=42))
    full = tape + [dict(name="PauliRot_skip", wires=[0], params=[], inverse=False)][:0]
    sv2, r2 = SV(n), ref.StateVector(n)
    for o in tape:
        getattr(sv2, o["name"])(o["wires"], o["inverse"], o["params"])
    r2.apply_ops(tape)
    o_struct = ops.algorithms.create_ops_listC128([o["name"] for o in tape], [o["params"] for o in tape],
                                                  [o["wires"] for o in tape], [o["inverse"] for o in tape],
                                                  [np.array([], dtype=dt)] * len(tape), [[]] * len(tape),
                                                  [[]] * len(tape))
    jac = ops.algorithms.AdjointJacobianC128()(sv2, [ob], o_struct, tp)
    np.testing.assert_allclose(jac, r2.adjoint_jacobian([rob], tape, tp).ravel(), rtol=0, atol=1e-12)


def test_errors_surface_as_runtime_error():
    sv = ops.StateVectorC128(2)
    with pytest.raises(RuntimeError, match="Error in PennyLane Lightning"):
        sv.RX([0], [True], [0], False, [0.1])
    with pytest.raises(RuntimeError, match="size of matrix"):
        sv.applyMatrix(np.eye(2, dtype=np.complex128), [0, 1], False)


@pytest.mark.parametrize("p", PREC)
def test_shot_based_api_matches_reference(ref, p):
    """MeasurementsBase shot API (MeasurementsBase.hpp:159-521) on top of bit-exact samples: identical
    estimates to lightning.qubit under a shared seed."""
    SV, M, dt = classes(p)
    N = getattr(ops.observables, f"NamedObsC{p}")
    T = getattr(ops.observables, f"TensorProdObsC{p}")
    Ham = getattr(ops.observables, f"HamiltonianC{p}")
    n = 6
    tape = circuits.random_circuit(n, 3, 77)
    sv, r = SV(n), ref.StateVector(n, dt)
    for o in tape:
        getattr(sv, o["name"])(o["wires"], o["inverse"], o["params"])
    r.apply_ops(tape)
    m = M(sv)
    m.set_random_seed(123)
    RN, RT, RH = ref.Observable.named, ref.Observable.tensor, ref.Observable.hamiltonian
    cases = [
        (N("PauliZ", [0]), RN("PauliZ", [0], dtype=dt)),
        (N("PauliX", [2]), RN("PauliX", [2], dtype=dt)),
        (N("PauliY", [4]), RN("PauliY", [4], dtype=dt)),
        (N("Hadamard", [1]), RN("Hadamard", [1], dtype=dt)),
        (T([N("PauliX", [0]), N("PauliY", [3]), N("PauliZ", [5])]),
         RT([RN("PauliX", [0], dtype=dt), RN("PauliY", [3], dtype=dt), RN("PauliZ", [5], dtype=dt)])),
        (Ham(np.array([0.4, -1.3]), [N("PauliZ", [1]), T([N("PauliX", [2]), N("PauliX", [4])])]),
         RH([0.4, -1.3], [RN("PauliZ", [1], dtype=dt), RT([RN("PauliX", [2], dtype=dt), RN("PauliX", [4], dtype=dt)])])),
    ]
    tol = 1e-12 if p == "128" else 1e-5
    for a, b in cases:
        assert abs(m.expval_shots(a, 500, []) - r.expval_shots(b, 500, 123)) < tol
        assert abs(m.expval_shots(a, 500, [1, 5, 7, 400]) - r.expval_shots(b, 500, 123, [1, 5, 7, 400])) < tol
        assert abs(m.var_shots(a, 500) - r.var_shots(b, 500, 123)) < 10 * tol
    np.testing.assert_allclose(m.probs_shots([3, 0], 400), r.probs_shots([3, 0], 400, 123), atol=tol)
    assert sum(m.counts(300).values()) == 300
    with pytest.raises(RuntimeError, match="do not support samples"):
        m.sample_obs(cases[-1][0], 10)


@pytest.mark.parametrize("p", PREC)
def test_lazy_gate_queue(p, ref):
    """Per-gate calls are validated at call time, queued, and applied as one fused tape on the first
    read: same state as the reference, errors raised where the reference raises them,
    full-state preparations drop the queue."""
    SV, M, dt = classes(p)
    n = 14
    tape = circuits.random_circuit(n, 4, 5)
    sv = SV(n)
    for o in tape:
        getattr(sv, o["name"])(o["wires"], o["inverse"], o["params"])
    assert sv.pendingOps() == len(tape)  # nothing has touched the device yet
    with pytest.raises(Exception):
        sv.RX([n + 3], False, [0.1])  # invalid wire: raised by the call, not at flush time
    with pytest.raises(Exception):
        sv.RX([0], False, [])  # missing parameter
    assert sv.pendingOps() == len(tape)
    got = state(sv, dt)  # first read: one fused tape
    assert sv.pendingOps() == 0
    r = ref.StateVector(n, dt)
    r.apply_ops(tape)
    np.testing.assert_allclose(got, r.get_state(), rtol=0, atol=1e-12 if p == "128" else 1e-5)
    # measurements flush as well
    sv.Hadamard([0], False, [])
    r.apply("Hadamard", [0])
    m = M(sv)
    assert sv.pendingOps() == 1
    pr = np.asarray(m.probs([0, 1]))
    assert sv.pendingOps() == 0
    np.testing.assert_allclose(pr, r.probs([0, 1]), atol=1e-12 if p == "128" else 1e-5)
    # a full-state preparation discards pending gates; small matrices and Pauli rotations JOIN the queue in program
    # order (validated at call time), larger matrices flush it and run at once
    sv.PauliX([1], False, [])
    sv.resetStateVector()
    assert sv.pendingOps() == 0 and np.isclose(state(sv, dt)[0], 1.0)
    sv.PauliX([n - 1], False, [])  # |0..01>
    sv.applyMatrix(np.array([[0, 1], [1, 0]], dtype=dt), [n - 2], False)  # |0..11>
    sv.PauliX([n - 1], False, [])  # |0..10>
    sv.applyPauliRot([0, n - 1], False, [np.pi], "XI")  # -i X on wire 0
    assert sv.pendingOps() == 4
    with pytest.raises(RuntimeError):  # a bad call still fails when it is made, not at the flush
        sv.applyMatrix(np.eye(2, dtype=dt), [n + 3], False)
    assert sv.pendingOps() == 4
    out = state(sv, dt)
    assert sv.pendingOps() == 0
    assert np.isclose(abs(out[2 + (1 << (n - 1))]), 1.0) and np.isclose(np.abs(out).sum(), 1.0)
    big = np.eye(32, dtype=dt)[::-1].copy()  # 5 wires: flushes, then runs at once (tensor-core path)
    sv.PauliX([0], False, [])
    sv.applyMatrix(big, [1, 2, 3, 4, 5], False)
    assert sv.pendingOps() == 0
