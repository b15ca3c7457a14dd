"""The fused (cache-blocked) tape executor must reproduce the un-fused kernels and the reference."""
import numpy as np
import pytest

from conftest import TOL, random_state
from pennylane_lightning_b200 import circuits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n", [14, 17, 20])
def test_fused_random_circuit_matches_reference(plb, ref, n, dtype):
    ops = circuits.random_circuit(n, 6, 1234 + n)
    st = random_state(n, dtype, n)
    a, u, b = plb.StateVector(n, dtype), plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    for sv in (a, u, b):
        sv.set_state(st)
    a.apply_ops(ops, fuse=True)
    gates, passes = a.last_apply_stats()
    u.apply_ops(ops, fuse=False)
    b.apply_ops(ops)
    tol = TOL[np.dtype(dtype)]
    np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=tol)
    np.testing.assert_allclose(u.get_state(), b.get_state(), rtol=0, atol=tol)
    assert gates == len(ops) and passes < len(ops) / 3, (gates, passes)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_fused_mixed_gate_set(plb, ref, dtype):
    """Fusable ops interleaved with ops that must run stand-alone (SWAP, IsingXX, Toffoli, matrices,
    PauliRot, multi-controlled gates): ordering/commutation of the scheduler."""
    n = 15
    rng = np.random.default_rng(3)
    names1 = ["Hadamard", "PauliX", "PauliY", "PauliZ", "S", "T", "SX", "RX", "RY", "RZ", "PhaseShift", "Rot"]
    names2 = ["CNOT", "CZ", "CY", "SWAP", "CRX", "CRY", "CRZ", "CRot", "ControlledPhaseShift", "IsingXX", "IsingZZ",
              "IsingXY", "SingleExcitationPlus"]
    npar = {"RX": 1, "RY": 1, "RZ": 1, "PhaseShift": 1, "Rot": 3, "CRX": 1, "CRY": 1, "CRZ": 1, "CRot": 3,
            "ControlledPhaseShift": 1, "IsingXX": 1, "IsingZZ": 1, "IsingXY": 1, "SingleExcitationPlus": 1}
    ops = []
    for _ in range(250):
        r = rng.random()
        if r < 0.5:
            nm = names1[int(rng.integers(len(names1)))]
            ops.append(circuits.op(nm, [int(rng.integers(n))], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.85:
            nm = names2[int(rng.integers(len(names2)))]
            ops.append(circuits.op(nm, [int(x) for x in rng.permutation(n)[:2]], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.9:
            ops.append(circuits.op("Toffoli", [int(x) for x in rng.permutation(n)[:3]]))
        elif r < 0.95:
            p = [int(x) for x in rng.permutation(n)[:4]]
            ops.append(circuits.op("RY", p[:1], [rng.uniform(0, 6)], ctrl_wires=p[1:4], ctrl_values=[True, False, True]))
        else:
            ops.append(circuits.op("MultiRZ", [int(x) for x in rng.permutation(n)[:3]], [rng.uniform(0, 6)]))
        if rng.random() < 0.05:
            ops.append(circuits.op("GlobalPhase", [0], [0.3]))
    st = random_state(n, dtype, 5)
    a, b = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    a.set_state(st), b.set_state(st)
    a.apply_ops(ops, fuse=True)
    b.apply_ops(ops)
    np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=5 * TOL[np.dtype(dtype)])


def test_fused_qft_and_sel(plb, ref):
    for ops, n in ((circuits.qft(16), 16), (circuits.strongly_entangling_layers(16, 3, 1)[0], 16)):
        a, b = plb.StateVector(n), ref.StateVector(n)
        a.set_basis_state([1, 0, 1] + [0] * (n - 3), list(range(n)))
        b.set_basis_state([1, 0, 1] + [0] * (n - 3), list(range(n)))
        a.apply_ops(ops, fuse=True)
        b.apply_ops(ops)
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-12)


def test_size_independent_properties_large(plb):
    """BASELINE.json configs[1] family at 26 qubits (1 GiB): norm preservation and U^dagger U = 1
    through the fused path, properties that need no oracle."""
    n = 26
    ops = circuits.random_circuit(n, 4, 1234)
    sv = plb.StateVector(n)
    sv.apply_ops(ops, fuse=True)
    assert abs(sv.norm2() - 1.0) < 1e-12
    inv = [dict(o, inverse=not o["inverse"]) for o in reversed(ops)]
    sv.apply_ops(inv, fuse=True)
    head = sv.get_state(16)
    assert abs(head[0] - 1.0) < 1e-12 and np.max(np.abs(head[1:])) < 1e-12
    assert abs(sv.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_fused_adjoint_matches_reference(plb, ref, dtype):
    """Single-observable adjoint runs as two-state tile passes: every fusable generator kind (X, Y,
    Z-parity, controlled, projector), interleaved with ops/generators that must run stand-alone."""
    n = 15
    rng = np.random.default_rng(21)
    ops = []
    for layer in range(3):
        for w in range(n):
            g = ("RX", "RY", "RZ", "PhaseShift")[int(rng.integers(4))]
            ops.append(circuits.op(g, [w], [rng.uniform(0, 6)], inverse=bool(rng.integers(2))))
        perm = [int(x) for x in rng.permutation(n)]
        for i in range(0, n - 1, 2):
            g = ("CNOT", "CRX", "CRY", "CRZ", "ControlledPhaseShift", "IsingZZ", "IsingXX", "CZ", "SWAP",
                 "SingleExcitation")[int(rng.integers(10))]
            npar = 0 if g in ("CNOT", "CZ", "SWAP") else 1
            ops.append(circuits.op(g, perm[i:i + 2], rng.uniform(0, 6, npar), inverse=bool(rng.integers(2))))
        p = [int(x) for x in rng.permutation(n)]
        ops.append(circuits.op("MultiRZ", p[:3], [rng.uniform(0, 6)]))
        ops.append(circuits.op("RY", p[3:4], [rng.uniform(0, 6)], ctrl_wires=p[4:6], ctrl_values=[True, False]))
        ops.append(circuits.op("GlobalPhase", [0], [rng.uniform(0, 6)], ctrl_wires=p[6:7], ctrl_values=[True]))
        ops.append(circuits.op("Hadamard", p[7:8]))
    n_par = sum(1 for o in ops if o["params"])
    tp = sorted(int(x) for x in rng.choice(n_par, size=n_par - 7, replace=False))
    co, words, wires = circuits.pauli_hamiltonian(n, 12, 5)
    a, b = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    ham_a = circuits.hamiltonian_observable(plb, co, words, wires)
    ham_b = circuits.hamiltonian_observable(ref, co, words, wires, dtype=dtype)
    ja = a.adjoint_jacobian([ham_a], ops, tp, apply_ops=True)
    jb = b.adjoint_jacobian([ham_b], ops, tp, apply_ops=True)
    tol = 1e-11 if dtype == np.complex128 else 2e-4
    np.testing.assert_allclose(ja, jb, rtol=0, atol=tol)
    # the overlaps are accumulated in a fixed order (per warp, per CTA, CTAs in index order): bit-identical reruns
    for _ in range(3):
        c = plb.StateVector(n, dtype)
        np.testing.assert_array_equal(c.adjoint_jacobian([ham_a], ops, tp, apply_ops=True), ja)
    # the un-fused sweep (env switch) gives the same numbers
    import os
    os.environ["PLB200_ADJOINT_UNFUSED"] = "1"
    try:
        ju = a.adjoint_jacobian([ham_a], ops, tp, apply_ops=True)
    finally:
        del os.environ["PLB200_ADJOINT_UNFUSED"]
    np.testing.assert_allclose(ja, ju, rtol=0, atol=tol)
