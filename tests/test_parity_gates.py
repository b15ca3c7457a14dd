"""Parity of every gate / matrix / generator kernel against the unmodified reference
lightning.qubit core (pattern (b) of the reference's own kernel tests: cross-kernel equivalence
on random states, Test_GateImplementations_CompareKernels.cpp:168-245).  Tolerances are the
north_star's: 1e-12 relative for c128, 1e-5 for c64 (states are normalised, so absolute == relative
to the state norm)."""
import itertools

import zlib

import numpy as np
import pytest

from conftest import CONTROLLED_GATES, CONTROLLED_GENERATORS, GATES, GENERATORS, TOL, random_state

pytestmark = pytest.mark.gpu
DTYPES = [np.complex128, np.complex64]


def _params(name, npar, rng, k):
    if name == "PCPhase":
        return [float(rng.uniform(0, 2 * np.pi)), float(rng.integers(0, 2**k + 1))]
    return [float(x) for x in rng.uniform(0, 2 * np.pi, size=npar)]


def _pair(plb, ref, n, dtype, seed):
    st = random_state(n, dtype, seed)
    a = plb.StateVector(n, dtype)
    a.set_state(st)
    b = ref.StateVector(n, dtype)
    b.set_state(st)
    return a, b


def _close(a, b, dtype):
    tol = TOL[np.dtype(dtype)]
    np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=tol)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", list(GATES))
@pytest.mark.parametrize("inverse", [False, True])
def test_gate_all_wire_choices(plb, ref, name, inverse, dtype):
    n = 6
    nw, npar = GATES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    wire_sets = []
    if nw == -1:
        wire_sets = [[2], [0, 5], [4, 1, 3], [5, 4, 3, 2, 1, 0]]
    elif nw <= 2:
        wire_sets = [list(p) for p in itertools.permutations(range(n), nw)]
    else:
        allp = [list(p) for p in itertools.permutations(range(n), nw)]
        idx = rng.choice(len(allp), size=24, replace=False)
        wire_sets = [allp[i] for i in idx]
    for wires in wire_sets:
        params = _params(name, npar, rng, len(wires))
        a, b = _pair(plb, ref, n, dtype, 11)
        a.apply(name, wires, inverse, params)
        b.apply(name, wires, inverse, params)
        _close(a, b, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", CONTROLLED_GATES)
def test_controlled_gate(plb, ref, name, dtype):
    n = 7
    nw, npar = GATES[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    for trial in range(10):
        k = nw if nw > 0 else int(rng.integers(1, 4))
        nc = int(rng.integers(1, min(3, n - k) + 1))
        perm = rng.permutation(n)
        wires = [int(x) for x in perm[:k]]
        cw = [int(x) for x in perm[k:k + nc]]
        cv = [bool(x) for x in rng.integers(0, 2, size=nc)]
        params = _params(name, npar, rng, k)
        inverse = bool(trial % 2)
        a, b = _pair(plb, ref, n, dtype, 100 + trial)
        a.apply(name, wires, inverse, params, cw, cv)
        b.apply(name, wires, inverse, params, cw, cv)
        _close(a, b, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_apply_matrix(plb, ref, k, dtype):
    n = 8
    rng = np.random.default_rng(k)
    for trial in range(6):
        m = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
        q, _ = np.linalg.qr(m)
        perm = rng.permutation(n)
        wires = [int(x) for x in perm[:k]]
        nc = trial % 3 if k + 2 <= n else 0
        cw = [int(x) for x in perm[k:k + nc]]
        cv = [bool(x) for x in rng.integers(0, 2, size=nc)]
        inverse = bool(trial % 2)
        a, b = _pair(plb, ref, n, dtype, 7 + trial)
        a.apply_matrix(q, wires, inverse, cw, cv)
        b.apply_matrix(q, wires, inverse, cw, cv)
        _close(a, b, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", list(GENERATORS))
def test_generator(plb, ref, name, dtype):
    n = 6
    nw = GENERATORS[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 2)
    for trial in range(8):
        k = nw if nw > 0 else int(rng.integers(1, 4))
        wires = [int(x) for x in rng.permutation(n)[:k]]
        a, b = _pair(plb, ref, n, dtype, 31 + trial)
        sa = a.apply_generator(name, wires, bool(trial % 2))
        sb = b.apply_generator(name, wires, bool(trial % 2))
        assert sa == sb
        _close(a, b, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", CONTROLLED_GENERATORS)
def test_controlled_generator(plb, ref, name, dtype):
    n = 7
    nw = GENERATORS[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 3)
    for trial in range(8):
        k = nw if nw > 0 else int(rng.integers(1, 4))
        nc = int(rng.integers(1, min(3, n - k) + 1))
        perm = rng.permutation(n)
        wires = [int(x) for x in perm[:k]]
        cw = [int(x) for x in perm[k:k + nc]]
        cv = [bool(x) for x in rng.integers(0, 2, size=nc)]
        a, b = _pair(plb, ref, n, dtype, 51 + trial)
        sa = a.apply_generator(name, wires, False, cw, cv)
        sb = b.apply_generator(name, wires, False, cw, cv)
        assert sa == sb
        _close(a, b, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_pauli_rot(plb, ref, dtype):
    n = 7
    rng = np.random.default_rng(5)
    # words without "I": the Python layer strips identities before calling applyPauliRot
    # (lightning_qubit/_state_vector.py:262-269); the C++ kernel treats any non-Z letter as X/Y.
    words = ["X", "Y", "Z", "XX", "XY", "YZ", "ZZ", "ZXZ", "XYZ", "YYYY", "XZZY", "ZZZZZ", "XYZXYZX", "ZYZ", "YXY"]
    for i, word in enumerate(words):
        wires = [int(x) for x in rng.permutation(n)[: len(word)]]
        theta = float(rng.uniform(0, 2 * np.pi))
        a, b = _pair(plb, ref, n, dtype, 71 + i)
        a.apply_pauli_rot(wires, bool(i % 2), theta, word)
        b.apply_pauli_rot(wires, bool(i % 2), theta, word)
        _close(a, b, dtype)


def test_pauli_rot_identity_letters(plb):
    """'I' letters act as identity: same result as the stripped word on the remaining wires."""
    st = random_state(6, np.complex128, 3)
    for word, wires in (("XIZ", [4, 1, 0]), ("IYI", [2, 5, 3]), ("III", [0, 1, 2]), ("ZIIZ", [5, 0, 1, 2])):
        a = plb.StateVector(6)
        a.set_state(st)
        a.apply_pauli_rot(wires, False, 0.81, word)
        b = plb.StateVector(6)
        b.set_state(st)
        kept = [(c, w) for c, w in zip(word, wires) if c != "I"]
        if kept:
            b.apply_pauli_rot([w for _, w in kept], False, 0.81, "".join(c for c, _ in kept))
        else:
            b.apply("GlobalPhase", [0], False, [0.81 / 2])
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-14)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 3, 12, 17])
def test_sizes_and_low_bits(plb, ref, n, dtype):
    """Edge sizes (1-3 qubits) and larger states where every index-bit class (0-3, 4-11, >=12) is hit."""
    rng = np.random.default_rng(n)
    a, b = _pair(plb, ref, n, dtype, n)
    for w in range(n):
        th = float(rng.uniform(0, 2 * np.pi))
        for sv in (a, b):
            sv.apply("RX", [w], False, [th])
            sv.apply("RZ", [w], False, [th / 3])
            sv.apply("Hadamard", [w])
            if n > 1:
                sv.apply("CNOT", [w, (w + 1) % n])
                sv.apply("CRZ", [(w + 1) % n, w], False, [th / 2])
                sv.apply("SWAP", [w, (w + n // 2 + 1) % n] if (w + n // 2 + 1) % n != w else [w, (w + 1) % n])
    _close(a, b, dtype)


def test_error_messages(plb):
    sv = plb.StateVector(3)
    with pytest.raises(plb.B200Error, match="must be disjoint"):
        sv.apply("RX", [0], False, [0.1], [0], [True])
    with pytest.raises(plb.B200Error, match="same size"):
        sv.apply("RX", [0], False, [0.1], [1, 2], [True])
    with pytest.raises(plb.B200Error, match="does not exist"):
        sv.apply("NotAGate", [0])
    with pytest.raises(plb.B200Error, match="size of matrix"):
        sv.apply_matrix(np.eye(2), [0, 1])
