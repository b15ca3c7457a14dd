"""Measurements, state preparation, sampling and adjoint Jacobian of the GPU engine against the
unmodified reference lightning.qubit core on the same seeded inputs (1e-12 c128 / 1e-5 c64;
samples bit-exact under a shared seed)."""
import itertools

import numpy as np
import pytest

from conftest import GENERATORS, TOL, random_state
from pennylane_lightning_b200 import circuits

pytestmark = pytest.mark.gpu
DTYPES = [np.complex128, np.complex64]


def _pair(plb, ref, n, dtype, seed):
    st = random_state(n, dtype, seed)
    a = plb.StateVector(n, dtype)
    a.set_state(st)
    b = ref.StateVector(n, dtype)
    b.set_state(st)
    return a, b


def _herm(rng, k):
    m = rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k))
    return (m + m.conj().T) / 2


@pytest.mark.parametrize("dtype", DTYPES)
def test_probs(plb, ref, dtype):
    n = 6
    a, b = _pair(plb, ref, n, dtype, 1)
    tol = TOL[np.dtype(dtype)]
    np.testing.assert_allclose(a.probs(), b.probs(), rtol=0, atol=tol)
    rng = np.random.default_rng(0)
    for k in range(1, n + 1):
        for _ in range(6):
            wires = [int(x) for x in rng.permutation(n)[:k]]
            np.testing.assert_allclose(a.probs(wires), b.probs(wires), rtol=0, atol=tol, err_msg=str(wires))


@pytest.mark.parametrize("dtype", DTYPES)
def test_probs_many_wires(plb, ref, dtype):
    n = 15
    a, b = _pair(plb, ref, n, dtype, 2)
    tol = TOL[np.dtype(dtype)]
    rng = np.random.default_rng(1)
    for k in (3, 11, 12, 13, 15):
        wires = [int(x) for x in rng.permutation(n)[:k]]
        np.testing.assert_allclose(a.probs(wires), b.probs(wires), rtol=0, atol=tol)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [2, 4, 5, 9, 20])
def test_probs_marginals_deterministic(plb, dtype, n):
    """Marginals over <= 11 wires (probs_marginal_kernel: one warp per (outcome, chunk), fixed-order sums, no
    atomics): every kind of target placement — on the lane bits (last wires), above them, mixed, all wires of a
    tiny state — against numpy on the gathered state, and bit-identical when repeated."""
    a = plb.StateVector(n, dtype)
    a.apply_ops(circuits.random_circuit(n, 3, 8) if n >= 2 else [], fuse=False)
    psi = a.get_state().astype(np.complex128)
    p_full = (np.abs(psi) ** 2).reshape((2,) * n)
    rng = np.random.default_rng(n)
    cases = [[n - 1], [0], list(range(n))[-min(n, 5):], list(range(min(n, 3))), [n - 1, 0]] if n > 1 else [[0]]
    for _ in range(6):
        k = int(rng.integers(1, min(n, 11) + 1))
        cases.append([int(x) for x in rng.permutation(n)[:k]])
    tol = 1e-13 if dtype == np.complex128 else 1e-6
    for wires in cases:
        if len(set(wires)) != len(wires) or len(wires) > 11:
            continue
        got = np.asarray(a.probs(wires))
        rest = tuple(w for w in range(n) if w not in wires)
        want = p_full.sum(axis=rest) if rest else p_full
        # axes of `want` are the kept wires in ascending order: bring them into the caller's order
        kept = sorted(wires)
        want = np.transpose(want, [kept.index(w) for w in wires]).reshape(-1)
        np.testing.assert_allclose(got, want, rtol=0, atol=tol, err_msg=str(wires))
        np.testing.assert_array_equal(np.asarray(a.probs(wires)), got)


@pytest.mark.parametrize("dtype", DTYPES)
def test_expval_var_named_and_matrix(plb, ref, dtype):
    n = 7
    a, b = _pair(plb, ref, n, dtype, 3)
    tol = 10 * TOL[np.dtype(dtype)]
    for name in ("Identity", "PauliX", "PauliY", "PauliZ", "Hadamard"):
        for w in range(n):
            assert abs(a.expval_named(name, [w]) - b.expval_named(name, [w])) < tol
            assert abs(a.var_named(name, [w]) - b.var_named(name, [w])) < tol
    rng = np.random.default_rng(4)
    for k in (1, 2, 3, 4, 5):
        for _ in range(4):
            wires = [int(x) for x in rng.permutation(n)[:k]]
            m = _herm(rng, k)
            assert abs(a.expval_matrix(m, wires) - b.expval_matrix(m, wires)) < 20 * tol
            assert abs(a.var_matrix(m, wires) - b.var_matrix(m, wires)) < 200 * tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_observable_classes(plb, ref, dtype):
    n = 6
    a, b = _pair(plb, ref, n, dtype, 5)
    tol = 20 * TOL[np.dtype(dtype)]
    rng = np.random.default_rng(6)
    h1, h2 = _herm(rng, 1), _herm(rng, 2)

    def build(mod, **kw):
        N, Hm, T, Ham = mod.Observable.named, mod.Observable.hermitian, mod.Observable.tensor, mod.Observable.hamiltonian
        obs = [
            N("PauliX", [2], **kw), N("Hadamard", [5], **kw), Hm(h1, [3], **kw), Hm(h2, [4, 0], **kw),
            T([N("PauliX", [0], **kw), N("PauliY", [3], **kw), N("PauliZ", [5], **kw)]),
            T([N("Hadamard", [1], **kw), N("PauliZ", [2], **kw)]),
            T([Hm(h2, [1, 2], **kw), N("PauliY", [4], **kw)]),
            Ham([0.3, -1.2, 0.7], [N("PauliZ", [0], **kw), T([N("PauliX", [1], **kw), N("PauliX", [2], **kw)]),
                                   N("Hadamard", [4], **kw)]),
            Ham([0.5, 2.0], [Hm(h1, [5], **kw), T([N("PauliY", [0], **kw), Hm(h1, [1], **kw)])]),
        ]
        return obs

    oa, ob = build(plb), build(ref, dtype=dtype)
    for x, y in zip(oa, ob):
        assert abs(a.expval(x) - b.expval(y)) < tol
        assert abs(a.var(x) - b.var(y)) < 10 * tol
        a2, b2 = _pair(plb, ref, n, dtype, 5)
        a2.apply_observable(x)
        b2.apply_observable(y)
        np.testing.assert_allclose(a2.get_state(), b2.get_state(), rtol=0, atol=tol)


@pytest.mark.parametrize("dtype", DTYPES)
def test_pauli_words_fused(plb, ref, dtype):
    n = 8
    a, b = _pair(plb, ref, n, dtype, 7)
    co, words, wires = circuits.pauli_hamiltonian(n, 40, 3)
    each = a.expval_pauli_words_each(words, wires)
    ham = circuits.hamiltonian_observable(ref, co, words, wires, dtype=dtype)
    tol = 50 * TOL[np.dtype(dtype)]
    assert abs(a.expval_pauli_words(words, wires, co) - b.expval(ham)) < tol
    for k in range(0, 40, 7):
        term = circuits.hamiltonian_observable(ref, [1.0], [words[k]], [wires[k]], dtype=dtype)
        assert abs(each[k] - b.expval(term)) < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_diagonal_words_single_pass(plb, ref, dtype):
    """Z-only words take the one-read kernel (8 words per sweep): <Z_w> for every wire, ZZ / ZZZ words,
    19 words = two full batches + a ragged one."""
    n = 11
    a, b = _pair(plb, ref, n, dtype, 17)
    words = ["Z"] * n + ["ZZ"] * 5 + ["ZZZ"] * 3
    wires = [[w] for w in range(n)] + [[w, (w + 3) % n] for w in range(5)] + [[w, w + 2, w + 5] for w in range(3)]
    each = a.expval_pauli_words_each(words, wires)
    tol = 10 * TOL[np.dtype(dtype)]
    for k, (word, ws) in enumerate(zip(words, wires)):
        term = circuits.hamiltonian_observable(ref, [1.0], [word], [ws], dtype=dtype)
        assert abs(each[k] - b.expval(term)) < tol, (k, word, ws)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [9, 13, 21, 24])
def test_all_z_expvals_one_sweep(plb, ref, dtype, n):
    """<Z_w> on every wire (+ the norm through an identity word) takes the one-sweep kernel: the low index bits
    are constant per thread, only the high ones get accumulators.  Sizes below / at / above the 2^19-thread grid."""
    a = plb.StateVector(n, dtype)
    ops = circuits.random_circuit(n, 2, 31)
    a.apply_ops(ops)
    # the reference evaluates the SAME amplitudes in double precision (lightning.qubit's c64 reductions accumulate
    # in float: at 2^24 amplitudes its own <Z> is only good to ~1e-4, which would be the error measured here)
    b = ref.StateVector(n, np.complex128)
    b.set_state(a.get_state().astype(np.complex128))
    order = [int(w) for w in np.random.default_rng(n).permutation(n)]
    words, wires = ["Z"] * n + ["I"], [[w] for w in order] + [[0]]
    l0 = a.kernel_launches
    each = a.expval_pauli_words_each(words, wires)
    assert a.kernel_launches - l0 <= 2  # one sweep + the final reduction
    tol = 1e-12 if dtype == np.complex128 else 1e-6
    for k, w in enumerate(order):
        assert abs(each[k] - b.expval_named("PauliZ", [w])) < tol, (k, w)
    assert abs(each[n] - 1.0) < (1e-12 if dtype == np.complex128 else 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_state_preparation(plb, ref, dtype):
    n = 6
    tol = TOL[np.dtype(dtype)]
    rng = np.random.default_rng(8)
    a, b = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    for sv in (a, b):
        sv.set_basis_state([1, 0, 1], [4, 0, 2])
    np.testing.assert_array_equal(a.get_state(), b.get_state())
    for wires in ([0, 1, 2], [5, 2, 3, 0], [1], list(range(n))):
        vals = rng.normal(size=2 ** len(wires)) + 1j * rng.normal(size=2 ** len(wires))
        # LGPU semantics (StateVectorCudaManaged.hpp:2424-2427): the engine zero-initialises the
        # other amplitudes; LQ only overwrites the addressed subspace, so start it from |0>.
        b.reset()
        a.set_state_vector(vals, wires), b.set_state_vector(vals, wires)
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=tol)
    a.reset(), b.reset()
    np.testing.assert_array_equal(a.get_state(), b.get_state())
    for wire, branch in itertools.product(range(n), (0, 1)):
        a, b = _pair(plb, ref, n, dtype, 9)
        a.collapse(wire, branch), b.collapse(wire, branch)
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=tol)
    a.set_state_indices([3, 17, 40], [0.5, 0.5j, -0.5])
    exp = np.zeros(2**n, dtype=dtype)
    exp[[3, 17, 40]] = [0.5, 0.5j, -0.5]
    np.testing.assert_array_equal(a.get_state(), exp)
    z = plb.StateVector(3, dtype)
    z.set_state(np.zeros(8, dtype=dtype))
    with pytest.raises(plb.B200Error, match="norm close to zero"):
        z.normalize()


@pytest.mark.parametrize("dtype", DTYPES)
def test_samples_bit_exact(plb, ref, dtype):
    n = 9
    ops = circuits.random_circuit(n, 3, 5)
    a, b = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    a.apply_ops(ops, fuse=False), b.apply_ops(ops)
    for seed in (0, 1, 12345):
        np.testing.assert_array_equal(a.generate_samples(200, seed=seed), b.generate_samples(200, seed=seed))
    s = a.generate_samples(500, seed=7)
    assert s.shape == (500, n) and s.dtype == np.uint64 and set(np.unique(s)) <= {0, 1}
    # histogram vs exact probabilities (Test_MeasurementsBase.cpp:1303-1308 uses margin 0.05)
    idx = (s * (1 << np.arange(n - 1, -1, -1, dtype=np.uint64))).sum(axis=1)
    counts = np.bincount(idx.astype(np.int64), minlength=2**n) / 500
    assert np.max(np.abs(counts - a.probs())) < 0.05


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 10, 15])
def test_device_sampler_distribution(plb, dtype, n):
    """plb200_generate_samples_device (chunk masses -> scan -> one warp per shot): same distribution as the exact
    probabilities (total-variation distance and per-wire marginals within sampling error), reproducible under a
    seed, different under another, wire subsets in the caller's order, sparse states never yield a zero-probability
    outcome."""
    ops = circuits.random_circuit(n, 3, 5)
    a = plb.StateVector(n, dtype)
    a.apply_ops(ops, fuse=False)
    shots = 200000
    s = a.generate_samples(shots, seed=3, device=True)
    assert s.shape == (shots, n) and s.dtype == np.uint64 and set(np.unique(s)) <= {0, 1}
    np.testing.assert_array_equal(s, a.generate_samples(shots, seed=3, device=True))
    assert not np.array_equal(s, a.generate_samples(shots, seed=4, device=True))
    p = np.asarray(a.probs(), dtype=np.float64)
    idx = (s * (1 << np.arange(n - 1, -1, -1, dtype=np.uint64))).sum(axis=1).astype(np.int64)
    emp = np.bincount(idx, minlength=2**n) / shots
    # E[TV] ~ sqrt(K / (2 pi shots)) for K outcomes: 0.16 at K = 2^15; three times that is a wrong distribution
    assert 0.5 * np.abs(emp - p).sum() < 0.02 + 1.5 * np.sqrt(2.0**n / (2 * np.pi * shots))
    for w in range(n):
        exact = a.expval_pauli_words_each(["Z"], [[w]])[0]
        assert abs((1.0 - 2.0 * s[:, w].mean()) - exact) < 5.0 / np.sqrt(shots)
    # a wire subset in a scrambled order = the marginal, columns in that order
    sub = [n - 1, 0, n // 2] if n >= 3 else [0]
    t = a.generate_samples(shots, wires=sub, seed=9, device=True)
    assert t.shape == (shots, len(sub))
    pm = np.asarray(a.probs(sub))
    ti = (t * (1 << np.arange(len(sub) - 1, -1, -1, dtype=np.uint64))).sum(axis=1).astype(np.int64)
    assert np.max(np.abs(np.bincount(ti, minlength=2 ** len(sub)) / shots - pm)) < 5.0 / np.sqrt(shots)
    # a sparse state: only outcomes with non-zero probability may appear
    b = plb.StateVector(n, dtype)
    amp = np.zeros(2**n, dtype=dtype)
    support = [0, 2**n - 1, 5 % 2**n]
    amp[support] = [0.6, 0.0, 0.8] if len(set(support)) == 3 else 1.0
    amp[support[1]] = 0.0
    amp /= np.linalg.norm(amp)
    b.set_state(amp)
    u = b.generate_samples(5000, seed=1, device=True)
    ui = (u * (1 << np.arange(n - 1, -1, -1, dtype=np.uint64))).sum(axis=1).astype(np.int64)
    assert set(np.unique(ui)) <= {i for i in support if abs(amp[i]) > 0}


@pytest.mark.parametrize("dtype", DTYPES)
def test_linear_algebra(plb, dtype):
    n = 10
    x, y = random_state(n, dtype, 1), random_state(n, dtype, 2)
    a, b = plb.StateVector(n, dtype), plb.StateVector(n, dtype)
    a.set_state(x), b.set_state(y)
    tol = 10 * TOL[np.dtype(dtype)]
    assert abs(a.dot(b) - np.vdot(x.astype(complex), y.astype(complex))) < tol
    assert abs(a.norm2() - 1.0) < tol
    a.axpy(0.3 - 0.2j, b)
    np.testing.assert_allclose(a.get_state(), x + (0.3 - 0.2j) * y, rtol=0, atol=tol)
    a.scale(2j)
    np.testing.assert_allclose(a.get_state(), 2j * (x + (0.3 - 0.2j) * y), rtol=0, atol=tol)


def _tape(rng, n):
    ops = []
    for name, nw in GENERATORS.items():
        k = nw if nw > 0 else 3
        wires = [int(x) for x in rng.permutation(n)[:k]]
        ops.append(circuits.op(name, wires, [rng.uniform(0, 6)], inverse=bool(rng.integers(0, 2))))
        ops.append(circuits.op("CNOT", [int(x) for x in rng.permutation(n)[:2]]))
        ops.append(circuits.op("Hadamard", [int(rng.integers(0, n))]))
    # N-controlled parametric ops (ControlledGeneratorOperation)
    for name in ("RX", "RZ", "PhaseShift", "IsingXX", "SingleExcitation", "DoubleExcitationPlus", "MultiRZ",
                 "GlobalPhase"):
        k = GENERATORS[name] if GENERATORS[name] > 0 else 2
        perm = [int(x) for x in rng.permutation(n)]
        ops.append(circuits.op(name, perm[:k], [rng.uniform(0, 6)], ctrl_wires=perm[k:k + 2],
                               ctrl_values=[True, False]))
    return ops


@pytest.mark.parametrize("dtype", DTYPES)
def test_adjoint_jacobian_every_generator(plb, ref, dtype):
    n = 7
    rng = np.random.default_rng(10)
    ops = _tape(rng, n)
    n_par = sum(1 for o in ops if o["params"])
    tp = sorted(int(x) for x in rng.choice(n_par, size=n_par - 5, replace=False))
    h = _herm(rng, 2)

    def obs(mod, **kw):
        N, T = mod.Observable.named, mod.Observable.tensor
        return [N("PauliZ", [0], **kw), T([N("PauliX", [1], **kw), N("PauliY", [4], **kw)]),
                mod.Observable.hermitian(h, [2, 6], **kw),
                mod.Observable.hamiltonian([0.4, -0.9], [N("PauliZ", [3], **kw), T([N("PauliX", [5], **kw),
                                                                                  N("PauliZ", [6], **kw)])])]

    a, b = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    ja = a.adjoint_jacobian(obs(plb), ops, tp, apply_ops=True)
    jb = b.adjoint_jacobian(obs(ref, dtype=dtype), ops, tp, apply_ops=True)
    np.testing.assert_allclose(ja, jb, rtol=0, atol=100 * TOL[np.dtype(dtype)])
    # single observable path + pre-applied state
    a.apply_ops(ops, fuse=False), b.apply_ops(ops)
    ja = a.adjoint_jacobian(obs(plb)[:1], ops, tp)
    jb = b.adjoint_jacobian(obs(ref, dtype=dtype)[:1], ops, tp)
    np.testing.assert_allclose(ja, jb, rtol=0, atol=100 * TOL[np.dtype(dtype)])


def test_adjoint_config1_sel20(plb, ref):
    """BASELINE.json configs[0]: 20-qubit 4-layer StronglyEntanglingLayers, c128, <Z0> + Jacobian."""
    ops, tp = circuits.strongly_entangling_layers(20, 4, 42)
    a, b = plb.StateVector(20), ref.StateVector(20)
    a.apply_ops(ops), b.apply_ops(ops)
    np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=0, atol=1e-12)
    oa, ob = plb.Observable.named("PauliZ", [0]), ref.Observable.named("PauliZ", [0])
    assert abs(a.expval(oa) - b.expval(ob)) < 1e-12
    np.testing.assert_allclose(a.adjoint_jacobian([oa], ops, tp), b.adjoint_jacobian([ob], ops, tp), rtol=0,
                               atol=1e-12)


def test_adjoint_matrix_ops_and_errors(plb, ref):
    n = 4
    rng = np.random.default_rng(11)
    q, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    ops = [circuits.op("RX", [0], [0.3]), dict(name="QubitUnitary", wires=[1, 2], params=[], matrix=q),
           circuits.op("RY", [2], [0.7]), dict(name="QubitUnitary", wires=[3, 0], params=[], matrix=q, inverse=True),
           circuits.op("CRZ", [0, 3], [1.1])]
    oa, ob = plb.Observable.named("PauliZ", [3]), ref.Observable.named("PauliZ", [3])
    a, b = plb.StateVector(n), ref.StateVector(n)
    np.testing.assert_allclose(a.adjoint_jacobian([oa], ops, [0, 1, 2], True),
                               b.adjoint_jacobian([ob], ops, [0, 1, 2], True), rtol=0, atol=1e-12)
    bad = [circuits.op("Rot", [0], [0.1, 0.2, 0.3])]
    with pytest.raises(plb.B200Error, match="not supported using the adjoint"):
        a.adjoint_jacobian([oa], bad, [0], True)
