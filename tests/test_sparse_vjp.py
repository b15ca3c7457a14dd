"""SparseHamiltonian (CSR expval / var / applyInPlace), Hermitian observables measured with
shots (own eigen-solver instead of the reference's run-time LAPACK), and VectorJacobianProduct.
Reference parity: SparseHamiltonian and VJP against the unmodified lightning.qubit core (oracle/_ref); the
eigen-solver against numpy (CPU test)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import TOL, random_state
from pennylane_lightning_b200 import circuits


def _random_sparse_hermitian(n, density, seed):
    rng = np.random.default_rng(seed)
    dim = 1 << n
    a = sp.random(dim, dim, density=density, random_state=rng, dtype=np.float64).astype(np.complex128)
    a = a + 1j * sp.random(dim, dim, density=density, random_state=rng, dtype=np.float64)
    h = (a + a.conj().T).tocsr()
    h.sort_indices()
    return h


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("dim", [2, 4, 8, 32])
def test_hermitian_eigh_matches_numpy(plb, dim):
    rng = np.random.default_rng(dim)
    a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    h = a + a.conj().T
    ev, u = plb.hermitian_eigh(h)
    np.testing.assert_allclose(ev, np.linalg.eigvalsh(h), rtol=0, atol=1e-11)
    np.testing.assert_allclose(u @ u.conj().T, np.eye(dim), rtol=0, atol=1e-12)  # unitary
    np.testing.assert_allclose(u @ h @ u.conj().T, np.diag(ev), rtol=0, atol=1e-10)  # rotates into the eigenbasis


def test_hermitian_eigh_degenerate_and_rejects_non_hermitian(plb):
    z = np.diag([1.0, -1.0]).astype(complex)
    zz = np.kron(z, z)
    ev, u = plb.hermitian_eigh(zz)
    np.testing.assert_allclose(ev, [-1, -1, 1, 1], atol=1e-14)
    np.testing.assert_allclose(u @ zz @ u.conj().T, np.diag(ev), atol=1e-13)
    with pytest.raises(plb.B200Error, match="not a Hermitian matrix"):
        plb.hermitian_eigh(np.array([[1, 2], [3, 4]], dtype=complex))


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,density", [(6, 0.2), (10, 0.01), (12, 0.002)])
def test_sparse_expval_var_apply(plb, ref, dtype, n, density):
    h = _random_sparse_hermitian(n, density, n)
    st = random_state(n, dtype, 3)
    a, r = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    a.set_state(st), r.set_state(st)
    tol = 1e-11 if dtype == np.complex128 else 2e-4
    psi = st.astype(np.complex128)
    e_np = float(np.real(np.vdot(psi, h @ psi)))
    v_np = float(np.real(np.vdot(h @ psi, h @ psi))) - e_np ** 2
    assert abs(a.expval_sparse(h.indptr, h.indices, h.data) - e_np) < tol * max(1, abs(e_np))
    assert abs(a.var_sparse(h.indptr, h.indices, h.data) - v_np) < tol * max(1, abs(v_np)) * 10
    oa = plb.Observable.sparse(h.indptr, h.indices, h.data, list(range(n)))
    orf = ref.Observable.sparse(h.indptr, h.indices, h.data, list(range(n)), dtype=dtype)
    assert abs(a.expval(oa) - r.expval(orf)) < tol * max(1, abs(e_np))
    assert abs(a.var(oa) - r.var(orf)) < tol * max(1, abs(v_np)) * 10
    a.apply_observable(oa)
    np.testing.assert_allclose(a.get_state(), (h @ psi), rtol=0, atol=tol * 10)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_sparse_hamiltonian_in_adjoint(plb, ref, dtype):
    n = 8
    ops, tp = circuits.strongly_entangling_layers(n, 2, 5)
    h = _random_sparse_hermitian(n, 0.05, 9)
    a, r = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    oa = plb.Observable.sparse(h.indptr, h.indices, h.data, list(range(n)))
    orf = ref.Observable.sparse(h.indptr, h.indices, h.data, list(range(n)), dtype=dtype)
    ja = a.adjoint_jacobian([oa], ops, tp, apply_ops=True)
    jr = r.adjoint_jacobian([orf], ops, tp, apply_ops=True)
    np.testing.assert_allclose(ja, jr, rtol=0, atol=1e-11 if dtype == np.complex128 else 5e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_vjp_matches_reference(plb, ref, dtype):
    n = 9
    rng = np.random.default_rng(4)
    ops, _ = circuits.strongly_entangling_layers(n, 2, 11)
    ops += [circuits.op("CRX", [0, 3], [0.4]), circuits.op("IsingXX", [1, 2], [0.7], inverse=True),
            circuits.op("RY", [4], [0.2], ctrl_wires=[5, 6], ctrl_values=[True, False]),
            circuits.op("SingleExcitation", [7, 8], [1.1]), circuits.op("Hadamard", [2])]
    n_par = sum(1 for o in ops if o["params"])
    tp = sorted(int(x) for x in rng.choice(n_par, size=n_par - 5, replace=False))
    dy = random_state(n, np.complex128, 8)
    a, r = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    va = a.vjp(ops, dy, tp, apply_ops=True)
    vr = r.vjp(ops, dy, tp, apply_ops=True)
    np.testing.assert_allclose(va, vr, rtol=0, atol=1e-12 if dtype == np.complex128 else 2e-5)
    assert np.max(np.abs(vr)) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("p", ["64", "128"])
def test_bindings_sparse_vjp_hermitian_shots(p):
    """The Python-visible surface (LGPUBindings.hpp:65-203, LQubitBindings.hpp:413-462): SparseHamiltonianC*,
    the CSR expval/var overloads, VectorJacobianProductC*, and a Hermitian observable measured with shots."""
    ops_mod = pytest.importorskip("pennylane_lightning_b200.lightning_b200_ops")
    dt = np.complex64 if p == "64" else np.complex128
    SV, M = getattr(ops_mod, f"StateVectorC{p}"), getattr(ops_mod, f"MeasurementsC{p}")
    n = 5
    sv = SV(n)
    for w in range(n):
        sv.RY([w], False, [0.3 + 0.2 * w])
        sv.RX([w], False, [0.9 - 0.1 * w])
    sv.CNOT([0, 1], False, [])
    psi = np.zeros(1 << n, dtype=dt)
    sv.getState(psi)
    psi = psi.astype(np.complex128)
    m = M(sv)
    h = _random_sparse_hermitian(n, 0.2, 3)
    tol = 1e-10 if p == "128" else 1e-4
    e_np = float(np.real(np.vdot(psi, h @ psi)))
    assert abs(m.expval(h.indptr, h.indices, h.data.astype(dt)) - e_np) < tol
    v_np = float(np.real(np.vdot(h @ psi, h @ psi))) - e_np ** 2
    assert abs(m.var(h.indptr, h.indices, h.data.astype(dt)) - v_np) < 10 * tol
    SpH = getattr(ops_mod.observables, f"SparseHamiltonianC{p}")
    o = SpH(h.data.astype(dt), h.indices, h.indptr, list(range(n)))
    assert abs(m.expval(o) - e_np) < tol and o.get_wires() == list(range(n))
    # Hermitian observable with shots: eigenbasis rotation by the engine's own solver
    rng = np.random.default_rng(1)
    a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    hm = (a + a.conj().T).astype(dt)
    HObs = getattr(ops_mod.observables, f"HermitianObsC{p}")
    ho = HObs(hm.ravel(), [1, 3])
    exact = m.expval(ho)
    m.set_random_seed(7)
    est = m.expval_shots(ho, 200000, [])
    assert abs(est - exact) < 0.03 * max(1.0, np.max(np.abs(np.linalg.eigvalsh(hm))))
    # VJP
    alg = ops_mod.algorithms
    names, params, wires = ["RX", "RY", "CNOT", "RZ"], [[0.3], [0.5], [], [0.7]], [[0], [1], [0, 1], [1]]
    opsl = getattr(alg, f"create_ops_listC{p}")(names, params, wires, [False] * 4, [np.zeros(0, dtype=dt)] * 4, [[]] * 4, [[]] * 4)
    sv2 = SV(2)
    sv2.RX([0], False, [0.3]); sv2.RY([1], False, [0.5]); sv2.CNOT([0, 1], False, []); sv2.RZ([1], False, [0.7])
    dy = (rng.normal(size=4) + 1j * rng.normal(size=4)).astype(dt)
    vjp = getattr(alg, f"VectorJacobianProductC{p}")()(sv2, opsl, dy, [0, 1, 2])
    # finite differences of psi(theta) with the numpy oracle
    from oracle import np_oracle

    def psi_of(th):
        o = np_oracle.StateVector(2, np.complex128)
        o.apply_ops([circuits.op("RX", [0], [th[0]]), circuits.op("RY", [1], [th[1]]), circuits.op("CNOT", [0, 1]),
                     circuits.op("RZ", [1], [th[2]])])
        return o.get_state()

    th0 = np.array([0.3, 0.5, 0.7])
    for k in range(3):
        d = np.zeros(3)
        d[k] = 1e-6
        fd = np.vdot(dy.astype(np.complex128), (psi_of(th0 + d) - psi_of(th0 - d)) / 2e-6)
        assert abs(vjp[k] - fd) < (1e-7 if p == "128" else 1e-4)
