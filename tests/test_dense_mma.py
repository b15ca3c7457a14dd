"""Tensor-core path for dense matrices on 5..7 wires (csrc/dense_mma.cu: DMMA for c128, 3xTF32 for c64) against
the reference lightning.qubit applyMatrix / applyControlledMatrix, and against the engine's own mat-vec kernel."""
import os

import numpy as np
import pytest

from conftest import TOL, random_state

pytestmark = pytest.mark.gpu


def _unitary(k, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


@pytest.mark.parametrize("dtype,k", [(np.complex128, 5), (np.complex128, 6), (np.complex64, 5), (np.complex64, 6),
                                     (np.complex64, 7)])
@pytest.mark.parametrize("ctrl", [0, 2])
def test_dense_matrix_on_tensor_cores(plb, ref, dtype, k, ctrl):
    n = 14
    rng = np.random.default_rng(100 * k + ctrl)
    # wires whose index bits are all >= 3 (bit = n-1-wire): the tensor-core path applies
    wires = [int(w) for w in rng.permutation(n - 3)[: k + ctrl]]
    tw, cw = wires[:k], wires[k:]
    cv = [bool(rng.integers(2)) for _ in cw]
    u = _unitary(k, k)
    st = random_state(n, dtype, 7)
    a, r = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    a.set_state(st), r.set_state(st)
    l0 = a.kernel_launches
    a.apply_matrix(u, tw, False, cw, cv)
    r.apply_matrix(u, tw, False, cw, cv)
    assert a.kernel_launches == l0 + 1
    tol = TOL[np.dtype(dtype)]
    np.testing.assert_allclose(a.get_state(), r.get_state(), rtol=0, atol=tol)
    # the inverse brings the state back (adjoint matrix through the same path)
    a.apply_matrix(u, tw, True, cw, cv)
    np.testing.assert_allclose(a.get_state(), st, rtol=0, atol=10 * tol)
    # and the scalar mat-vec kernel agrees
    os.environ["PLB200_DENSE_MMA"] = "0"
    try:
        b = plb.StateVector(n, dtype)
        b.set_state(st)
        b.apply_matrix(u, tw, False, cw, cv)
    finally:
        del os.environ["PLB200_DENSE_MMA"]
    np.testing.assert_allclose(b.get_state(), r.get_state(), rtol=0, atol=tol)


def test_tensor_core_instructions_are_in_the_library(plb):
    import subprocess

    sass = subprocess.run(["cuobjdump", "-sass", plb.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in sass and "HMMA.1688.F32.TF32" in sass
