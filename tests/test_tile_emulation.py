"""CPU tests of the fused tile path's HOST-visible logic: scheduler + pass encoder + the per-thread tile
interpreter (csrc/tile_exec.cuh), executed thread-by-thread on host memory by the test-only library
libplb200_emu.so (csrc/emu_abi.cu) and compared with the numpy oracle.  The product library has no CPU
path; this only makes the encoder checkable without a GPU (the `-m gpu` tests check the kernel itself)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import TOL, random_state
from oracle import np_oracle
from pennylane_lightning_b200 import circuits

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "pennylane-lightning_b200", "csrc")
EMU = os.path.join(HERE, "_emu", "libplb200_emu.so")


@pytest.fixture(scope="module")
def emu():
    res = subprocess.run(["make", "-C", CSRC, "-j8", "emu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lib = C.CDLL(EMU)
    lib.plb200_emu_last_error.restype = C.c_char_p
    return lib


def emu_apply(emu, plb, n, ops, state, scaled=True):
    dtype = state.dtype
    st = np.ascontiguousarray(state.copy())
    blob = plb.OpsBlob(ops)
    stats = (C.c_int64 * 4)()
    rc = emu.plb200_emu_apply_ops(C.c_int64(n), 64 if dtype == np.complex128 else 32, blob.ptr(),
                                  st.ctypes.data_as(C.c_void_p), int(scaled), stats)
    assert rc == 0, emu.plb200_emu_last_error()
    return st, list(stats)


def oracle_apply(n, ops, state):
    sv = np_oracle.StateVector(n, np.complex128)
    sv.set_state(state.astype(np.complex128))
    sv.apply_ops(ops)
    return sv.get_state()


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("scaled", [True, False])
def test_random_circuit_through_tile_interpreter(emu, plb, dtype, scaled):
    n = 15
    ops = circuits.random_circuit(n, 6, 1234)
    st = random_state(n, dtype, 1)
    out, stats = emu_apply(emu, plb, n, ops, st, scaled)
    assert stats[0] >= 1 and stats[3] + stats[1] == len(ops)
    np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_mixed_gate_set_through_tile_interpreter(emu, plb, dtype):
    """Every op kind of the interpreter: scaled rotations, Hadamard, general/real/RX-like 2x2, X,
    phase gates, controlled forms with the control on register / thread / outside bits, control
    value 0, parity diagonals over several bits, global phases, and stand-alone ops in between."""
    n = 15
    rng = np.random.default_rng(3)
    names1 = ["Hadamard", "PauliX", "PauliY", "PauliZ", "S", "SX", "T", "RX", "RY", "RZ", "PhaseShift", "Rot"]
    names2 = ["CNOT", "CZ", "CY", "SWAP", "CRX", "CRY", "CRZ", "CRot", "ControlledPhaseShift", "IsingXX", "IsingZZ",
              "IsingXY", "SingleExcitationPlus"]
    npar = {"RX": 1, "RY": 1, "RZ": 1, "PhaseShift": 1, "Rot": 3, "CRX": 1, "CRY": 1, "CRZ": 1, "CRot": 3,
            "ControlledPhaseShift": 1, "IsingXX": 1, "IsingZZ": 1, "IsingXY": 1, "SingleExcitationPlus": 1}
    ops = []
    for _ in range(400):
        r = rng.random()
        if r < 0.45:
            nm = names1[int(rng.integers(len(names1)))]
            ops.append(circuits.op(nm, [int(rng.integers(n))], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.8:
            nm = names2[int(rng.integers(len(names2)))]
            ops.append(circuits.op(nm, [int(x) for x in rng.permutation(n)[:2]], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.85:
            ops.append(circuits.op("Toffoli", [int(x) for x in rng.permutation(n)[:3]]))
        elif r < 0.93:
            p = [int(x) for x in rng.permutation(n)[:4]]
            nm = ["RY", "RZ", "PauliX", "PhaseShift", "Hadamard"][int(rng.integers(5))]
            ops.append(circuits.op(nm, p[:1], rng.uniform(0, 6, npar.get(nm, 0)), ctrl_wires=p[1:4],
                                   ctrl_values=[bool(b) for b in rng.integers(0, 2, 3)]))
        else:
            k = int(rng.integers(2, 5))
            ops.append(circuits.op("MultiRZ", [int(x) for x in rng.permutation(n)[:k]], [rng.uniform(0, 6)]))
        if rng.random() < 0.05:
            ops.append(circuits.op("GlobalPhase", [0], [0.3]))
    st = random_state(n, dtype, 5)
    hist = (C.c_int64 * 32)()
    emu.plb200_emu_kind_histogram(hist, 1)
    out, stats = emu_apply(emu, plb, n, ops, st)
    assert stats[0] >= 1
    np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=5 * TOL[np.dtype(dtype)])
    emu.plb200_emu_kind_histogram(hist, 1)
    missing = [k for k in range(20) if hist[k] == 0]  # kinds 0..19 = every forward op kind
    assert not missing, (missing, list(hist))


def test_angles_near_multiples_of_pi_pick_the_stable_normalisation(emu, plb):
    """tan(theta/2) blows up at theta = pi: the encoder must switch to the off-diagonal normalisation."""
    n = 13
    ops = []
    for i, th in enumerate([0.0, np.pi, np.pi - 1e-9, np.pi / 2, 2 * np.pi, 3 * np.pi / 2, 1e-12, np.pi + 1e-7]):
        ops += [circuits.op("RX", [i % n], [th]), circuits.op("RY", [(i + 3) % n], [th]),
                circuits.op("RZ", [(i + 5) % n], [th]), circuits.op("CNOT", [i % n, (i + 1) % n])]
    st = random_state(n, np.complex128, 9)
    out, stats = emu_apply(emu, plb, n, ops, st)
    np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=1e-12)


def test_qft_and_sel_from_basis_state(emu, plb):
    n = 14
    for ops in (circuits.qft(n), circuits.strongly_entangling_layers(n, 3, 1)[0]):
        st = np.zeros(1 << n, dtype=np.complex128)
        st[5] = 1.0
        out, stats = emu_apply(emu, plb, n, ops, st)
        np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=1e-12)


def test_long_scaled_tape_keeps_the_pending_scalar_bounded_c64(emu, plb):
    """Hundreds of scaled rotations in c64: the host-side scalar is multiplied back before it can
    leave the float range."""
    n = 14
    rng = np.random.default_rng(11)
    ops = [circuits.op(("RX", "RY", "Hadamard")[int(rng.integers(3))], [int(rng.integers(n))],
                       [np.pi / 2] if True else []) for _ in range(600)]
    for o in ops:
        if o["name"] == "Hadamard":
            o["params"] = []
    st = random_state(n, np.complex64, 2)
    out, stats = emu_apply(emu, plb, n, ops, st)
    assert np.all(np.isfinite(out))
    np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=2e-5)


@pytest.mark.parametrize("specialised", [False, True])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_adjoint_sweep_through_tile_interpreter(emu, plb, dtype, specialised, monkeypatch):
    """Two-state passes with in-register generator overlaps (the fused adjoint) against the oracle's
    adjoint loop (AdjointJacobianLQubit.hpp:347-491); `specialised`: through the generated pass code
    (jit_codegen.hpp, compiled with g++) instead of the interpreter."""
    if specialised:
        monkeypatch.setenv("PLB200_EMU_JIT", "1")
        emu.plb200_emu_jit_passes.restype = C.c_int64
    before = emu.plb200_emu_jit_passes() if specialised else 0
    n = 13
    rng = np.random.default_rng(4)
    ops = []
    for layer in range(3):
        for w in range(n):
            ops.append(circuits.op(("RX", "RY", "RZ")[int(rng.integers(3))], [w], [rng.uniform(0, 6)]))
        for w in range(0, n - 1, 2):
            nm = ("CNOT", "CRZ", "CRX", "IsingZZ", "ControlledPhaseShift")[int(rng.integers(5))]
            ops.append(circuits.op(nm, [w, (w + 1 + layer) % n] if (w + 1 + layer) % n != w else [w, (w + 1) % n],
                                   [rng.uniform(0, 6)] if nm != "CNOT" else []))
    n_par = sum(1 for o in ops if o["params"])
    tp = sorted(rng.choice(n_par, size=n_par * 2 // 3, replace=False).tolist())
    obs = np_oracle.Observable.hamiltonian([0.7, -0.4], [np_oracle.Observable.named("PauliZ", [0]),
                                                          np_oracle.Observable.named("PauliX", [3])])
    ref = np_oracle.StateVector(n, np.complex128)
    expect = ref.adjoint_jacobian([obs], ops, tp, apply_ops=True)
    lam = np_oracle.StateVector(n, np.complex128)
    lam.apply_ops(ops)
    hl = np_oracle.StateVector(n, np.complex128)
    hl.set_state(lam.get_state())
    obs.apply(hl)
    a = np.ascontiguousarray(lam.get_state().astype(dtype))
    b = np.ascontiguousarray(hl.get_state().astype(dtype))
    blob = plb.OpsBlob(ops)
    tpa = np.asarray(tp, dtype=np.int64)
    jac = np.zeros(len(tp))
    stats = (C.c_int64 * 4)()
    rc = emu.plb200_emu_adjoint_sweep(C.c_int64(n), 64 if dtype == np.complex128 else 32, blob.ptr(),
                                      tpa.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(tp)),
                                      a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 1,
                                      jac.ctypes.data_as(C.POINTER(C.c_double)), stats)
    assert rc == 0, emu.plb200_emu_last_error()
    assert stats[0] >= 1
    np.testing.assert_allclose(jac, np.asarray(expect).ravel(), rtol=0, atol=TOL[np.dtype(dtype)] * 20)
    if specialised:
        assert emu.plb200_emu_jit_passes() > before


def _fuzz_tape(n, rng, m, style):
    """Random tapes over the whole gate set: mixed (1-/2-/3-qubit gates, multi-controlled gates with
    random control values, MultiRZ, global phases, angles at multiples of pi/2) or QFT-like ladders."""
    names1 = ["Hadamard", "PauliX", "PauliY", "PauliZ", "S", "SX", "T", "RX", "RY", "RZ", "PhaseShift", "Rot"]
    names2 = ["CNOT", "CZ", "CY", "SWAP", "CRX", "CRY", "CRZ", "CRot", "ControlledPhaseShift", "IsingXX", "IsingZZ",
              "IsingXY", "IsingYY", "SingleExcitation", "SingleExcitationPlus", "SingleExcitationMinus", "PSWAP"]
    npar = {"RX": 1, "RY": 1, "RZ": 1, "PhaseShift": 1, "Rot": 3, "CRX": 1, "CRY": 1, "CRZ": 1, "CRot": 3,
            "ControlledPhaseShift": 1, "IsingXX": 1, "IsingZZ": 1, "IsingXY": 1, "IsingYY": 1, "SingleExcitation": 1,
            "SingleExcitationPlus": 1, "SingleExcitationMinus": 1, "PSWAP": 1}
    ops = []
    for _ in range(m):
        r = rng.random()
        if style == "ladder":
            t = int(rng.integers(n))
            ops.append(circuits.op("Hadamard", [t]))
            for c in rng.permutation(n)[: int(rng.integers(1, n))]:
                if int(c) != t:
                    ops.append(circuits.op("ControlledPhaseShift", [int(c), t], [rng.uniform(0, 6)]))
            if rng.random() < 0.3:
                ops.append(circuits.op("SWAP", [int(x) for x in rng.permutation(n)[:2]]))
        elif r < 0.4:
            nm = names1[int(rng.integers(len(names1)))]
            ang = rng.uniform(0, 6, npar.get(nm, 0))
            if rng.random() < 0.15:
                ang = np.array([np.pi * int(rng.integers(0, 5)) / 2] * len(ang))
            ops.append(circuits.op(nm, [int(rng.integers(n))], ang, inverse=bool(rng.integers(2))))
        elif r < 0.75:
            nm = names2[int(rng.integers(len(names2)))]
            ops.append(circuits.op(nm, [int(x) for x in rng.permutation(n)[:2]], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.8:
            ops.append(circuits.op(["Toffoli", "CSWAP"][int(rng.integers(2))], [int(x) for x in rng.permutation(n)[:3]]))
        elif r < 0.92:
            k = int(rng.integers(1, 4))
            p = [int(x) for x in rng.permutation(n)[: k + 2]]
            nm = ["RY", "RZ", "PauliX", "PhaseShift", "Hadamard", "RX", "SWAP", "IsingZZ", "GlobalPhase"][int(rng.integers(9))]
            nt = 2 if nm in ("SWAP", "IsingZZ") else 1
            ops.append(circuits.op(nm, p[:nt], rng.uniform(0, 6, npar.get(nm, 1 if nm == "GlobalPhase" else 0)),
                                   ctrl_wires=p[nt:nt + k], ctrl_values=[bool(b) for b in rng.integers(0, 2, k)]))
        elif r < 0.97:
            k = int(rng.integers(2, 6))
            ops.append(circuits.op("MultiRZ", [int(x) for x in rng.permutation(n)[:k]], [rng.uniform(0, 6)]))
        else:
            ops.append(circuits.op("GlobalPhase", [0], [rng.uniform(0, 6)]))
    return ops


@pytest.mark.parametrize("multistart", [False, True])
def test_fuzz_random_tapes(emu, plb, multistart, monkeypatch):
    """30 random tapes per scheduler variant (the multi-start tile choice is normally reserved for
    states >= 1 GiB; PLB200_SCHED_MULTISTART forces it here)."""
    if multistart:
        monkeypatch.setenv("PLB200_SCHED_MULTISTART", "1")
    else:
        monkeypatch.setenv("PLB200_SCHED_GREEDY1", "1")
    for seed in range(30):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(12, 16))
        dtype = [np.complex128, np.complex64][seed % 2]
        style = "ladder" if seed % 5 == 0 else "mixed"
        m = int(rng.integers(40, 70)) if style == "ladder" else int(rng.integers(100, 500))
        ops = _fuzz_tape(n, rng, m, style)
        st = random_state(n, dtype, seed)
        out, stats = emu_apply(emu, plb, n, ops, st)
        tol = 5e-12 if dtype == np.complex128 else 3e-4
        err = float(np.max(np.abs(out - oracle_apply(n, ops, st))))
        assert err < tol, (seed, n, style, len(ops), err, stats)


@pytest.mark.parametrize("dtype,n", [(np.complex128, 12), (np.complex64, 14)])
def test_edge_case_tapes(emu, plb, dtype, n):
    """Smallest tiled sizes (n = M + 1) and degenerate tapes: only global phases (a pass that only carries
    the scalar), only diagonal gates, gates only on the always-resident low bits, stand-alone ops between
    passes, a 2000-gate single-qubit tape (pass / record caps, repeated scalar flushes), many controls."""
    rng = np.random.default_rng(0)
    op = circuits.op
    tol = 5e-12 if dtype == np.complex128 else 3e-4

    def check(ops, min_passes=1):
        st = random_state(n, dtype, 3)
        out, stats = emu_apply(emu, plb, n, ops, st)
        assert stats[0] >= min_passes, stats
        assert float(np.max(np.abs(out - oracle_apply(n, ops, st)))) < tol

    check([op("GlobalPhase", [0], [0.3]) for _ in range(5)])
    check([op("RZ", [int(rng.integers(n))], [rng.uniform(0, 6)]) for _ in range(60)]
          + [op("CZ", [int(x) for x in rng.permutation(n)[:2]]) for _ in range(40)])
    check([op("RX", [n - 1 - int(rng.integers(3))], [rng.uniform(0, 6)]) for _ in range(80)])
    check([op("RX", [0], [0.4])], min_passes=0)
    check([op("RX", [0], [0.4]), op("IsingXX", [0, 1], [0.2]), op("RY", [1], [0.1]),
           op("DoubleExcitation", [0, 1, 2, 3], [0.3]), op("RZ", [2], [0.5]), op("RX", [3], [1.0])])
    long_tape = [op(("RX", "RY", "RZ")[int(rng.integers(3))], [int(rng.integers(n))], [rng.uniform(0, 6)])
                 for _ in range(2000)]
    check(long_tape, min_passes=10)
    many = []
    for _ in range(40):
        k = int(rng.integers(3, 7))
        p = [int(x) for x in rng.permutation(n)[: k + 1]]
        many.append(op(("RY", "PhaseShift", "RZ")[int(rng.integers(3))], p[:1], [rng.uniform(0, 6)], ctrl_wires=p[1:],
                       ctrl_values=[bool(b) for b in rng.integers(0, 2, k)]))
        many.append(op("RX", [int(rng.integers(n))], [rng.uniform(0, 6)]))
    check(many)
