"""Pass specialisation (csrc/jit_codegen.hpp + jit_runtime.cpp): the NVRTC-compiled kernel of a fused pass must
give the interpreter kernel's and the reference's amplitudes.

CPU (`-m "not gpu"`): every pass of a tape has a specialised source, NVRTC compiles it for sm_100a (no device
needed), and the SAME generated text, compiled with g++ under -DPLB_JIT_HOST by the test-only emulation
library, run thread by thread on host memory, matches the numpy oracle.
GPU (`-m gpu`): compiled kernels vs interpreter vs lightning.qubit."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import TOL, random_state
from pennylane_lightning_b200 import circuits


def _mixed_tape(n, seed, count=160):
    rng = np.random.default_rng(seed)
    names1 = ["Hadamard", "PauliX", "PauliY", "PauliZ", "S", "SX", "T", "RX", "RY", "RZ", "PhaseShift", "Rot"]
    names2 = ["CNOT", "CZ", "CY", "SWAP", "CRX", "CRY", "CRZ", "CRot", "ControlledPhaseShift", "IsingZZ"]
    npar = {"RX": 1, "RY": 1, "RZ": 1, "PhaseShift": 1, "Rot": 3, "CRX": 1, "CRY": 1, "CRZ": 1, "CRot": 3,
            "ControlledPhaseShift": 1, "IsingZZ": 1}
    ops = []
    for _ in range(count):
        r = rng.random()
        if r < 0.45:
            nm = names1[int(rng.integers(len(names1)))]
            ops.append(circuits.op(nm, [int(rng.integers(n))], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.8:
            nm = names2[int(rng.integers(len(names2)))]
            ops.append(circuits.op(nm, [int(x) for x in rng.permutation(n)[:2]], rng.uniform(0, 6, npar.get(nm, 0)),
                                   inverse=bool(rng.integers(2))))
        elif r < 0.9:
            p = [int(x) for x in rng.permutation(n)]
            ops.append(circuits.op(("RX", "RY", "RZ", "PauliX", "PhaseShift")[int(rng.integers(5))], p[:1],
                                   rng.uniform(0, 6, 1) if rng.random() < 2 else (), ctrl_wires=p[1:3],
                                   ctrl_values=[bool(rng.integers(2)), bool(rng.integers(2))]))
            if ops[-1]["name"] == "PauliX":
                ops[-1]["params"] = []
        else:
            p = [int(x) for x in rng.permutation(n)]
            ops.append(circuits.op("MultiRZ", p[:3], [rng.uniform(0, 6)]))
    return ops


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("prec", [64, 32])
@pytest.mark.parametrize("kind", ["random", "qft"])
def test_every_pass_compiles_with_nvrtc(plb, prec, kind):
    if not plb.jit_available() and not os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12"):
        pytest.skip("NVRTC not present")
    n = 20
    ops = circuits.random_circuit(n, 3, 1234) if kind == "random" else circuits.qft(n)
    blob = plb.OpsBlob(ops)
    npass, nok = C.c_int64(), C.c_int64()
    rc = plb.lib().plb200_jit_compile_check(C.c_int64(n), prec, blob.ptr(), C.byref(npass), C.byref(nok))
    assert rc == 0, plb.lib().plb200_last_error()
    assert npass.value >= 1 and nok.value == npass.value


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_generated_code_on_host_matches_oracle(plb, dtype, monkeypatch):
    from test_tile_emulation import emu_apply, oracle_apply, emu as _emu_fixture  # noqa: F401

    emu = _emu_fixture.__wrapped__() if hasattr(_emu_fixture, "__wrapped__") else None
    if emu is None:
        import subprocess

        from test_tile_emulation import CSRC, EMU

        res = subprocess.run(["make", "-C", CSRC, "-j8", "emu"], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        emu = C.CDLL(EMU)
        emu.plb200_emu_last_error.restype = C.c_char_p
    emu.plb200_emu_jit_passes.restype = C.c_int64
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    n = 15 if dtype == np.complex128 else 16
    before = emu.plb200_emu_jit_passes()
    for ops in (circuits.random_circuit(n, 3, 99), _mixed_tape(n, 5, 90), circuits.qft(n)):
        st = random_state(n, dtype, 2)
        out, stats = emu_apply(emu, plb, n, ops, st, True)
        np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=TOL[np.dtype(dtype)])
    assert emu.plb200_emu_jit_passes() > before


def _pair2_tape(n, seed, count=120):
    """Rotations + the two-qubit gates whose action is 2x2 blocks on amplitude pairs of a 4-dimensional subspace."""
    rng = np.random.default_rng(seed)
    names2 = ["IsingXX", "IsingXY", "IsingYY", "SingleExcitation", "SingleExcitationPlus", "SingleExcitationMinus", "PSWAP",
              "CNOT", "IsingZZ"]
    ops = []
    for _ in range(count):
        if rng.random() < 0.12:  # four-bit pair op (K_PAIR4): a Givens rotation on |0011>, |1100>, also controlled
            w = [int(x) for x in rng.permutation(n)[:5]]
            ctrl = w[4:5] if rng.random() < 0.3 else []
            ops.append(circuits.op("DoubleExcitation", w[:4], [rng.uniform(0, 6)], inverse=bool(rng.integers(2)),
                                   ctrl_wires=ctrl, ctrl_values=[bool(rng.integers(2))] * len(ctrl)))
        elif rng.random() < 0.45:
            ops.append(circuits.op(("RX", "RY", "RZ", "Hadamard")[int(rng.integers(4))], [int(rng.integers(n))],
                                   [rng.uniform(0, 6)] if rng.random() < 2 else []))
            if ops[-1]["name"] == "Hadamard":
                ops[-1]["params"] = []
        else:
            nm = names2[int(rng.integers(len(names2)))]
            w = [int(x) for x in rng.permutation(n)[:4]]
            ctrl = w[2:3] if (rng.random() < 0.25 and nm not in ("CNOT",)) else []
            ops.append(circuits.op(nm, w[:2], [rng.uniform(0, 6)] if nm != "CNOT" else [], inverse=bool(rng.integers(2)),
                                   ctrl_wires=ctrl, ctrl_values=[bool(rng.integers(2))] * len(ctrl)))
    return ops


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_two_bit_pair_ops_fused_on_host(plb, dtype, monkeypatch):
    """PLB200_FUSE_PAIR2=1: IsingXX / XY / YY, SingleExcitation(+-), PSWAP (also controlled) run INSIDE the tile
    passes as K_PAIR2 ops of the generated code, DoubleExcitation as K_PAIR4: no stand-alone kernels, oracle's
    amplitudes."""
    import subprocess

    from test_tile_emulation import CSRC, EMU, emu_apply, oracle_apply

    res = subprocess.run(["make", "-C", CSRC, "-j8", "emu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    emu = C.CDLL(EMU)
    emu.plb200_emu_last_error.restype = C.c_char_p
    monkeypatch.setenv("PLB200_FUSE_PAIR2", "1")
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    n = 14
    ops = _pair2_tape(n, 3)
    st = random_state(n, dtype, 9)
    out, stats = emu_apply(emu, plb, n, ops, st, True)
    assert stats[1] == 0, stats  # nothing ran stand-alone
    np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=TOL[np.dtype(dtype)])
    hist = (C.c_int64 * 32)()
    emu.plb200_emu_kind_histogram(hist, 1)
    assert hist[26] > 0 and hist[31] > 0  # K_PAIR2 and K_PAIR4 records were emitted
    # without the switch the same tape leaves those gates stand-alone
    monkeypatch.setenv("PLB200_FUSE_PAIR2", "0")
    _, stats0 = emu_apply(emu, plb, n, ops, st, True)
    assert stats0[1] > 0


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture
def jit_sync(plb, monkeypatch):
    if not plb.jit_available():
        pytest.fail("NVRTC / libcuda could not be loaded on a GPU box")
    monkeypatch.setenv("PLB200_JIT_MIN_QUBITS", "12")
    plb.jit_set_mode(2)
    yield
    plb.jit_set_mode(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("kind", ["random", "mixed", "qft", "sel"])
def test_jit_kernels_match_interpreter_and_reference(plb, ref, jit_sync, dtype, kind):
    n = 18
    ops = {"random": lambda: circuits.random_circuit(n, 6, 1234), "mixed": lambda: _mixed_tape(n, 7, 300),
           "qft": lambda: circuits.qft(n), "sel": lambda: circuits.strongly_entangling_layers(n, 3, 42)[0]}[kind]()
    st = random_state(n, dtype, 3)
    before = plb.jit_stats()
    a = plb.StateVector(n, dtype)
    a.set_state(st)
    a.apply_ops(ops, fuse=True)
    got = a.get_state()
    after = plb.jit_stats()
    assert after["jit_launches"] > before["jit_launches"], (before, after)
    assert after["failed"] == before["failed"]
    plb.jit_set_mode(0)
    b = plb.StateVector(n, dtype)
    b.set_state(st)
    b.apply_ops(ops, fuse=True)
    interp = b.get_state()
    plb.jit_set_mode(2)
    r = ref.StateVector(n, dtype)
    r.set_state(st)
    r.apply_ops(ops)
    tol = TOL[np.dtype(dtype)]
    np.testing.assert_allclose(got, r.get_state(), rtol=0, atol=tol)
    np.testing.assert_allclose(got, interp, rtol=0, atol=tol)


@pytest.mark.gpu
def test_jit_async_tiering(plb, monkeypatch):
    """Default mode: first sighting runs the interpreter, the second queues a background compile, after
    jit_wait() the compiled kernel runs — and all three give the same state."""
    if not plb.jit_available():
        pytest.fail("NVRTC / libcuda could not be loaded on a GPU box")
    monkeypatch.setenv("PLB200_JIT_MIN_QUBITS", "12")
    plb.jit_set_mode(1)
    try:
        n = 17
        ops = circuits.random_circuit(n, 5, 4321)
        states = []
        s0 = plb.jit_stats()
        for it in range(3):
            sv = plb.StateVector(n)
            sv.apply_ops(ops, fuse=True)
            states.append(sv.get_state())
            if it == 1:
                plb.jit_wait()
        s1 = plb.jit_stats()
        assert s1["compiled"] + s1["from_disk_cache"] > s0["compiled"] + s0["from_disk_cache"]
        assert s1["jit_launches"] > s0["jit_launches"] and s1["interpreter_launches"] > s0["interpreter_launches"]
        np.testing.assert_allclose(states[2], states[0], rtol=0, atol=1e-12)
    finally:
        plb.jit_set_mode(-1)


@pytest.mark.gpu
def test_jit_same_kernel_other_angles(plb, ref, jit_sync):
    """Angles are run-time arguments: a second parameter set reuses the compiled kernels (no new compile)."""
    n = 18
    ops1, _ = circuits.strongly_entangling_layers(n, 2, 1)
    ops2, _ = circuits.strongly_entangling_layers(n, 2, 2)
    a = plb.StateVector(n)
    a.apply_ops(ops1, fuse=True)
    s1 = plb.jit_stats()
    b = plb.StateVector(n)
    b.apply_ops(ops2, fuse=True)
    s2 = plb.jit_stats()
    assert s2["compiled"] + s2["from_disk_cache"] == s1["compiled"] + s1["from_disk_cache"]
    assert s2["jit_launches"] > s1["jit_launches"]
    r = ref.StateVector(n)
    r.apply_ops(ops2)
    np.testing.assert_allclose(b.get_state(), r.get_state(), rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_jit_adjoint_passes_match_reference(plb, ref, jit_sync, dtype):
    """The fused adjoint sweep (two-state passes with in-register generator overlaps) through specialised kernels."""
    n = 16
    ops, tp = circuits.hardware_efficient_ansatz(n, 120, 5)
    co, words, wires = circuits.pauli_hamiltonian(n, 10, 5)
    a, r = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    ham_a = circuits.hamiltonian_observable(plb, co, words, wires)
    ham_r = circuits.hamiltonian_observable(ref, co, words, wires, dtype=dtype)
    before = plb.jit_stats()
    ja = a.adjoint_jacobian([ham_a], ops, tp, apply_ops=True)
    after = plb.jit_stats()
    assert after["jit_launches"] > before["jit_launches"] and after["failed"] == before["failed"]
    jr = r.adjoint_jacobian([ham_r], ops, tp, apply_ops=True)
    np.testing.assert_allclose(ja, jr, rtol=0, atol=1e-11 if dtype == np.complex128 else 2e-4)
    np.testing.assert_array_equal(a.adjoint_jacobian([ham_a], ops, tp, apply_ops=True), ja)  # fixed-order sums
    plb.jit_set_mode(0)
    ji = a.adjoint_jacobian([ham_a], ops, tp, apply_ops=True)
    plb.jit_set_mode(2)
    np.testing.assert_allclose(ja, ji, rtol=0, atol=1e-11 if dtype == np.complex128 else 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_jit_two_bit_pair_ops_fused(plb, ref, jit_sync, dtype):
    """In the compile-at-first-sight regime the two-bit pair gates are tile ops: one launch per pass, no stand-alone
    kernels in between, reference amplitudes."""
    n = 17
    ops = _pair2_tape(n, 4, 200)
    st = random_state(n, dtype, 5)
    a, r = plb.StateVector(n, dtype), ref.StateVector(n, dtype)
    a.set_state(st), r.set_state(st)
    blob = plb.OpsBlob(ops)
    stats = (C.c_int64 * 4)()
    assert plb.lib().plb200_schedule_stats(C.c_int64(n), 64 if dtype == np.complex128 else 32, blob.ptr(), stats) == 0
    assert stats[1] == 0 and stats[0] >= 1, list(stats)
    a.apply_ops(blob, fuse=True)
    r.apply_ops(ops)
    assert a.last_apply_stats()[1] == stats[0]
    np.testing.assert_allclose(a.get_state(), r.get_state(), rtol=0, atol=TOL[np.dtype(dtype)])


def _emu_lib():
    import subprocess

    from test_tile_emulation import CSRC, EMU

    res = subprocess.run(["make", "-C", CSRC, "-j8", "emu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    emu = C.CDLL(EMU)
    emu.plb200_emu_last_error.restype = C.c_char_p
    emu.plb200_emu_jit_passes.restype = C.c_int64
    return emu


K_LADDER, K_SROT_R, K_SROT_I, K_SROK_R, K_SROK_I = 22, 27, 28, 29, 30


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_jit_forms_scaled_rotations_and_in_stream_ladders(plb, dtype, monkeypatch):
    """The encoder's "jit forms" (fusion.cu build_schedule, jit_forms = true): uncontrolled RX / RY as scaled
    rotations (tangent form, cotangent form at half turns), controlled two-valued diagonals split into controlled
    phases, phases sharing a register bit merged into in-stream ladders.  The generated code on host memory must
    give the oracle's amplitudes, and the same tape in the interpreter's forms must not contain those kinds."""
    from test_tile_emulation import emu_apply, oracle_apply

    emu = _emu_lib()
    n = 15 if dtype == np.complex128 else 16
    rng = np.random.default_rng(17)
    special = [np.pi, -np.pi, 3 * np.pi, np.pi - 1e-7, np.pi + 1e-9, 0.0, 2 * np.pi, np.pi / 2, 1e-9]
    ops = []
    for layer in range(8):
        for w in range(n):
            th = special[int(rng.integers(len(special)))] if rng.random() < 0.3 else rng.uniform(0, 2 * np.pi)
            ops.append(circuits.op(("RX", "RY", "RZ")[int(rng.integers(3))], [w], [th], inverse=bool(rng.integers(2))))
        p = rng.permutation(n)
        for i in range(0, n - 1, 2):
            nm = ("CNOT", "CRZ", "CRZ", "ControlledPhaseShift", "CZ")[int(rng.integers(5))]
            ops.append(circuits.op(nm, [int(p[i]), int(p[i + 1])], [rng.uniform(0, 6)] if nm in ("CRZ", "ControlledPhaseShift") else []))
    st = random_state(n, dtype, 3)
    expect = oracle_apply(n, ops, st)
    hist = (C.c_int64 * 32)()
    # the interpreter's forms: none of the specialised kinds, no in-stream ladders of fewer than six entries
    monkeypatch.delenv("PLB200_EMU_JIT", raising=False)
    emu.plb200_emu_kind_histogram(hist, 1)
    out, _ = emu_apply(emu, plb, n, ops, st)
    emu.plb200_emu_kind_histogram(hist, 1)
    np.testing.assert_allclose(out, expect, rtol=0, atol=4 * TOL[np.dtype(dtype)])
    assert hist[K_SROT_R] + hist[K_SROT_I] + hist[K_SROK_R] + hist[K_SROK_I] == 0
    # the specialised forms
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    before = emu.plb200_emu_jit_passes()
    out, stats = emu_apply(emu, plb, n, ops, st)
    emu.plb200_emu_kind_histogram(hist, 1)
    assert emu.plb200_emu_jit_passes() > before and stats[1] == 0
    np.testing.assert_allclose(out, expect, rtol=0, atol=4 * TOL[np.dtype(dtype)])
    assert hist[K_SROT_R] > 0 and hist[K_SROT_I] > 0, list(hist)
    assert hist[K_SROK_R] + hist[K_SROK_I] > 0, list(hist)  # the exact half turns
    assert hist[K_LADDER] > 0, list(hist)
    # PLB200_JIT_FORMS=0 keeps the specialised kernels on the interpreter's encoding
    monkeypatch.setenv("PLB200_JIT_FORMS", "0")
    out, _ = emu_apply(emu, plb, n, ops, st)
    emu.plb200_emu_kind_histogram(hist, 1)
    np.testing.assert_allclose(out, expect, rtol=0, atol=4 * TOL[np.dtype(dtype)])
    assert hist[K_SROT_R] + hist[K_SROT_I] == 0


def test_jit_forms_growth_stays_bounded_c64(plb, monkeypatch):
    """600 rotations with angles a hair away from a half turn (tangents up to 2^16): the stored amplitudes of the
    tangent form grow by 1 / cos per rotation; the encoder must fold the carried scalar back / switch to the
    cotangent form before fp32 overflows."""
    from test_tile_emulation import emu_apply, oracle_apply

    emu = _emu_lib()
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    n = 16
    rng = np.random.default_rng(5)
    ops = []
    for _ in range(600):
        th = np.pi + rng.choice([-1, 1]) * 10 ** rng.uniform(-6, -1)
        ops.append(circuits.op(("RX", "RY")[int(rng.integers(2))], [int(rng.integers(n))], [th]))
        if rng.random() < 0.2:
            ops.append(circuits.op("CNOT", [int(x) for x in rng.permutation(n)[:2]]))
    st = random_state(n, np.complex64, 8)
    out, _ = emu_apply(emu, plb, n, ops, st)
    assert np.all(np.isfinite(out.view(np.float32)))
    np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=5e-5)


@pytest.mark.parametrize("prec", [64, 32])
def test_pass_sources_do_not_depend_on_angles_or_on_refusals(plb, prec, monkeypatch, tmp_path):
    """The structure key of a compiled pass is its generated text.  It must be the same (a) for every parameter set
    of a variational circuit — the carried scalar has one slot per pass whether it is folded back there or not,
    tangent-form rotations do not switch form with the angle — and (b) whether earlier passes ran in the
    specialised forms or were refused (asynchronous tier, first sightings) and ran in the interpreter's."""
    n = 26

    def sources(ops, sub, refuse):
        out = tmp_path / sub
        out.mkdir()
        if refuse:
            monkeypatch.setenv("PLB200_DUMP_REFUSE", "1")
        else:
            monkeypatch.delenv("PLB200_DUMP_REFUSE", raising=False)
        blob = plb.OpsBlob(ops)
        npass = C.c_int64()
        rc = plb.lib().plb200_jit_dump_sources(C.c_int64(n), prec, blob.ptr(), str(out).encode(), C.byref(npass))
        assert rc == 0, plb.lib().plb200_last_error()
        return [(out / f"pass_{i}.cu").read_text() for i in range(npass.value)]

    ops = circuits.random_circuit(n, 12, 77)
    rng = np.random.default_rng(3)
    ops2 = [dict(o, params=[float(rng.uniform(0, 2 * np.pi)) for _ in o["params"]]) for o in ops]
    a = sources(ops, "a", False)
    assert len(a) >= 5 and all("srot_" in t for t in a[:3])
    assert sources(ops, "b", True) == a
    assert sources(ops2, "c", False) == a
    assert sources(ops2, "d", True) == a


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_jit_forms_special_angles_on_device(plb, ref, jit_sync, dtype):
    """Scaled rotations on the device: half turns (cotangent form), angles a hair away from them (tangents up to
    2^16, growth bounded per pass), zero angles, mixed with CRZ / CNOT / controlled phases (split diagonals,
    in-stream ladders) — against lightning.qubit."""
    n = 20
    rng = np.random.default_rng(23)
    special = [np.pi, -np.pi, 3 * np.pi, np.pi - 1e-6, np.pi + 1e-4, 0.0, 2 * np.pi, np.pi / 2]
    ops = []
    for layer in range(6):
        for w in range(n):
            th = special[int(rng.integers(len(special)))] if rng.random() < 0.4 else rng.uniform(0, 2 * np.pi)
            ops.append(circuits.op(("RX", "RY", "RZ")[int(rng.integers(3))], [w], [th], inverse=bool(rng.integers(2))))
        p = rng.permutation(n)
        for i in range(0, n - 1, 2):
            nm = ("CNOT", "CRZ", "ControlledPhaseShift", "CZ")[int(rng.integers(4))]
            ops.append(circuits.op(nm, [int(p[i]), int(p[i + 1])], [rng.uniform(0, 6)] if nm in ("CRZ", "ControlledPhaseShift") else []))
    s0 = plb.jit_stats()
    a = plb.StateVector(n, dtype)
    a.apply_ops(ops, fuse=True)
    s1 = plb.jit_stats()
    assert s1["jit_launches"] > s0["jit_launches"]
    r = ref.StateVector(n, dtype)
    r.apply_ops(ops)
    np.testing.assert_allclose(a.get_state(), r.get_state(), rtol=0, atol=4 * TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_refused_passes_fall_back_to_the_interpreter(plb, dtype, monkeypatch):
    """Asynchronous tier, first sightings: a consumer without the pass's kernel refuses it (PLB200_EMU_REFUSE=1).
    A pass in the jit forms comes back in the interpreter's encoding; a pass holding two- / four-bit pair ops is
    cut at those ops — interpreter passes over the same tile in between, the pair ops stand-alone.  Same
    amplitudes as the oracle either way, and nothing ran through generated code."""
    from test_tile_emulation import emu_apply, oracle_apply

    emu = _emu_lib()
    monkeypatch.setenv("PLB200_FUSE_PAIR2", "1")
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    monkeypatch.setenv("PLB200_EMU_REFUSE", "1")
    n = 14
    for ops in (_pair2_tape(n, 4, 150), circuits.random_circuit(n, 5, 3)):
        st = random_state(n, dtype, 6)
        before = emu.plb200_emu_jit_passes()
        out, stats = emu_apply(emu, plb, n, ops, st, True)
        assert emu.plb200_emu_jit_passes() == before  # every pass ran through the interpreter emulation
        np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=2 * TOL[np.dtype(dtype)])
        n_pair = sum(1 for o in ops if o["name"] in ("IsingXX", "IsingXY", "IsingYY", "SingleExcitation", "SingleExcitationPlus",
                                                     "SingleExcitationMinus", "PSWAP", "DoubleExcitation"))
        assert stats[1] >= n_pair and stats[0] >= 1, stats  # the pair ops ran stand-alone, the rest in tile passes


@pytest.mark.gpu
def test_pair_ops_fused_in_the_default_tier(plb, ref, monkeypatch):
    """Default (asynchronous) tier: a tape with IsingXX / SingleExcitation / DoubleExcitation ... is scheduled with
    those gates INSIDE the tile passes.  First sightings have no kernel: the passes are cut at the pair ops
    (interpreter pieces + stand-alone kernels).  After the background compile the whole tape runs as specialised
    passes with no stand-alone kernel — and every run gives lightning.qubit's state."""
    if not plb.jit_available():
        pytest.fail("NVRTC / libcuda could not be loaded on a GPU box")
    monkeypatch.setenv("PLB200_JIT_MIN_QUBITS", "12")
    plb.jit_set_mode(1)
    try:
        n = 17
        ops = _pair2_tape(n, 8, 160)
        r = ref.StateVector(n)
        r.apply_ops(ops)
        expect = r.get_state()
        blob = plb.OpsBlob(ops)
        sched = (C.c_int64 * 4)()
        assert plb.lib().plb200_schedule_stats(C.c_int64(n), 64, blob.ptr(), sched) == 0
        assert sched[1] == 0 and sched[0] >= 2, list(sched)  # every gate of the tape sits in a tile pass
        launches = []
        for it in range(3):
            sv = plb.StateVector(n)
            sv.apply_ops(ops, fuse=True)
            np.testing.assert_allclose(sv.get_state(), expect, rtol=0, atol=1e-12)
            launches.append(sv.last_apply_stats()[1])
            if it == 1:
                plb.jit_wait()
        # first sighting: interpreter pieces + stand-alone pair ops; with the kernels: one launch per pass
        assert launches[0] > sched[0] and launches[2] == sched[0], (launches, list(sched))
    finally:
        plb.jit_set_mode(-1)


def _unitary_tape(n, seed, count=90):
    """Rotations / CNOTs mixed with Haar-ish random unitaries on two wires (QubitUnitary), some controlled."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(count):
        r = rng.random()
        if r < 0.4:
            ops.append(circuits.op(("RX", "RY", "RZ")[int(rng.integers(3))], [int(rng.integers(n))], [rng.uniform(0, 6)]))
        elif r < 0.55:
            ops.append(circuits.op("CNOT", [int(x) for x in rng.permutation(n)[:2]]))
        else:
            a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            q, rr = np.linalg.qr(a)
            u = q * (np.diag(rr) / np.abs(np.diag(rr)))
            w = [int(x) for x in rng.permutation(n)[:4]]
            ctrl = w[2:3] if rng.random() < 0.3 else []
            o = circuits.op("QubitUnitary", w[:2], [], inverse=bool(rng.integers(2)), ctrl_wires=ctrl,
                            ctrl_values=[bool(rng.integers(2))] * len(ctrl))
            o["matrix"] = u
            ops.append(o)
    return ops


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_two_wire_unitaries_fused_on_host(plb, dtype, monkeypatch):
    """QubitUnitary on two wires (also controlled, also inverted) as K_DENSE2 tile ops of the generated code: no
    stand-alone kernel, the oracle's amplitudes; refused (no kernel yet) they run stand-alone between interpreter
    pieces with the same result."""
    from test_tile_emulation import emu_apply, oracle_apply

    emu = _emu_lib()
    monkeypatch.setenv("PLB200_FUSE_PAIR2", "1")
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    n = 14
    ops = _unitary_tape(n, 12)
    n_u = sum(1 for o in ops if o["name"] == "QubitUnitary")
    st = random_state(n, dtype, 4)
    expect = oracle_apply(n, ops, st)
    out, stats = emu_apply(emu, plb, n, ops, st, True)
    assert stats[1] == 0 and stats[3] == len(ops), stats
    np.testing.assert_allclose(out, expect, rtol=0, atol=2 * TOL[np.dtype(dtype)])
    monkeypatch.setenv("PLB200_EMU_REFUSE", "1")
    out, stats = emu_apply(emu, plb, n, ops, st, True)
    assert stats[1] >= n_u, stats
    np.testing.assert_allclose(out, expect, rtol=0, atol=2 * TOL[np.dtype(dtype)])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_two_wire_unitaries_fused_on_device(plb, ref, jit_sync, dtype):
    """QubitUnitary on two wires inside the specialised passes (K_DENSE2) against lightning.qubit; one launch per
    pass, nothing stand-alone."""
    n = 17
    ops = _unitary_tape(n, 21, 140)
    blob = plb.OpsBlob(ops)
    sched = (C.c_int64 * 4)()
    assert plb.lib().plb200_schedule_stats(C.c_int64(n), 64 if dtype == np.complex128 else 32, blob.ptr(), sched) == 0
    assert sched[1] == 0, list(sched)
    a = plb.StateVector(n, dtype)
    a.apply_ops(blob, fuse=True)
    assert a.last_apply_stats()[1] == sched[0]
    r = ref.StateVector(n, dtype)
    r.apply_ops(ops)
    np.testing.assert_allclose(a.get_state(), r.get_state(), rtol=0, atol=2 * TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_non_parity_diagonals_expand_into_controlled_phases(plb, dtype, monkeypatch):
    """fusion.cu expand_for_fusion: the residual 16-entry diagonal of DoubleExcitationPlus / Minus and diagonal
    QubitUnitaries whose entries are all equal but a few become a scalar + controlled phases, which every pass can
    hold — with pair ops fused nothing of such a tape runs stand-alone; both through the generated code and through
    the interpreter (refused passes)."""
    from test_tile_emulation import emu_apply, oracle_apply

    emu = _emu_lib()
    monkeypatch.setenv("PLB200_FUSE_PAIR2", "1")
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    n = 14
    rng = np.random.default_rng(31)
    ops = []
    for _ in range(70):
        r = rng.random()
        w = [int(x) for x in rng.permutation(n)[:5]]
        if r < 0.35:
            ops.append(circuits.op(("RX", "RY", "RZ", "Hadamard")[int(rng.integers(4))], w[:1], [rng.uniform(0, 6)]))
            if ops[-1]["name"] == "Hadamard":
                ops[-1]["params"] = []
        elif r < 0.6:
            ops.append(circuits.op(("DoubleExcitationPlus", "DoubleExcitationMinus")[int(rng.integers(2))], w[:4],
                                   [rng.uniform(0, 6)], inverse=bool(rng.integers(2)),
                                   ctrl_wires=w[4:5] if rng.random() < 0.3 else [],
                                   ctrl_values=[True] if False else []))
            ops[-1]["ctrl_values"] = [bool(rng.integers(2))] * len(ops[-1]["ctrl_wires"])
        elif r < 0.8:  # diagonal unitary on 2 or 3 wires with one or two exceptional entries
            k = int(rng.integers(2, 4))
            d = np.full(1 << k, np.exp(1j * rng.uniform(0, 6)))
            for i in rng.permutation(1 << k)[: int(rng.integers(1, 3))]:
                d[i] = np.exp(1j * rng.uniform(0, 6))
            o = circuits.op("QubitUnitary", w[:k], [], inverse=bool(rng.integers(2)))
            o["matrix"] = np.diag(d)
            ops.append(o)
        else:
            ops.append(circuits.op("CNOT", w[:2]))
    st = random_state(n, dtype, 2)
    expect = oracle_apply(n, ops, st)
    out, stats = emu_apply(emu, plb, n, ops, st, True)
    assert stats[1] == 0 and stats[0] >= 1, stats
    np.testing.assert_allclose(out, expect, rtol=0, atol=2 * TOL[np.dtype(dtype)])
    monkeypatch.setenv("PLB200_EMU_REFUSE", "1")
    out, stats = emu_apply(emu, plb, n, ops, st, True)
    np.testing.assert_allclose(out, expect, rtol=0, atol=2 * TOL[np.dtype(dtype)])


@pytest.mark.gpu
def test_excitation_variants_and_diagonal_unitaries_fused_on_device(plb, ref, jit_sync):
    """DoubleExcitationPlus / Minus (K_PAIR4 + expanded residual phases) and nearly-uniform diagonal unitaries inside
    specialised passes against lightning.qubit: no stand-alone kernel."""
    n = 17
    rng = np.random.default_rng(41)
    ops = []
    for _ in range(80):
        r = rng.random()
        w = [int(x) for x in rng.permutation(n)[:5]]
        if r < 0.4:
            ops.append(circuits.op(("RX", "RY", "RZ")[int(rng.integers(3))], w[:1], [rng.uniform(0, 6)]))
        elif r < 0.65:
            ops.append(circuits.op(("DoubleExcitationPlus", "DoubleExcitationMinus", "DoubleExcitation")[int(rng.integers(3))],
                                   w[:4], [rng.uniform(0, 6)], inverse=bool(rng.integers(2))))
        elif r < 0.8:
            d = np.full(8, np.exp(1j * rng.uniform(0, 6)))
            d[int(rng.integers(8))] = np.exp(1j * rng.uniform(0, 6))
            o = circuits.op("QubitUnitary", w[:3], [])
            o["matrix"] = np.diag(d)
            ops.append(o)
        else:
            ops.append(circuits.op("CNOT", w[:2]))
    blob = plb.OpsBlob(ops)
    sched = (C.c_int64 * 4)()
    assert plb.lib().plb200_schedule_stats(C.c_int64(n), 64, blob.ptr(), sched) == 0
    assert sched[1] == 0, list(sched)
    a = plb.StateVector(n)
    a.apply_ops(blob, fuse=True)
    assert a.last_apply_stats()[1] == sched[0]
    r_ = ref.StateVector(n)
    r_.apply_ops(ops)
    np.testing.assert_allclose(a.get_state(), r_.get_state(), rtol=0, atol=1e-12)


def _pauli_rot_tape(n, seed, count=60):
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(count):
        r = rng.random()
        w = [int(x) for x in rng.permutation(n)[:4]]
        if r < 0.45:
            k = int(rng.integers(1, 5))
            word = "".join("XYZ"[int(rng.integers(3))] for _ in range(k))
            ops.append(circuits.op(f"PauliRot[{word}]", w[:k], [rng.uniform(0, 6)], inverse=bool(rng.integers(2))))
        elif r < 0.8:
            ops.append(circuits.op(("RX", "RY", "RZ")[int(rng.integers(3))], w[:1], [rng.uniform(0, 6)]))
        else:
            ops.append(circuits.op("CNOT", w[:2]))
    return ops


def _apply_with_pauli_rot(n, ops, state):
    from oracle import np_oracle

    sv = np_oracle.StateVector(n, np.complex128)
    sv.set_state(state.astype(np.complex128))
    for o in ops:
        if o["name"].startswith("PauliRot["):
            sv.apply_pauli_rot(o["wires"], o["inverse"], o["params"][0], o["name"][9:-1])
        else:
            sv.apply_ops([o])
    return sv.get_state()


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("specialised", [False, True])
def test_pauli_rotations_fused_through_basis_changes(plb, dtype, specialised, monkeypatch):
    """fusion.cu expand_for_fusion: exp(-i theta/2 P) with X / Y letters runs inside tile passes as H / S basis
    changes around a parity diagonal — nothing stand-alone — and gives GateImplementationsLM.hpp:575-629's state
    (the oracle's apply_pauli_rot), through the interpreter and through the generated code."""
    from test_tile_emulation import emu_apply

    emu = _emu_lib()
    if specialised:
        monkeypatch.setenv("PLB200_EMU_JIT", "1")
    n = 14
    ops = _pauli_rot_tape(n, 6)
    st = random_state(n, dtype, 3)
    out, stats = emu_apply(emu, plb, n, ops, st, True)
    assert stats[1] == 0 and stats[0] >= 1, stats
    np.testing.assert_allclose(out, _apply_with_pauli_rot(n, ops, st), rtol=0, atol=2 * TOL[np.dtype(dtype)])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["interpreter", "specialised"])
def test_pauli_rotations_fused_on_device(plb, ref, mode, monkeypatch):
    """Pauli rotations with X / Y letters inside a tape: no stand-alone kernel (basis changes + a parity diagonal
    inside the tile passes), lightning.qubit's applyPauliRot state."""
    monkeypatch.setenv("PLB200_JIT_MIN_QUBITS", "12")
    plb.jit_set_mode(2 if mode == "specialised" else 0)
    try:
        n = 17
        ops = _pauli_rot_tape(n, 9, 80)
        blob = plb.OpsBlob(ops)
        sched = (C.c_int64 * 4)()
        assert plb.lib().plb200_schedule_stats(C.c_int64(n), 64, blob.ptr(), sched) == 0
        assert sched[1] == 0, list(sched)
        a = plb.StateVector(n)
        a.apply_ops(blob, fuse=True)
        assert a.last_apply_stats()[1] == sched[0]
        r = ref.StateVector(n)
        for o in ops:
            if o["name"].startswith("PauliRot["):
                r.apply_pauli_rot(o["wires"], o["inverse"], o["params"][0], o["name"][9:-1])
            else:
                r.apply_ops([o])
        np.testing.assert_allclose(a.get_state(), r.get_state(), rtol=0, atol=1e-12)
    finally:
        plb.jit_set_mode(-1)


def test_fuzz_generated_code_all_forms(plb, monkeypatch):
    """Random tapes over the whole gate set — plus DoubleExcitation(+-), two-wire unitaries, nearly-uniform diagonal
    unitaries — through the encoder's jit forms and the g++-compiled generated code, with pair ops fused
    (PLB200_FUSE_PAIR2=1); every third tape with all such passes refused (interpreter pieces + stand-alone pair ops).
    Amplitudes against the oracle."""
    from test_tile_emulation import _fuzz_tape, emu_apply, oracle_apply

    emu = _emu_lib()
    monkeypatch.setenv("PLB200_FUSE_PAIR2", "1")
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    for seed in range(12):
        rng = np.random.default_rng(7000 + seed)
        n = int(rng.integers(12, 15)) + (seed % 2) * 2  # c64 tiles need 14 qubits
        dtype = [np.complex128, np.complex64][seed % 2]
        ops = _fuzz_tape(n, rng, int(rng.integers(60, 160)), "ladder" if seed % 6 == 5 else "mixed")
        extra = []
        for _ in range(int(rng.integers(3, 10))):
            w = [int(x) for x in rng.permutation(n)[:5]]
            r = rng.random()
            if r < 0.4:
                nm = ("DoubleExcitation", "DoubleExcitationPlus", "DoubleExcitationMinus")[int(rng.integers(3))]
                extra.append(circuits.op(nm, w[:4], [rng.uniform(0, 6)], inverse=bool(rng.integers(2))))
            elif r < 0.7:
                a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
                q, rr = np.linalg.qr(a)
                o = circuits.op("QubitUnitary", w[:2], [], inverse=bool(rng.integers(2)),
                                ctrl_wires=w[2:3] if rng.random() < 0.3 else [])
                o["ctrl_values"] = [bool(rng.integers(2))] * len(o["ctrl_wires"])
                o["matrix"] = q * (np.diag(rr) / np.abs(np.diag(rr)))
                extra.append(o)
            else:
                d = np.full(8, np.exp(1j * rng.uniform(0, 6)))
                d[int(rng.integers(8))] = np.exp(1j * rng.uniform(0, 6))
                o = circuits.op("QubitUnitary", w[:3], [])
                o["matrix"] = np.diag(d)
                extra.append(o)
        for o in extra:
            ops.insert(int(rng.integers(len(ops) + 1)), o)
        if seed % 3 == 2:
            monkeypatch.setenv("PLB200_EMU_REFUSE", "1")
        else:
            monkeypatch.delenv("PLB200_EMU_REFUSE", raising=False)
        st = random_state(n, dtype, seed)
        out, stats = emu_apply(emu, plb, n, ops, st)
        tol = 5e-12 if dtype == np.complex128 else 3e-4
        err = float(np.max(np.abs(out - oracle_apply(n, ops, st))))
        assert err < tol, (seed, n, len(ops), err, stats)
