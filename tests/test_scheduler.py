"""CPU tests of the fusion scheduler (host logic only, through plb200_schedule_stats): every gate of
the tape is scheduled exactly once, passes are far fewer than gates, un-fusable ops run stand-alone."""
import ctypes as C

import numpy as np
import pytest

from pennylane_lightning_b200 import circuits


def stats(plb, n, ops, precision=64):
    blob = plb.OpsBlob(ops)
    out = (C.c_int64 * 4)()
    rc = plb.lib().plb200_schedule_stats(C.c_int64(n), precision, blob.ptr(), out)
    assert rc == 0, plb.lib().plb200_last_error()
    return list(out)


@pytest.mark.parametrize("precision", [64, 32])
def test_random_circuit_30q_is_blocked_into_few_passes(plb, precision):
    ops = circuits.random_circuit(30, 20, 1234)
    passes, alone, rounds, fused = stats(plb, 30, ops, precision)
    assert fused + alone == len(ops) == 900
    assert alone == 0
    assert passes <= 30 and rounds <= 120


def test_small_states_are_not_tiled(plb):
    ops = circuits.random_circuit(10, 3, 1)
    passes, alone, rounds, fused = stats(plb, 10, ops)
    assert passes == 0 and alone == len(ops)


def test_unfusable_ops_run_alone_and_everything_is_scheduled_once(plb):
    n = 20
    ops = circuits.qft(n)  # H + ControlledPhaseShift ladders + SWAP: all inside tile passes
    passes, alone, rounds, fused = stats(plb, n, ops)
    assert alone == 0 and fused == len(ops) and passes <= 6
    # IsingXX / DoubleExcitation have no tile form: they run stand-alone between the passes
    extra = [circuits.op("IsingXX", [3, 11], [0.3]), circuits.op("DoubleExcitation", [0, 5, 9, 14], [0.2])]
    ops2 = ops[:100] + extra + ops[100:]
    passes, alone, rounds, fused = stats(plb, n, ops2)
    assert alone >= 2 and fused + alone >= len(ops2)  # (a lowered gate may be more than one canonical op)
    ops = circuits.strongly_entangling_layers(20, 4, 42)[0]
    passes, alone, rounds, fused = stats(plb, 20, ops)
    assert fused + alone == len(ops) and passes < len(ops) // 10


def test_lowering_errors_are_reported(plb):
    blob = plb.OpsBlob([dict(name="NoSuchGate", wires=[0], params=[])])
    out = (C.c_int64 * 4)()
    assert plb.lib().plb200_schedule_stats(C.c_int64(20), 64, blob.ptr(), out) != 0
    assert b"does not exist" in plb.lib().plb200_last_error()


def test_plan_cache_reuses_the_schedule_for_new_angles(plb, monkeypatch):
    """fusion.cu build_schedule: the plan (passes, tile bits, rounds) depends on the items' scheduling signature, not
    on their angles.  A second parameter set of the same circuit must come from the cache (no search) and still give
    the oracle's state; a different circuit must not; PLB200_SCHED_CACHE=0 switches the cache off."""
    import ctypes as C

    import numpy as np
    from conftest import TOL, random_state
    from pennylane_lightning_b200 import circuits
    from test_tile_emulation import CSRC, EMU, emu_apply, oracle_apply
    import subprocess

    res = subprocess.run(["make", "-C", CSRC, "-j8", "emu"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    emu = C.CDLL(EMU)
    emu.plb200_emu_last_error.restype = C.c_char_p
    emu.plb200_emu_plan_hits.restype = C.c_int64
    n = 14
    ops = circuits.random_circuit(n, 6, 4242)
    rng = np.random.default_rng(1)

    def reangle(tape):
        return [dict(o, params=[float(rng.uniform(0, 2 * np.pi)) for _ in o["params"]]) for o in tape]

    for dtype in (np.complex128, np.complex64):
        st = random_state(n, dtype, 1)
        h0 = emu.plb200_emu_plan_hits()
        out, stats_a = emu_apply(emu, plb, n, ops, st)
        np.testing.assert_allclose(out, oracle_apply(n, ops, st), rtol=0, atol=TOL[np.dtype(dtype)])
        h1 = emu.plb200_emu_plan_hits()
        for _ in range(3):
            tape = reangle(ops)
            out, stats_b = emu_apply(emu, plb, n, tape, st)
            np.testing.assert_allclose(out, oracle_apply(n, tape, st), rtol=0, atol=TOL[np.dtype(dtype)])
            assert stats_b == stats_a  # the very same schedule
        h2 = emu.plb200_emu_plan_hits()
        assert h2 - h1 == 3, (h0, h1, h2)
        other = circuits.random_circuit(n, 6, 4243)  # another structure: a miss
        emu_apply(emu, plb, n, other, st)
        assert emu.plb200_emu_plan_hits() == h2
        monkeypatch.setenv("PLB200_SCHED_CACHE", "0")
        out, _ = emu_apply(emu, plb, n, reangle(ops), st)
        assert emu.plb200_emu_plan_hits() == h2
        monkeypatch.delenv("PLB200_SCHED_CACHE")
