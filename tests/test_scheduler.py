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
