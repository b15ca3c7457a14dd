"""CPU check of the routed ("swap-out") pass of the sharded mode: the generated last round stores every amplitude
into the slab and position it has AFTER exchanging k local index bits with k global (rank) bits.  The generated
code (jit_codegen.hpp, Route) runs on host memory through the test-only emulation library for every rank of a
2^g-rank world in turn; the assembled result must equal the numpy oracle's state."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import random_state
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


def _tape(kind, nloc):
    if kind == "random":
        return circuits.random_circuit(nloc, 3, 77)
    if kind == "tail":  # ends with gates the scheduler runs stand-alone: an op-less pass carries the route
        return circuits.random_circuit(nloc, 2, 78) + [circuits.op("IsingXX", [1, 4], [0.3]),
                                                       circuits.op("DoubleExcitation", [0, 2, 5, 7], [0.9])]
    if kind == "one":  # a single gate: nothing to fuse
        return [circuits.op("RX", [3], [0.4])]
    return []  # "empty": a pure exchange


@pytest.mark.parametrize("dtype,g,nloc,lbits,kind", [
    (np.complex128, 1, 13, [9], "random"),           # one bit, inside or outside the tile depending on the schedule
    (np.complex128, 2, 13, [12, 5], "random"),       # two bits
    (np.complex128, 3, 13, [11, 7, 4], "random"),    # three bits: 8-way all-to-all
    (np.complex64, 2, 15, [14, 6], "random"),
    (np.complex128, 2, 13, [10, 3], "tail"),
    (np.complex128, 1, 13, [6], "one"),
    (np.complex128, 2, 13, [12, 8], "empty"),
    # swapped bits OUTSIDE the tile, unsorted: the tile walk is permuted (route_tile) so that consecutive tiles hit
    # every destination; every tile must still be visited exactly once
    (np.complex128, 3, 15, [14, 12, 13], "empty"),
    (np.complex128, 3, 15, [13, 11, 14], "random"),
])
def test_routed_pass_equals_apply_then_swap(emu, plb, dtype, g, nloc, lbits, kind, monkeypatch):
    monkeypatch.setenv("PLB200_EMU_JIT", "1")
    world, n = 1 << g, nloc + g
    k = len(lbits)
    # the swapped global bits: rank bits 0..k-1  <->  local bits lbits[i]
    ops = _tape(kind, nloc)  # the same local tape on every rank (targets all local)
    full = random_state(n, dtype, 5)
    slabs = [full[r << nloc:(r + 1) << nloc].copy() for r in range(world)]  # the tape runs in place on these
    alts = [np.zeros(1 << nloc, dtype=dtype) for _ in range(world)]
    blob = plb.OpsBlob(ops)
    lb = (C.c_int64 * k)(*lbits)
    n_routed = 0
    for r in range(world):
        my_value = r & ((1 << k) - 1)
        dst = (C.c_void_p * (1 << k))()
        for p in range(1 << k):
            peer = (r & ~((1 << k) - 1)) | p
            dst[p] = alts[peer].ctypes.data
        routed = C.c_int(0)
        rc = emu.plb200_emu_apply_ops_route(C.c_int64(nloc), 64 if dtype == np.complex128 else 32, blob.ptr(),
                                            slabs[r].ctypes.data_as(C.c_void_p), C.c_int64(k), lb, C.c_int64(my_value),
                                            dst, C.byref(routed))
        assert rc == 0, emu.plb200_emu_last_error()
        n_routed += routed.value
    assert n_routed == world  # every rank always routes (an op-less pass if its schedule has no final tile pass)
    got = np.concatenate(alts)
    # oracle: apply the local tape to every slab (= the tape on the low nloc qubits of the full state) ...
    ref = np_oracle.StateVector(n, np.complex128)
    ref.set_state(full.astype(np.complex128))
    ref.apply_ops([dict(o, wires=[w + g for w in o["wires"]]) for o in ops])
    st = ref.get_state()
    # ... then exchange index bit (nloc + i) with index bit lbits[i]
    idx = np.arange(1 << n, dtype=np.int64)
    src = idx.copy()
    for i, lbit in enumerate(lbits):
        gb = nloc + i
        a, b = (src >> gb) & 1, (src >> lbit) & 1
        src = (src & ~((1 << gb) | (1 << lbit))) | (b << gb) | (a << lbit)
    want = st[src]
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    np.testing.assert_allclose(got, want, rtol=0, atol=tol)
