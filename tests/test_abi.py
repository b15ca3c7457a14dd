"""CPU tests: libplb200.so loads and exports every symbol include/plb200.h declares; the host-side
argument checks that need no device behave like the reference's."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "plb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(plb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    assert len(syms) >= 45
    for s in ("plb200_sv_apply", "plb200_sv_apply_matrix", "plb200_adjoint_jacobian", "plb200_probs",
              "plb200_expval_pauli_words", "plb200_generate_samples", "plb200_sv_swap_bit_peer", "plb200_sv_swap_bits_peer"):
        assert s in syms


def test_library_exports_every_declared_symbol(plb):
    lib = plb.lib()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_error_channel(plb):
    lib = plb.lib()
    assert b"sm_100a" in lib.plb200_version()
    assert isinstance(lib.plb200_last_error(), bytes)


def test_no_cpu_fallback_without_device(plb):
    from conftest import HAS_GPU

    if HAS_GPU:
        pytest.skip("a device is present")
    with pytest.raises(plb.B200Error):
        plb.StateVector(3)


def test_built_for_sm100a_only():
    so = os.path.join(ROOT, "pennylane-lightning_b200", "lib", "libplb200.so")
    import subprocess

    try:
        out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True, timeout=60).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
