"""GPU engine against the committed golden fixtures (generated from the unmodified reference) and
the literals transcribed from the reference's own C++ tests."""
import numpy as np
import pytest

import golden_checks as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_gates(plb, tag):
    gc.check_gates(plb, tag)


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_generators(plb, tag):
    gc.check_generators(plb, tag)


@pytest.mark.parametrize("tag", ["c128", "c64"])
def test_circuits_measurements_adjoint_samples(plb, tag):
    gc.check_circuits(plb, tag)


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
def test_reference_known_answers(plb, dt):
    gc.check_reference_kats(plb, dt)
