"""Engine-agnostic checks against tests/golden/ (fixtures generated from the unmodified reference
by tests/golden/make_golden.py + literals transcribed from the reference's C++ tests).  `mod` is any
module exposing StateVector / Observable with the shared call surface: oracle.np_oracle,
oracle.lq_ref, or pennylane_lightning_b200."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "lq_golden.npz"))
META = json.load(open(os.path.join(HERE, "golden", "lq_golden_meta.json")))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
DT = {"c128": np.complex128, "c64": np.complex64}
TOL = {"c128": 1e-12, "c64": 1e-5}


def _obs_kw(mod, dt):
    return {"dtype": dt} if mod.__name__.endswith("lq_ref") else {}


def check_gates(mod, tag):
    n = META["n"]
    st = G[f"init_{tag}"]
    for e in META["gates"]:
        if e["dtype"] != tag:
            continue
        sv = mod.StateVector(n, DT[tag])
        sv.set_state(st)
        sv.apply(e["name"], e["wires"], e["inverse"], e["params"])
        np.testing.assert_allclose(sv.get_state(), G[e["key"]], rtol=0, atol=TOL[tag], err_msg=e["key"])
    for e in META["ctrl_gates"]:
        if e["dtype"] != tag:
            continue
        sv = mod.StateVector(n, DT[tag])
        sv.set_state(st)
        sv.apply(e["name"], e["wires"], False, e["params"], e["ctrl_wires"], e["ctrl_values"])
        np.testing.assert_allclose(sv.get_state(), G[e["key"]], rtol=0, atol=TOL[tag], err_msg=e["key"])


def check_generators(mod, tag):
    n = META["n"]
    st = G[f"init_{tag}"]
    for e in META["generators"]:
        if e["dtype"] != tag:
            continue
        sv = mod.StateVector(n, DT[tag])
        sv.set_state(st)
        scale = sv.apply_generator(e["name"], e["wires"])
        assert scale == e["scale"], e["key"]
        np.testing.assert_allclose(sv.get_state(), G[e["key"]], rtol=0, atol=TOL[tag], err_msg=e["key"])


def check_circuits(mod, tag, samples=True):
    from pennylane_lightning_b200 import circuits

    dt, tol = DT[tag], TOL[tag]
    kw = _obs_kw(mod, dt)
    ops, tp = circuits.strongly_entangling_layers(6, 2, 42)
    sv = mod.StateVector(6, dt)
    sv.apply_ops(ops)
    np.testing.assert_allclose(sv.get_state(), G[f"sel6_state_{tag}"], rtol=0, atol=tol)
    ob = mod.Observable.named("PauliZ", [0], **kw)
    np.testing.assert_allclose(sv.expval(ob), G[f"sel6_expval_{tag}"][0], rtol=0, atol=tol)
    np.testing.assert_allclose(sv.adjoint_jacobian([ob], ops, tp), G[f"sel6_jac_{tag}"], rtol=0, atol=10 * tol)

    ops = circuits.random_circuit(8, 4, 1234)
    sv = mod.StateVector(8, dt)
    sv.apply_ops(ops)
    np.testing.assert_allclose(sv.get_state(), G[f"rand8_state_{tag}"], rtol=0, atol=tol)
    np.testing.assert_allclose(sv.probs([5, 1, 3]), G[f"rand8_probs_{tag}"], rtol=0, atol=tol)
    if samples:  # bit-exact under the shared seed
        np.testing.assert_array_equal(sv.generate_samples(64, seed=37), G[f"rand8_samples_{tag}"])
        np.testing.assert_array_equal(sv.generate_samples(64, wires=[6, 0, 2], seed=11), G[f"rand8_samples_w_{tag}"])

    sv = mod.StateVector(7, dt)
    sv.set_basis_state([1, 0, 1, 1, 0, 0, 1], list(range(7)))
    sv.apply_ops(circuits.qft(7))
    np.testing.assert_allclose(sv.get_state(), G[f"qft7_state_{tag}"], rtol=0, atol=tol)

    ops, tp = circuits.hardware_efficient_ansatz(6, 40, 99)
    co, words, wires = circuits.pauli_hamiltonian(6, 12, 99)
    ham = circuits.hamiltonian_observable(mod, co, words, wires, **kw)
    sv = mod.StateVector(6, dt)
    sv.apply_ops(ops)
    np.testing.assert_allclose(sv.expval(ham), G[f"hea6_expval_{tag}"][0], rtol=0, atol=20 * tol)
    np.testing.assert_allclose(sv.var(ham), G[f"hea6_var_{tag}"][0], rtol=0, atol=100 * tol)
    np.testing.assert_allclose(sv.adjoint_jacobian([ham], ops, tp), G[f"hea6_jac_{tag}"], rtol=0, atol=50 * tol)


def nontrivial_state(mod, dt=np.complex128, n=3):
    sv = mod.StateVector(n, dt)
    ph = 0.7
    for k in range(n):
        sv.apply("RX", [k], False, [ph])
        sv.apply("RY", [k], False, [ph])
        ph -= 0.2
    return sv


def check_reference_kats(mod, dt=np.complex128):
    """Literals from Test_MeasurementsBase.cpp / Test_AdjointJacobian.cpp (see reference_kats.json)."""
    kw = _obs_kw(mod, dt)
    sv = nontrivial_state(mod, dt)
    for wires, expected in KATS["probs"]["cases"]:
        np.testing.assert_allclose(sv.probs(wires), expected, rtol=0, atol=KATS["probs"]["tol"])
    for name in ("PauliX", "PauliY", "PauliZ"):
        for w in range(3):
            assert abs(sv.expval_named(name, [w]) - KATS["expval"][name][w]) < KATS["expval"]["tol"]
            assert abs(sv.expval(mod.Observable.named(name, [w], **kw)) - KATS["expval"][name][w]) < 1e-6
            assert abs(sv.var_named(name, [w]) - KATS["var"][name][w]) < KATS["var"]["tol"]
    k = KATS["adjoint_rx3_zzz"]
    ops = [dict(name="RX", wires=[i], params=[k["params"][i]]) for i in range(3)]
    obs = mod.Observable.tensor([mod.Observable.named("PauliZ", [i], **kw) for i in range(3)])
    psi = mod.StateVector(3, dt)
    jac = psi.adjoint_jacobian([obs], ops, [0, 1, 2], apply_ops=True)
    np.testing.assert_allclose(jac.ravel(), k["jacobian"], rtol=0, atol=k["tol"])
