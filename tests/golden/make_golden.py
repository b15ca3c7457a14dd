"""Generates tests/golden/lq_golden.npz by running the UNMODIFIED reference lightning.qubit core
(oracle/_ref/liblq_ref.so, built from /root/reference by oracle/Makefile) on seeded inputs.
Run here (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures travel to the GPU box; the reference sources do not."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CONTROLLED_GATES, GATES, GENERATORS, random_state  # noqa: E402

from oracle import lq_ref as R  # noqa: E402
from pennylane_lightning_b200 import circuits  # noqa: E402

N = 5
out, meta = {}, {"n": N, "gates": [], "ctrl_gates": [], "generators": [], "circuits": []}
rng = np.random.default_rng(2024)


def params_for(name, npar, k):
    if name == "PCPhase":
        return [float(rng.uniform(0, 2 * np.pi)), float(rng.integers(0, 2**k + 1))]
    return [float(x) for x in rng.uniform(0, 2 * np.pi, size=npar)]


for dt, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
    st = random_state(N, dt, 1)
    out[f"init_{tag}"] = st
    for name, (nw, npar) in GATES.items():
        for inv in (False, True):
            k = nw if nw > 0 else 3
            wires = [int(x) for x in rng.permutation(N)[:k]]
            p = params_for(name, npar, k)
            sv = R.StateVector(N, dt)
            sv.set_state(st)
            sv.apply(name, wires, inv, p)
            key = f"gate_{name}_{int(inv)}_{tag}"
            out[key] = sv.get_state()
            meta["gates"].append(dict(key=key, name=name, wires=wires, params=p, inverse=inv, dtype=tag))
    for name in CONTROLLED_GATES:
        nw, npar = GATES[name]
        k = nw if nw > 0 else 2
        perm = [int(x) for x in rng.permutation(N)]
        wires, cw = perm[:k], perm[k:k + 1]
        cv = [bool(rng.integers(0, 2))]
        p = params_for(name, npar, k)
        sv = R.StateVector(N, dt)
        sv.set_state(st)
        sv.apply(name, wires, False, p, cw, cv)
        key = f"cgate_{name}_{tag}"
        out[key] = sv.get_state()
        meta["ctrl_gates"].append(dict(key=key, name=name, wires=wires, params=p, ctrl_wires=cw, ctrl_values=cv,
                                       dtype=tag))
    for name, nw in GENERATORS.items():
        k = nw if nw > 0 else 3
        wires = [int(x) for x in rng.permutation(N)[:k]]
        sv = R.StateVector(N, dt)
        sv.set_state(st)
        scale = sv.apply_generator(name, wires)
        key = f"gen_{name}_{tag}"
        out[key] = sv.get_state()
        meta["generators"].append(dict(key=key, name=name, wires=wires, scale=scale, dtype=tag))

# circuits: config-1 family at reduced size (state, <Z0>, Jacobian), random circuit, QFT, HEA+Hamiltonian
for dt, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
    ops, tp = circuits.strongly_entangling_layers(6, 2, 42)
    sv = R.StateVector(6, dt)
    sv.apply_ops(ops)
    out[f"sel6_state_{tag}"] = sv.get_state()
    ob = R.Observable.named("PauliZ", [0], dtype=dt)
    out[f"sel6_expval_{tag}"] = np.array([sv.expval(ob)])
    out[f"sel6_jac_{tag}"] = sv.adjoint_jacobian([ob], ops, tp)
    ops = circuits.random_circuit(8, 4, 1234)
    sv = R.StateVector(8, dt)
    sv.apply_ops(ops)
    out[f"rand8_state_{tag}"] = sv.get_state()
    out[f"rand8_probs_{tag}"] = sv.probs([5, 1, 3])
    out[f"rand8_samples_{tag}"] = sv.generate_samples(64, seed=37)
    out[f"rand8_samples_w_{tag}"] = sv.generate_samples(64, wires=[6, 0, 2], seed=11)
    ops = circuits.qft(7)
    sv = R.StateVector(7, dt)
    sv.set_basis_state([1, 0, 1, 1, 0, 0, 1], list(range(7)))
    sv.apply_ops(ops)
    out[f"qft7_state_{tag}"] = sv.get_state()
    ops, tp = circuits.hardware_efficient_ansatz(6, 40, 99)
    co, words, wires = circuits.pauli_hamiltonian(6, 12, 99)
    ham = circuits.hamiltonian_observable(R, co, words, wires, dtype=dt)
    sv = R.StateVector(6, dt)
    sv.apply_ops(ops)
    out[f"hea6_expval_{tag}"] = np.array([sv.expval(ham)])
    out[f"hea6_var_{tag}"] = np.array([sv.var(ham)])
    out[f"hea6_jac_{tag}"] = sv.adjoint_jacobian([ham], ops, tp)

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lq_golden.npz"), **out)
json.dump(meta, open(os.path.join(ROOT, "tests", "golden", "lq_golden_meta.json"), "w"))
print("wrote", len(out), "arrays")
