"""Synthetic workloads of BASELINE.json's configs, as plain op lists
(dicts with name / wires / params / inverse / ctrl_wires / ctrl_values) that both the engine and
the reference oracle consume.  RNG = numpy.random.default_rng(seed); angles U[0, 2pi)."""
from __future__ import annotations

import numpy as np


def op(name, wires, params=(), inverse=False, ctrl_wires=(), ctrl_values=()):
    return dict(name=name, wires=[int(w) for w in wires], params=[float(p) for p in params], inverse=bool(inverse),
                ctrl_wires=[int(w) for w in ctrl_wires], ctrl_values=[bool(v) for v in ctrl_values])


def random_circuit(n, depth=20, seed=1234):
    """Config 2/4: per layer one of {RX,RY,RZ}(theta) on every qubit, then a random perfect matching
    with CNOT or CRZ(theta) (p=1/2, random orientation).  n=30, depth=20 -> 900 gates."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(depth):
        kinds = rng.integers(0, 3, size=n)
        thetas = rng.uniform(0, 2 * np.pi, size=n)
        for q in range(n):
            ops.append(op(("RX", "RY", "RZ")[kinds[q]], [q], [thetas[q]]))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            a, b = int(perm[i]), int(perm[i + 1])
            if rng.integers(0, 2):
                a, b = b, a
            if rng.integers(0, 2):
                ops.append(op("CNOT", [a, b]))
            else:
                ops.append(op("CRZ", [a, b], [rng.uniform(0, 2 * np.pi)]))
    return ops


def strongly_entangling_layers(n=20, layers=4, seed=42):
    """Config 1: StronglyEntanglingLayers with Rot decomposed RZ.RY.RZ as the serializer does
    (lightning_base/_serialize.py:526-529); CNOT(w,(w+r)%n), r=(l mod (n-1))+1."""
    rng = np.random.default_rng(seed)
    weights = rng.uniform(0, 2 * np.pi, size=(layers, n, 3))
    ops = []
    for l in range(layers):
        for w in range(n):
            phi, theta, omega = weights[l, w]
            ops += [op("RZ", [w], [phi]), op("RY", [w], [theta]), op("RZ", [w], [omega])]
        r = (l % (n - 1)) + 1
        for w in range(n):
            ops.append(op("CNOT", [w, (w + r) % n]))
    trainable = list(range(3 * n * layers))
    return ops, trainable


def qft(n):
    """Config 3: PennyLane's QFT decomposition: H(i); CPS(pi/2^(j-i), [j,i]) for j>i; final SWAPs."""
    ops = []
    for i in range(n):
        ops.append(op("Hadamard", [i]))
        for j in range(i + 1, n):
            ops.append(op("ControlledPhaseShift", [j, i], [np.pi / 2 ** (j - i)]))
    for i in range(n // 2):
        ops.append(op("SWAP", [i, n - 1 - i]))
    return ops


def hardware_efficient_ansatz(n=24, n_params=1000, seed=99):
    """Config 5: layers of RY on all wires, RZ on all wires, CNOT ring; truncated to n_params
    trainable parameters (n=24: 20 full layers + RY on all + RZ on wires 0..15)."""
    rng = np.random.default_rng(seed)
    ops, count = [], 0
    while count < n_params:
        for g in ("RY", "RZ"):
            for w in range(n):
                if count < n_params:
                    ops.append(op(g, [w], [rng.uniform(0, 2 * np.pi)]))
                    count += 1
        if count < n_params:
            for w in range(n):
                ops.append(op("CNOT", [w, (w + 1) % n]))
    return ops, list(range(n_params))


def pauli_hamiltonian(n=24, terms=100, seed=99):
    """Config 5: `terms` Pauli words of weight U{1..4} on distinct wires, coeff ~ N(0,1).
    Returns (coeffs, words, wires)."""
    rng = np.random.default_rng(seed + 1)
    coeffs, words, wires = [], [], []
    for _ in range(terms):
        w = int(rng.integers(1, 5))
        ws = [int(x) for x in rng.permutation(n)[:w]]
        word = "".join("XYZ"[int(rng.integers(0, 3))] for _ in range(w))
        coeffs.append(float(rng.normal()))
        words.append(word)
        wires.append(ws)
    return coeffs, words, wires


_PAULI_NAME = {"X": "PauliX", "Y": "PauliY", "Z": "PauliZ", "I": "Identity"}


def hamiltonian_observable(mod, coeffs, words, wires, **kw):
    """Build Hamiltonian(coeffs, [TensorProdObs(NamedObs...)]) with module `mod`'s Observable class
    (pennylane_lightning_b200 or oracle.lq_ref)."""
    terms = []
    for word, ws in zip(words, wires):
        named = [mod.Observable.named(_PAULI_NAME[c], [w], **kw) for c, w in zip(word, ws)]
        terms.append(named[0] if len(named) == 1 else mod.Observable.tensor(named))
    return mod.Observable.hamiltonian(coeffs, terms)


def algorithmic_bytes(o, n, elem_bytes=16):
    """SURVEY.md §8(d) byte table for one un-fused gate on an n-qubit state."""
    S = (1 << n) * elem_bytes
    name = o["name"]
    c = len(o.get("ctrl_wires", ()))
    if name in ("CNOT", "CY", "CRX", "CRY", "CRZ", "CRot"):
        b = S
    elif name == "Toffoli":
        b = S / 2
    elif name in ("PhaseShift", "PauliZ", "S", "T", "SWAP", "SingleExcitation", "IsingXY", "PSWAP"):
        b = S
    elif name in ("CZ", "ControlledPhaseShift", "CSWAP"):
        b = S / 2
    elif name in ("DoubleExcitation",):
        b = S / 4
    elif name == "Identity":
        b = 0
    else:
        b = 2 * S
    return b / (2 ** c)
