"""TEST INFRASTRUCTURE — numpy restatement of lightning.qubit's state-vector hot path.

A CPU "port" oracle for small qubit counts (n <~ 20): gate application, generators, measurements,
alias-method sampling and the adjoint Jacobian, restated from the reference sources cited per
function (paths relative to /root/reference/pennylane_lightning/core/).  It is pinned against
(a) oracle/_ref (the reference itself, compiled here) and (b) the committed golden vectors in
tests/golden/ — see tests/test_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it; the product never does.

Conventions (GateImplementationsLM.hpp:690-699, StateVectorBase.hpp:195): wire 0 is the most
significant index bit; matrices are row-major with wires[0] the most significant matrix bit.
"""
from __future__ import annotations

import numpy as np

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
PAULI = {"I": I2, "X": X, "Y": Y, "Z": Z}
NAMED_OBS = {"Identity": I2, "PauliX": X, "PauliY": Y, "PauliZ": Z, "Hadamard": H}


def _ctrl(u):
    d = u.shape[0]
    m = np.eye(2 * d, dtype=complex)
    m[d:, d:] = u
    return m


def _rot(phi, theta, omega):
    # gates/Gates.hpp:390-410 getRot = RZ(omega) RY(theta) RZ(phi)
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    return np.array([[np.exp(-0.5j * (phi + omega)) * c, -np.exp(0.5j * (phi - omega)) * s],
                     [np.exp(-0.5j * (phi - omega)) * s, np.exp(0.5j * (phi + omega)) * c]])


def gate_matrix(name, params=(), n_wires=None):
    """Dense matrices of gates/Gates.hpp:38-1385 (getPauliX ... getPSWAP, getMultiRZ, getPCPhase)."""
    p = [float(x) for x in params]
    c = np.cos(p[0] / 2) if p else None
    s = np.sin(p[0] / 2) if p else None
    if name == "Identity":
        return np.eye(2 ** (n_wires or 1), dtype=complex)
    if name in ("PauliX", "PauliY", "PauliZ", "Hadamard"):
        return NAMED_OBS[name]
    if name == "S":
        return np.diag([1, 1j])
    if name == "T":
        return np.diag([1, np.exp(0.25j * np.pi)])
    if name == "SX":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
    if name == "PhaseShift":
        return np.diag([1, np.exp(1j * p[0])])
    if name == "RX":  # Gates.hpp:329
        return np.array([[c, -1j * s], [-1j * s, c]])
    if name == "RY":  # Gates.hpp:346
        return np.array([[c, -s], [s, c]], dtype=complex)
    if name == "RZ":  # Gates.hpp:363
        return np.diag([np.exp(-0.5j * p[0]), np.exp(0.5j * p[0])])
    if name == "Rot":
        return _rot(*p)
    if name == "CNOT":
        return _ctrl(X)
    if name == "CY":
        return _ctrl(Y)
    if name == "CZ":
        return _ctrl(Z)
    if name == "SWAP":
        return np.eye(4, dtype=complex)[[0, 2, 1, 3]]
    if name == "IsingXX":
        return c * np.eye(4) - 1j * s * np.kron(X, X)
    if name == "IsingYY":
        return c * np.eye(4) - 1j * s * np.kron(Y, Y)
    if name == "IsingZZ":
        return np.diag(np.exp(-0.5j * p[0] * np.array([1, -1, -1, 1])))
    if name == "IsingXY":
        m = np.eye(4, dtype=complex)
        m[1:3, 1:3] = [[c, 1j * s], [1j * s, c]]
        return m
    if name == "ControlledPhaseShift":
        return np.diag([1, 1, 1, np.exp(1j * p[0])])
    if name in ("CRX", "CRY", "CRZ"):
        return _ctrl(gate_matrix(name[1:], p))
    if name == "CRot":
        return _ctrl(_rot(*p))
    if name in ("SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus"):
        e = {"SingleExcitation": 1.0, "SingleExcitationMinus": np.exp(-0.5j * p[0]),
             "SingleExcitationPlus": np.exp(0.5j * p[0])}[name]
        m = np.diag([e, 0, 0, e]).astype(complex)
        m[1:3, 1:3] = [[c, -s], [s, c]]
        return m
    if name == "PSWAP":
        m = np.zeros((4, 4), dtype=complex)
        m[0, 0] = m[3, 3] = 1
        m[1, 2] = m[2, 1] = np.exp(1j * p[0])
        return m
    if name == "Toffoli":
        return _ctrl(_ctrl(X))
    if name == "CSWAP":
        return _ctrl(gate_matrix("SWAP"))
    if name in ("DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus"):
        e = {"DoubleExcitation": 1.0, "DoubleExcitationMinus": np.exp(-0.5j * p[0]),
             "DoubleExcitationPlus": np.exp(0.5j * p[0])}[name]
        m = np.eye(16, dtype=complex) * e
        m[3, 3] = m[12, 12] = c
        m[3, 12], m[12, 3] = -s, s
        return m
    if name == "MultiRZ":  # GateImplementationsLM.hpp:1988-2010
        k = n_wires
        par = np.array([bin(i).count("1") & 1 for i in range(2 ** k)])
        return np.diag(np.exp(-0.5j * p[0] * (1 - 2 * par)))
    if name == "GlobalPhase":  # GateImplementationsLM.hpp:2044-2111
        return np.exp(-1j * p[0]) * np.eye(2 ** (n_wires or 1), dtype=complex)
    if name == "PCPhase":  # GateImplementationsLM.hpp:2113-2162
        k = n_wires
        d = int(round(p[1]))
        return np.diag([np.exp(1j * p[0])] * d + [np.exp(-1j * p[0])] * (2 ** k - d))
    raise ValueError(f"unknown gate {name}")


def generator_matrix(name, n_wires=None):
    """(G, scale) with the kernels' semantics, GateImplementationsLM.hpp:2181-2952."""
    P1 = np.diag([0, 1]).astype(complex)

    def pair(k, a, b, mab, mba, rest):
        m = np.diag(np.full(2 ** k, rest, dtype=complex))
        m[a, a] = m[b, b] = 0
        m[a, b], m[b, a] = mab, mba
        return m

    table = {
        "PhaseShift": (P1, 1.0), "RX": (X, -0.5), "RY": (Y, -0.5), "RZ": (Z, -0.5),
        "IsingXX": (np.kron(X, X), -0.5), "IsingYY": (np.kron(Y, Y), -0.5), "IsingZZ": (np.kron(Z, Z), -0.5),
        "IsingXY": (pair(2, 1, 2, 1, 1, 0), 0.5), "PSWAP": (pair(2, 1, 2, 1, 1, 0), 1.0),
        "CRX": (np.kron(P1, X), -0.5), "CRY": (np.kron(P1, Y), -0.5), "CRZ": (np.kron(P1, Z), -0.5),
        "ControlledPhaseShift": (np.kron(P1, P1), 1.0),
        "SingleExcitation": (pair(2, 1, 2, -1j, 1j, 0), -0.5),
        "SingleExcitationMinus": (pair(2, 1, 2, -1j, 1j, 1), -0.5),
        "SingleExcitationPlus": (pair(2, 1, 2, -1j, 1j, -1), -0.5),
        "DoubleExcitation": (pair(4, 3, 12, -1j, 1j, 0), -0.5),
        "DoubleExcitationMinus": (pair(4, 3, 12, -1j, 1j, 1), -0.5),
        "DoubleExcitationPlus": (pair(4, 3, 12, 1j, -1j, 1), 0.5),
    }
    if name in table:
        return table[name]
    if name == "MultiRZ":
        par = np.array([bin(i).count("1") & 1 for i in range(2 ** n_wires)])
        return np.diag((1 - 2 * par).astype(complex)), -0.5
    if name == "GlobalPhase":
        return np.eye(2 ** (n_wires or 1), dtype=complex), -1.0
    raise ValueError(f"unknown generator {name}")


class StateVector:
    """numpy mirror of the call surface used by the parity tests."""

    def __init__(self, num_qubits, dtype=np.complex128):
        self.n = self.num_qubits = num_qubits
        self.dtype = np.dtype(dtype)
        self.state = np.zeros(2 ** num_qubits, dtype=self.dtype)
        self.state[0] = 1

    # -- data ------------------------------------------------------------------------------
    def get_state(self):
        return self.state.copy()

    def set_state(self, v):
        self.state = np.asarray(v, dtype=self.dtype).copy()

    def reset(self):
        self.state[:] = 0
        self.state[0] = 1

    def set_basis_state(self, bits, wires):  # StateVectorLQubit.hpp:943-953
        idx = 0
        for b, w in zip(bits, wires):
            idx |= int(b) << (self.n - 1 - w)
        self.state[:] = 0
        self.state[idx] = 1

    def set_state_vector(self, values, wires):  # StateVectorLQubit.hpp:1005-1036
        t = np.zeros((2,) * self.n, dtype=self.dtype)
        idx = [0] * self.n
        v = np.asarray(values, dtype=self.dtype).reshape((2,) * len(wires))
        sl = [0] * self.n
        for w in wires:
            sl[w] = slice(None)
        order = np.argsort(wires)
        t[tuple(sl)] = np.transpose(v, order)
        self.state = t.reshape(-1)

    def collapse(self, wire, branch):  # StateVectorLQubit.hpp:881-903
        t = self.state.reshape((2,) * self.n)
        sl = [slice(None)] * self.n
        sl[wire] = 0 if branch else 1
        t[tuple(sl)] = 0
        self.normalize()

    def normalize(self):
        self.state = (self.state / np.linalg.norm(self.state)).astype(self.dtype)

    # -- gates -----------------------------------------------------------------------------
    def _apply_dense(self, m, wires, ctrl_wires=(), ctrl_values=()):
        """applyNCN semantics (GateImplementationsLM.hpp:407-440): dense matrix on `wires`
        inside the subspace selected by the controls."""
        n, k = self.n, len(wires)
        t = self.state.reshape((2,) * n)
        sl = [slice(None)] * n
        for w, v in zip(ctrl_wires, ctrl_values):
            sl[w] = 1 if v else 0
        sub = t[tuple(sl)]
        rem = [w for w in range(n) if w not in ctrl_wires]
        axes = [rem.index(w) for w in wires]
        mt = np.asarray(m, dtype=complex).reshape((2,) * (2 * k))
        out = np.tensordot(mt, sub, axes=(list(range(k, 2 * k)), axes))
        out = np.moveaxis(out, list(range(k)), axes)
        t = t.copy()
        t[tuple(sl)] = out
        self.state = t.reshape(-1).astype(self.dtype)

    def apply(self, name, wires, inverse=False, params=(), ctrl_wires=(), ctrl_values=()):
        if name == "Identity":
            return
        m = gate_matrix(name, params, len(wires))
        if inverse:
            m = m.conj().T
        self._apply_dense(m, list(wires), list(ctrl_wires), list(ctrl_values))

    def apply_matrix(self, matrix, wires, inverse=False, ctrl_wires=(), ctrl_values=()):
        m = np.asarray(matrix, dtype=complex).reshape(2 ** len(wires), -1)
        if inverse:  # GateImplementationsLM.hpp:287-293
            m = m.conj().T
        self._apply_dense(m, list(wires), list(ctrl_wires), list(ctrl_values))

    def apply_pauli_rot(self, wires, inverse, theta, word):  # GateImplementationsLM.hpp:575-629
        m = np.array([[1.0]], dtype=complex)
        for ch in word:
            m = np.kron(m, PAULI[ch])
        th = -theta if inverse else theta
        u = np.cos(th / 2) * np.eye(m.shape[0]) - 1j * np.sin(th / 2) * m
        self._apply_dense(u, list(wires))

    def apply_generator(self, name, wires, adj=False, ctrl_wires=(), ctrl_values=()):
        g, scale = generator_matrix(name, len(wires))
        if ctrl_wires:  # applyNCGenerator*: projector (x) G
            t = self.state.reshape((2,) * self.n).copy()
            keep = np.zeros_like(t)
            sl = [slice(None)] * self.n
            for w, v in zip(ctrl_wires, ctrl_values):
                sl[w] = 1 if v else 0
            keep[tuple(sl)] = t[tuple(sl)]
            self.state = keep.reshape(-1)
        self._apply_dense(g, list(wires), list(ctrl_wires), list(ctrl_values))
        return scale

    def apply_ops(self, ops, fuse=False):
        for o in ops:
            if o.get("matrix") is not None and len(np.ravel(o["matrix"])) and o["name"] not in _KNOWN:
                self.apply_matrix(o["matrix"], o["wires"], o.get("inverse", False), o.get("ctrl_wires", ()),
                                  o.get("ctrl_values", ()))
            else:
                self.apply(o["name"], o["wires"], o.get("inverse", False), o.get("params", ()),
                           o.get("ctrl_wires", ()), o.get("ctrl_values", ()))

    # -- measurements ----------------------------------------------------------------------
    def probs(self, wires=None):  # MeasurementsLQubit.hpp:90-163
        p = (self.state.real.astype(self.state.real.dtype) ** 2 + self.state.imag ** 2)
        if wires is None:
            return p.astype(np.float64)
        t = p.reshape((2,) * self.n)
        rest = tuple(w for w in range(self.n) if w not in wires)
        t = t.sum(axis=rest) if rest else t
        kept = sorted(wires)
        t = np.transpose(t, [kept.index(w) for w in wires])
        return t.reshape(-1).astype(np.float64)

    def expval_matrix(self, matrix, wires):  # MeasurementsLQubit.hpp:213-238
        tmp = StateVector(self.n, np.complex128)
        tmp.state = self.state.astype(np.complex128)
        tmp._apply_dense(np.asarray(matrix).reshape(2 ** len(wires), -1), list(wires))
        return float(np.real(np.vdot(self.state.astype(np.complex128), tmp.state)))

    def expval_named(self, name, wires):  # MeasurementsLQubit.hpp:247-305, ExpValFunc.hpp:69-95
        return self.expval_matrix(NAMED_OBS[name], wires)

    def var_matrix(self, matrix, wires):
        tmp = StateVector(self.n, np.complex128)
        tmp.state = self.state.astype(np.complex128)
        tmp._apply_dense(np.asarray(matrix).reshape(2 ** len(wires), -1), list(wires))
        e = np.real(np.vdot(self.state.astype(np.complex128), tmp.state))
        return float(np.real(np.vdot(tmp.state, tmp.state)) - e * e)

    def var_named(self, name, wires):
        return self.var_matrix(NAMED_OBS[name], wires)

    def expval_pauli_word(self, word, wires):
        tmp = StateVector(self.n, np.complex128)
        tmp.state = self.state.astype(np.complex128)
        for ch, w in zip(word, wires):
            tmp._apply_dense(PAULI[ch], [w])
        return float(np.real(np.vdot(self.state.astype(np.complex128), tmp.state)))

    def apply_observable(self, obs):
        obs.apply(self)

    def expval(self, obs):  # MeasurementsLQubit.hpp:372-394
        tmp = StateVector(self.n, np.complex128)
        tmp.state = self.state.astype(np.complex128)
        obs.apply(tmp)
        return float(np.real(np.vdot(tmp.state, self.state.astype(np.complex128))))

    def var(self, obs):  # MeasurementsLQubit.hpp:431-457
        tmp = StateVector(self.n, np.complex128)
        tmp.state = self.state.astype(np.complex128)
        obs.apply(tmp)
        e = np.real(np.vdot(self.state.astype(np.complex128), tmp.state))
        return float(np.real(np.vdot(tmp.state, tmp.state)) - e * e)

    def generate_samples(self, shots, wires=None, seed=0):
        """MeasurementsLQubit.hpp:646-679 + DiscreteRandomVariable (MeasurementKernels.hpp:308-381):
        alias method driven by std::mt19937(seed) and std::uniform_real_distribution<PrecisionT>."""
        wires = list(range(self.n)) if wires is None else list(wires)
        fp = np.float64 if self.dtype == np.complex128 else np.float32
        probs = self.probs(wires if len(wires) != self.n or wires != list(range(self.n)) else None).astype(fp)
        n = len(probs)
        first = np.zeros(n, dtype=np.float64)
        second = np.full(n, -1, dtype=np.int64)
        under, over = [], []
        for i in range(n):
            first[i] = float(fp(n) * probs[i])
            (under if first[i] < 1.0 else over).append(i)
        while under and over:
            i, j = over.pop(), under.pop()
            second[j] = i
            first[i] += first[j] - 1.0
            (under if first[i] < 1.0 else over).append(i)
        raw = _mt19937_raw(seed, shots * (4 if fp == np.float64 else 2))
        u = _uniform_real(raw, fp)
        out = np.zeros((shots, len(wires)), dtype=np.uint64)
        for s in range(shots):
            idx = int(fp(u[2 * s]) * fp(n)) if fp == np.float32 else int(u[2 * s] * n)
            if u[2 * s + 1] >= first[idx] and second[idx] >= 0:
                idx = int(second[idx])
            for j in range(len(wires)):
                out[s, len(wires) - 1 - j] = (idx >> j) & 1
        return out

    # -- adjoint Jacobian ------------------------------------------------------------------
    def adjoint_jacobian(self, observables, ops, trainable, apply_ops=False):
        """AdjointJacobianLQubit.hpp:347-491 (loop structure, scaling, -2*s*Im<H lambda|mu>)."""
        lam = StateVector(self.n, np.complex128)
        lam.state = self.state.astype(np.complex128)
        if apply_ops:
            lam.apply_ops(ops)
        hl = []
        for ob in observables:
            h = StateVector(self.n, np.complex128)
            h.state = lam.state.copy()
            ob.apply(h)
            hl.append(h)
        tp = list(trainable)
        jac = np.zeros((len(observables), len(tp)))
        n_par_ops = sum(1 for o in ops if len(o.get("params", ())))
        cur = n_par_ops - 1
        t = len(tp) - 1
        for o in reversed(ops):
            if len(o.get("params", ())) > 1:
                raise RuntimeError("The operation is not supported using the adjoint differentiation method")
            if o["name"] in ("StatePrep", "BasisState"):
                continue
            if t < 0:
                break
            inv = o.get("inverse", False)
            cw, cv = o.get("ctrl_wires", ()), o.get("ctrl_values", ())
            if len(o.get("params", ())):
                if cur == tp[t]:
                    mu = StateVector(self.n, np.complex128)
                    mu.state = lam.state.copy()
                    s = mu.apply_generator(o["name"], o["wires"], not inv, cw, cv) * (-1 if inv else 1)
                    for i, h in enumerate(hl):
                        jac[i, t] = -2 * s * np.imag(np.vdot(h.state, mu.state))
                    t -= 1
                cur -= 1
            for sv in [lam] + hl:
                if o.get("matrix") is not None and len(np.ravel(o["matrix"])) and o["name"] not in _KNOWN:
                    sv.apply_matrix(o["matrix"], o["wires"], not inv, cw, cv)
                else:
                    sv.apply(o["name"], o["wires"], not inv, o.get("params", ()), cw, cv)
        return jac


_KNOWN = {"Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T", "PhaseShift", "RX", "RY", "RZ",
          "Rot", "CNOT", "CY", "CZ", "SWAP", "IsingXX", "IsingXY", "IsingYY", "IsingZZ", "ControlledPhaseShift",
          "CRX", "CRY", "CRZ", "CRot", "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus",
          "PSWAP", "Toffoli", "CSWAP", "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus",
          "MultiRZ", "GlobalPhase", "PCPhase"}


def _mt19937_raw(seed, count):
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed))  # init_genrand(seed) == std::mt19937(seed)
    return bg.random_raw(count).astype(np.uint64)


def _uniform_real(raw, fp):
    """libstdc++ std::generate_canonical for mt19937 (32-bit words): double = 2 words, float = 1."""
    if fp == np.float64:
        lo, hi = raw[0::2].astype(np.float64), raw[1::2].astype(np.float64)
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        r[r >= 1.0] = np.nextafter(1.0, 0.0)
        return r
    r = raw.astype(np.float32) / np.float32(4294967296.0)
    r[r >= 1.0] = np.nextafter(np.float32(1.0), np.float32(0.0))
    return r


class Observable:
    """NamedObs / HermitianObs / TensorProdObs / Hamiltonian (observables/Observables.hpp:128-584)."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    @classmethod
    def named(cls, name, wires, dtype=None):
        return cls("named", name=name, wires=list(wires))

    @classmethod
    def hermitian(cls, matrix, wires, dtype=None):
        return cls("hermitian", matrix=np.asarray(matrix, dtype=complex), wires=list(wires))

    @classmethod
    def tensor(cls, terms):
        return cls("tensor", terms=list(terms))

    @classmethod
    def hamiltonian(cls, coeffs, terms):
        return cls("hamiltonian", coeffs=list(coeffs), terms=list(terms))

    def apply(self, sv):
        if self.kind == "named":
            sv.apply(self.name, self.wires)
        elif self.kind == "hermitian":
            sv._apply_dense(self.matrix.reshape(2 ** len(self.wires), -1), self.wires)
        elif self.kind == "tensor":
            for t in self.terms:
                t.apply(sv)
        else:  # ObservablesLQubit.hpp:156-199
            acc = np.zeros_like(sv.state)
            orig = sv.state.copy()
            for c, t in zip(self.coeffs, self.terms):
                sv.state = orig.copy()
                t.apply(sv)
                acc = acc + c * sv.state
            sv.state = acc
