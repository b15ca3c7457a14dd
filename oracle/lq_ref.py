"""TEST INFRASTRUCTURE — ctypes view of oracle/_ref/liblq_ref.so (the UNMODIFIED reference
lightning.qubit core, built by oracle/Makefile from /root/reference).  Interface mirrors
pennylane_lightning_b200._capi.StateVector so parity tests can drive both with the same calls.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liblq_ref.so")
_lib = None

_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


class _OpsT(C.Structure):
    _fields_ = [
        ("n_ops", C.c_int64), ("names", C.POINTER(C.c_char_p)), ("wires", _i64p), ("wires_off", _i64p),
        ("ctrl_wires", _i64p), ("ctrl_off", _i64p), ("ctrl_values", _u8p), ("params", _f64p),
        ("params_off", _i64p), ("inverses", _u8p), ("mats", _f64p), ("mats_off", _i64p),
    ]


class RefError(RuntimeError):
    pass


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build(ref="/root/reference"):
    """Compile the reference LQ core (only possible where /root/reference exists)."""
    if not os.path.isdir(ref):
        return None
    res = subprocess.run(["make", "-C", _HERE, "-j8", f"REF={ref}"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building liblq_ref.so failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RefError(f"{LIB_PATH} missing (build it where /root/reference exists: make -C oracle)")
        _lib = C.CDLL(LIB_PATH)
        _lib.lqref_last_error.restype = C.c_char_p
        for sfx in ("c64", "c128"):
            for f in ("sv_create", "sv_data", "obs_named", "obs_hermitian", "obs_tensor", "obs_hamiltonian"):
                getattr(_lib, f"lqref_{f}_{sfx}").restype = C.c_void_p
            getattr(_lib, f"lqref_sv_length_{sfx}").restype = C.c_int64
    return _lib


def _check(rc):
    if rc != 0:
        raise RefError(lib().lqref_last_error().decode())


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def _u8(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=bool).astype(np.uint8))
    return a, a.ctypes.data_as(_u8p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


def _c128(a):
    a = np.ascontiguousarray(a, dtype=np.complex128).ravel()
    return a, a.ctypes.data_as(_f64p)


def num_threads():
    return lib().lqref_num_threads()


def set_num_threads(n):
    lib().lqref_set_num_threads(int(n))


class OpsBlob:
    def __init__(self, ops):
        names, wires, woff, cw, coff, cv, params, poff, inv, mats, moff = [], [], [0], [], [0], [], [], [0], [], [], [0]
        for op in ops:
            names.append(op["name"].encode())
            wires += list(op["wires"])
            woff.append(len(wires))
            cw += list(op.get("ctrl_wires", ()))
            cv += [bool(v) for v in op.get("ctrl_values", ())]
            coff.append(len(cw))
            params += [float(p) for p in op.get("params", ())]
            poff.append(len(params))
            inv.append(bool(op.get("inverse", False)))
            m = op.get("matrix", None)
            if m is not None and len(np.asarray(m).ravel()) > 0:
                mats += list(np.asarray(m, dtype=np.complex128).ravel())
            moff.append(len(mats))
        self.n = len(ops)
        self._names = (C.c_char_p * max(self.n, 1))(*names)
        self._keep = [_i64(wires), _i64(woff), _i64(cw), _i64(coff), _u8(cv), _f64(params), _i64(poff), _u8(inv),
                      _c128(mats), _i64(moff)]
        k = self._keep
        self.struct = _OpsT(self.n, C.cast(self._names, C.POINTER(C.c_char_p)), k[0][1], k[1][1], k[2][1], k[3][1],
                            k[4][1], k[5][1], k[6][1], k[7][1], k[8][1], k[9][1])

    def ptr(self):
        return C.byref(self.struct)


class Observable:
    def __init__(self, handle, sfx, keep=()):
        self._h, self._sfx, self._keep = handle, sfx, keep

    @classmethod
    def named(cls, name, wires, dtype=np.complex128):
        sfx = "c128" if np.dtype(dtype) == np.complex128 else "c64"
        w, wp = _i64(wires)
        h = getattr(lib(), f"lqref_obs_named_{sfx}")(name.encode(), wp, C.c_int64(len(w)))
        if not h:
            raise RefError(lib().lqref_last_error().decode())
        return cls(C.c_void_p(h), sfx)

    @classmethod
    def hermitian(cls, matrix, wires, dtype=np.complex128):
        sfx = "c128" if np.dtype(dtype) == np.complex128 else "c64"
        w, wp = _i64(wires)
        m, mp = _c128(matrix)
        h = getattr(lib(), f"lqref_obs_hermitian_{sfx}")(mp, wp, C.c_int64(len(w)))
        if not h:
            raise RefError(lib().lqref_last_error().decode())
        return cls(C.c_void_p(h), sfx)

    @classmethod
    def sparse(cls, indptr, indices, data, wires, dtype=np.complex128):
        sfx = "c128" if np.dtype(dtype) == np.complex128 else "c64"
        ip, ipp = _i64(indptr)
        ix, ixp = _i64(indices)
        d, dp = _c128(data)
        w, wp = _i64(wires)
        f = getattr(lib(), f"lqref_obs_sparse_{sfx}")
        f.restype = C.c_void_p
        h = f(ipp, ixp, dp, C.c_int64(len(ip) - 1), wp, C.c_int64(len(w)))
        if not h:
            raise RefError(lib().lqref_last_error().decode())
        return cls(C.c_void_p(h), sfx)

    @classmethod
    def tensor(cls, terms):
        sfx = terms[0]._sfx
        arr = (C.c_void_p * len(terms))(*[t._h for t in terms])
        h = getattr(lib(), f"lqref_obs_tensor_{sfx}")(arr, C.c_int64(len(terms)))
        if not h:
            raise RefError(lib().lqref_last_error().decode())
        return cls(C.c_void_p(h), sfx, keep=terms)

    @classmethod
    def hamiltonian(cls, coeffs, terms):
        sfx = terms[0]._sfx
        c, cp = _f64(coeffs)
        arr = (C.c_void_p * len(terms))(*[t._h for t in terms])
        h = getattr(lib(), f"lqref_obs_hamiltonian_{sfx}")(cp, arr, C.c_int64(len(terms)))
        if not h:
            raise RefError(lib().lqref_last_error().decode())
        return cls(C.c_void_p(h), sfx, keep=terms)

    def __del__(self):
        try:
            if self._h:
                getattr(lib(), f"lqref_obs_destroy_{self._sfx}")(self._h)
                self._h = None
        except Exception:
            pass


class StateVector:
    """lightning.qubit StateVectorLQubitManaged<float|double> behind the shim."""

    def __init__(self, num_qubits, dtype=np.complex128):
        self.dtype = np.dtype(dtype)
        self.sfx = "c128" if self.dtype == np.complex128 else "c64"
        self.num_qubits = num_qubits
        h = self._f("sv_create")(C.c_int64(num_qubits))
        if not h:
            raise RefError(lib().lqref_last_error().decode())
        self._h = C.c_void_p(h)

    def _f(self, name):
        return getattr(lib(), f"lqref_{name}_{self.sfx}")

    def __del__(self):
        try:
            if self._h:
                self._f("sv_destroy")(self._h)
                self._h = None
        except Exception:
            pass

    def __len__(self):
        return 1 << self.num_qubits

    def _view(self):
        ptr = self._f("sv_data")(self._h)
        n = len(self)
        ctype = C.c_double if self.sfx == "c128" else C.c_float
        buf = (ctype * (2 * n)).from_address(ptr)
        return np.frombuffer(buf, dtype=self.dtype, count=n)

    def get_state(self):
        return self._view().copy()

    def set_state(self, host):
        self._view()[:] = np.asarray(host, dtype=self.dtype)

    def reset(self):
        _check(self._f("sv_reset")(self._h))

    def set_basis_state(self, state, wires):
        s, sp = _i64(state)
        w, wp = _i64(wires)
        _check(self._f("sv_set_basis_state")(self._h, sp, wp, C.c_int64(len(w))))

    def set_state_vector(self, values, wires):
        w, wp = _i64(wires)
        v, vp = _c128(values)
        _check(self._f("sv_set_state_vector")(self._h, vp, wp, C.c_int64(len(w))))

    def collapse(self, wire, branch):
        _check(self._f("sv_collapse")(self._h, C.c_int64(wire), int(bool(branch))))

    def normalize(self):
        _check(self._f("sv_normalize")(self._h))

    def apply(self, name, wires, inverse=False, params=(), ctrl_wires=(), ctrl_values=()):
        w, wp = _i64(wires)
        cw, cwp = _i64(ctrl_wires)
        cv, cvp = _u8(ctrl_values)
        p, pp = _f64(params)
        _check(self._f("sv_apply")(self._h, name.encode(), cwp, cvp, C.c_int64(len(cw)), wp, C.c_int64(len(w)),
                                   int(bool(inverse)), pp, C.c_int64(len(p))))

    def apply_matrix(self, matrix, wires, inverse=False, ctrl_wires=(), ctrl_values=()):
        w, wp = _i64(wires)
        cw, cwp = _i64(ctrl_wires)
        cv, cvp = _u8(ctrl_values)
        m, mp = _c128(matrix)
        _check(self._f("sv_apply_matrix")(self._h, mp, cwp, cvp, C.c_int64(len(cw)), wp, C.c_int64(len(w)),
                                          int(bool(inverse))))

    def apply_pauli_rot(self, wires, inverse, theta, word):
        w, wp = _i64(wires)
        _check(self._f("sv_apply_pauli_rot")(self._h, wp, C.c_int64(len(w)), int(bool(inverse)), C.c_double(theta),
                                             word.encode()))

    def apply_generator(self, name, wires, adj=False, ctrl_wires=(), ctrl_values=()):
        w, wp = _i64(wires)
        cw, cwp = _i64(ctrl_wires)
        cv, cvp = _u8(ctrl_values)
        s = C.c_double()
        _check(self._f("sv_apply_generator")(self._h, name.encode(), cwp, cvp, C.c_int64(len(cw)), wp,
                                             C.c_int64(len(w)), int(bool(adj)), C.byref(s)))
        return s.value

    def apply_ops(self, ops, fuse=False):
        blob = ops if isinstance(ops, OpsBlob) else OpsBlob(ops)
        _check(self._f("apply_ops")(self._h, blob.ptr()))

    def apply_observable(self, obs):
        _check(self._f("obs_apply")(obs._h, self._h))

    def probs(self, wires=None):
        if wires is None:
            out = np.empty(len(self), dtype=np.float64)
            _check(self._f("probs")(self._h, None, C.c_int64(-1), out.ctypes.data_as(_f64p)))
            return out
        w, wp = _i64(wires)
        out = np.empty(1 << len(w), dtype=np.float64)
        _check(self._f("probs")(self._h, wp, C.c_int64(len(w)), out.ctypes.data_as(_f64p)))
        return out

    def _scalar(self, fname, *args):
        out = C.c_double()
        _check(self._f(fname)(self._h, *args, C.byref(out)))
        return out.value

    def expval_named(self, name, wires):
        w, wp = _i64(wires)
        return self._scalar("expval_named", name.encode(), wp, C.c_int64(len(w)))

    def var_named(self, name, wires):
        w, wp = _i64(wires)
        return self._scalar("var_named", name.encode(), wp, C.c_int64(len(w)))

    def expval_matrix(self, matrix, wires):
        w, wp = _i64(wires)
        m, mp = _c128(matrix)
        return self._scalar("expval_matrix", mp, wp, C.c_int64(len(w)))

    def var_matrix(self, matrix, wires):
        w, wp = _i64(wires)
        m, mp = _c128(matrix)
        return self._scalar("var_matrix", mp, wp, C.c_int64(len(w)))

    def expval(self, obs):
        return self._scalar("expval_obs", obs._h)

    def var(self, obs):
        return self._scalar("var_obs", obs._h)

    def expval_shots(self, obs, shots, seed, shot_range=()):
        r, rp = _i64(shot_range)
        return self._scalar("expval_shots", obs._h, C.c_int64(shots), C.c_int64(seed), rp, C.c_int64(len(r)))

    def var_shots(self, obs, shots, seed):
        return self._scalar("var_shots", obs._h, C.c_int64(shots), C.c_int64(seed))

    def probs_shots(self, wires, shots, seed):
        w, wp = _i64(wires)
        out = np.empty(1 << len(w), dtype=np.float64)
        _check(self._f("probs_shots")(self._h, wp, C.c_int64(len(w)), C.c_int64(shots), C.c_int64(seed),
                                      out.ctypes.data_as(_f64p)))
        return out

    def generate_samples(self, shots, wires=None, seed=-1):
        if wires is None:
            nw, wp, k = -1, None, self.num_qubits
        else:
            w, wp = _i64(wires)
            nw = k = len(w)
        out = np.empty((shots, k), dtype=np.uint64)
        _check(self._f("generate_samples")(self._h, wp, C.c_int64(nw), C.c_int64(shots), C.c_int64(seed),
                                           out.ctypes.data_as(_u64p)))
        return out

    def vjp(self, ops, dy, trainable, apply_ops=False):
        blob = ops if isinstance(ops, OpsBlob) else OpsBlob(ops)
        tp, tpp = _i64(trainable)
        d, dp = _c128(dy)
        out = np.zeros(len(tp), dtype=np.complex128)
        _check(self._f("vjp")(self._h, dp, blob.ptr(), tpp, C.c_int64(len(tp)), int(bool(apply_ops)),
                              out.ctypes.data_as(_f64p)))
        return out

    def adjoint_jacobian(self, observables, ops, trainable, apply_ops=False):
        blob = ops if isinstance(ops, OpsBlob) else OpsBlob(ops)
        tp, tpp = _i64(trainable)
        arr = (C.c_void_p * max(len(observables), 1))(*[o._h for o in observables])
        jac = np.zeros(len(observables) * len(tp), dtype=np.float64)
        _check(self._f("adjoint_jacobian")(self._h, arr, C.c_int64(len(observables)), blob.ptr(), tpp,
                                           C.c_int64(len(tp)), int(bool(apply_ops)), jac.ctypes.data_as(_f64p)))
        return jac.reshape(len(observables), len(tp))
