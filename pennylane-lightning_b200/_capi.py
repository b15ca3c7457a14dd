"""ctypes view of libplb200.so (include/plb200.h).

Mirrors the call shapes of the reference's binding layer (core/bindings/Bindings.hpp:223-276,
650-729): ``sv.apply(name, wires, inverse, params)``, controlled overloads, ``applyMatrix``,
measurements and the adjoint Jacobian.  The product path is the shared library; this file only
marshals arguments.  It raises if the library is missing — there is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLB200_LIB_PATH") or os.path.join(_HERE, "lib", "libplb200.so")

_lib = None


class B200Error(RuntimeError):
    """Mirror of Pennylane::Util::LightningException surfaced by the C ABI."""


def build(verbose: bool = False) -> str:
    """Compile libplb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "all", "emu"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libplb200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


class _OpsT(C.Structure):
    _fields_ = [
        ("n_ops", C.c_int64),
        ("names", C.POINTER(C.c_char_p)),
        ("wires", _i64p),
        ("wires_off", _i64p),
        ("ctrl_wires", _i64p),
        ("ctrl_off", _i64p),
        ("ctrl_values", _u8p),
        ("params", _f64p),
        ("params_off", _i64p),
        ("inverses", _u8p),
        ("mats", _f64p),
        ("mats_off", _i64p),
    ]


def lib():
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} is missing: run __graft_entry__.build() / make -C pennylane-lightning_b200/csrc; "
                "there is no fallback implementation"
            )
        _lib = C.CDLL(LIB_PATH)
        _lib.plb200_last_error.restype = C.c_char_p
        _lib.plb200_version.restype = C.c_char_p
        _lib.plb200_sv_device_ptr.restype = C.c_void_p
        for f in ("plb200_sv_num_qubits", "plb200_sv_length", "plb200_sv_kernel_launches"):
            getattr(_lib, f).restype = C.c_int64
    return _lib


def _check(rc):
    if rc != 0:
        raise B200Error(lib().plb200_last_error().decode())


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


# ---- pass specialisation (NVRTC) controls, include/plb200.h
def jit_available() -> bool:
    return bool(lib().plb200_jit_available())


def jit_mode() -> int:
    """0 = off, 1 = background compilation from the second sighting of a pass structure, 2 = blocking"""
    return int(lib().plb200_jit_mode())


def jit_enabled() -> bool:
    return jit_mode() != 0 and jit_available()


def jit_set_mode(mode: int) -> None:
    lib().plb200_jit_set_mode(int(mode))


def jit_wait() -> None:
    """Block until every queued pass kernel is compiled (warm-up helper)."""
    _check(lib().plb200_jit_wait())


def jit_stats() -> dict:
    out = (C.c_int64 * 8)()
    lib().plb200_jit_stats(out)
    keys = ("compiled", "from_disk_cache", "jit_launches", "interpreter_launches", "failed", "compile_us", "pending",
            "structures_seen")
    return dict(zip(keys, [int(x) for x in out]))


def hermitian_eigh(matrix):
    """Host eigen-decomposition of a Hermitian matrix -> (eigvals ascending, unitary = V^dagger)."""
    m = np.ascontiguousarray(np.asarray(matrix, dtype=np.complex128))
    dim = m.shape[0]
    ev = np.empty(dim, dtype=np.float64)
    u = np.empty((dim, dim), dtype=np.complex128)
    _check(lib().plb200_hermitian_eigh(m.ctypes.data_as(_f64p), C.c_int64(dim), ev.ctypes.data_as(_f64p),
                                       u.ctypes.data_as(_f64p)))
    return ev, u


class OpsBlob:
    """Flattened tape (plb200_ops_t): the C image of OpsData (JacobianData.hpp:39-253).

    ``ops`` is a list of dicts/tuples with keys name, wires, params, inverse, ctrl_wires,
    ctrl_values, matrix."""

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
        self._keep = [
            _i64(wires), _i64(woff), _i64(cw), _i64(coff), _u8(cv), _f64(params), _i64(poff), _u8(inv),
            _c128(mats), _i64(moff),
        ]
        k = self._keep
        self.struct = _OpsT(
            self.n, C.cast(self._names, C.POINTER(C.c_char_p)), k[0][1], k[1][1], k[2][1], k[3][1], k[4][1],
            k[5][1], k[6][1], k[7][1], k[8][1], k[9][1],
        )

    def ptr(self):
        return C.byref(self.struct)


class Observable:
    """Handle to a plb200_obs tree (NamedObs / HermitianObs / TensorProdObs / Hamiltonian)."""

    def __init__(self, handle, keep=()):
        self._h = handle
        self._keep = keep

    @classmethod
    def named(cls, name, wires, params=()):
        h = C.c_void_p()
        w, wp = _i64(wires)
        p, pp = _f64(params)
        _check(lib().plb200_obs_named(C.byref(h), name.encode(), wp, C.c_int64(len(w)), pp, C.c_int64(len(p))))
        return cls(h)

    @classmethod
    def hermitian(cls, matrix, wires):
        h = C.c_void_p()
        w, wp = _i64(wires)
        m, mp = _c128(matrix)
        if len(m) != 4 ** len(w):
            raise B200Error("The size of matrix does not match with the given number of wires")
        _check(lib().plb200_obs_hermitian(C.byref(h), mp, wp, C.c_int64(len(w))))
        return cls(h)

    @classmethod
    def sparse(cls, indptr, indices, data, wires=None):
        """SparseHamiltonian: CSR matrix over the full 2^n index space (`wires` kept for signature parity)."""
        h = C.c_void_p()
        ip, ipp = _i64(indptr)
        ix, ixp = _i64(indices)
        d, dp = _c128(data)
        _check(lib().plb200_obs_sparse(C.byref(h), ipp, ixp, dp, C.c_int64(len(ip) - 1)))
        return cls(h)

    @classmethod
    def tensor(cls, terms):
        h = C.c_void_p()
        arr = (C.c_void_p * len(terms))(*[t._h for t in terms])
        _check(lib().plb200_obs_tensor(C.byref(h), arr, C.c_int64(len(terms))))
        return cls(h)

    @classmethod
    def hamiltonian(cls, coeffs, terms):
        h = C.c_void_p()
        c, cp = _f64(coeffs)
        arr = (C.c_void_p * len(terms))(*[t._h for t in terms])
        _check(lib().plb200_obs_hamiltonian(C.byref(h), cp, arr, C.c_int64(len(terms))))
        return cls(h)

    def __del__(self):
        try:
            if self._h:
                lib().plb200_obs_destroy(self._h)
                self._h = None
        except Exception:
            pass


class StateVector:
    """Device-resident state vector driven through the C ABI."""

    def __init__(self, num_qubits, dtype=np.complex128, device=0, stream=0, device_ptr=None):
        self.dtype = np.dtype(dtype)
        self.precision = 64 if self.dtype == np.complex128 else 32
        self._h = C.c_void_p()
        if device_ptr is None:
            _check(lib().plb200_sv_create(C.byref(self._h), C.c_int64(num_qubits), self.precision, device,
                                          C.c_void_p(stream)))
        else:
            _check(lib().plb200_sv_create_external(C.byref(self._h), C.c_int64(num_qubits), self.precision, device,
                                                   C.c_void_p(stream), C.c_void_p(device_ptr)))
        self.num_qubits = num_qubits

    def __del__(self):
        try:
            if self._h:
                lib().plb200_sv_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ data movement
    def __len__(self):
        return 1 << self.num_qubits

    @property
    def device_ptr(self):
        return lib().plb200_sv_device_ptr(self._h)

    @property
    def kernel_launches(self):
        return lib().plb200_sv_kernel_launches(self._h)

    def sync(self):
        _check(lib().plb200_sv_sync(self._h))

    def set_state(self, host, n_elems=None, async_=False):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        n = host.size if n_elems is None else n_elems
        _check(lib().plb200_sv_h2d(self._h, host.ctypes.data_as(C.c_void_p), C.c_int64(n), int(async_)))
        self._pin = host

    def h2d_ptr(self, ptr, n_elems, async_=True):
        _check(lib().plb200_sv_h2d(self._h, C.c_void_p(ptr), C.c_int64(n_elems), int(async_)))

    def d2h_ptr(self, ptr, n_elems, async_=True):
        _check(lib().plb200_sv_d2h(self._h, C.c_void_p(ptr), C.c_int64(n_elems), int(async_)))

    def get_state(self, n_elems=None):
        n = len(self) if n_elems is None else n_elems
        out = np.empty(n, dtype=self.dtype)
        _check(lib().plb200_sv_d2h(self._h, out.ctypes.data_as(C.c_void_p), C.c_int64(n), 0))
        return out

    def copy_from(self, other):
        _check(lib().plb200_sv_d2d(self._h, other._h))

    def reset(self):
        _check(lib().plb200_sv_reset(self._h))

    def set_basis_state(self, state, wires):
        s, sp = _i64(state)
        w, wp = _i64(wires)
        _check(lib().plb200_sv_set_basis_state(self._h, sp, wp, C.c_int64(len(w))))

    def set_state_vector(self, values, wires):
        w, wp = _i64(wires)
        v, vp = _c128(values)
        if len(v) != 2 ** len(w):
            raise B200Error("Inconsistent state and wires dimensions.")
        _check(lib().plb200_sv_set_state_vector(self._h, vp, wp, C.c_int64(len(w))))

    def set_state_indices(self, indices, values):
        i, ip = _i64(indices)
        v, vp = _c128(values)
        if len(i) != len(v):
            raise B200Error("Indices and values length must match")
        _check(lib().plb200_sv_set_state_indices(self._h, ip, vp, C.c_int64(len(i))))

    def collapse(self, wire, branch):
        _check(lib().plb200_sv_collapse(self._h, C.c_int64(wire), int(bool(branch))))

    def normalize(self):
        _check(lib().plb200_sv_normalize(self._h))

    # ------------------------------------------------------------------ gates
    def apply(self, name, wires, inverse=False, params=(), ctrl_wires=(), ctrl_values=()):
        w, wp = _i64(wires)
        cw, cwp = _i64(ctrl_wires)
        cv, cvp = _u8(ctrl_values)
        if len(cw) != len(cv):
            raise B200Error("`controlled_wires` must have the same size as `controlled_values`.")
        p, pp = _f64(params)
        _check(lib().plb200_sv_apply(self._h, name.encode(), cwp, cvp, C.c_int64(len(cw)), wp, C.c_int64(len(w)),
                                     int(bool(inverse)), pp, C.c_int64(len(p))))

    def apply_matrix(self, matrix, wires, inverse=False, ctrl_wires=(), ctrl_values=()):
        w, wp = _i64(wires)
        cw, cwp = _i64(ctrl_wires)
        cv, cvp = _u8(ctrl_values)
        m, mp = _c128(matrix)
        if len(w) == 0:
            raise B200Error("Number of wires must be larger than 0")
        if len(m) != 4 ** len(w):
            raise B200Error("The size of matrix does not match with the given number of wires")
        if len(cw) != len(cv):
            raise B200Error("`controlled_wires` must have the same size as `controlled_values`.")
        _check(lib().plb200_sv_apply_matrix(self._h, mp, cwp, cvp, C.c_int64(len(cw)), wp, C.c_int64(len(w)),
                                            int(bool(inverse))))

    def apply_pauli_rot(self, wires, inverse, theta, word):
        w, wp = _i64(wires)
        _check(lib().plb200_sv_apply_pauli_rot(self._h, wp, C.c_int64(len(w)), int(bool(inverse)),
                                               C.c_double(theta), word.encode()))

    def apply_generator(self, name, wires, adj=False, ctrl_wires=(), ctrl_values=()):
        w, wp = _i64(wires)
        cw, cwp = _i64(ctrl_wires)
        cv, cvp = _u8(ctrl_values)
        s = C.c_double()
        _check(lib().plb200_sv_apply_generator(self._h, name.encode(), cwp, cvp, C.c_int64(len(cw)), wp,
                                               C.c_int64(len(w)), int(bool(adj)), C.byref(s)))
        return s.value

    def apply_ops(self, ops, fuse=True):
        blob = ops if isinstance(ops, OpsBlob) else OpsBlob(ops)
        _check(lib().plb200_sv_apply_ops(self._h, blob.ptr(), int(bool(fuse))))

    def last_apply_stats(self):
        s = (C.c_int64 * 2)()
        lib().plb200_sv_last_apply_stats(self._h, s)
        return int(s[0]), int(s[1])

    # ------------------------------------------------------------------ linear algebra
    def dot(self, other):
        out = (C.c_double * 2)()
        _check(lib().plb200_sv_dot(self._h, other._h, out))
        return complex(out[0], out[1])

    def axpy(self, alpha, x):
        a = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        _check(lib().plb200_sv_axpy(self._h, a, x._h))

    def scale(self, alpha):
        a = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        _check(lib().plb200_sv_scale(self._h, a))

    def norm2(self):
        out = C.c_double()
        _check(lib().plb200_sv_norm2(self._h, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ measurements
    def probs(self, wires=None):
        if wires is None:
            out = np.empty(len(self), dtype=np.float64)
            _check(lib().plb200_probs(self._h, None, C.c_int64(-1), out.ctypes.data_as(_f64p)))
            return out
        w, wp = _i64(wires)
        out = np.empty(1 << len(w), dtype=np.float64)
        _check(lib().plb200_probs(self._h, wp, C.c_int64(len(w)), out.ctypes.data_as(_f64p)))
        return out

    def expval_named(self, name, wires):
        w, wp = _i64(wires)
        out = C.c_double()
        _check(lib().plb200_expval_named(self._h, name.encode(), wp, C.c_int64(len(w)), C.byref(out)))
        return out.value

    def var_named(self, name, wires):
        w, wp = _i64(wires)
        out = C.c_double()
        _check(lib().plb200_var_named(self._h, name.encode(), wp, C.c_int64(len(w)), C.byref(out)))
        return out.value

    def expval_matrix(self, matrix, wires):
        w, wp = _i64(wires)
        m, mp = _c128(matrix)
        if len(m) != 4 ** len(w):
            raise B200Error("The size of matrix does not match with the given number of wires")
        out = C.c_double()
        _check(lib().plb200_expval_matrix(self._h, mp, wp, C.c_int64(len(w)), C.byref(out)))
        return out.value

    def var_matrix(self, matrix, wires):
        w, wp = _i64(wires)
        m, mp = _c128(matrix)
        if len(m) != 4 ** len(w):
            raise B200Error("The size of matrix does not match with the given number of wires")
        out = C.c_double()
        _check(lib().plb200_var_matrix(self._h, mp, wp, C.c_int64(len(w)), C.byref(out)))
        return out.value

    def _words(self, words, wires):
        arr = (C.c_char_p * max(len(words), 1))(*[w.encode() for w in words])
        flat, off = [], [0]
        for w in wires:
            flat += list(w)
            off.append(len(flat))
        return arr, _i64(flat), _i64(off)

    def expval_pauli_words(self, words, wires, coeffs):
        arr, (f, fp), (o, op) = self._words(words, wires)
        c, cp = _f64(coeffs)
        out = C.c_double()
        _check(lib().plb200_expval_pauli_words(self._h, arr, fp, op, cp, C.c_int64(len(words)), C.byref(out)))
        return out.value

    def expval_pauli_words_each(self, words, wires):
        arr, (f, fp), (o, op) = self._words(words, wires)
        out = np.empty(len(words), dtype=np.float64)
        _check(lib().plb200_expval_pauli_words_each(self._h, arr, fp, op, C.c_int64(len(words)),
                                                    out.ctypes.data_as(_f64p)))
        return out

    def expval_sparse(self, indptr, indices, data):
        ip, ipp = _i64(indptr)
        ix, ixp = _i64(indices)
        d, dp = _c128(data)
        out = C.c_double()
        _check(lib().plb200_expval_sparse(self._h, ipp, ixp, dp, C.c_int64(len(ip) - 1), C.byref(out)))
        return out.value

    def var_sparse(self, indptr, indices, data):
        ip, ipp = _i64(indptr)
        ix, ixp = _i64(indices)
        d, dp = _c128(data)
        out = C.c_double()
        _check(lib().plb200_var_sparse(self._h, ipp, ixp, dp, C.c_int64(len(ip) - 1), C.byref(out)))
        return out.value

    def expval(self, obs: Observable):
        out = C.c_double()
        _check(lib().plb200_expval_obs(self._h, obs._h, C.byref(out)))
        return out.value

    def var(self, obs: Observable):
        out = C.c_double()
        _check(lib().plb200_var_obs(self._h, obs._h, C.byref(out)))
        return out.value

    def apply_observable(self, obs: Observable):
        _check(lib().plb200_obs_apply(obs._h, self._h))

    def generate_samples(self, shots, wires=None, seed=-1, device=None):
        """device=None: alias table on the host up to 24 wires (the reference's random stream), the table-free
        device sampler above; True / False force one of them."""
        if wires is None:
            nw, wp, k = -1, None, self.num_qubits
        else:
            w, wp = _i64(wires)
            nw = k = len(w)
        out = np.empty((shots, k), dtype=np.uint64)
        fn = lib().plb200_generate_samples_device if device else lib().plb200_generate_samples
        if device is False and k > 24:
            raise ValueError("the host alias table is limited to 24 wires")
        _check(fn(self._h, wp, C.c_int64(nw), C.c_int64(shots), C.c_int64(seed), out.ctypes.data_as(_u64p)))
        return out

    def vjp(self, ops, dy, trainable, apply_ops=False):
        """VectorJacobianProduct: vjp[k] = sum_i conj(dy_i) d psi_i / d theta_k (dy: 2^n complex cotangent)."""
        blob = ops if isinstance(ops, OpsBlob) else OpsBlob(ops)
        tp, tpp = _i64(trainable)
        d, dp = _c128(dy)
        out = np.zeros(len(tp), dtype=np.complex128)
        _check(lib().plb200_vjp(self._h, dp, blob.ptr(), tpp, C.c_int64(len(tp)), int(bool(apply_ops)),
                                out.ctypes.data_as(_f64p)))
        return out

    # ------------------------------------------------------------------ adjoint Jacobian
    def adjoint_jacobian(self, observables, ops, trainable, apply_ops=False):
        blob = ops if isinstance(ops, OpsBlob) else OpsBlob(ops)
        tp, tpp = _i64(trainable)
        arr = (C.c_void_p * max(len(observables), 1))(*[o._h for o in observables])
        jac = np.zeros(len(observables) * len(tp), dtype=np.float64)
        _check(lib().plb200_adjoint_jacobian(self._h, arr, C.c_int64(len(observables)), blob.ptr(), tpp,
                                             C.c_int64(len(tp)), int(bool(apply_ops)), jac.ctypes.data_as(_f64p)))
        return jac.reshape(len(observables), len(tp))
