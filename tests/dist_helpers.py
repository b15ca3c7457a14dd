"""Helpers for the multi-rank tests: a numpy local engine (test infrastructure, built on the oracle)
so that the sharding/permutation/swap-scheduling host logic of dist.py runs on CPU under gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


class NumpyEngine:
    def __init__(self, nloc, dtype):
        import torch
        from oracle import np_oracle

        self.torch = torch
        self.nloc = nloc
        self.sv = np_oracle.StateVector(nloc, dtype)
        self.sv.state[:] = 0
        self.launches = 0

    def zero(self):
        self.sv.state[:] = 0

    def set_amp(self, idx, val):
        self.sv.state[idx] = val

    def apply_ops(self, ops, fuse=True):
        self.sv.apply_ops(ops)
        self.launches += len(ops)

    def buffers(self):
        half = 1 << (self.nloc - 1)
        dt = self.torch.complex128 if self.sv.dtype == np.complex128 else self.torch.complex64
        return self.torch.empty(half, dtype=dt), self.torch.empty(half, dtype=dt)

    def _sel(self, bit, keep):
        idx = np.arange(1 << self.nloc)
        return idx[((idx >> bit) & 1) != keep]

    def pack_bit(self, bit, keep, buf):
        buf.copy_(self.torch.from_numpy(self.sv.state[self._sel(bit, keep)].copy()))

    def unpack_bit(self, bit, keep, buf):
        self.sv.state[self._sel(bit, keep)] = buf.numpy()

    def z_sums(self, local_wires):
        p = np.abs(self.sv.state.astype(np.complex128)) ** 2
        idx = np.arange(1 << self.nloc)
        out = [float(np.sum(p * (1 - 2 * ((idx >> (self.nloc - 1 - w)) & 1)))) for w in local_wires]
        return out + [float(p.sum())]

    def pauli_sums(self, words, local_wires):
        return np.array([self.sv.expval_pauli_word(w, ws) if w.strip("I") else float(np.vdot(self.sv.state, self.sv.state).real)
                         for w, ws in zip(words, local_wires)])

    def matrix_sum(self, matrix, local_wires):
        return self.sv.expval_matrix(matrix, local_wires)

    def probs(self, local_wires=None):
        return self.sv.probs(local_wires)

    def host_state(self):
        return self.sv.state.copy()

    def sample_bits(self, shots, seed):
        p = np.abs(self.sv.state.astype(np.complex128)) ** 2
        cdf = np.cumsum(p)
        draws = np.searchsorted(cdf, np.random.default_rng(seed).random(shots) * cdf[-1], side="right").clip(0, len(cdf) - 1)
        return np.stack([(draws >> (self.nloc - 1 - w)) & 1 for w in range(self.nloc)], axis=1).astype(np.uint64).reshape(shots, self.nloc)

    def copy_from(self, other):
        self.sv.state[:] = other.sv.state

    def axpy(self, alpha, other):
        self.sv.state += alpha * other.sv.state

    def dot(self, other):
        return complex(np.vdot(self.sv.state, other.sv.state))

    def apply_generator(self, name, wires, adj, ctrl_wires, ctrl_values):
        self.launches += 1
        return self.sv.apply_generator(name, wires, adj, ctrl_wires, ctrl_values)

    @property
    def kernel_launches(self):
        return self.launches


def mixed_circuit(n, seed, depth=4):
    """config-4 family plus gates that exercise every classification branch of dist.py."""
    from pennylane_lightning_b200 import circuits

    rng = np.random.default_rng(seed)
    ops = circuits.random_circuit(n, depth, seed)
    extra = []
    for _ in range(3 * n):
        r = rng.random()
        p = [int(x) for x in rng.permutation(n)]
        if r < 0.15:
            extra.append(circuits.op("Toffoli", p[:3]))
        elif r < 0.3:
            extra.append(circuits.op("ControlledPhaseShift", p[:2], [rng.uniform(0, 6)]))
        elif r < 0.45:
            extra.append(circuits.op("IsingZZ", p[:2], [rng.uniform(0, 6)], inverse=True))
        elif r < 0.55:
            extra.append(circuits.op("MultiRZ", p[:3], [rng.uniform(0, 6)]))
        elif r < 0.65:
            extra.append(circuits.op("SWAP", p[:2]))
        elif r < 0.75:
            extra.append(circuits.op("RY", p[:1], [rng.uniform(0, 6)], ctrl_wires=p[1:3], ctrl_values=[True, False]))
        elif r < 0.85:
            extra.append(circuits.op("GlobalPhase", p[:1], [rng.uniform(0, 6)]))
        elif r < 0.92:
            extra.append(circuits.op("CSWAP", p[:3]))
        else:
            extra.append(circuits.op("Hadamard", p[:1]))
    out = []
    for i, o in enumerate(ops):
        out.append(o)
        if i % 3 == 0 and extra:
            out.append(extra.pop())
    return out + extra


def worker(rank, world, n, seed, port, backend, q, swap="auto"):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if swap == "peer-multi":  # the k-bit all-to-all exchange kernel, stand-alone (no routed passes)
        os.environ["PLB200_SWAP_MULTI"] = "1"
        os.environ["PLB200_SWAP_FUSED"] = "0"
        swap = "peer"
    elif swap == "peer-nofuse":  # chained single-bit exchange kernels
        os.environ["PLB200_SWAP_FUSED"] = "0"
        swap = "peer"
    elif swap == "peer":  # default: the swap rides on the last pass's store phase (routed passes)
        os.environ["PLB200_SWAP_MULTI"] = "1"
        os.environ["PLB200_JIT_MIN_QUBITS"] = "12"
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pennylane_lightning_b200.dist import DistStateVector

    try:
        factory = NumpyEngine if backend == "gloo" else None
        if backend == "nccl":
            torch.cuda.set_device(rank)
        sv = DistStateVector(n, np.complex128, engine_factory=factory, swap=swap)
        ops = mixed_circuit(n, seed)
        sv.apply_ops(ops, fuse=True)
        z = sv.expval_z_all()
        full = sv.gather_state()
        if rank == 0:
            q.put(dict(state=full, z=z, norm2=sv.last_norm2, swaps=sv.n_swaps,
                       fused_swaps=getattr(sv, "n_fused_swaps", 0)))
    finally:
        dist.destroy_process_group()


def multicall_tapes(n, seed):
    """Several apply_ops calls whose controlled gates sit on (possibly) global wires with no later use of
    the target: the case where a rank-dependent control skip used to make the ranks' schedules diverge."""
    from pennylane_lightning_b200 import circuits

    rng = np.random.default_rng(seed)
    tapes = [[circuits.op("RY", [w], [0.3 + 0.1 * w]) for w in range(n)]]
    for _ in range(4):
        t = []
        for _ in range(n):
            p = [int(x) for x in rng.permutation(n)]
            r = rng.random()
            if r < 0.4:
                t.append(circuits.op("CNOT", p[:2]))
            elif r < 0.6:
                t.append(circuits.op("RX", p[:1], [rng.uniform(0, 6)]))
            elif r < 0.75:
                t.append(circuits.op("CRY", p[:2], [rng.uniform(0, 6)]))
            elif r < 0.9:
                t.append(circuits.op("Toffoli", p[:3]))
            else:
                t.append(circuits.op("RX", p[:1], [rng.uniform(0, 6)], ctrl_wires=p[1:3], ctrl_values=[False, True]))
        tapes.append(t)
    # the advisor's reproducer (n >= 6): controlled gate with control and target both global, then more ops
    tapes.append([circuits.op("CNOT", [2 % n, 3 % n]), circuits.op("RX", [2 % n], [0.7]),
                  circuits.op("CNOT", [4 % n, 5 % n]), circuits.op("RX", [3 % n], [0.9]),
                  circuits.op("CNOT", [0, 4 % n]), circuits.op("RX", [1], [1.1])])
    return tapes


def worker_multicall(rank, world, n, seed, port, backend, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pennylane_lightning_b200.dist import DistStateVector

    try:
        sv = DistStateVector(n, np.complex128, engine_factory=NumpyEngine)
        for t in multicall_tapes(n, seed):
            sv.apply_ops(t, fuse=True)
        phys = [None] * world
        dist.all_gather_object(phys, list(sv.phys))
        full = sv.gather_state()
        if rank == 0:
            q.put(dict(state=full, phys=phys, swaps=sv.n_swaps))
    finally:
        dist.destroy_process_group()


def run_ranks_multicall(world, n, seed, port):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker_multicall, args=(r, world, n, seed, port, "gloo", q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    return res


def worker_measure(rank, world, n, seed, port, backend, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pennylane_lightning_b200 import circuits
    from pennylane_lightning_b200.dist import DistStateVector

    try:
        if backend == "nccl":
            torch.cuda.set_device(rank)
            os.environ["PLB200_JIT_MIN_QUBITS"] = "12"
        sv = DistStateVector(n, np.complex128, engine_factory=NumpyEngine if backend == "gloo" else None)
        sv.apply_ops(circuits.random_circuit(n, 3, seed), fuse=True)
        co, words, wires = circuits.pauli_hamiltonian(n, 12, seed)
        each = sv.expval_pauli_words(words, wires)
        total = sv.expval_pauli_words(words, wires, co)
        pw = [int(x) for x in np.random.default_rng(seed).permutation(n)[:4]]
        p_sub = sv.probs(pw)
        p_all = sv.probs()
        samples = sv.generate_samples(4000, seed=11)
        samples2 = sv.generate_samples(4000, seed=11)
        var = sv.var_pauli_hamiltonian(co, words, wires)
        rngm = np.random.default_rng(seed + 5)
        mw = [int(x) for x in rngm.permutation(n)[:2]]
        hm = rngm.normal(size=(4, 4)) + 1j * rngm.normal(size=(4, 4))
        hm = hm + hm.conj().T
        em = sv.expval_matrix(hm, mw)
        full = sv.gather_state()
        if rank == 0:
            q.put(dict(each=each, total=total, pw=pw, p_sub=p_sub, p_all=p_all, samples=samples, var=var, em=em, hm=hm, mw=mw,
                       same=bool(np.array_equal(samples, samples2)), state=full, words=words, wires=wires, co=co))
    finally:
        dist.destroy_process_group()


def check_measurements(res, n):
    from oracle import np_oracle

    psi = res["state"]
    ref = np_oracle.StateVector(n)
    ref.set_state(psi)
    for k, (w, ws) in enumerate(zip(res["words"], res["wires"])):
        assert abs(res["each"][k] - ref.expval_pauli_word(w, ws)) < 1e-12
    assert abs(res["total"] - sum(c * ref.expval_pauli_word(w, ws) for c, w, ws in zip(res["co"], res["words"], res["wires"]))) < 1e-12
    # variance of the Hamiltonian: ||H psi||^2 - <H>^2 with H psi from the oracle's Pauli-word application
    hpsi = np.zeros_like(psi, dtype=np.complex128)
    for c, w, ws in zip(res["co"], res["words"], res["wires"]):
        t = np_oracle.StateVector(n)
        t.set_state(psi)
        t.apply_ops([dict(name={"X": "PauliX", "Y": "PauliY", "Z": "PauliZ"}[ch], wires=[q], params=[], inverse=False,
                          ctrl_wires=[], ctrl_values=[]) for ch, q in zip(w, ws)])
        hpsi += c * t.get_state()
    want_var = float(np.vdot(hpsi, hpsi).real - np.vdot(psi, hpsi).real ** 2)
    assert abs(res["var"] - want_var) < 1e-11, (res["var"], want_var)
    assert abs(res["em"] - ref.expval_matrix(res["hm"], res["mw"])) < 1e-11
    np.testing.assert_allclose(res["p_sub"], ref.probs(res["pw"]), rtol=0, atol=1e-13)
    np.testing.assert_allclose(res["p_all"], ref.probs(), rtol=0, atol=1e-13)
    s = res["samples"]
    assert s.shape == (4000, n) and res["same"]  # reproducible under a seed
    idx = np.zeros(len(s), dtype=np.int64)
    for w in range(n):
        idx |= s[:, w].astype(np.int64) << (n - 1 - w)
    if n <= 8:  # total-variation distance of 4000 draws over 2^n outcomes: well below 0.25 for the right distribution
        emp = np.bincount(idx, minlength=1 << n) / len(s)
        assert 0.5 * np.abs(emp - ref.probs()).sum() < 0.25
    for w in range(n):  # every single-wire marginal (4 sigma of 4000 shots = 0.063)
        zw = 1.0 - 2.0 * s[:, w].mean()
        assert abs(zw - ref.expval_pauli_word("Z", [w])) < 0.08, w
    # and a two-wire correlation, which a wrong wire map would break
    zz = np.mean((1.0 - 2.0 * s[:, 0]) * (1.0 - 2.0 * s[:, n - 1]))
    assert abs(zz - ref.expval_pauli_word("ZZ", [0, n - 1])) < 0.08

def run_ranks_measure(world, n, seed, backend="gloo", port=29680):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker_measure, args=(r, world, n, seed, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    return res


def run_ranks(world, n, seed, backend="gloo", port=29611, swap="auto"):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, n, seed, port, backend, q, swap)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    return res


def adjoint_case(n, seed):
    """Tape (one-parameter gates incl. controlled ones on every wire position), trainable subset, two Hamiltonians."""
    from pennylane_lightning_b200 import circuits

    rng = np.random.default_rng(seed)
    ops = []
    for layer in range(3):
        for w in range(n):
            ops.append(circuits.op(("RX", "RY", "RZ", "PhaseShift")[int(rng.integers(4))], [w], [rng.uniform(0, 6)],
                                   inverse=bool(rng.integers(2))))
        p = [int(x) for x in rng.permutation(n)]
        for i in range(0, n - 1, 2):
            nm = ("CNOT", "CRX", "CRY", "CRZ", "IsingXX", "IsingZZ", "ControlledPhaseShift", "CZ")[int(rng.integers(8))]
            ops.append(circuits.op(nm, [p[i], p[i + 1]], [rng.uniform(0, 6)] if nm not in ("CNOT", "CZ") else []))
        ops.append(circuits.op("RY", [p[0]], [rng.uniform(0, 6)], ctrl_wires=[p[1], p[2]], ctrl_values=[True, False]))
    n_par = sum(1 for o in ops if o["params"])
    tp = sorted(int(x) for x in rng.choice(n_par, size=(2 * n_par) // 3, replace=False))
    obs = []
    for k in range(2):
        co, words, wires = circuits.pauli_hamiltonian(n, 5, seed + 10 * k)
        obs.append((co, words, wires))
    return ops, tp, obs


def worker_adjoint(rank, world, n, seed, port, backend, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pennylane_lightning_b200.dist import DistStateVector

    try:
        if backend == "nccl":
            torch.cuda.set_device(rank)
            os.environ["PLB200_JIT_MIN_QUBITS"] = "12"
        sv = DistStateVector(n, np.complex128, engine_factory=NumpyEngine if backend == "gloo" else None)
        ops, tp, obs = adjoint_case(n, seed)
        jac = sv.adjoint_jacobian(ops, tp, obs)
        if rank == 0:
            q.put(dict(jac=jac))
    finally:
        dist.destroy_process_group()


def run_ranks_adjoint(world, n, seed, backend="gloo", port=29720):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker_adjoint, args=(r, world, n, seed, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=60)
    return res


def check_adjoint(res, n, seed):
    from oracle import np_oracle

    ops, tp, obs = adjoint_case(n, seed)
    name = {"X": "PauliX", "Y": "PauliY", "Z": "PauliZ"}
    hams = []
    for co, words, wires in obs:
        terms = []
        for word, ws in zip(words, wires):
            named = [np_oracle.Observable.named(name[c], [w]) for c, w in zip(word, ws)]
            terms.append(named[0] if len(named) == 1 else np_oracle.Observable.tensor(named))
        hams.append(np_oracle.Observable.hamiltonian(co, terms))
    expect = np.asarray(np_oracle.StateVector(n, np.complex128).adjoint_jacobian(hams, ops, tp, apply_ops=True))
    np.testing.assert_allclose(res["jac"], expect.reshape(len(hams), len(tp)), rtol=0, atol=1e-11)
