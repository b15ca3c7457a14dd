"""Distributed state vector: index bits sharded over the GPUs of one box, one process per GPU.

Replaces the reference's MPI backends (lightning_gpu/StateVectorCudaMPI.hpp:1936-2243: swap — apply —
swap back around every gate touching a global wire; lightning_kokkos/StateVectorKokkosMPI.hpp:747-1097:
lazy wire permutation) with:

* rank r holds the slab of amplitudes whose top g = log2(world) *physical* index bits equal r;
* a persistent logical-wire -> physical-bit permutation kept on the host (never swapped back);
* per op: all non-diagonal targets local -> plain local kernel(s); a control on a global bit ->
  ranks whose bit mismatches skip, the others drop the control (no communication); a diagonal factor
  on a global bit -> rank-dependent phase (no communication); otherwise an index-bit swap
  (pack kernel -> NCCL send/recv between the two partner ranks -> unpack kernel; half a slab each way);
* ops are collected into maximal local batches (ops on disjoint wires commute) which run through the
  fused tile executor; swaps are chosen for the wires blocked ops wait for, evicting the local wires
  whose next non-diagonal use is farthest away;
* reductions: local kernel -> all_reduce(SUM) of fp64 scalars.

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is only plumbing; the local engine is
libplb200 through the C ABI (`_capi.StateVector` on a torch-owned slab).
"""
from __future__ import annotations

import os

import numpy as np

# gates that are "control (x) base": name -> (base gate, number of leading control wires)
IMPLICIT_CTRL = {
    "CNOT": ("PauliX", 1), "CY": ("PauliY", 1), "CZ": ("PauliZ", 1), "CRX": ("RX", 1), "CRY": ("RY", 1),
    "CRZ": ("RZ", 1), "CRot": ("Rot", 1), "Toffoli": ("PauliX", 2), "CSWAP": ("SWAP", 1),
    "ControlledPhaseShift": ("PhaseShift", 1),
}


def _diag_of(name, params, k, inverse):
    """Diagonal of a diagonal base gate on k wires (wires[0] = most significant), or None."""
    p = params
    if name == "PauliZ":
        d = np.array([1, -1], dtype=complex)
    elif name == "S":
        d = np.array([1, 1j])
    elif name == "T":
        d = np.array([1, np.exp(0.25j * np.pi)])
    elif name == "PhaseShift":
        d = np.array([1, np.exp(1j * p[0])])
    elif name == "RZ":
        d = np.array([np.exp(-0.5j * p[0]), np.exp(0.5j * p[0])])
    elif name == "IsingZZ":
        a, b = np.exp(-0.5j * p[0]), np.exp(0.5j * p[0])
        d = np.array([a, b, b, a])
    elif name == "MultiRZ":
        par = np.array([bin(i).count("1") & 1 for i in range(1 << k)])
        d = np.exp(-0.5j * p[0] * (1 - 2 * par))
    elif name == "GlobalPhase":
        d = np.full(1 << k, np.exp(-1j * p[0]))
    elif name == "Identity":
        d = np.ones(1 << k, dtype=complex)
    else:
        return None
    return d.conj() if inverse else d


_DIAG_BASES = frozenset(("PauliZ", "S", "T", "PhaseShift", "RZ", "IsingZZ", "MultiRZ", "GlobalPhase", "Identity"))


def normalize_op(o):
    """-> dict(base, targets, ctrl_wires, ctrl_values, params, inverse, matrix, diag)"""
    name = o["name"]
    wires = list(o["wires"])
    cw = list(o.get("ctrl_wires", ()))
    cv = [bool(v) for v in o.get("ctrl_values", ())]
    if name in IMPLICIT_CTRL:
        base, nc = IMPLICIT_CTRL[name]
        cw = cw + wires[:nc]
        cv = cv + [True] * nc
        wires = wires[nc:]
        name = base
    m = o.get("matrix", None)
    return dict(base=name, targets=wires, ctrl_wires=cw, ctrl_values=cv, params=list(o.get("params", ())),
                inverse=bool(o.get("inverse", False)), matrix=m, diag=(m is None and name in _DIAG_BASES),
                wires=frozenset(wires) | frozenset(cw))


class LocalEngine:
    """GPU local engine: libplb200 on its own cudaMalloc'ed slab (IPC-exportable for the peer path)."""

    def __init__(self, nloc, dtype, device):
        import torch

        from . import _capi

        self.torch = torch
        self.nloc = nloc
        self.dtype = np.dtype(dtype)
        self.device = device
        self.tdt = torch.complex128 if self.dtype == np.complex128 else torch.complex64
        self.sv = _capi.StateVector(nloc, dtype, device.index or 0, torch.cuda.current_stream(device).cuda_stream)
        self._buf = None
        self.peers = {}

    def zero(self):
        self.sv.set_state_indices([], [])

    def set_amp(self, idx, val):
        self.sv.set_state_indices([idx], [val])

    # ---- peer (CUDA IPC over NVLink) plumbing
    def ipc_handle(self):
        import ctypes as C

        from . import _capi

        buf = (C.c_ubyte * 64)()
        _capi._check(_capi.lib().plb200_sv_ipc_handle(self.sv._h, buf))
        return bytes(buf)

    def open_peer(self, rank, handle, alt=False):
        import ctypes as C

        from . import _capi

        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        _capi._check(_capi.lib().plb200_ipc_open(buf, self.device.index or 0, C.byref(ptr)))
        if alt:
            if not hasattr(self, "peers_alt"):
                self.peers_alt = {}
            self.peers_alt[rank] = ptr.value
        else:
            self.peers[rank] = ptr.value

    def flip_slabs(self):
        """After a routed pass every rank's ping-pong slab is its state: the peer maps follow."""
        self.peers, self.peers_alt = self.peers_alt, self.peers

    # ---- ping-pong slab + routed tapes (the index-bit swap rides on the last pass's store phase)
    def alloc_alt(self):
        from . import _capi

        _capi._check(_capi.lib().plb200_sv_alloc_alt(self.sv._h))

    def ipc_handle_alt(self):
        import ctypes as C

        from . import _capi

        buf = (C.c_ubyte * 64)()
        _capi._check(_capi.lib().plb200_sv_ipc_handle_alt(self.sv._h, buf))
        return bytes(buf)

    def apply_ops_route(self, ops, lbits, my_value, dst_ptrs):
        """Apply `ops`; the last fused pass stores through the swap of local bits `lbits` with the ranks' global
        bits (dst_ptrs[p]: peer slab for local-bit value p).  Returns True when the state left through the route."""
        import ctypes as C

        from . import _capi

        blob = _capi.OpsBlob(ops)
        k = len(lbits)
        arr = (C.c_void_p * (1 << k))(*[None if p == my_value else dst_ptrs[p] for p in range(1 << k)])
        routed = C.c_int(0)
        _capi._check(_capi.lib().plb200_sv_apply_ops_route(self.sv._h, blob.ptr(), C.c_int64(k), (C.c_int64 * k)(*lbits),
                                                           C.c_int64(my_value), arr, C.byref(routed)))
        return bool(routed.value)

    def close_peers(self):
        """Unmap the peers' slabs (their owners cannot release the memory while a mapping exists)."""
        import ctypes as C

        from . import _capi

        for ptrs in (self.peers, getattr(self, "peers_alt", {})):
            for r, ptr in list(ptrs.items()):
                _capi.lib().plb200_ipc_close(C.c_void_p(ptr), self.device.index or 0)
        self.peers = {}
        self.peers_alt = {}

    def swap_bit_peer(self, bit, keep, partner, half):
        import ctypes as C

        from . import _capi

        _capi._check(_capi.lib().plb200_sv_swap_bit_peer(self.sv._h, C.c_int64(bit), int(keep),
                                                         C.c_void_p(self.peers[partner]), int(half)))

    def swap_bits_peer(self, bits, my_value, partner_ranks):
        """all-to-all exchange of len(bits) local bits; partner_ranks[p] = rank holding value p."""
        import ctypes as C

        from . import _capi

        k = len(bits)
        arr = (C.c_void_p * (1 << k))()
        for p, r in enumerate(partner_ranks):
            arr[p] = None if p == my_value else self.peers[r]
        b = (C.c_int64 * k)(*bits)
        _capi._check(_capi.lib().plb200_sv_swap_bits_peer(self.sv._h, b, C.c_int64(k), C.c_int64(my_value), arr))

    def sync(self):
        self.sv.sync()

    def apply_ops(self, ops, fuse=True):
        if ops:
            self.sv.apply_ops(ops, fuse=fuse)

    def apply_matrix(self, matrix, wires, ctrl_wires=(), ctrl_values=()):
        self.sv.apply_matrix(matrix, wires, False, ctrl_wires, ctrl_values)

    def buffers(self):
        if self._buf is None:
            half = 1 << (self.nloc - 1)
            self._buf = (self.torch.empty(half, dtype=self.tdt, device=self.device),
                         self.torch.empty(half, dtype=self.tdt, device=self.device))
        return self._buf

    def pack_bit(self, bit, keep, buf):
        from . import _capi
        import ctypes as C

        _capi._check(_capi.lib().plb200_sv_pack_bit(self.sv._h, C.c_int64(bit), int(keep), C.c_void_p(buf.data_ptr())))

    def unpack_bit(self, bit, keep, buf):
        from . import _capi
        import ctypes as C

        _capi._check(_capi.lib().plb200_sv_unpack_bit(self.sv._h, C.c_int64(bit), int(keep),
                                                      C.c_void_p(buf.data_ptr())))

    def z_sums(self, local_wires):
        """un-normalised sum_i (+-)|a_i|^2 for Z on each local wire, plus the local norm^2 (last)."""
        words = ["Z"] * len(local_wires) + ["I"]
        wires = [[w] for w in local_wires] + [[0]]
        return self.sv.expval_pauli_words_each(words, wires)

    def pauli_sums(self, words, local_wires):
        """un-normalised <slab| P_k |slab> for Pauli words over LOCAL wires (engine wire numbering)"""
        return np.asarray(self.sv.expval_pauli_words_each(words, local_wires))

    def matrix_sum(self, matrix, local_wires):
        """un-normalised <slab| M |slab> for a matrix on LOCAL wires"""
        return self.sv.expval_matrix(matrix, local_wires)

    def probs(self, local_wires=None):
        """un-normalised |a_i|^2 marginal over the given local wires (all local wires when None)"""
        return self.sv.probs(local_wires)

    def host_state(self):
        return self.sv.get_state()

    def sample_bits(self, shots, seed):
        """`shots` basis states of this slab's (un-normalised) distribution: (shots, nloc) bits, engine wire order"""
        if shots == 0:
            return np.zeros((0, self.nloc), dtype=np.uint64)
        return self.sv.generate_samples(shots, seed=int(seed) & 0x7FFFFFFFFFFFFFFF, device=True)

    # ---- linear algebra between slabs with the same layout (sharded adjoint: lambda, H lambda, mu)
    def copy_from(self, other):
        self.sv.copy_from(other.sv)

    def axpy(self, alpha, other):
        self.sv.axpy(alpha, other.sv)

    def dot(self, other):
        """local part of <self|other>"""
        return self.sv.dot(other.sv)

    def apply_generator(self, name, wires, adj, ctrl_wires, ctrl_values):
        return self.sv.apply_generator(name, wires, adj, ctrl_wires, ctrl_values)

    @property
    def kernel_launches(self):
        return self.sv.kernel_launches


class DistStateVector:
    def __init__(self, num_qubits, dtype=np.complex128, engine_factory=None, group=None, swap="auto"):
        import torch
        import torch.distributed as dist

        self.dist, self.torch = dist, torch
        self.group = group
        self._ctor = (engine_factory, swap)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.g = int(np.log2(self.world))
        if (1 << self.g) != self.world:
            raise ValueError("world size must be a power of two")
        self.n = num_qubits
        self.nloc = num_qubits - self.g
        if self.nloc < 1:
            raise ValueError("too few qubits for this many ranks")
        self.dtype = np.dtype(dtype)
        if engine_factory is None:
            device = torch.device("cuda", torch.cuda.current_device())
            self.engine = LocalEngine(self.nloc, dtype, device)
        else:
            self.engine = engine_factory(self.nloc, dtype)
        self.n_swaps = 0
        self.swap_bytes = 0
        # swap transport: "peer" = in-place exchange over NVLink peer memory (CUDA IPC, no staging buffers:
        # a 33-qubit c128 slab is 128 GiB and leaves no room for them); "nccl" = pack -> send/recv -> unpack.
        self.swap_mode = swap
        if swap == "auto":
            self.swap_mode = "peer" if hasattr(self.engine, "ipc_handle") else "nccl"
        if self.swap_mode == "peer":
            handles = [None] * self.world
            dist.all_gather_object(handles, self.engine.ipc_handle(), group=group)
            # PLB200_SWAP_MULTI=1: k global bits in ONE all-to-all exchange (needs every rank mapped).
            # Off by default: correct on 4 and 8 GPUs at 18 qubits (tests/test_dist_gpu.py), but its first
            # run at 33 local qubits did not finish inside the round's remaining GPU budget, so the
            # chained single-bit swaps (measured at 36 qubits) stay the default until that is understood.
            self.multi_swap = os.environ.get("PLB200_SWAP_MULTI", "0") == "1"
            # PLB200_SWAP_FUSED (default on when a second slab fits): the swap is folded into the store phase of
            # the last pass before it (plb200_sv_apply_ops_route): amplitudes go straight to their owner's
            # ping-pong slab over NVLink, no separate sweep.  Needs 2 x slab bytes of HBM per GPU.
            self.fused_swap = False
            if os.environ.get("PLB200_SWAP_FUSED", "1") != "0" and hasattr(self.engine, "alloc_alt"):
                free, _ = torch.cuda.mem_get_info()
                ok = torch.tensor([1.0 if (self.dtype.itemsize << self.nloc) * 1.05 < free else 0.0], device="cuda")
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                self.fused_swap = bool(ok.item() >= 1.0)
            every = self.multi_swap or self.fused_swap
            for r in range(self.world):
                if r != self.rank and (every or bin(r ^ self.rank).count("1") == 1):
                    self.engine.open_peer(r, handles[r])
            if self.fused_swap:
                self.engine.alloc_alt()
                alt_handles = [None] * self.world
                dist.all_gather_object(alt_handles, self.engine.ipc_handle_alt(), group=group)
                for r in range(self.world):
                    if r != self.rank:
                        self.engine.open_peer(r, alt_handles[r], alt=True)
            self.n_fused_swaps = 0
        self.reset()

    def close(self):
        """Collective: unmap every peer slab, then release the local one."""
        if hasattr(self.engine, "close_peers"):
            self.engine.sync()
            self.dist.barrier(group=self.group)
            self.engine.close_peers()
            self.dist.barrier(group=self.group)
        self.engine = None

    # ------------------------------------------------------------------ bookkeeping
    def reset(self):
        # wire w -> physical bit (wire 0 = most significant physical bit)
        self.phys = [self.n - 1 - w for w in range(self.n)]
        self.engine.zero()
        if self.rank == 0:
            self.engine.set_amp(0, 1.0)

    def relabel(self):
        """Forget the wire permutation WITHOUT moving amplitudes: the slabs are reinterpreted with wire w on
        physical bit n-1-w again (a different, equally valid state).  Benchmarks use it so that every timed
        step schedules the same tape from the same wire map (as a run from reset() does)."""
        self.phys = [self.n - 1 - w for w in range(self.n)]

    def _is_global(self, w):
        return self.phys[w] >= self.nloc

    def _rank_bit(self, w):
        return (self.rank >> (self.phys[w] - self.nloc)) & 1

    def _lw(self, w):
        """local wire index (engine convention: wire 0 = most significant local bit)"""
        return self.nloc - 1 - self.phys[w]

    @property
    def kernel_launches(self):
        return self.engine.kernel_launches

    # ------------------------------------------------------------------ op classification
    def _localize(self, op):
        """Try to express a normalised op on the local slab without communication.
        Returns ('skip', None) | ('ops', [engine op dicts]) | ('matrix', (matrix, wires, cw, cv)) |
        ('blocked', set of global wires that must become local)."""
        # Scheduling decisions must be identical on every rank: whether the op is blocked depends only on
        # the (rank-independent) wire -> bit map, so it is decided BEFORE the rank-dependent control test.
        tg = [w for w in op["targets"] if self._is_global(w)]
        d = None
        if tg or op["base"] == "GlobalPhase":
            d = None if op["matrix"] is not None else _diag_of(op["base"], op["params"], len(op["targets"]),
                                                               op["inverse"])
            if d is None:
                return "blocked", set(tg)
        cw, cv = [], []
        for w, v in zip(op["ctrl_wires"], op["ctrl_values"]):
            if self._is_global(w):
                if self._rank_bit(w) != int(v):
                    return "skip", None
            else:
                cw.append(w), cv.append(v)
        if not tg and op["base"] != "GlobalPhase":
            o = dict(name=op["base"], wires=[self._lw(w) for w in op["targets"]], params=op["params"],
                     inverse=op["inverse"], ctrl_wires=[self._lw(w) for w in cw], ctrl_values=cv)
            if op["matrix"] is not None:
                o["matrix"] = op["matrix"]
            return "ops", [o]
        # diagonal: fix the global target bits to this rank's values
        k = len(op["targets"])
        d = d.reshape((2,) * k) if k else d.reshape(())
        sl = tuple(self._rank_bit(w) if self._is_global(w) else slice(None) for w in op["targets"])
        d = np.asarray(d[sl] if k else d).reshape(-1)
        lt = [w for w in op["targets"] if not self._is_global(w)]
        if op["base"] == "GlobalPhase":
            lt, d = [], d[:1]
        if not lt:
            # scalar on the (controlled) subspace: 2x2 scalar matrix on any free local wire
            free = next(w for w in range(self.n) if not self._is_global(w) and w not in cw)
            if abs(d[0] - 1.0) == 0.0:
                return "skip", None
            return "matrix", (np.diag([d[0], d[0]]), [self._lw(free)], [self._lw(w) for w in cw], cv)
        return "matrix", (np.diag(d), [self._lw(w) for w in lt], [self._lw(w) for w in cw], cv)

    @staticmethod
    def _all_wires(op):
        return op["wires"]

    def _nondiag_targets(self, op):
        return () if op["diag"] else op["targets"]

    # ------------------------------------------------------------------ index-bit swap
    def _swap(self, gw, lw):
        """Exchange global wire gw with local wire lw (persistent: only the permutation remembers)."""
        dist, torch = self.dist, self.torch
        gb, lb = self.phys[gw], self.phys[lw]
        r = gb - self.nloc
        partner = self.rank ^ (1 << r)
        keep = (self.rank >> r) & 1  # this rank keeps local bit == its rank bit
        if self.swap_mode == "peer":
            # both partners exchange half of the pair range each, in place, with 128-bit peer loads/stores.
            # Device-side fences instead of host barriers: a stream-ordered all_reduce completes only when every
            # rank has finished what precedes it, so the partner's slab is quiescent before the exchange touches it
            # and complete before the next pass reads it — and the host keeps scheduling ahead
            self._fence_all()
            self.engine.swap_bit_peer(lb, keep, partner, 1 + keep)
            self._fence_all()
            self.phys[gw], self.phys[lw] = lb, gb
            self.n_swaps += 1
            self.swap_bytes += (1 << (self.nloc - 1)) * self.dtype.itemsize
            return
        send, recv = self.engine.buffers()
        self.engine.pack_bit(lb, keep, send)
        ops = [dist.P2POp(dist.isend, send, partner, group=self.group),
               dist.P2POp(dist.irecv, recv, partner, group=self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.engine.unpack_bit(lb, keep, recv)
        self.phys[gw], self.phys[lw] = lb, gb
        self.n_swaps += 1
        self.swap_bytes += send.numel() * send.element_size()

    def _fence_all(self):
        """Stream-ordered barrier over the ranks (no host synchronisation)."""
        if self.dist.get_backend(self.group) != "nccl":
            self.engine.sync()
            self.dist.barrier(group=self.group)
            return
        if not hasattr(self, "_fence"):
            self._fence = self.torch.zeros(1, device="cuda")
        self.dist.all_reduce(self._fence, group=self.group)

    def _swap_multi(self, pairs):
        """Exchange several (global wire, local wire) pairs in ONE all-to-all over peer memory:
        (1 - 2^-k) S per GPU on the links instead of k S/2 for k chained swaps."""
        dist = self.dist
        gbits = [self.phys[gw] for gw, _ in pairs]
        lbits = [self.phys[lw] for _, lw in pairs]
        k = len(pairs)
        my_value = sum(((self.rank >> (gb - self.nloc)) & 1) << i for i, gb in enumerate(gbits))
        partner_ranks = []
        for p in range(1 << k):
            r = self.rank
            for i, gb in enumerate(gbits):
                bit = 1 << (gb - self.nloc)
                r = (r | bit) if (p >> i) & 1 else (r & ~bit)
            partner_ranks.append(r)
        self._fence_all()
        self.engine.swap_bits_peer(lbits, my_value, partner_ranks)
        self._fence_all()
        for (gw, lw), gb, lb in zip(pairs, gbits, lbits):
            self.phys[gw], self.phys[lw] = lb, gb
        self.n_swaps += k
        self.swap_bytes += ((1 << self.nloc) - (1 << (self.nloc - k))) * self.dtype.itemsize

    # ------------------------------------------------------------------ tape execution
    def apply_ops(self, ops, fuse=True):
        # the same tape object applied again (a benchmark loop, a variational iteration re-using its list) is
        # normalised once
        key = (id(ops), len(ops))
        if getattr(self, "_norm_key", None) == key and len(ops) and self._norm_first is ops[0]:
            pending = list(self._norm_ops)
        else:
            pending = [normalize_op(o) for o in ops]
            if len(ops):
                self._norm_key, self._norm_ops, self._norm_first = key, list(pending), ops[0]
        while pending:
            batch, rest, blocked, need = [], [], set(), []
            for op in pending:
                wires = self._all_wires(op)
                if wires & blocked:
                    blocked |= wires
                    rest.append(op)
                    continue
                kind, payload = self._localize(op)
                if kind == "blocked":
                    blocked |= wires
                    rest.append(op)
                    for w in sorted(payload):
                        if w not in need:
                            need.append(w)
                elif kind == "ops":
                    batch += payload
                elif kind == "matrix":
                    m, w, cw, cv = payload
                    batch.append(dict(name="Matrix", wires=w, params=[], inverse=False, ctrl_wires=cw, ctrl_values=cv,
                                      matrix=m))
            pairs = self._choose_pairs(need, rest) if rest else []
            if pairs and fuse and getattr(self, "fused_swap", False) and len(pairs) <= 3:
                self._apply_and_swap_fused(batch, pairs)
            else:
                self.engine.apply_ops(batch, fuse=fuse)
                if pairs:
                    self._swap_pairs(pairs)
            pending = rest

    def _choose_pairs(self, need, rest):
        """(global wire, local wire) pairs to exchange: the needed global wires come in, the local wires with the
        farthest next non-diagonal use go out.  Depends only on the tape and the wire map: same on every rank."""
        nxt = {}
        for i, op in enumerate(rest):
            for w in self._nondiag_targets(op):
                nxt.setdefault(w, i)
        # at most PLB200_SWAP_MAX_BITS wires per exchange (default: all g global bits).  A routed 1-bit exchange sends
        # S/2 per GPU and hides behind the pass that carries it; a 3-bit one sends 7S/8 and does not — fewer, larger
        # exchanges are not always cheaper
        need = need[: max(1, min(self.g, int(os.environ.get("PLB200_SWAP_MAX_BITS", "3"))))]
        cand = [w for w in range(self.n) if not self._is_global(w) and w not in need]
        # the wires on the lowest local bits are evicted last (the sort key puts them behind every other
        # candidate): swapping a bit below the 128-B line splits every line (323 vs 692 GB/s measured)
        low = 3 if self.dtype == np.complex128 else 4
        cand.sort(key=lambda w: (self.phys[w] >= low, nxt.get(w, 1 << 30), self.phys[w]), reverse=True)
        return list(zip(need, cand))

    def _swap_pairs(self, pairs):
        if (self.swap_mode == "peer" and getattr(self, "multi_swap", False) and 1 < len(pairs) <= 3
                and hasattr(self.engine, "swap_bits_peer")):
            self._swap_multi(pairs)
        else:
            for gw, lw in pairs:
                self._swap(gw, lw)

    def _route_tables(self, pairs):
        gbits = [self.phys[gw] for gw, _ in pairs]
        lbits = [self.phys[lw] for _, lw in pairs]
        k = len(pairs)
        my_value = sum(((self.rank >> (gb - self.nloc)) & 1) << i for i, gb in enumerate(gbits))
        partner_ranks = []
        for p in range(1 << k):
            r = self.rank
            for i, gb in enumerate(gbits):
                bit = 1 << (gb - self.nloc)
                r = (r | bit) if (p >> i) & 1 else (r & ~bit)
            partner_ranks.append(r)
        return gbits, lbits, my_value, partner_ranks

    def _apply_and_swap_fused(self, batch, pairs):
        """The batch's last pass stores every amplitude into its owner's ping-pong slab (own or a peer's, over
        NVLink peer memory): the exchange costs no sweep of its own.  Falls back to the stand-alone swap kernels
        when the engine could not route (identical decision on every rank: it depends on the schedule only)."""
        dist = self.dist
        gbits, lbits, my_value, partner_ranks = self._route_tables(pairs)
        dst = [None if p == my_value else self.engine.peers_alt[r] for p, r in enumerate(partner_ranks)]
        routed = self.engine.apply_ops_route(batch, lbits, my_value, dst)
        if not routed:
            # NVRTC missing or a slab smaller than a tile: properties of the installation / the state size, the same
            # on every rank, so every rank takes this branch together
            self._swap_pairs(pairs)
            return
        # Device-side barrier, no host synchronisation: the reduction is stream-ordered after this rank's routed
        # kernel and completes only when every rank has issued its own, and the next pass on this stream waits for
        # it — so every peer's stores into this rank's ping-pong slab have landed before anything reads it, while
        # the host runs ahead and schedules the next segment (a host sync here idles the GPU for the ~10 ms of
        # Python that classify the rest of the tape: measured 8 ms per exchange at N = 8)
        self._fence_all()
        self.engine.flip_slabs()
        for (gw, lw), gb, lb in zip(pairs, gbits, lbits):
            self.phys[gw], self.phys[lw] = lb, gb
        k = len(pairs)
        self.n_swaps += k
        self.n_fused_swaps += k
        self.swap_bytes += ((1 << self.nloc) - (1 << (self.nloc - k))) * self.dtype.itemsize

    # ------------------------------------------------------------------ measurements
    def _allreduce(self, arr):
        t = self.torch.as_tensor(np.asarray(arr, dtype=np.float64))
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def expval_z_all(self):
        """<Z_w> for every wire (and the norm as a by-product): local reduction + all_reduce."""
        local = [w for w in range(self.n) if not self._is_global(w)]
        sums = np.asarray(self.engine.z_sums([self._lw(w) for w in local]))
        norm_local = sums[-1]
        out = np.zeros(self.n + 1)
        for w, s in zip(local, sums[:-1]):
            out[w] = s
        for w in range(self.n):
            if self._is_global(w):
                out[w] = (1 - 2 * self._rank_bit(w)) * norm_local
        out[self.n] = norm_local
        out = self._allreduce(out)
        self.last_norm2 = float(out[self.n])
        return out[: self.n]

    def norm2(self):
        self.expval_z_all()
        return self.last_norm2

    # -- general measurements on the sharded state (local reduction -> all_reduce of fp64 vectors;
    #    replaces MeasurementsGPUMPI.hpp / MeasurementsKokkosMPI: expval / probs / generate_samples)
    def _make_wires_local(self, wires):
        """Bring the given wires onto local bits (one exchange), evicting wires outside the set."""
        need = [w for w in wires if self._is_global(w)]
        if not need:
            return
        if len(set(wires)) > self.nloc:
            raise ValueError("more wires than local qubits: cannot be made local at once")
        cand = [w for w in range(self.n) if not self._is_global(w) and w not in wires]
        low = 3 if self.dtype == np.complex128 else 4
        cand.sort(key=lambda w: (self.phys[w] >= low, self.phys[w]), reverse=True)
        pairs = list(zip(need, cand))
        if getattr(self, "fused_swap", False) and len(pairs) <= 3 and self.nloc >= 13:
            self._apply_and_swap_fused([], pairs)
        else:
            self._swap_pairs(pairs)

    def expval_pauli_words(self, words, wires, coeffs=None):
        """<P_k> for Pauli words (strings over I/X/Y/Z) on the given wires; with coeffs: sum_k c_k <P_k>.
        X / Y letters must act on local wires (they are swapped in, word by word if needed); Z / I letters on
        global wires become a rank-dependent sign.  One local reduction per batch of words + ONE all_reduce."""
        out = np.zeros(len(words) + 1)
        batch = []  # (index, local word, local wires, sign)

        def flush():
            if not batch:
                return
            lw = [b[1] for b in batch] + ["I"]
            ww = [b[2] for b in batch] + [[0]]
            sums = self.engine.pauli_sums(lw, ww)
            for (k, _, _, sign), s_ in zip(batch, sums[:-1]):
                out[k] += sign * s_
            out[len(words)] = sums[-1]
            batch.clear()

        for k, (word, ws) in enumerate(zip(words, wires)):
            xy = [w for c, w in zip(word, ws) if c in "XY"]
            if any(self._is_global(w) for w in xy):
                flush()  # the layout changes: evaluate what was collected under the old one first
                self._make_wires_local(xy)
            sign, lword, lws = 1.0, "", []
            for c, w in zip(word, ws):
                if self._is_global(w):
                    if c == "Z" and self._rank_bit(w):
                        sign = -sign
                else:
                    lword += c
                    lws.append(self._lw(w))
            if not lws:
                free = next(w for w in range(self.n) if not self._is_global(w))
                lword, lws = "I", [self._lw(free)]
            batch.append((k, lword, lws, sign))
        flush()
        if out[len(words)] == 0.0:  # no word evaluated the norm slot (empty list)
            out[len(words)] = self.engine.pauli_sums(["I"], [[0]])[0]
        out = self._allreduce(out)
        self.last_norm2 = float(out[len(words)])
        vals = out[: len(words)]
        return float(np.dot(coeffs, vals)) if coeffs is not None else vals

    def expval_matrix(self, matrix, wires):
        """<psi| M |psi> for a Hermitian matrix on `wires` (MeasurementsGPUMPI::expval(matrix, wires)): the wires are
        made local (one exchange at most), every rank evaluates its slab, one all_reduce."""
        wires = list(wires)
        self._make_wires_local(wires)
        loc = self.engine.matrix_sum(np.asarray(matrix, dtype=np.complex128), [self._lw(w) for w in wires])
        nrm = self.engine.pauli_sums(["I"], [[0]])[0]
        out = self._allreduce([loc, nrm])
        self.last_norm2 = float(out[1])
        return float(out[0] / out[1])

    def probs(self, wires=None):
        """Marginal probabilities over `wires` (all wires when None), ordered by the given wire order
        (MeasurementsLQubit.hpp:90-163 semantics): local marginal over the local target wires, placed by this
        rank's values on the global target wires, then one all_reduce of the 2^k vector."""
        wires = list(range(self.n)) if wires is None else list(wires)
        k = len(wires)
        if k > 26:
            raise ValueError("probs over more than 26 wires is not gathered")
        loc = [w for w in wires if not self._is_global(w)]
        p_loc = np.asarray(self.engine.probs([self._lw(w) for w in loc])) if loc else np.array(
            [self.engine.pauli_sums(["I"], [[0]])[0]])
        out = np.zeros(1 << k)
        # index of a local outcome inside the full 2^k table: local wires keep their relative order
        pos = {w: k - 1 - i for i, w in enumerate(wires)}  # wire -> bit position in the output index
        base = 0
        for w in wires:
            if self._is_global(w) and self._rank_bit(w):
                base |= 1 << pos[w]
        idx = np.zeros(1 << len(loc), dtype=np.int64)
        for j, w in enumerate(loc):
            bit = (np.arange(1 << len(loc)) >> (len(loc) - 1 - j)) & 1
            idx |= bit << pos[w]
        np.add.at(out, idx + base, p_loc)
        return self._allreduce(out)

    def generate_samples(self, shots, seed=0):
        """Computational-basis samples of all wires, shape (shots, n), wire 0 first.  Per-rank mass -> the same
        multinomial split on every rank (shared seed, no communication) -> each rank draws its share from its
        slab's distribution -> local bit strings are translated through the wire map and gathered.  (Not the
        reference's alias-table stream: a sharded state has no single table; the distribution is the same.)"""
        mass = float(self.engine.pauli_sums(["I"], [[0]])[0])  # this slab's norm^2
        masses = self._allreduce(np.eye(self.world)[self.rank] * mass)
        masses = masses / masses.sum()
        counts = np.random.default_rng(seed).multinomial(shots, masses)
        mine = int(counts[self.rank])
        # the slab is sampled where it lives (LocalEngine: the table-free device sampler, no 2^nloc host table)
        local = self.engine.sample_bits(mine, seed * 1000003 + self.rank)  # (mine, nloc), engine wire order
        bits = np.zeros((mine, self.n), dtype=np.uint64)
        for w in range(self.n):
            if self._is_global(w):
                bits[:, w] = self._rank_bit(w)
            else:
                bits[:, w] = local[:, self._lw(w)]
        parts = [None] * self.world
        self.dist.all_gather_object(parts, bits, group=self.group)
        return np.concatenate(parts, axis=0)

    # -- adjoint Jacobian on the sharded state (AdjointJacobianGPUMPI.hpp / AdjointJacobianLQubit.hpp:347-491:
    #    lambda = U psi, H lambda per observable, reverse sweep with jac = -2 s Im<H lambda| G |lambda>)
    def clone(self, copy_state=True):
        """Collective: another sharded vector with the same wire map (and, with copy_state, the same amplitudes)."""
        factory, swap = self._ctor
        other = DistStateVector(self.n, self.dtype, engine_factory=factory, group=self.group, swap=swap)
        other.phys = list(self.phys)
        if copy_state:
            other.engine.copy_from(self.engine)
        else:
            other.engine.zero()
        return other

    def _assign(self, src):
        """self <- src (same world, same size): amplitudes and wire map"""
        self.phys = list(src.phys)
        self.engine.copy_from(src.engine)

    def inner(self, other):
        """<self|other> of two vectors with the same wire map: local dot + all_reduce"""
        if self.phys != other.phys:
            raise ValueError("inner product of sharded vectors with different wire maps")
        z = complex(self.engine.dot(other.engine))
        out = self._allreduce([z.real, z.imag])
        return complex(out[0], out[1])

    def _apply_pauli_word(self, word, wires):
        gates = [dict(name={"X": "PauliX", "Y": "PauliY", "Z": "PauliZ"}[c], wires=[w], params=[], inverse=False,
                      ctrl_wires=[], ctrl_values=[]) for c, w in zip(word, wires) if c != "I"]
        self.apply_ops(gates, fuse=False)

    def _apply_hamiltonian(self, coeffs, words, wires, scratch, followers=()):
        """H|self> for H = sum_k c_k P_k as a new sharded vector on self's wire map.  `scratch` (a clone) holds one
        term at a time; `followers` are vectors that must stay on the same map: every exchange this needs (X / Y
        letters on global wires) is made on all of them."""
        out = self.clone(copy_state=False)
        for c, word, ws in zip(coeffs, words, wires):
            xy = [w for ch, w in zip(word, ws) if ch in "XY"]
            for v in [self] + list(followers) + [out]:  # same call on every vector: the maps stay equal
                v._make_wires_local(xy)
            scratch._assign(self)
            scratch._apply_pauli_word(word, ws)
            assert scratch.phys == out.phys == self.phys
            out.engine.axpy(complex(c), scratch.engine)
        return out

    def var_pauli_hamiltonian(self, coeffs, words, wires):
        """Variance <H^2> - <H>^2 of H = sum_k c_k P_k on the sharded state (MeasurementsGPUMPI::var for a
        Hamiltonian: ||H psi||^2 - <psi|H psi>^2 with H psi built slab by slab).  The state's wire map may change
        (exchanges for X / Y letters on global wires); its amplitudes do not."""
        scratch = self.clone(copy_state=False)
        hpsi = self._apply_hamiltonian(coeffs, words, wires, scratch)
        h2 = hpsi.inner(hpsi).real
        h1 = self.inner(hpsi).real
        nrm = self.inner(self).real
        for v in (scratch, hpsi):
            v.close()
        return h2 / nrm - (h1 / nrm) ** 2

    def adjoint_jacobian(self, ops, trainable_params, observables):
        """Jacobian d<H_j>/d theta_p of the tape `ops` from |0...0>, for Hamiltonians of Pauli words
        `observables` = [(coeffs, words, wires), ...] and the trainable parameter indices `trainable_params`
        (indices into the tape's parametrised ops, as JacobianData's; one-parameter ops only).  Returns an array
        (len(observables), len(trainable_params)), identical on every rank.

        Every vector involved (lambda, one H lambda per observable, the scratch mu) is a sharded vector of this
        world; they are kept on ONE wire map by giving them the same sequence of operations — exchange decisions
        depend only on the tape and the map — so slab-local axpy / dot are meaningful.  Before a trainable op's
        generator is applied all its target wires are made local; a control on a global wire is a rank test."""
        tp = sorted(int(t) for t in trainable_params)
        n_par_ops = sum(1 for o in ops if len(o.get("params", ())) > 0)
        if any(t < 0 or t >= n_par_ops for t in tp):
            raise ValueError("trainable parameter index out of range")
        for o in ops:
            if len(o.get("params", ())) > 1:
                raise ValueError("The operation is not supported using the adjoint differentiation method")
        lam = self.clone(copy_state=False)
        lam.reset()
        lam.apply_ops(ops, fuse=True)
        mu = lam.clone(copy_state=False)
        hls = []
        for coeffs, words, wires in observables:
            hls.append(lam._apply_hamiltonian(coeffs, words, wires, mu, hls))
        jac = np.zeros((len(observables), len(tp)))
        tpi, cur = len(tp) - 1, n_par_ops - 1
        for o in reversed(ops):
            if o["name"] in ("StatePrep", "BasisState"):
                continue
            if tpi < 0:
                break
            if len(o.get("params", ())) > 0:
                if cur == tp[tpi]:
                    nop = normalize_op(o)
                    if nop["matrix"] is not None:
                        raise ValueError("The operation is not supported using the adjoint differentiation method")
                    for v in [lam] + hls:
                        v._make_wires_local(list(nop["targets"]))
                    mu._assign(lam)
                    cw, cv, dead = [], [], False
                    for w, val in zip(nop["ctrl_wires"], nop["ctrl_values"]):
                        if mu._is_global(w):
                            dead = dead or mu._rank_bit(w) != int(val)
                        else:
                            cw.append(mu._lw(w)), cv.append(bool(val))
                    # the generator of "control (x) base" is projector (x) generator(base), same scale factor
                    scale = mu.engine.apply_generator(nop["base"], [mu._lw(w) for w in nop["targets"]], False, cw, cv)
                    if dead:
                        mu.engine.zero()
                    sign = -1.0 if nop["inverse"] else 1.0
                    for j, hl in enumerate(hls):
                        jac[j, tpi] = -2.0 * sign * scale * hl.inner(mu).imag
                    tpi -= 1
                cur -= 1
            if tpi < 0:
                break
            inv = dict(o, inverse=not o.get("inverse", False))
            for v in [lam] + hls:
                v.apply_ops([inv], fuse=False)
        for v in [lam, mu] + hls:
            v.close()
        return jac

    def gather_state(self):
        """Full state in logical wire order on every rank (tests, small n only)."""
        loc = np.ascontiguousarray(self.engine.host_state())
        parts = [None] * self.world
        self.dist.all_gather_object(parts, loc, group=self.group)
        full = np.zeros(1 << self.n, dtype=loc.dtype)
        idx = np.arange(1 << self.n, dtype=np.int64)
        pidx = np.zeros_like(idx)
        for w in range(self.n):
            bit = (idx >> (self.n - 1 - w)) & 1
            pidx |= bit << self.phys[w]
        stacked = np.concatenate(parts)
        full[:] = stacked[pidx]
        return full
