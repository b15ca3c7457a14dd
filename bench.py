#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N=1 workload (config.workload): BASELINE.json configs[1] — 30-qubit random circuit
(RX/RY/RZ/CNOT/CRZ, depth 20 = 900 gates, seed 1234), c128, one B200.  A *step* is one pass of
the whole 900-gate tape over the 16 GiB state.  `value` = gates/s with the state resident in HBM
(device-timed, CUDA events on the launching stream); `e2e` = the same metric through the
reference-facing call sequence (reset -> applyOperations(host tape) -> expval(PauliZ(w)) for all
w -> host floats), host<->device copies inside the timed region.
N>1: weak scaling, 30 local qubits per GPU (n = 30 + log2 N), index bits sharded over ranks
(pennylane-lightning_b200/dist.py), one process per GPU under torchrun.

`--impl reference` times the reference's own lightning.qubit (oracle/_ref/liblq_ref.so, compiled
unmodified from the reference sources; OpenMP + AVX2/AVX-512 kernels) on the host cores, each
step a bounded sample of the same workload (the first layer = 45 gates of the same 30-qubit tape).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gates/sec, 30q c128 random circuit (RX/RY/RZ/CNOT/CRZ depth 20)"
UNIT = "gates/s"
LOCAL_QUBITS = int(os.environ.get("PLB200_BENCH_QUBITS", "30"))
DEPTH = 20
SEED = 1234
CPU_SAMPLE_GATES = 45  # one layer of the tape


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def nvlink_tx_rx_bytes(index=0):
    """NVLink data-payload counters of one GPU (sum over its links), from `nvidia-smi nvlink -gt d`; None when the
    tool or the counters are unavailable."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
        tx = rx = 0
        seen = False
        for line in out.splitlines():
            f = line.replace(":", " ").split()
            if "Tx" in f and "KiB" in f:
                tx += int(f[f.index("KiB") - 1]) * 1024
                seen = True
            if "Rx" in f and "KiB" in f:
                rx += int(f[f.index("KiB") - 1]) * 1024
        return (tx, rx) if seen else None
    except Exception:
        return None


# ------------------------------------------------------------------------------ reference arm
def time_reference(n, ops_sample, steps, warmup, threads=None):
    from oracle import lq_ref

    # torchrun exports OMP_NUM_THREADS=1 to its workers: the host-core count is set explicitly
    lq_ref.set_num_threads(threads or host_cores())
    cores = lq_ref.num_threads()
    sv = lq_ref.StateVector(n, np.complex128)
    blob = lq_ref.OpsBlob(ops_sample)
    times = []
    for it in range(warmup + steps):
        sv.reset()
        t0 = time.perf_counter()
        sv.apply_ops(blob)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return len(ops_sample) / (sum(times) / len(times)), cores, float(np.mean(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pennylane_lightning_b200 import circuits

    n = LOCAL_QUBITS
    ops = circuits.random_circuit(n, DEPTH, SEED)[:CPU_SAMPLE_GATES]
    steps = max(1, args.steps)
    gps, cores, sec = time_reference(n, ops, steps, max(1, min(args.warmup, 1)))
    sample = (f"first {CPU_SAMPLE_GATES} gates (layer 1) of the same {n}-qubit tape per step, lightning.qubit "
              f"LM/AVX kernels with OpenMP over the amplitude loop")
    line = {
        "impl": "reference", "metric": METRIC, "value": gps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 1)), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": f"{n}q random circuit RX/RY/RZ/CNOT/CRZ depth {DEPTH} seed {SEED}, c128",
                   "qubits": n, "gates": CPU_SAMPLE_GATES, "sample_of_gates": DEPTH * (n + n // 2)},
        "cpu_baseline": {"value": gps, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------- parity at benchmark size
def check_config2(plb, lq_ref, circuits, n, stream, fuse=True):
    """Outside the timed region: the config-2 tape at n qubits from |0..0>, norm and <Z_w> for every wire
    on the engine vs the reference's lightning.qubit (oracle/_ref) on the same tape."""
    import torch

    ops = circuits.random_circuit(n, DEPTH, SEED)
    t0 = time.perf_counter()
    sv = plb.StateVector(n, np.complex128, torch.cuda.current_device(), stream)
    sv.apply_ops(plb.OpsBlob(ops), fuse=fuse)
    ours = np.asarray(sv.expval_pauli_words_each(["Z"] * n + ["I"], [[w] for w in range(n)] + [[0]]))
    del sv
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    lq_ref.set_num_threads(host_cores())
    r = lq_ref.StateVector(n, np.complex128)
    r.apply_ops(lq_ref.OpsBlob(ops))
    ref = np.array([r.expval_named("PauliZ", [w]) for w in range(n)])
    view = r._view()
    nrm = 0.0
    step = 1 << 24
    for i in range(0, len(view), step):  # chunked: no 16 GiB temporary
        c = view[i:i + step]
        nrm += float(np.vdot(c, c).real)
    del r, view
    t_ref = time.perf_counter() - t0
    err = float(np.max(np.abs(ours[:n] - ref)))
    return {"qubits": n, "gates": len(ops), "max_abs_err_expval_z_all_wires": err,
            "norm2_minus_1": float(ours[n] - 1.0), "ref_norm2_minus_1": nrm - 1.0,
            "expval_z0": float(ours[0]), "ref_expval_z0": float(ref[0]), "tolerance": 1e-12,
            "pass": bool(err <= 1e-12 and abs(ours[n] - 1.0) <= 1e-12), "reference": "oracle/_ref (lightning.qubit)",
            "gpu_s": t_gpu, "ref_s": t_ref, "ref_cores": lq_ref.num_threads()}


def check_sharded(plb, lq_ref, circuits, dist_mod, DistStateVector, rank, world, n=24):
    """N>1, before timing: the same circuit family at n qubits sharded over the ranks vs one GPU vs the
    reference, every amplitude (the one-GPU test lease skips tests/test_dist_gpu.py)."""
    ops = circuits.random_circuit(n, DEPTH, SEED)
    d = DistStateVector(n, np.complex128)
    d.apply_ops(ops, fuse=True)
    z = d.expval_z_all()
    full = d.gather_state()
    swaps = d.n_swaps
    d.close()
    del d
    out = None
    if rank == 0:
        single = plb.StateVector(n, np.complex128, 0)
        single.apply_ops(plb.OpsBlob(ops), fuse=True)
        s1 = single.get_state()
        del single
        out = {"qubits": n, "ranks": world, "swaps": swaps,
               "max_abs_err_vs_single_gpu": float(np.max(np.abs(full - s1))), "tolerance": 1e-12}
        try:
            lq_ref.set_num_threads(host_cores())
            r = lq_ref.StateVector(n, np.complex128)
            r.apply_ops(lq_ref.OpsBlob(ops))
            out["max_abs_err_vs_reference"] = float(np.max(np.abs(full - r.get_state())))
            out["max_abs_err_expval_z"] = float(max(abs(z[w] - r.expval_named("PauliZ", [w])) for w in range(n)))
        except Exception as exc:
            out["reference_unavailable"] = str(exc)
        errs = [v for k, v in out.items() if k.startswith("max_abs_err")]
        out["pass"] = bool(max(errs) <= 1e-12)
    return out


# ----------------------------------------------------- configs 3 and 5 inside the driver-run line
def extra_qft(plb, circuits, n, dtype, tag, stream, peak):
    import torch

    free, _ = torch.cuda.mem_get_info()
    need = (1 << n) * (16 if dtype == np.complex128 else 8)
    if need > free * 0.97:
        return {"skipped": f"state needs {need / 2**30:.0f} GiB, {free / 2**30:.0f} GiB free"}
    rng = np.random.default_rng(7)
    x = int(rng.integers(0, 1 << 62)) % (1 << n)
    bits = [(x >> (n - 1 - w)) & 1 for w in range(n)]
    ops = circuits.qft(n)
    blob = plb.OpsBlob(ops)
    sv = plb.StateVector(n, dtype, torch.cuda.current_device(), stream)
    best = None
    for it in range(4):  # runs 1-2 let the pass kernels be compiled (second sighting), 3-4 use them
        if it == 2 and plb.jit_enabled():
            torch.cuda.synchronize()
            plb.jit_wait()
        sv.set_basis_state(bits, list(range(n)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sv.apply_ops(blob, fuse=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    passes = sv.last_apply_stats()[1]
    head = sv.get_state(1 << 16)
    del sv
    ph = np.array([((x * kk) % (1 << n)) / float(1 << n) for kk in range(1 << 16)])
    exact = 2.0 ** (-n / 2) * np.exp(2j * np.pi * ph)
    err = float(np.max(np.abs(head - exact)) / 2.0 ** (-n / 2))
    tol = 1e-12 if dtype == np.complex128 else 1e-5
    return {"qubits": n, "dtype": tag, "gates": len(ops), "seconds": best * 1e-3, "gates_per_s": len(ops) / (best * 1e-3),
            "hbm_passes": passes, "roofline_frac": passes * 2 * need / (best * 1e-3) / 1e9 / peak,
            "roofline_def": "passes x 2S / time / peak", "max_rel_err_vs_analytic_2^16_amplitudes": err,
            "tolerance": tol, "pass": bool(err <= tol)}


def extra_c64(plb, circuits, stream, peak, n=30, steps=5):
    """The headline tape in single precision (c64): device-resident fused rate and per-pass HBM fraction, with
    <Z_w> on every wire compared against the c128 run of the same tape (tolerance 1e-5, north_star's for c64)."""
    import torch

    ops = circuits.random_circuit(n, DEPTH, SEED)
    blob = plb.OpsBlob(ops)
    zw = [[w] for w in range(n)]
    ref = plb.StateVector(n, np.complex128, torch.cuda.current_device(), stream)
    ref.apply_ops(blob, fuse=True)
    z128 = np.asarray(ref.expval_pauli_words_each(["Z"] * n, zw))
    del ref
    torch.cuda.empty_cache()
    sv = plb.StateVector(n, np.complex64, torch.cuda.current_device(), stream)
    for it in range(3):  # sightings 1-2 let the kernels be compiled, 3 loads them
        if it == 2 and plb.jit_enabled():
            torch.cuda.synchronize()
            plb.jit_wait()
        sv.reset()
        sv.apply_ops(blob, fuse=True)
    z64 = np.asarray(sv.expval_pauli_words_each(["Z"] * n, zw))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sv.apply_ops(blob, fuse=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    passes = sv.last_apply_stats()[1]
    S = (1 << n) * 8
    err = float(np.max(np.abs(z64 - z128)))
    del sv
    return {"qubits": n, "dtype": "c64", "gates": len(ops), "ms_per_step": ms, "gates_per_s": len(ops) / (ms * 1e-3),
            "hbm_passes": passes, "roofline_frac": passes * 2 * S / (ms * 1e-3) / 1e9 / peak,
            "roofline_def": "passes x 2S / time / peak", "max_abs_err_expval_z_vs_c128": err, "tolerance": 1e-5,
            "pass": bool(err <= 1e-5)}


def extra_sampling(plb, circuits, stream, peak, n=30, shots=100000):
    """Computational-basis samples of the config-2 state with the table-free device sampler (one sweep for the chunk
    masses, a scan, one warp per shot); checked through every single-wire marginal against <Z_w>."""
    import torch

    sv = plb.StateVector(n, np.complex128, torch.cuda.current_device(), stream)
    sv.apply_ops(plb.OpsBlob(circuits.random_circuit(n, DEPTH, SEED)), fuse=True)
    sv.sync()
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        s = sv.generate_samples(shots, seed=5, device=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    z = np.asarray(sv.expval_pauli_words_each(["Z"] * n, [[w] for w in range(n)]))
    err = float(np.max(np.abs((1.0 - 2.0 * s.mean(axis=0)) - z)))
    S = (1 << n) * 16
    del sv
    return {"qubits": n, "shots": shots, "seconds": best, "shots_per_s": shots / best,
            "roofline_frac_of_one_sweep": S / best / 1e9 / peak,
            "max_abs_err_single_wire_marginals_vs_expval_z": err, "tolerance": 5.0 / np.sqrt(shots),
            "pass": bool(err <= 5.0 / np.sqrt(shots)),
            "what": "plb200_generate_samples_device, samples returned to the host (shots x n uint64)"}


def extra_adjoint(plb, lq_ref, circuits, stream, peak, n=24, n_params=1000, ref_params=40):
    import torch

    ops, tp = circuits.hardware_efficient_ansatz(n, n_params, 99)
    co, words, wires = circuits.pauli_hamiltonian(n, 100, 99)
    out = {"qubits": n, "params": n_params, "hamiltonian_terms": 100}
    jac128 = None
    for dt, tag in ((np.complex128, "c128"), (np.complex64, "c64")):
        sv = plb.StateVector(n, dt, torch.cuda.current_device(), stream)
        ham = circuits.hamiltonian_observable(plb, co, words, wires)
        blob = plb.OpsBlob(ops)
        sv.apply_ops(blob)
        sv.sync()
        best = None
        for it in range(5):  # sweeps 1-2 let the two-state pass kernels be compiled, 3-5 use them
            if it == 2 and plb.jit_enabled():
                plb.jit_wait()
            t0 = time.perf_counter()
            jac = sv.adjoint_jacobian([ham], blob, tp)
            dt_s = time.perf_counter() - t0
            best = dt_s if best is None else min(best, dt_s)
        S = (1 << n) * (16 if dt == np.complex128 else 8)
        passes, alone = sv.last_apply_stats()  # two-state tile passes, stand-alone items of the last sweep
        out[tag] = {"adjoint_s": best, "two_state_passes": passes, "stand_alone_items": alone,
                    "roofline_frac": (passes * 4 * S + alone * 2 * S) / best / 1e9 / peak,
                    "roofline_def": "(passes x 4S [lambda and H lambda read + written] + stand-alone x 2S) / time / peak",
                    "effective_multiplier_vs_6S_per_parameter": 6 * S * n_params / max(1.0, passes * 4 * S + alone * 2 * S),
                    "expval": float(sv.expval(ham)), "jac_norm": float(np.linalg.norm(jac))}
        if dt == np.complex128:
            jac128 = jac
        else:
            out[tag]["max_rel_err_vs_c128"] = float(np.max(np.abs(jac - jac128)) / np.max(np.abs(jac128)))
        del sv
    try:  # bounded lightning.qubit sample beside it: the last `ref_params` trainable parameters of the same tape
        lq_ref.set_num_threads(host_cores())
        sub = tp[-ref_params:]
        rsv = lq_ref.StateVector(n, np.complex128)
        rham = circuits.hamiltonian_observable(lq_ref, co, words, wires, dtype=np.complex128)
        rblob = lq_ref.OpsBlob(ops)
        rsv.apply_ops(rblob)
        t0 = time.perf_counter()
        rj = rsv.adjoint_jacobian([rham], rblob, sub)
        t_r = time.perf_counter() - t0
        err = float(np.max(np.abs(rj.ravel() - jac128.ravel()[-ref_params:])) / np.max(np.abs(rj)))
        out["reference_sample"] = {"what": f"lightning.qubit adjoint of the same tape, last {ref_params} of the "
                                           f"{n_params} parameters trainable", "seconds": t_r,
                                   "cores": lq_ref.num_threads(), "max_rel_err_vs_ours": err, "tolerance": 1e-12,
                                   "pass": bool(err <= 1e-12),
                                   "ref_expval": float(rsv.expval(rham))}
    except Exception as exc:
        out["reference_sample"] = {"unavailable": str(exc)}
    return out


# ------------------------------------------------------------------------------------ our arm
def kernel_family(o):
    return "diag_kernel" if o["name"] in ("RZ", "CRZ") else "pairs_kernel"


def config4_block(plb, circuits, DistStateVector, dist, torch, rank, world, stream, nloc=33):
    """BASELINE config 4: weak scaling at 33 local qubits per GPU (36 qubits on 8 GPUs), ONE timed step after
    one warm-up step, next to a 33-qubit single-GPU step measured on rank 0 in the same job."""
    free, _ = torch.cuda.mem_get_info()
    need = (1 << nloc) * 16
    ok = torch.tensor([1.0 if need <= free * 0.97 else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() < 1:
        return {"skipped": f"a {nloc}-qubit slab needs {need / 2**30:.0f} GiB, {free / 2**30:.0f} GiB free"}
    g = int(np.log2(world))
    t1 = None
    if rank == 0:
        ops1 = circuits.random_circuit(nloc, DEPTH, SEED)
        one = plb.StateVector(nloc, np.complex128, torch.cuda.current_device(), stream)
        blob = plb.OpsBlob(ops1)
        for it in range(2):
            one.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            one.apply_ops(blob, fuse=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter() - t0
        del one
        torch.cuda.empty_cache()
    dist.barrier()
    n = nloc + g
    ops = circuits.random_circuit(n, DEPTH, SEED)
    d = DistStateVector(n, np.complex128)
    tN = None
    for it in range(2):
        d.reset()
        s0, b0 = d.n_swaps, d.swap_bytes
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        d.apply_ops(ops, fuse=True)
        torch.cuda.synchronize()
        tN = time.perf_counter() - t0
        swaps, sbytes = d.n_swaps - s0, d.swap_bytes - b0
    z = d.expval_z_all()
    norm2 = d.last_norm2
    t = torch.tensor([tN], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tN = float(t.item())
    d.close()
    del d
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"qubits": n, "local_qubits": nloc, "gates": len(ops), "wall_s": tN, "gates_per_s": len(ops) / tN,
            "index_bit_swaps": swaps, "nvlink_bytes_per_swap_per_gpu": sbytes // max(1, swaps),
            "single_gpu_33q": {"gates": len(ops1), "wall_s": t1},
            "weak_scaling_efficiency": t1 * (len(ops) / len(ops1)) / tN,
            "efficiency_def": "T(33q, 1 GPU) x gates_n / gates_33 / T(n, N GPUs) (SURVEY 8d row 4)",
            "norm2_minus_1": norm2 - 1.0, "expval_z0": float(z[0])}


def sv_kernel_name():
    try:
        import pennylane_lightning_b200 as plb

        return "jit_pass_kernel (NVRTC-specialised fused pass)" if plb.jit_enabled() else "tile_kernel (fused pass, interpreter)"
    except Exception:
        return "tile_kernel (fused pass)"


def run_ours(args):
    import torch

    import pennylane_lightning_b200 as plb
    from pennylane_lightning_b200 import circuits

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nloc = LOCAL_QUBITS
    g = int(np.log2(world))
    n = nloc + g
    ops = circuits.random_circuit(n, DEPTH, SEED)
    n_gates = len(ops)
    stream = torch.cuda.current_stream().cuda_stream
    fuse = not args.no_fuse

    if world > 1:
        from pennylane_lightning_b200.dist import DistStateVector

        sv = DistStateVector(n, np.complex128)

        def apply_tape():
            # every step starts from the identity wire map, as a circuit run from reset() does (the amplitudes are
            # left where the previous step put them: relabel() moves nothing), so the tape is scheduled into the
            # same passes / exchanges every step
            sv.relabel()
            sv.apply_ops(ops, fuse=fuse)

        launches = lambda: sv.kernel_launches
        expvals = lambda: sv.expval_z_all()
        reset = sv.reset
    else:
        sv = plb.StateVector(n, np.complex128, local_rank, stream)
        blob = plb.OpsBlob(ops)
        apply_tape = lambda: sv.apply_ops(blob, fuse=fuse)
        launches = lambda: sv.kernel_launches
        zw = [[w] for w in range(n)]
        expvals = lambda: sv.expval_pauli_words_each(["Z"] * n, zw)
        reset = sv.reset

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    checks = {}
    want_checks = os.environ.get("PLB200_BENCH_CHECKS", "1") != "0"
    if world > 1 and want_checks:
        from oracle import lq_ref

        try:
            checks["sharded_24q"] = check_sharded(plb, lq_ref, circuits, dist, DistStateVector, rank, world, 24)
        except Exception as exc:
            checks["sharded_24q"] = {"pass": False, "error": repr(exc)}
        barrier()

    # ---- device-resident timing -------------------------------------------------------
    reset()
    for _ in range(args.warmup):
        apply_tape()
    if plb.jit_enabled():
        # pass kernels are compiled in the background from the second sighting of a pass structure: wait for
        # the queue to drain, then one more untimed step loads the modules
        torch.cuda.synchronize()
        plb.jit_wait()
        apply_tape()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = launches()
    nv0 = nvlink_tx_rx_bytes(local_rank) if (world > 1 and rank == 0) else None
    swaps0 = sv.n_swaps if world > 1 else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        apply_tape()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    nvlink_measured = None
    if world > 1 and rank == 0:
        nv1 = nvlink_tx_rx_bytes(local_rank)
        if nv0 and nv1:
            dsw = max(1, sv.n_swaps - swaps0)
            nvlink_measured = {"tx_bytes_per_step": (nv1[0] - nv0[0]) / args.steps, "rx_bytes_per_step": (nv1[1] - nv0[1]) / args.steps,
                               "tx_bytes_per_swap": (nv1[0] - nv0[0]) / dsw, "source": "nvidia-smi nvlink -gt d, GPU of rank 0, "
                               "difference over the timed region (includes NCCL's barrier / reduction traffic)"}
    gpu_launches = launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    # units all ranks processed / time: every rank applies every gate of the tape to its own
    # 2^LOCAL_QUBITS-amplitude slab, so the job processes world * n_gates slab-gates per step
    # (identical to plain gates/s at N=1).
    value = world * n_gates * args.steps / (ms * 1e-3)

    # ---- end-to-end through the public call sequence, host buffers ----------------------
    e2e_steps = max(3, min(args.steps, 10))
    reset()  # one untimed end-to-end step (first-touch of the host-side buffers)
    (sv.apply_ops(ops, fuse=fuse) if world > 1 else sv.apply_ops(plb.OpsBlob(ops), fuse=fuse))
    expvals()
    tape_bytes = sum(8 * (len(o["wires"]) + len(o["params"]) + len(o["ctrl_wires"])) + len(o["name"]) + 2
                     for o in ops)
    barrier()
    t0 = time.perf_counter()
    e2e_step_ms = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        reset()
        if world > 1:
            sv.apply_ops(ops, fuse=fuse)
        else:
            sv.apply_ops(plb.OpsBlob(ops), fuse=fuse)  # marshals the host tape every step
        ez = np.asarray(expvals())
        e2e_step_ms.append((time.perf_counter() - ts) * 1e3)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n_gates * e2e_steps / e2e_s

    # ---- roofline of the dominant kernel, live CUDA events ---------------------------------
    # Fused mode: every launch of the timed region is tile_kernel (one HBM sweep for ~43 gates).
    #   achieved = algorithmic bytes of the gates a launch processes (SURVEY section 8(d) table:
    #   2S per full-touch gate, S per singly-controlled gate ...) / launch duration  -> "effective
    #   HBM GB/s"; it exceeds the HBM peak exactly because fusion avoids the per-gate sweeps.
    #   traffic  = dram bytes one launch really moves (ncu --set full capture in profiles/).
    # The un-fused per-gate kernels (one sweep per gate: the reference's execution model) are timed
    # gate by gate as well and reported under per_gate_kernels.
    roofline = None
    if rank == 0 and world == 1:
        peak, peak_src = measured_peak()
        fam_ms, fam_bytes, fam_n = {}, {}, {}
        evs = []
        for o in ops:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sv.apply(o["name"], o["wires"], o["inverse"], o["params"])
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        for o, (a, b) in zip(ops, evs):
            f = kernel_family(o)
            fam_ms[f] = fam_ms.get(f, 0.0) + a.elapsed_time(b)
            fam_bytes[f] = fam_bytes.get(f, 0.0) + circuits.algorithmic_bytes(o, n)
            fam_n[f] = fam_n.get(f, 0) + 1
        tot_ms = sum(fam_ms.values())
        per_gate = {f: {"achieved_GBps": fam_bytes[f] / (fam_ms[f] * 1e-3) / 1e9,
                        "frac_of_peak": fam_bytes[f] / (fam_ms[f] * 1e-3) / 1e9 / peak, "launches": fam_n[f],
                        "avg_launch_ms": fam_ms[f] / fam_n[f],
                        "avg_algorithmic_bytes_per_launch": fam_bytes[f] / fam_n[f]} for f in fam_ms}
        per_gate["unfused_gates_per_s"] = n_gates / (tot_ms * 1e-3)
        alg_total = sum(fam_bytes.values())
        if fuse:
            sv.apply_ops(blob, fuse=True)
            torch.cuda.synchronize()
            passes = sv.last_apply_stats()[1]
            S = (1 << n) * 16
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "tile_kernel_traffic.json")
            if os.path.exists(tpath):
                try:
                    tj = json.load(open(tpath))
                    traffic, traffic_src = tj.get(str(n)), tj.get("source")
                except Exception:
                    traffic = None
            # every launch of the timed region is one fused pass: it reads and writes the state once
            launch_ms = ms_per_step / max(1, passes)
            ach = 2 * S / (launch_ms * 1e-3) / 1e9
            eff = alg_total / (ms_per_step * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": sv_kernel_name(), "achieved": ach, "peak": peak,
                        "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": traffic,
                        "traffic_source": traffic_src,
                        "algorithmic_bytes_per_launch": 2 * S, "launches_per_step": passes, "avg_launch_ms": launch_ms,
                        "dram_GBps": (traffic / (launch_ms * 1e-3) / 1e9) if traffic else None,
                        "fusion_multiplier": eff / ach,
                        "effective_GBps_vs_one_sweep_per_gate": eff,
                        "note": "achieved = 2S (one read + one write of the state per fused pass) / average launch "
                                "time, live CUDA events over the timed region. fusion_multiplier = SURVEY 8d bytes of "
                                "the gates a pass applies (what one sweep per gate would move) / 2S: how many sweeps "
                                "a pass replaces; it is not a roofline fraction",
                        "per_gate_kernels": per_gate}
        else:
            dom = max(fam_ms, key=fam_ms.get)
            roofline = {"bound": "hbm", "kernel": dom, "achieved": per_gate[dom]["achieved_GBps"], "peak": peak,
                        "unit": "GB/s", "frac": per_gate[dom]["frac_of_peak"], "peak_source": peak_src,
                        "traffic": per_gate[dom]["avg_algorithmic_bytes_per_launch"],
                        "launches_per_step": fam_n[dom], "avg_launch_ms": per_gate[dom]["avg_launch_ms"],
                        "avg_algorithmic_bytes_per_launch": per_gate[dom]["avg_algorithmic_bytes_per_launch"],
                        "per_gate_kernels": per_gate}

    # ---- CPU baseline on the box's host cores (bounded sample) ---------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sample_ops = circuits.random_circuit(nloc, DEPTH, SEED)[:CPU_SAMPLE_GATES]
            gps, cores, sec = time_reference(nloc, sample_ops, 1, 1)
            cpu_baseline = {"value": gps, "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": f"first {CPU_SAMPLE_GATES} gates (layer 1) of the same {nloc}-qubit tape, "
                                      f"1 warm-up + 1 timed pass ({sec:.1f} s), lightning.qubit OpenMP+AVX"}
        except Exception as exc:  # oracle/_ref missing on this box
            cpu_baseline = {"value": None, "unit": UNIT, "cores": None, "kind": "reference",
                            "sample": f"unavailable: {exc}"}

    extra = None
    stats = sv.last_apply_stats() if world == 1 else (n_gates, None)
    n_swaps, swap_bytes = (sv.n_swaps, sv.swap_bytes) if world > 1 else (0, 0)

    class _Stats:
        n_fused_swaps = getattr(sv, "n_fused_swaps", 0)

    sv_ref_for_stats = _Stats
    def make_line(extra_value):
        return {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": f"{n}q random circuit RX/RY/RZ/CNOT/CRZ depth {DEPTH} seed {SEED}, c128, "
                                   f"{nloc} local qubits per GPU",
                       "qubits": n, "gates": n_gates, "fused": fuse, "hbm_passes_per_step": stats[1],
                       "l2": "state (16 GiB per GPU) is far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"index-bit sharding over {world} GPU(s)",
                       "value_definition": "ranks x gates / time: each rank applies every gate to its own "
                                           f"2^{nloc}-amplitude slab (= plain gates/s at N=1)",
                       "circuit_gates_per_s": n_gates * args.steps / (ms * 1e-3),
                       "fused_index_bit_swaps_total": getattr(sv_ref_for_stats, "n_fused_swaps", 0) if world > 1 else 0,
                       "index_bit_swaps_per_step": (n_swaps // max(1, args.warmup + args.steps + e2e_steps + 1))
                       if world > 1 else 0,
                       "nvlink_bytes_per_swap_per_gpu": (swap_bytes // max(1, n_swaps)) if world > 1 and
                       n_swaps else 0,
                       "nvlink_counters": nvlink_measured},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": tape_bytes,
                    "d2h_bytes_per_step": 8 * n, "steps": e2e_steps, "step_ms": e2e_step_ms,
                    "what": "reset + applyOperations(host tape) + expval(PauliZ(w)) for every wire -> host"},
            "gpu_launches": int(gpu_launches), "clocks": clocks, "jit": plb.jit_stats(),
            "checks": checks, "extra": extra_value,
        }

    if world > 1 and os.environ.get("PLB200_BENCH_CONFIG4", "1") != "0":
        sv.close()
        del sv
        torch.cuda.empty_cache()
        key4 = f"config4_{33 + g}q_{world}gpu"
        limit = float(os.environ.get("PLB200_BENCH_CONFIG4_TIMEOUT", "420"))

        def _watchdog():
            # a 128-GiB-slab run that does not come back must not cost the line of the run that did
            if rank == 0:
                print(json.dumps(make_line({key4: {"error": f"no result within {limit:.0f} s (watchdog)"}})), flush=True)
            os._exit(0)

        timer = threading.Timer(limit, _watchdog)
        timer.daemon = True
        timer.start()
        try:
            c4 = config4_block(plb, circuits, DistStateVector, dist, torch, rank, world, stream)
        except Exception as exc:
            c4 = {"error": repr(exc)}
        timer.cancel()
        if rank == 0:
            extra = {key4: c4}
        sv = None
    if rank == 0 and world == 1:
        peak, _ = measured_peak()
        del sv, blob
        sv = None
        torch.cuda.empty_cache()
        if want_checks:
            from oracle import lq_ref

            try:
                import psutil

                avail = psutil.virtual_memory().available
            except Exception:
                avail = 0
            nchk = int(os.environ.get("PLB200_BENCH_CHECK_QUBITS", "0")) or (nloc if avail >= (40 << 30) else min(nloc, 28))
            try:
                checks["config2_vs_lightning_qubit"] = check_config2(plb, lq_ref, circuits, nchk, stream, fuse)
            except Exception as exc:
                checks["config2_vs_lightning_qubit"] = {"pass": False, "error": repr(exc)}
        if os.environ.get("PLB200_BENCH_EXTRA", "1") != "0":
            from oracle import lq_ref

            extra = {}
            for key, fn in (("config3_qft33_c128", lambda: extra_qft(plb, circuits, 33, np.complex128, "c128", stream, peak)),
                            ("config3_qft33_c64", lambda: extra_qft(plb, circuits, 33, np.complex64, "c64", stream, peak)),
                            ("config5_adjoint_24q_1000", lambda: extra_adjoint(plb, lq_ref, circuits, stream, peak)),
                            ("config2_30q_c64", lambda: extra_c64(plb, circuits, stream, peak, nloc)),
                            ("sampling_30q_device", lambda: extra_sampling(plb, circuits, stream, peak, nloc))):
                try:
                    extra[key] = fn()
                except Exception as exc:
                    extra[key] = {"error": repr(exc)}
                torch.cuda.empty_cache()

    if rank == 0:
        print(json.dumps(make_line(extra)), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-fuse", action="store_true", help="one kernel per gate (per-gate roofline mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()
