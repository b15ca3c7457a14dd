/*
 * plb200.h — C ABI of the B200-native state-vector engine (libplb200.so).
 *
 * This is the drop-in boundary for pennylane-lightning's hot path: gate
 * application on a 2^n complex amplitude vector, the measurement reductions on it, and the
 * adjoint-Jacobian sweep.  Every entry point is `extern "C"`, takes plain pointers and
 * sizes only (no C++/torch types) and returns an int status: 0 = ok, non-zero = failure with
 * the message available from plb200_last_error() (thread-local).  The C++ classes in
 * pennylane-lightning_b200/host/ (StateVectorB200 / Measurements / AdjointJacobian, the
 * mirrors of the reference's CRTP interfaces) convert a non-zero status into the reference's
 * LightningException text, as PL_ABORT does (core/utils/Error.hpp:111-139).
 *
 * Conventions shared with the reference:
 *  - one contiguous array of interleaved (re,im); wire w <-> index bit (n-1-w), wire 0 = MSB
 *    (GateImplementationsLM.hpp:690-699);
 *  - gate matrices row-major, wires[0] = most-significant matrix bit (StateVectorBase.hpp:195);
 *  - params and matrices cross the ABI as double / interleaved complex128 for both
 *    precisions (the engine narrows to fp32 for a c64 state);
 *  - all device work is enqueued on the stream given at creation (one stream per state
 *    vector, as DevTag does in core/utils/cuda_utils/DevTag.hpp); calls that return host
 *    data synchronise that stream before returning.
 *
 * There is NO CPU fallback: every compute entry point fails with a CUDA error when no
 * sm_100 device is present.
 */
#ifndef PLB200_H
#define PLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plb200_sv plb200_sv;   /* opaque state vector (device resident) */
typedef struct plb200_obs plb200_obs; /* opaque observable tree (host resident) */

#define PLB200_C64 32  /* complex<float>  */
#define PLB200_C128 64 /* complex<double> */

/* Flattened tape of operations: the C image of OpsData<StateVectorT>
 * (core/algorithms/JacobianData.hpp:39-253) and of the argument lists of
 * StateVectorBase::applyOperations (core/simulators/base/StateVectorBase.hpp:116-160).
 * Arrays *_off have n_ops+1 entries (CSR style). mats_off counts complex elements. */
typedef struct plb200_ops_t {
    int64_t n_ops;
    const char *const *names;
    const int64_t *wires;
    const int64_t *wires_off;
    const int64_t *ctrl_wires;
    const int64_t *ctrl_off;
    const uint8_t *ctrl_values; /* indexed like ctrl_wires */
    const double *params;
    const int64_t *params_off;
    const uint8_t *inverses;
    const double *mats; /* interleaved complex128, row-major; may be empty per op */
    const int64_t *mats_off;
} plb200_ops_t;

/* ------------------------------------------------------------------ errors / info */
const char *plb200_last_error(void);
int plb200_device_count(int *count);
/* major*10+minor of `device`, e.g. 100 for B200 (BindingsCudaUtils.hpp:47-115 get_gpu_arch) */
int plb200_device_arch(int device, int *arch);
const char *plb200_version(void);

/* ------------------------------------------------------------------ lifecycle
 * replaces StateVectorCudaManaged(num_qubits, DevTag) — lightning_gpu/StateVectorCudaManaged.hpp
 * ctor family / LGPUBindings.hpp:283-287.  `stream` is a cudaStream_t (NULL = legacy default). */
int plb200_sv_create(plb200_sv **out, int64_t num_qubits, int precision, int device,
                     void *stream);
/* wraps device memory owned by the caller (e.g. a torch tensor / an NVLink-peer-mapped slab) */
int plb200_sv_create_external(plb200_sv **out, int64_t num_qubits, int precision, int device,
                              void *stream, void *device_ptr);
int plb200_sv_destroy(plb200_sv *sv);
int64_t plb200_sv_num_qubits(const plb200_sv *sv);
int64_t plb200_sv_length(const plb200_sv *sv);
int plb200_sv_precision(const plb200_sv *sv);
int plb200_sv_device(const plb200_sv *sv);
void *plb200_sv_device_ptr(const plb200_sv *sv);
int plb200_sv_sync(plb200_sv *sv);
/* number of CUDA kernels this library has launched for `sv` since creation */
int64_t plb200_sv_kernel_launches(const plb200_sv *sv);

/* ------------------------------------------------------------------ copies
 * DeviceToHost / HostToDevice / DeviceToDevice / updateData — LGPUBindings.hpp:289-349,
 * StateVectorCudaBase.hpp CopyHostDataToGpu/CopyGpuDataToHost.  n_elems = complex elements. */
int plb200_sv_h2d(plb200_sv *sv, const void *host, int64_t n_elems, int async);
int plb200_sv_d2h(plb200_sv *sv, void *host, int64_t n_elems, int async);
int plb200_sv_d2d(plb200_sv *dst, const plb200_sv *src);

/* ------------------------------------------------------------------ state preparation
 * StateVectorLQubit.hpp:881-1036 (resetStateVector, setBasisState, setStateVector, collapse,
 * normalize) and lightning_gpu/initSV.cu:76-136. */
int plb200_sv_reset(plb200_sv *sv);
int plb200_sv_set_basis_state_index(plb200_sv *sv, int64_t index);
int plb200_sv_set_basis_state(plb200_sv *sv, const int64_t *state, const int64_t *wires,
                              int64_t n_wires);
/* values: 2^n_wires interleaved complex128 on host; other wires are put in |0> */
int plb200_sv_set_state_vector(plb200_sv *sv, const double *values, const int64_t *wires,
                               int64_t n_wires);
/* zero the state, then sv[indices[i]] = values[i] (setStateVector(indices, values)) */
int plb200_sv_set_state_indices(plb200_sv *sv, const int64_t *indices, const double *values,
                                int64_t n);
int plb200_sv_collapse(plb200_sv *sv, int64_t wire, int branch);
int plb200_sv_normalize(plb200_sv *sv);

/* ------------------------------------------------------------------ gates
 * applyOperation(name, [ctrl_wires, ctrl_values,] wires, inverse, params) —
 * StateVectorLQubit.hpp:378-430; names from core/gates/Constant.hpp:30-135.
 * n_ctrl == 0 selects the uncontrolled overload. */
int plb200_sv_apply(plb200_sv *sv, const char *name, const int64_t *ctrl_wires,
                    const uint8_t *ctrl_values, int64_t n_ctrl, const int64_t *wires,
                    int64_t n_wires, int inverse, const double *params, int64_t n_params);
/* applyMatrix / applyControlledMatrix — StateVectorLQubit.hpp:571-710 */
int plb200_sv_apply_matrix(plb200_sv *sv, const double *matrix, const int64_t *ctrl_wires,
                           const uint8_t *ctrl_values, int64_t n_ctrl, const int64_t *wires,
                           int64_t n_wires, int inverse);
/* applyPauliRot — GateImplementationsLM.hpp:575-629 */
int plb200_sv_apply_pauli_rot(plb200_sv *sv, const int64_t *wires, int64_t n_wires, int inverse,
                              double theta, const char *word);
/* applyGenerator -> scaling factor — GateImplementationsLM.hpp:2181-2952 */
int plb200_sv_apply_generator(plb200_sv *sv, const char *name, const int64_t *ctrl_wires,
                              const uint8_t *ctrl_values, int64_t n_ctrl, const int64_t *wires,
                              int64_t n_wires, int adj, double *scale);
/* host-only validation of one gate call against an n-qubit state (same checks and messages as
 * plb200_sv_apply, no device work): lets a caller queue gates lazily and still fail at call time. */
int plb200_validate_op(int64_t num_qubits, const char *name, const int64_t *ctrl_wires,
                       const uint8_t *ctrl_values, int64_t n_ctrl, const int64_t *wires, int64_t n_wires,
                       int inverse, const double *params, int64_t n_params);
/* the same for a whole blob (ops with an explicit matrix, "PauliRot[word]" entries of the lazy queue) */
int plb200_validate_ops(int64_t num_qubits, const plb200_ops_t *ops);
/* applyOperations over a whole tape; `fuse` != 0 lets the engine schedule the tape into
 * cache-blocked passes (same arithmetic per gate, fewer HBM sweeps). */
int plb200_sv_apply_ops(plb200_sv *sv, const plb200_ops_t *ops, int fuse);
/* host-only dry run of the fusion scheduler for an n-qubit state (no device needed):
 * out4 = {tile passes, stand-alone kernels, register rounds, gates executed inside tile passes} */
int plb200_schedule_stats(int64_t num_qubits, int precision, const plb200_ops_t *ops, int64_t *out4);
/* Pass specialisation: a fused pass whose structure has been seen before runs as a kernel generated for that
 * structure and compiled with NVRTC for sm_100a (angles / phases / tile placement stay run-time arguments);
 * until the kernel exists the generic interpreter kernel runs the pass.  Replaces nothing in the reference
 * (its gate application is one precompiled kernel per gate, GateImplementationsLM.hpp:650-703).
 * mode: 0 = off, 1 = compile in the background from the second sighting (default, PLB200_JIT=async),
 * 2 = compile at first sight, blocking (PLB200_JIT=sync); set_mode(-1) returns to the environment's choice.
 * stats out8 = {compiled, loaded from the disk cache, launches of compiled kernels, launches left to the
 * interpreter, failed compiles, compile microseconds, queued + in flight, structures seen}. */
int plb200_jit_available(void);
int plb200_jit_mode(void);
void plb200_jit_set_mode(int mode);
int plb200_jit_wait(void);
void plb200_jit_stats(int64_t *out8);
/* host-only (no device needed): write / compile the specialised source of every tile pass of a tape */
int plb200_jit_dump_sources(int64_t num_qubits, int precision, const plb200_ops_t *ops, const char *dir,
                            int64_t *n_passes);
int plb200_jit_compile_check(int64_t num_qubits, int precision, const plb200_ops_t *ops, int64_t *n_passes,
                             int64_t *n_ok);
/* statistics of the last plb200_sv_apply_ops call: [0]=gates, [1]=HBM passes (kernel launches) */
int plb200_sv_last_apply_stats(const plb200_sv *sv, int64_t *stats2);

/* ------------------------------------------------------------------ linear algebra
 * innerProdC / scaleAndAdd / squaredNorm — lightning_qubit/utils/LinearAlgebra.hpp:87-159,322-400;
 * cuBLAS dotc/axpy/scal wrappers — cuda_utils/LinearAlg.hpp:97-280. */
int plb200_sv_dot(const plb200_sv *a, const plb200_sv *b, double *out_re_im); /* <a|b> */
int plb200_sv_axpy(plb200_sv *y, const double *alpha_re_im, const plb200_sv *x);
int plb200_sv_scale(plb200_sv *sv, const double *alpha_re_im);
int plb200_sv_norm2(const plb200_sv *sv, double *out);

/* ------------------------------------------------------------------ observables
 * NamedObs / HermitianObs / TensorProdObs / Hamiltonian — core/observables/Observables.hpp:128-584,
 * lightning_qubit/observables/ObservablesLQubit.hpp:48-370.  Children are copied. */
int plb200_obs_named(plb200_obs **out, const char *name, const int64_t *wires, int64_t n_wires,
                     const double *params, int64_t n_params);
int plb200_obs_hermitian(plb200_obs **out, const double *matrix, const int64_t *wires,
                         int64_t n_wires);
int plb200_obs_tensor(plb200_obs **out, const plb200_obs *const *terms, int64_t n_terms);
int plb200_obs_hamiltonian(plb200_obs **out, const double *coeffs,
                           const plb200_obs *const *terms, int64_t n_terms);
int plb200_obs_destroy(plb200_obs *obs);
/* Observable::applyInPlace(sv) — Observables.hpp:63 */
int plb200_obs_apply(const plb200_obs *obs, plb200_sv *sv);

/* ------------------------------------------------------------------ measurements
 * MeasurementsLQubit.hpp:90-163 (probs), :213-305 (expval matrix / named), :372-394 (expval obs),
 * :431-508 (var), :646-679 (generate_samples); MeasurementsGPU.hpp:449-530 (Pauli words). */
/* n_wires < 0 => all wires in order; out has 2^n_wires doubles (host) */
int plb200_probs(plb200_sv *sv, const int64_t *wires, int64_t n_wires, double *out);
int plb200_expval_named(plb200_sv *sv, const char *name, const int64_t *wires, int64_t n_wires,
                        double *out);
int plb200_var_named(plb200_sv *sv, const char *name, const int64_t *wires, int64_t n_wires,
                     double *out);
int plb200_expval_matrix(plb200_sv *sv, const double *matrix, const int64_t *wires,
                         int64_t n_wires, double *out);
int plb200_var_matrix(plb200_sv *sv, const double *matrix, const int64_t *wires,
                      int64_t n_wires, double *out);
/* sum_k coeffs[k] <word_k>: words[k] is a string over {I,X,Y,Z}; its wires are
 * wires[wires_off[k] .. wires_off[k+1]).  All words are evaluated in one fused launch. */
int plb200_expval_pauli_words(plb200_sv *sv, const char *const *words, const int64_t *wires,
                              const int64_t *wires_off, const double *coeffs, int64_t n_words,
                              double *out);
/* the same, returning every <word_k> separately in out[n_words] */
int plb200_expval_pauli_words_each(plb200_sv *sv, const char *const *words,
                                   const int64_t *wires, const int64_t *wires_off,
                                   int64_t n_words, double *out);
int plb200_expval_obs(plb200_sv *sv, const plb200_obs *obs, double *out);
int plb200_var_obs(plb200_sv *sv, const plb200_obs *obs, double *out);
/* alias-method sampling with std::mt19937(seed) exactly as MeasurementKernels.hpp:308-381;
 * seed < 0 => std::random_device.  out: shots x n_wires uint64 (wire order as given);
 * n_wires < 0 => all wires.  Above 24 wires (or PLB200_SAMPLES=device) it forwards to the device sampler. */
int plb200_generate_samples(plb200_sv *sv, const int64_t *wires, int64_t n_wires, int64_t shots,
                            int64_t seed, uint64_t *out);
/* Measurements::generate_samples (MeasurementsLQubit.hpp:646-679, MeasurementsGPU.hpp generate_samples) with the
 * table-free device sampler: chunk masses in one sweep, scan, one warp per shot (Philox).  Same distribution and
 * output layout as plb200_generate_samples, its own random stream; no 2^k host table. */
int plb200_generate_samples_device(plb200_sv *sv, const int64_t *wires, int64_t n_wires, int64_t shots,
                                   int64_t seed, uint64_t *out);

/* ------------------------------------------------------------------ adjoint Jacobian
 * AdjointJacobian::adjointJacobian — AdjointJacobianLQubit.hpp:347-491.  jac has
 * n_obs*n_tp doubles, observable-major (the layout returned to Python, Bindings.hpp:710-729).
 * `sv` is not modified. apply_ops != 0 applies the tape to a copy of sv first. */
int plb200_adjoint_jacobian(const plb200_sv *sv, const plb200_obs *const *obs, int64_t n_obs,
                            const plb200_ops_t *ops, const int64_t *trainable, int64_t n_tp,
                            int apply_ops, double *jac);

/* SparseHamiltonian (core/observables/Observables.hpp:592-699; ObservablesGPU.hpp SparseHamiltonian): a CSR
 * matrix over the full 2^n index space; data = nnz (re, im) pairs.  Applied with the engine's own CSR SpMV
 * kernel (replaces cusparseSpMV, lightning_gpu/utils/LinearAlg.hpp:378-620). */
int plb200_obs_sparse(plb200_obs **out, const int64_t *indptr, const int64_t *indices, const double *data,
                      int64_t n_rows);
/* Measurements::expval / var CSR overloads (lightning_gpu/bindings/LGPUBindings.hpp:65-150) */
int plb200_expval_sparse(plb200_sv *sv, const int64_t *indptr, const int64_t *indices, const double *data,
                         int64_t n_rows, double *out);
int plb200_var_sparse(plb200_sv *sv, const int64_t *indptr, const int64_t *indices, const double *data,
                      int64_t n_rows, double *out);
/* Host-only Hermitian eigen-decomposition for HermitianObs measured with shots (replaces the scipy-openblas
 * zheev of core/utils/UtilLinearAlg.hpp:59-117): eigvals ascending, unitary = V^dagger row-major (re, im). */
int plb200_hermitian_eigh(const double *matrix, int64_t dim, double *eigvals, double *unitary);
/* VectorJacobianProduct (lightning_qubit/algorithms/VectorJacobianProduct.hpp:43-163): dy = 2^n host complex
 * pairs (cotangent of the state), out = n_tp complex pairs. */
int plb200_vjp(const plb200_sv *sv, const double *dy, const plb200_ops_t *ops, const int64_t *trainable,
               int64_t n_tp, int apply_ops, double *out);

/* ------------------------------------------------------------------ distributed support
 * Local half of a global<->local index-bit swap (replaces custatevecSVSwapWorker /
 * StateVectorKokkosMPI::swapGlobalLocalWires, StateVectorKokkosMPI.hpp:747-889).
 * pack:   buf[j] = sv[insert(j, bit, 1-keep)]  for the 2^(n-1) amplitudes whose local index bit
 *         `bit` differs from `keep`;  unpack writes them back.  buf is device memory. */
int plb200_sv_pack_bit(const plb200_sv *sv, int64_t bit, int keep, void *buf);
int plb200_sv_unpack_bit(plb200_sv *sv, int64_t bit, int keep, const void *buf);
/* in-place exchange of the half with local bit `bit` != keep against a peer-mapped slab
 * (cudaDeviceEnablePeerAccess over NVLink); each GPU of the pair calls it with its own keep. */
int plb200_sv_swap_bit_peer(plb200_sv *sv, int64_t bit, int keep, void *peer_device_ptr,
                            int do_half);
/* k = 1..3 global bits <-> k local bits in ONE all-to-all exchange over peer memory (NVSwitch): with
 * my_value = this rank's value on the swapped global bits (bit i of it pairs with bits[i]), the block of
 * local-bit value p trades places with block my_value of the rank whose global bits read p;
 * peer_device_ptrs[p] = that rank's mapped slab (entry my_value unused).  Moves (1 - 2^-k) S per GPU
 * instead of k S/2 (replaces the chained custatevecSVSwapWorker exchanges, MPIWorker.hpp:1-311). */
int plb200_sv_swap_bits_peer(plb200_sv *sv, const int64_t *bits, int64_t k, int64_t my_value,
                             void *const *peer_device_ptrs);

/* CUDA IPC plumbing for the peer path (one process per GPU): export the slab of `sv` as a 64-byte
 * handle, map a peer's slab into this process (peer access over NVLink is enabled lazily), unmap. */
int plb200_sv_ipc_handle(const plb200_sv *sv, unsigned char *handle64);
/* Routed tape ("swap-out"): like plb200_sv_apply_ops(fuse = 1), but the LAST fused pass stores every amplitude
 * where it belongs after exchanging the k local index bits lbits[] with k global (rank) bits — into this rank's
 * own ping-pong slab, or straight into a peer's ping-pong slab over NVLink peer memory — so that the index-bit
 * swap rides on the pass's store phase instead of costing a sweep of its own (replaces the swap — apply — swap
 * sequence of StateVectorCudaMPI.hpp:1936-2243 and swapGlobalLocalWires, StateVectorKokkosMPI.hpp:747-889).
 * dst[p] (p = value of the swapped LOCAL bits, bit i <-> lbits[i]) = peer-mapped ping-pong slab of the rank whose
 * swapped global bits read p; dst[my_value] is ignored (own slab).  *routed = 1: the state now lives in the
 * ping-pong slabs (every rank must synchronise before reading); 0: the tape was applied in place and the caller
 * swaps with plb200_sv_swap_bit[s]_peer.  plb200_sv_alloc_alt allocates the second slab (same size). */
int plb200_sv_alloc_alt(plb200_sv *sv);
/* bandwidth probe of the peer path: device-to-device copy kernel on the state's stream (either side may be a peer
 * mapping: remote stores = push, remote loads = pull) */
int plb200_sv_peer_copy(plb200_sv *sv, void *dst, const void *src, int64_t n_bytes, int unroll);
void *plb200_sv_alt_ptr(const plb200_sv *sv);
int plb200_sv_ipc_handle_alt(const plb200_sv *sv, unsigned char *handle64);
int plb200_sv_apply_ops_route(plb200_sv *sv, const plb200_ops_t *ops, int64_t k, const int64_t *lbits,
                              int64_t my_value, void *const *dst, int *routed);
int plb200_ipc_open(const unsigned char *handle64, int device, void **peer_ptr);
int plb200_ipc_close(void *peer_ptr, int device);

#ifdef __cplusplus
}
#endif
#endif /* PLB200_H */
