// Device-side helpers and the state-vector object of the plb200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "common.hpp"

namespace plb200 {

#define PLB_CUDA(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess)                                                 \
            ::plb200::fail(std::string("CUDA error: ") + cudaGetErrorString(_e) + \
                           " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

template <typename T> struct Cplx;
template <> struct Cplx<float> {
    using type = float2;
};
template <> struct Cplx<double> {
    using type = double2;
};

template <typename T2> __host__ __device__ inline T2 mk(double re, double im) {
    T2 r;
    r.x = static_cast<decltype(r.x)>(re);
    r.y = static_cast<decltype(r.y)>(im);
    return r;
}
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b) {
    T2 r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename T2> __device__ __forceinline__ T2 cfma(T2 a, T2 b, T2 c) { // a*b + c
    T2 r;
    r.x = fma(a.x, b.x, fma(-a.y, b.y, c.x));
    r.y = fma(a.x, b.y, fma(a.y, b.x, c.y));
    return r;
}
template <typename T2> __device__ __forceinline__ T2 cneg(T2 a) {
    a.x = -a.x, a.y = -a.y;
    return a;
}

// Insert zero bits at the (ascending) positions pos[0..n) of x.
struct BitInsert {
    int n;
    uint64_t lowmask[40]; // (1<<pos)-1 for each ascending position
};
__host__ __device__ __forceinline__ uint64_t insert_bits(uint64_t x, const BitInsert &bi) {
#pragma unroll 4
    for (int i = 0; i < bi.n; i++) {
        const uint64_t lm = bi.lowmask[i];
        x = ((x & ~lm) << 1) | (x & lm);
    }
    return x;
}

// ---------------------------------------------------------------------------------------
struct StateVec {
    int64_t n = 0;       // qubits
    int precision = 64;  // 32 | 64
    int device = 0;
    cudaStream_t stream = nullptr;
    void *data = nullptr;
    bool owned = true;
    void *alt = nullptr; // ping-pong slab of the sharded mode's routed passes (same size as data), or null
    // scratch for reductions (partials + result), lazily allocated
    double *red = nullptr;
    size_t red_cap = 0;
    void *tbl = nullptr; // device table scratch (diag tables / dense matrices)
    size_t tbl_cap = 0;
    void *plan = nullptr; // device arena for fused-pass plans
    size_t plan_cap = 0;
    int64_t launches = 0;
    int64_t last_stats[2] = {0, 0};
    int sm_count = 148;

    size_t length() const { return size_t{1} << n; }
    size_t elem_bytes() const { return precision == 64 ? 16 : 8; }
    size_t bytes() const { return length() * elem_bytes(); }
    void set_device() const { PLB_CUDA(cudaSetDevice(device)); }
    double *reduce_buf(size_t n_doubles);
    void *table_buf(size_t bytes);
    void *plan_buf(size_t bytes);
    void sync() const { PLB_CUDA(cudaStreamSynchronize(stream)); }
};

// gate_kernels.cu
void launch_op(StateVec &sv, const COp &op);
void launch_ops(StateVec &sv, const std::vector<COp> &ops);

// dense_mma.cu: OP_DENSE on 5..7 wires through the tensor cores (DMMA / 3xTF32); false = not applicable
bool launch_dense_mma(StateVec &sv, const COp &op);

// measure_kernels.cu
double norm2(StateVec &sv);
void dot(const StateVec &a, const StateVec &b, StateVec &scratch_owner, double out[2]);
void scale(StateVec &sv, cd alpha);
void axpy(StateVec &y, cd alpha, const StateVec &x);
void probs_all(StateVec &sv, double *host_out);
void probs_wires(StateVec &sv, const std::vector<int> &bits_msb_first, double *host_out);
// sample_kernels.cu: computational-basis samples drawn on the device; bits[j] = index bit of reported wire j
void sample_device(StateVec &sv, const std::vector<int> &bits, int64_t shots, uint64_t seed, uint64_t *host_out);
// sum_i conj(a[i]) (P b)[i] for W words in one launch; out[2*W]
void pauli_inner(StateVec &a, const StateVec &b, const PauliWordMask *words, int64_t n_words,
                 double *out_re_im);
double expval_matrix_small(StateVec &sv, const std::vector<cd> &matrix, const std::vector<int> &tbits);
// out = sum_k coeff_k P_k in   (Hamiltonian of Pauli words applied out-of-place)
void pauli_sum_apply(StateVec &out, const StateVec &in, const PauliWordMask *words, const double *coeffs,
                     int64_t n_words);
// out = A in, A in CSR over the full index space (device arrays; values complex128)
void csr_apply(StateVec &out, const StateVec &in, const int64_t *d_indptr, const int64_t *d_indices, const void *d_vals,
               int64_t nnz);
void scatter_values(StateVec &sv, const int64_t *idx, const double *vals, int64_t n);
void set_state_on_wires(StateVec &sv, const double *vals, const std::vector<int> &tbits);
void collapse_zero(StateVec &sv, int bit, int keep_value);
void pack_bit(const StateVec &sv, int bit, int keep, void *buf);
void unpack_bit(StateVec &sv, int bit, int keep, const void *buf);
void peer_copy(StateVec &sv, void *dst, const void *src, uint64_t n16, int unroll);
void swap_bit_peer(StateVec &sv, int bit, int keep, void *peer, int do_half);
void swap_bits_peer(StateVec &sv, const int *bits, int k, int my_value, void *const *peers);

} // namespace plb200
