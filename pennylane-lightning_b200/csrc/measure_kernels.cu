// Reductions and state utilities of the plb200 engine (sm_100a): norm / dot / axpy, probs,
// Pauli-word inner products (expval + adjoint generator overlaps in ONE pass, many words per
// launch), small-matrix expval, Pauli-sum (Hamiltonian) application, state preparation helpers
// and the local half of the distributed index-bit swap.
// All sums accumulate in fp64 per thread, reduce by warp shuffle, then by a fixed-order second
// stage -> deterministic for a given launch shape: no reduction of the engine uses floating-point atomics (marginals
// over <= 11 wires: probs_marginal_kernel; the overlap accumulators of the fused adjoint passes: fusion.cu).
#include <type_traits>

#include "device.cuh"

#include <algorithm>
#include <cstring>

namespace plb200 {

double *StateVec::reduce_buf(size_t n_doubles) {
    if (n_doubles > red_cap) {
        set_device();
        if (red) PLB_CUDA(cudaFree(red));
        red_cap = std::max<size_t>(n_doubles, 1 << 16);
        PLB_CUDA(cudaMalloc(&red, red_cap * sizeof(double)));
    }
    return red;
}
void *StateVec::table_buf(size_t bytes) {
    if (bytes > tbl_cap) {
        set_device();
        PLB_CUDA(cudaStreamSynchronize(stream));
        if (tbl) PLB_CUDA(cudaFree(tbl));
        tbl_cap = std::max<size_t>(bytes, 1 << 16);
        PLB_CUDA(cudaMalloc(&tbl, tbl_cap));
    }
    return tbl;
}

void *StateVec::plan_buf(size_t bytes) {
    if (bytes > plan_cap) {
        set_device();
        PLB_CUDA(cudaStreamSynchronize(stream));
        if (plan) PLB_CUDA(cudaFree(plan));
        plan_cap = std::max<size_t>(bytes * 2, 1 << 20);
        PLB_CUDA(cudaMalloc(&plan, plan_cap));
    }
    return plan;
}

namespace {
constexpr int kThreads = 256;

template <int NV> __device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *out) {
    __shared__ double sm[NV][kThreads / 32];
#pragma unroll
    for (int q = 0; q < NV; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; q++) sm[q][warp] = v[q];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double s = 0;
            for (int w = 0; w < kThreads / 32; w++) s += sm[q][w];
            out[q] = s;
        }
    }
}

// partials laid out [row][nparts][NV]; out[row][NV]
template <int NV> __global__ void final_reduce_kernel(const double *partials, int nparts, double *out) {
    const int row = blockIdx.x;
    double v[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) v[q] = 0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x)
#pragma unroll
        for (int q = 0; q < NV; q++) v[q] += partials[(static_cast<size_t>(row) * nparts + i) * NV + q];
    block_reduce_store<NV>(v, out + static_cast<size_t>(row) * NV);
}

inline int reduce_blocks(const StateVec &sv, uint64_t items) {
    uint64_t b = (items + kThreads * 4 - 1) / (kThreads * 4);
    return static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>(b, uint64_t(sv.sm_count) * 16)));
}

template <typename T2>
__global__ void __launch_bounds__(kThreads) norm2_kernel(const T2 *__restrict__ a, uint64_t len, double *partials) {
    double v[1] = {0};
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const T2 x = a[i];
        v[0] += static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y;
    }
    block_reduce_store<1>(v, partials + blockIdx.x);
}

template <typename T2>
__global__ void __launch_bounds__(kThreads)
    dot_kernel(const T2 *__restrict__ a, const T2 *__restrict__ b, uint64_t len, double *partials) {
    double v[2] = {0, 0};
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const T2 x = a[i], y = b[i];
        v[0] += static_cast<double>(x.x) * y.x + static_cast<double>(x.y) * y.y;
        v[1] += static_cast<double>(x.x) * y.y - static_cast<double>(x.y) * y.x;
    }
    block_reduce_store<2>(v, partials + 2 * blockIdx.x);
}

template <typename T2>
__global__ void __launch_bounds__(kThreads) scale_kernel(T2 *__restrict__ a, uint64_t len, T2 alpha) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
        a[i] = cmul(a[i], alpha);
}
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    axpy_kernel(T2 *__restrict__ y, const T2 *__restrict__ x, uint64_t len, T2 alpha) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
        y[i] = cfma(alpha, x[i], y[i]);
}

template <typename T2>
__global__ void __launch_bounds__(kThreads) probs_all_kernel(const T2 *__restrict__ a, uint64_t len, double *out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const T2 x = a[i];
        // re*re + im*im without FMA contraction, as std::norm on the host
        if constexpr (sizeof(T2) == 16) out[i] = __dadd_rn(__dmul_rn(x.x, x.x), __dmul_rn(x.y, x.y));
        else out[i] = static_cast<double>(__fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y)));
    }
}

struct ProbsArgs {
    int k;
    int bits[40]; // output bit j (lsb-first) <- state bit bits[j]
};
// small k: deterministic marginals.  A warp owns the elements that share the values of all target bits above the
// lane bits and a chunk id over the non-target bits; its 32 lanes cover index bits 0-4 (one coalesced 512-byte access
// per step), every lane sums its own elements in step order, lanes are combined by xor-shuffles over the lane bits
// that are NOT targets, and each (outcome, chunk) partial is written by exactly one lane; probs_reduce_kernel adds an
// outcome's chunks by a fixed tree.  No atomics anywhere: the same bits run after run.
struct MargArgs {
    int k;           // output bit j (lsb-first) <- state bit bits[j]
    int bits[12];
    int nlane;       // lane bits = index bits 0 .. nlane-1 (5, or n for tiny states)
    int nhi;         // target bits above the lane bits, ascending: thi[]
    int thi[12];
    int nstep;       // non-target upper bits walked inside the warp (lowest ones), ascending: pstep[]
    int pstep[8];
    int nchunkbits;  // the remaining non-target upper bits: pchunk[] (ascending)
    int pchunk[64];
    unsigned lane_nontarget; // lane bits that are not targets
};
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    probs_marginal_kernel(const T2 *__restrict__ a, uint64_t nwarps, double *__restrict__ part, const __grid_constant__ MargArgs p) {
    const uint64_t w = (static_cast<uint64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const uint64_t nchunks = uint64_t{1} << p.nchunkbits;
    const uint64_t chunk = w & (nchunks - 1), ohi = w >> p.nchunkbits;
    uint64_t base = 0;
    for (int j = 0; j < p.nhi; j++) base |= ((ohi >> j) & 1) << p.thi[j];
    for (int j = 0; j < p.nchunkbits; j++) base |= ((chunk >> j) & 1) << p.pchunk[j];
    double s = 0;
    if (lane < (1u << p.nlane)) {
        // eight loads in flight per lane; the additions keep their order
        for (unsigned st0 = 0; st0 < (1u << p.nstep); st0 += 8) {
            T2 x[8];
#pragma unroll
            for (unsigned q = 0; q < 8; q++) {
                const unsigned st = st0 + q;
                uint64_t off = 0;
#pragma unroll
                for (int j = 0; j < 7; j++)
                    if (j < p.nstep) off |= static_cast<uint64_t>((st >> j) & 1u) << p.pstep[j];
                x[q] = st < (1u << p.nstep) ? a[base | off | lane] : T2{0, 0};
            }
#pragma unroll
            for (unsigned q = 0; q < 8; q++) s += static_cast<double>(x[q].x) * x[q].x + static_cast<double>(x[q].y) * x[q].y;
        }
    }
    for (int b = 0; b < 5; b++)
        if ((p.lane_nontarget >> b) & 1u) s += __shfl_xor_sync(0xffffffffu, s, 1 << b);
    if ((lane & p.lane_nontarget) == 0 && lane < (1u << p.nlane)) {
        const uint64_t idx = base | lane;
        unsigned o = 0;
        for (int j = 0; j < p.k; j++) o |= static_cast<unsigned>((idx >> p.bits[j]) & 1) << j;
        part[static_cast<uint64_t>(o) * nchunks + chunk] = s;
    }
}
// out[o] = sum of part[o * nchunks + c] over c, by a fixed tree (one block per outcome)
__global__ void __launch_bounds__(256) probs_reduce_kernel(const double *__restrict__ part, uint64_t nchunks, double *__restrict__ out) {
    __shared__ double sm[8];
    const double *col = part + static_cast<uint64_t>(blockIdx.x) * nchunks;
    double s = 0;
    for (uint64_t c = threadIdx.x; c < nchunks; c += 256) s += col[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += sm[w];
        out[blockIdx.x] = t;
    }
}
// large k: one thread per output, ordered sum over the remaining bits
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    probs_gather_kernel(const T2 *__restrict__ a, uint64_t nout, uint64_t nrest, double *out,
                        const __grid_constant__ ProbsArgs p, const __grid_constant__ BitInsert rest_ins) {
    const uint64_t o = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o >= nout) return;
    uint64_t base = 0;
    for (int j = 0; j < p.k; j++) base |= ((o >> j) & 1) << p.bits[j];
    double s = 0;
    for (uint64_t r = 0; r < nrest; r++) {
        // rest_ins inserts zeros at the *measured* bit positions
        const T2 x = a[insert_bits(r, rest_ins) | base];
        s += static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y;
    }
    out[o] = s;
}

// ------------------------------------------------------------------ Pauli-word inner products
struct WordDev {
    uint64_t x, z, cmask, cval;
    int ny;
    int pad;
};

template <typename T2>
__global__ void __launch_bounds__(kThreads)
    pauli_inner_kernel(const T2 *__restrict__ a, const T2 *__restrict__ b, uint64_t len,
                       const WordDev *__restrict__ words, double *partials) {
    const WordDev w = words[blockIdx.x];
    double v[2] = {0, 0};
    const uint64_t stride = static_cast<uint64_t>(gridDim.y) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.y) * blockDim.x + threadIdx.x; i < len; i += stride) {
        if ((i & w.cmask) != w.cval) continue;
        const uint64_t j = i ^ w.x;
        const T2 x = a[i], y = b[j];
        // conj(x) * y
        double re = static_cast<double>(x.x) * y.x + static_cast<double>(x.y) * y.y;
        double im = static_cast<double>(x.x) * y.y - static_cast<double>(x.y) * y.x;
        if (__popcll(j & w.z) & 1) re = -re, im = -im;
        v[0] += re;
        v[1] += im;
    }
    // multiply by i^ny once per thread
    double re = v[0], im = v[1];
    switch (w.ny & 3) {
    case 1:
        v[0] = -im, v[1] = re;
        break;
    case 2:
        v[0] = -re, v[1] = -im;
        break;
    case 3:
        v[0] = im, v[1] = -re;
        break;
    default:
        break;
    }
    block_reduce_store<2>(v, partials + 2 * (static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y));
}

// <psi| Z-word |psi> for up to W diagonal words in ONE read of the state:
// sum_i |a_i|^2 (-1)^{popcount(i & z_q)}, q < W  (expval of PauliZ on every wire, ZZ terms of an Ising
// Hamiltonian, ...).  partials laid out [block][W].
// <Z_b> for EVERY index bit b and the norm in ONE read of the state (e2e: expval(PauliZ(w)) for all wires).  The
// grid has a power-of-two number of threads 2^s and strides by it, so the low s index bits are constant per thread:
// for those a thread needs only its total, the sign is applied once at the end; only the n - s high bits vary inside
// the loop (they are the loop counter's bits) and get an accumulator each: total - 2 * (sum over set bit).
constexpr int kZAllMaxBits = 40, kZAllMaxHigh = 16;
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    zall_kernel(const T2 *__restrict__ a, int n, int s, double *partials) {
    const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t iters = uint64_t{1} << (n - s);
    double total = 0.0, hi[kZAllMaxHigh];
#pragma unroll
    for (int j = 0; j < kZAllMaxHigh; j++) hi[j] = 0.0;
#pragma unroll 4
    for (uint64_t k = 0; k < iters; k++) {
        const T2 x = a[(k << s) | gtid];
        const double p = static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y;
        total += p;
#pragma unroll
        for (int j = 0; j < kZAllMaxHigh; j++)
            if ((k >> j) & 1) hi[j] += p;
    }
    double v[kZAllMaxBits + 1];
#pragma unroll
    for (int b = 0; b <= kZAllMaxBits; b++) {
        double r = 0.0;
        if (b < s) r = ((gtid >> b) & 1) ? -total : total;
        else if (b < n) r = total - 2.0 * hi[(b - s) < kZAllMaxHigh ? (b - s) : 0];
        else if (b == kZAllMaxBits) r = total; // the norm travels in the last slot
        v[b] = r;
    }
    block_reduce_store<kZAllMaxBits + 1>(v, partials + static_cast<size_t>(blockIdx.x) * (kZAllMaxBits + 1));
}

template <int W> struct ZMasks {
    uint64_t z[W];
};
template <typename T2, int W>
__global__ void __launch_bounds__(kThreads)
    zwords_kernel(const T2 *__restrict__ a, uint64_t len, const __grid_constant__ ZMasks<W> m, double *partials) {
    // the sign masks sit in the constant bank (uniform loads): W accumulators per thread, ONE read of the state
    double v[W];
#pragma unroll
    for (int q = 0; q < W; q++) v[q] = 0;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        const T2 x = a[i];
        const double p = static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y;
#pragma unroll
        for (int q = 0; q < W; q++) v[q] += (__popcll(i & m.z[q]) & 1) ? -p : p;
    }
    block_reduce_store<W>(v, partials + static_cast<size_t>(blockIdx.x) * W);
}

template <typename T2>
__global__ void __launch_bounds__(kThreads)
    pauli_sum_apply_kernel(T2 *__restrict__ out, const T2 *__restrict__ in, uint64_t len,
                           const WordDev *__restrict__ words, const double *__restrict__ coeffs, int nwords) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride) {
        double re = 0, im = 0;
        for (int k = 0; k < nwords; k++) {
            const WordDev w = words[k];
            const uint64_t j = i ^ w.x;
            const T2 y = in[j];
            double c = coeffs[k];
            if (__popcll(j & w.z) & 1) c = -c;
            double yr = y.x, yi = y.y;
            switch (w.ny & 3) { // times i^ny
            case 1: {
                const double t = yr;
                yr = -yi, yi = t;
            } break;
            case 2:
                yr = -yr, yi = -yi;
                break;
            case 3: {
                const double t = yr;
                yr = yi, yi = -t;
            } break;
            default:
                break;
            }
            re = fma(c, yr, re);
            im = fma(c, yi, im);
        }
        out[i] = mk<T2>(re, im);
    }
}

// --------------------------------------------------------------- <psi|M|psi>, M on k<=4 wires
template <typename T2> struct MatExpArgs {
    BitInsert ins;
    uint64_t ngroups;
    uint64_t off[16];
    double2 m[256];
};
template <typename T2, int K>
__global__ void __launch_bounds__(128)
    expval_matrix_kernel(const T2 *__restrict__ sv, const MatExpArgs<T2> *__restrict__ pp, double *partials) {
    constexpr int D = 1 << K;
    __shared__ double2 sm[D * D];
    const MatExpArgs<T2> &p = *pp;
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sm[i] = p.m[i];
    __syncthreads();
    double acc = 0;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < p.ngroups; g += stride) {
        const uint64_t base = insert_bits(g, p.ins);
        double2 v[D];
#pragma unroll
        for (int c = 0; c < D; c++) {
            const T2 t = sv[base + p.off[c]];
            v[c] = make_double2(t.x, t.y);
        }
#pragma unroll
        for (int r = 0; r < D; r++) {
            double2 s = make_double2(0, 0);
#pragma unroll
            for (int c = 0; c < D; c++) s = cfma(sm[r * D + c], v[c], s);
            acc += v[r].x * s.x + v[r].y * s.y; // Re(conj(v_r) s)
        }
    }
    // block reduce (128 threads)
    __shared__ double red[4];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
}

// ------------------------------------------------------------------------- state helpers
template <typename T2>
__global__ void scatter_kernel(T2 *sv, const int64_t *idx, const double2 *vals, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) sv[idx[i]] = mk<T2>(vals[i].x, vals[i].y);
}
struct WiresArgs {
    int k;
    int tbits[40];
};
template <typename T2>
__global__ void set_on_wires_kernel(T2 *sv, const double2 *vals, uint64_t nvals, const __grid_constant__ WiresArgs p) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nvals) return;
    uint64_t o = 0;
    for (int j = 0; j < p.k; j++) o |= ((i >> j) & 1) << p.tbits[j];
    sv[o] = mk<T2>(vals[i].x, vals[i].y);
}
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    collapse_kernel(T2 *__restrict__ sv, uint64_t half, const __grid_constant__ BitInsert ins, uint64_t zero_bit) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < half; g += stride)
        sv[insert_bits(g, ins) | zero_bit] = mk<T2>(0.0, 0.0);
}
template <typename T2, bool PACK>
__global__ void __launch_bounds__(kThreads)
    pack_kernel(T2 *__restrict__ sv, T2 *__restrict__ buf, uint64_t half, const __grid_constant__ BitInsert ins,
                uint64_t sel_bit) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < half; g += stride) {
        const uint64_t i = insert_bits(g, ins) | sel_bit;
        if constexpr (PACK) buf[g] = sv[i];
        else sv[i] = buf[g];
    }
}
// In-place exchange with a peer slab over NVLink: this GPU owns the pairs g in [lo, hi) and swaps
// its amplitude (bit != keep) with the peer's amplitude (bit == keep... i.e. peer's own non-kept
// half), using 128-bit peer loads/stores.
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    swap_peer_kernel(T2 *__restrict__ mine, T2 *__restrict__ peer, uint64_t lo, uint64_t hi,
                     const __grid_constant__ BitInsert ins, uint64_t my_bit, uint64_t peer_bit) {
    // 4 independent local + 4 independent NVLink loads in flight per thread before any store
    constexpr int U = 4;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t g0 = lo + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g0 < hi; g0 += U * stride) {
        uint64_t b[U];
        T2 a[U], c[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t g = g0 + u * stride;
            b[u] = insert_bits(g < hi ? g : g0, ins);
            c[u] = peer[b[u] | peer_bit];
            a[u] = mine[b[u] | my_bit];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (g0 + u * stride < hi) {
                mine[b[u] | my_bit] = c[u];
                peer[b[u] | peer_bit] = a[u];
            }
        }
    }
}

// k-bit all-to-all exchange (k global bits <-> k local bits at once): with r = this rank's value on the
// swapped global bits, the block of local-bit value p (p != r) trades places with block r of the rank
// whose global bits read p — one in-place transpose over 2^k GPUs that moves (1 - 2^-k) S per GPU
// instead of k S/2 for k single-bit swaps.  Each pair of ranks splits its block in halves so both
// directions of every link carry data; grid.y = partner.
struct MultiSwapArgs {
    BitInsert ins; // zeros at the k local bit positions
    uint64_t nrest;
    int npartners;
    void *peer[7];
    uint64_t my_off[7]; // local-bit deposit of the partner's value p
    uint64_t peer_off;  // local-bit deposit of my value r
    int lower[7];       // 1: this rank handles the first half of the pair's elements
};
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    swap_multi_peer_kernel(T2 *__restrict__ mine, const __grid_constant__ MultiSwapArgs a) {
    constexpr int U = 4;
    const int pi = blockIdx.y;
    T2 *__restrict__ peer = static_cast<T2 *>(a.peer[pi]);
    const uint64_t half = a.nrest >> 1;
    const uint64_t lo = a.lower[pi] ? 0 : half, hi = a.lower[pi] ? half : a.nrest;
    const uint64_t my_off = a.my_off[pi], peer_off = a.peer_off;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t g0 = lo + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g0 < hi; g0 += U * stride) {
        uint64_t b[U];
        T2 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t g = g0 + u * stride;
            b[u] = insert_bits(g < hi ? g : g0, a.ins);
            y[u] = peer[b[u] | peer_off];
            x[u] = mine[b[u] | my_off];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (g0 + u * stride < hi) {
                mine[b[u] | my_off] = y[u];
                peer[b[u] | peer_off] = x[u];
            }
        }
    }
}

// y = A x for a CSR matrix over the full index space (SparseHamiltonian; replaces the cuSPARSE SpMV of
// lightning_gpu/utils/LinearAlg.hpp:378-620 and LQ's apply_Sparse_Matrix).  One sub-warp of G lanes per row:
// the lanes stride over the row's entries (x gathered through L2), partial sums meet in a shuffle reduction.
template <typename T2, int G>
__global__ void __launch_bounds__(kThreads)
    csr_apply_kernel(T2 *__restrict__ y, const T2 *__restrict__ x, const int64_t *__restrict__ indptr,
                     const int64_t *__restrict__ indices, const double2 *__restrict__ vals, uint64_t nrows) {
    const uint64_t gtid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t row = gtid / G;
    const int lane = static_cast<int>(gtid % G);
    double re = 0.0, im = 0.0;
    if (row < nrows) {
        const int64_t b = indptr[row], e = indptr[row + 1];
        for (int64_t k = b + lane; k < e; k += G) {
            const double2 a = vals[k];
            const T2 v = x[indices[k]];
            re += a.x * static_cast<double>(v.x) - a.y * static_cast<double>(v.y);
            im += a.x * static_cast<double>(v.y) + a.y * static_cast<double>(v.x);
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (row < nrows && lane == 0) y[row] = mk<T2>(re, im);
}

// Bandwidth probe of the NVLink peer path (tools/peer_bw.py): dst[i] = src[i], 128-bit accesses, U in flight
// per thread; one of the two pointers is a peer mapping (push: remote stores, pull: remote loads).
template <int U>
__global__ void __launch_bounds__(kThreads) peer_copy_kernel(double2 *__restrict__ dst, const double2 *__restrict__ src, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i0 = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (i0 + u * stride < n) v[u] = src[i0 + u * stride];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (i0 + u * stride < n) dst[i0 + u * stride] = v[u];
    }
}
BitInsert single_insert(int bit) {
    BitInsert bi;
    bi.n = 1;
    bi.lowmask[0] = (uint64_t{1} << bit) - 1;
    return bi;
}

template <int NV> void finish_reduce(StateVec &sv, double *partials, int rows, int nparts, double *host_out) {
    double *res = partials + static_cast<size_t>(rows) * nparts * NV;
    final_reduce_kernel<NV><<<rows, kThreads, 0, sv.stream>>>(partials, nparts, res);
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
    PLB_CUDA(cudaMemcpyAsync(host_out, res, sizeof(double) * rows * NV, cudaMemcpyDeviceToHost, sv.stream));
    PLB_CUDA(cudaStreamSynchronize(sv.stream));
}

} // namespace

#define DISPATCH(sv, call64, call32)                                                                     \
    do {                                                                                                 \
        if ((sv).precision == 64) {                                                                      \
            using T2 = double2;                                                                          \
            call64;                                                                                      \
        } else {                                                                                         \
            using T2 = float2;                                                                           \
            call32;                                                                                      \
        }                                                                                                \
    } while (0)

double norm2(StateVec &sv) {
    sv.set_device();
    const int nb = reduce_blocks(sv, sv.length());
    double *part = sv.reduce_buf(static_cast<size_t>(nb) + 8);
    DISPATCH(sv, (norm2_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), sv.length(), part)),
             (norm2_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), sv.length(), part)));
    sv.launches++;
    double out;
    finish_reduce<1>(sv, part, 1, nb, &out);
    return out;
}

void dot(const StateVec &a, const StateVec &b, StateVec &owner, double out[2]) {
    PLB_CHECK(a.n == b.n && a.precision == b.precision, "dot: incompatible state vectors");
    owner.set_device();
    const int nb = reduce_blocks(owner, a.length());
    double *part = owner.reduce_buf(2 * static_cast<size_t>(nb) + 8);
    DISPATCH(a,
             (dot_kernel<T2><<<nb, kThreads, 0, owner.stream>>>(static_cast<const T2 *>(a.data),
                                                                static_cast<const T2 *>(b.data), a.length(), part)),
             (dot_kernel<T2><<<nb, kThreads, 0, owner.stream>>>(static_cast<const T2 *>(a.data),
                                                                static_cast<const T2 *>(b.data), a.length(), part)));
    owner.launches++;
    finish_reduce<2>(owner, part, 1, nb, out);
}

void scale(StateVec &sv, cd alpha) {
    sv.set_device();
    const int nb = reduce_blocks(sv, sv.length());
    DISPATCH(sv,
             (scale_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), sv.length(),
                                                               mk<T2>(alpha.real(), alpha.imag()))),
             (scale_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), sv.length(),
                                                               mk<T2>(alpha.real(), alpha.imag()))));
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
}

void axpy(StateVec &y, cd alpha, const StateVec &x) {
    PLB_CHECK(x.n == y.n && x.precision == y.precision, "axpy: incompatible state vectors");
    y.set_device();
    const int nb = reduce_blocks(y, y.length());
    DISPATCH(y,
             (axpy_kernel<T2><<<nb, kThreads, 0, y.stream>>>(static_cast<T2 *>(y.data), static_cast<const T2 *>(x.data),
                                                             y.length(), mk<T2>(alpha.real(), alpha.imag()))),
             (axpy_kernel<T2><<<nb, kThreads, 0, y.stream>>>(static_cast<T2 *>(y.data), static_cast<const T2 *>(x.data),
                                                             y.length(), mk<T2>(alpha.real(), alpha.imag()))));
    y.launches++;
    PLB_CUDA(cudaGetLastError());
}

void probs_all(StateVec &sv, double *host_out) {
    sv.set_device();
    double *dout;
    PLB_CUDA(cudaMalloc(&dout, sv.length() * sizeof(double)));
    const int nb = reduce_blocks(sv, sv.length());
    DISPATCH(sv,
             (probs_all_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), sv.length(), dout)),
             (probs_all_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), sv.length(), dout)));
    sv.launches++;
    cudaError_t e = cudaMemcpyAsync(host_out, dout, sv.length() * sizeof(double), cudaMemcpyDeviceToHost, sv.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv.stream);
    cudaFree(dout);
    PLB_CUDA(e);
}

void probs_wires(StateVec &sv, const std::vector<int> &bits_msb_first, double *host_out) {
    sv.set_device();
    const int k = static_cast<int>(bits_msb_first.size());
    PLB_CHECK(k <= 40, "too many wires");
    ProbsArgs p;
    p.k = k;
    uint64_t mmask = 0;
    for (int j = 0; j < k; j++) {
        p.bits[j] = bits_msb_first[k - 1 - j];
        mmask |= uint64_t{1} << p.bits[j];
    }
    const uint64_t nout = uint64_t{1} << k;
    double *dout;
    PLB_CUDA(cudaMalloc(&dout, nout * sizeof(double)));
    cudaError_t e = cudaSuccess;
    double *dpart = nullptr;
    if (k <= 11) {
        MargArgs m;
        std::memset(&m, 0, sizeof(m));
        m.k = k;
        for (int j = 0; j < k; j++) m.bits[j] = p.bits[j];
        m.nlane = sv.n < 5 ? static_cast<int>(sv.n) : 5;
        for (int b = 0; b < m.nlane; b++)
            if (!(mmask >> b & 1)) m.lane_nontarget |= 1u << b;
        for (int b = m.nlane; b < sv.n; b++) {
            if (mmask >> b & 1) m.thi[m.nhi++] = b;
            else if (m.nstep < 7) m.pstep[m.nstep++] = b;
            else m.pchunk[m.nchunkbits++] = b;
        }
        const uint64_t nchunks = uint64_t{1} << m.nchunkbits;
        const uint64_t nwarps = nchunks << m.nhi;
        e = cudaMalloc(&dpart, nout * nchunks * sizeof(double));
        if (e == cudaSuccess) {
            const unsigned nb = static_cast<unsigned>((nwarps * 32 + kThreads - 1) / kThreads);
            DISPATCH(sv,
                     (probs_marginal_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), nwarps, dpart, m)),
                     (probs_marginal_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), nwarps, dpart, m)));
            probs_reduce_kernel<<<static_cast<unsigned>(nout), 256, 0, sv.stream>>>(dpart, nchunks, dout);
            sv.launches++;
        }
    } else {
        BitInsert ins;
        ins.n = 0;
        for (int b = 0; b < 64; b++)
            if (mmask >> b & 1) ins.lowmask[ins.n++] = (uint64_t{1} << b) - 1;
        const uint64_t nrest = uint64_t{1} << (sv.n - k);
        const unsigned nb = static_cast<unsigned>((nout + kThreads - 1) / kThreads);
        DISPATCH(sv,
                 (probs_gather_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), nout,
                                                                          nrest, dout, p, ins)),
                 (probs_gather_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<const T2 *>(sv.data), nout,
                                                                          nrest, dout, p, ins)));
    }
    sv.launches++;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, dout, nout * sizeof(double), cudaMemcpyDeviceToHost, sv.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv.stream);
    cudaFree(dout);
    if (dpart) cudaFree(dpart);
    PLB_CUDA(e);
}

static std::vector<WordDev> to_dev_words(const PauliWordMask *w, int64_t n) {
    std::vector<WordDev> h(n);
    for (int64_t i = 0; i < n; i++) h[i] = {w[i].x, w[i].z, w[i].cmask, w[i].cval, w[i].ny, 0};
    return h;
}

void pauli_inner(StateVec &a, const StateVec &b, const PauliWordMask *words, int64_t n_words, double *out) {
    PLB_CHECK(a.n == b.n && a.precision == b.precision, "pauli_inner: incompatible state vectors");
    if (n_words == 0) return;
    a.set_device();
    // expectation values of diagonal (Z-only) words: 8 words per read of the state
    bool all_diag = a.data == b.data;
    for (int64_t k = 0; k < n_words && all_diag; k++)
        all_diag = words[k].x == 0 && words[k].cmask == 0 && words[k].ny == 0;
    if (all_diag && n_words > 8 && a.n <= kZAllMaxBits) {
        // every word a single Z (or the identity): all <Z_b> + the norm from ONE sweep (zall_kernel)
        bool singles = true;
        for (int64_t k = 0; k < n_words && singles; k++) singles = __builtin_popcountll(words[k].z) <= 1;
        int s = static_cast<int>(std::min<int64_t>(a.n, 19)); // 2^19 threads
        while (a.n - s > kZAllMaxHigh) s++;
        if (singles && s >= 8) {
            const int nb = 1 << (s - 8); // kThreads = 2^8
            constexpr int NV = kZAllMaxBits + 1;
            double *part = a.reduce_buf(static_cast<size_t>(NV) * nb + NV + 8);
            DISPATCH(a, (zall_kernel<T2><<<nb, kThreads, 0, a.stream>>>(static_cast<const T2 *>(a.data), static_cast<int>(a.n), s, part)),
                     (zall_kernel<T2><<<nb, kThreads, 0, a.stream>>>(static_cast<const T2 *>(a.data), static_cast<int>(a.n), s, part)));
            a.launches++;
            PLB_CUDA(cudaGetLastError());
            double r[NV];
            finish_reduce<NV>(a, part, 1, nb, r);
            for (int64_t k = 0; k < n_words; k++) {
                out[2 * k] = words[k].z ? r[__builtin_ctzll(words[k].z)] : r[kZAllMaxBits];
                out[2 * k + 1] = 0.0;
            }
            return;
        }
    }
    if (all_diag && n_words > 1) {
        // 8 words per read of the state, or 32 when there are many (<Z_w> on every wire of a 30-qubit register
        // is ONE sweep instead of four)
        const int nb = reduce_blocks(a, a.length());
        auto run = [&](auto wtag, int64_t w0, int nw) {
            constexpr int W = decltype(wtag)::value;
            ZMasks<W> m;
            for (int q = 0; q < W; q++) m.z[q] = q < nw ? words[w0 + q].z : 0;
            double *part = a.reduce_buf(static_cast<size_t>(W) * nb + W + 8);
            DISPATCH(a,
                     (zwords_kernel<T2, W><<<nb, kThreads, 0, a.stream>>>(static_cast<const T2 *>(a.data), a.length(), m, part)),
                     (zwords_kernel<T2, W><<<nb, kThreads, 0, a.stream>>>(static_cast<const T2 *>(a.data), a.length(), m, part)));
            a.launches++;
            PLB_CUDA(cudaGetLastError());
            double r[W];
            finish_reduce<W>(a, part, 1, nb, r);
            for (int q = 0; q < nw; q++) out[2 * (w0 + q)] = r[q], out[2 * (w0 + q) + 1] = 0.0;
        };
        int64_t w0 = 0;
        while (w0 < n_words) {
            const int64_t left = n_words - w0;
            if (left > 8) {
                const int nw = static_cast<int>(std::min<int64_t>(32, left));
                run(std::integral_constant<int, 32>{}, w0, nw);
                w0 += nw;
            } else {
                run(std::integral_constant<int, 8>{}, w0, static_cast<int>(left));
                w0 += left;
            }
        }
        return;
    }
    const int64_t kMaxBatch = 4096;
    for (int64_t w0 = 0; w0 < n_words; w0 += kMaxBatch) {
        const int64_t W = std::min(kMaxBatch, n_words - w0);
        int nchunks = reduce_blocks(a, a.length());
        // keep the partial buffer bounded when many words are batched
        nchunks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nchunks, (int64_t{1} << 22) / W)));
        nchunks = std::min(nchunks, 65535);
        auto h = to_dev_words(words + w0, W);
        const size_t wbytes = h.size() * sizeof(WordDev);
        WordDev *dw = static_cast<WordDev *>(a.table_buf(wbytes));
        PLB_CUDA(cudaMemcpyAsync(dw, h.data(), wbytes, cudaMemcpyHostToDevice, a.stream));
        PLB_CUDA(cudaStreamSynchronize(a.stream));
        double *part = a.reduce_buf(2 * static_cast<size_t>(W) * nchunks + 2 * W + 8);
        dim3 grid(static_cast<unsigned>(W), static_cast<unsigned>(nchunks));
        DISPATCH(a,
                 (pauli_inner_kernel<T2><<<grid, kThreads, 0, a.stream>>>(
                     static_cast<const T2 *>(a.data), static_cast<const T2 *>(b.data), a.length(), dw, part)),
                 (pauli_inner_kernel<T2><<<grid, kThreads, 0, a.stream>>>(
                     static_cast<const T2 *>(a.data), static_cast<const T2 *>(b.data), a.length(), dw, part)));
        a.launches++;
        PLB_CUDA(cudaGetLastError());
        finish_reduce<2>(a, part, static_cast<int>(W), nchunks, out + 2 * w0);
    }
}

void pauli_sum_apply(StateVec &out, const StateVec &in, const PauliWordMask *words, const double *coeffs,
                     int64_t n_words) {
    PLB_CHECK(out.n == in.n && out.precision == in.precision && out.data != in.data,
              "pauli_sum_apply: needs distinct compatible state vectors");
    out.set_device();
    auto h = to_dev_words(words, n_words);
    const size_t wbytes = h.size() * sizeof(WordDev);
    const size_t cbytes = n_words * sizeof(double);
    unsigned char *buf = static_cast<unsigned char *>(out.table_buf(wbytes + cbytes));
    PLB_CUDA(cudaMemcpyAsync(buf, h.data(), wbytes, cudaMemcpyHostToDevice, out.stream));
    PLB_CUDA(cudaMemcpyAsync(buf + wbytes, coeffs, cbytes, cudaMemcpyHostToDevice, out.stream));
    PLB_CUDA(cudaStreamSynchronize(out.stream));
    const int nb = reduce_blocks(out, out.length());
    DISPATCH(out,
             (pauli_sum_apply_kernel<T2><<<nb, kThreads, 0, out.stream>>>(
                 static_cast<T2 *>(out.data), static_cast<const T2 *>(in.data), out.length(),
                 reinterpret_cast<const WordDev *>(buf), reinterpret_cast<const double *>(buf + wbytes),
                 static_cast<int>(n_words))),
             (pauli_sum_apply_kernel<T2><<<nb, kThreads, 0, out.stream>>>(
                 static_cast<T2 *>(out.data), static_cast<const T2 *>(in.data), out.length(),
                 reinterpret_cast<const WordDev *>(buf), reinterpret_cast<const double *>(buf + wbytes),
                 static_cast<int>(n_words))));
    out.launches++;
    PLB_CUDA(cudaGetLastError());
}

template <typename T2> static double expval_matrix_typed(StateVec &sv, const std::vector<cd> &matrix, const std::vector<int> &tbits) {
    const int k = static_cast<int>(tbits.size());
    const int D = 1 << k;
    std::vector<MatExpArgs<T2>> hv(1);
    MatExpArgs<T2> &a = hv[0];
    uint64_t tmask = 0;
    for (int b : tbits) tmask |= uint64_t{1} << b;
    a.ins.n = 0;
    for (int b = 0; b < 64; b++)
        if (tmask >> b & 1) a.ins.lowmask[a.ins.n++] = (uint64_t{1} << b) - 1;
    a.ngroups = uint64_t{1} << (sv.n - k);
    for (int c = 0; c < D; c++) {
        uint64_t o = 0;
        for (int j = 0; j < k; j++)
            if (c >> j & 1) o |= uint64_t{1} << tbits[j];
        a.off[c] = o;
    }
    for (int i = 0; i < D * D; i++) a.m[i] = make_double2(matrix[i].real(), matrix[i].imag());
    auto *dargs = static_cast<MatExpArgs<T2> *>(sv.table_buf(sizeof(MatExpArgs<T2>)));
    PLB_CUDA(cudaMemcpyAsync(dargs, &a, sizeof(a), cudaMemcpyHostToDevice, sv.stream));
    PLB_CUDA(cudaStreamSynchronize(sv.stream));
    const int nb = static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>((a.ngroups + 127) / 128, uint64_t(sv.sm_count) * 16)));
    double *part = sv.reduce_buf(static_cast<size_t>(nb) + 8);
    const T2 *d = static_cast<const T2 *>(sv.data);
    switch (k) {
    case 1:
        expval_matrix_kernel<T2, 1><<<nb, 128, 0, sv.stream>>>(d, dargs, part);
        break;
    case 2:
        expval_matrix_kernel<T2, 2><<<nb, 128, 0, sv.stream>>>(d, dargs, part);
        break;
    case 3:
        expval_matrix_kernel<T2, 3><<<nb, 128, 0, sv.stream>>>(d, dargs, part);
        break;
    default:
        expval_matrix_kernel<T2, 4><<<nb, 128, 0, sv.stream>>>(d, dargs, part);
        break;
    }
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
    double out;
    finish_reduce<1>(sv, part, 1, nb, &out);
    return out;
}

double expval_matrix_small(StateVec &sv, const std::vector<cd> &matrix, const std::vector<int> &tbits) {
    PLB_CHECK(tbits.size() >= 1 && tbits.size() <= 4, "expval_matrix_small: 1..4 wires");
    sv.set_device();
    return sv.precision == 64 ? expval_matrix_typed<double2>(sv, matrix, tbits)
                              : expval_matrix_typed<float2>(sv, matrix, tbits);
}

void scatter_values(StateVec &sv, const int64_t *idx, const double *vals, int64_t n) {
    sv.set_device();
    PLB_CUDA(cudaMemsetAsync(sv.data, 0, sv.bytes(), sv.stream));
    if (n == 0) return;
    const size_t ib = n * sizeof(int64_t), vb = n * 2 * sizeof(double);
    unsigned char *buf = static_cast<unsigned char *>(sv.table_buf(ib + vb));
    PLB_CUDA(cudaMemcpyAsync(buf, vals, vb, cudaMemcpyHostToDevice, sv.stream));
    PLB_CUDA(cudaMemcpyAsync(buf + vb, idx, ib, cudaMemcpyHostToDevice, sv.stream));
    PLB_CUDA(cudaStreamSynchronize(sv.stream));
    const unsigned nb = static_cast<unsigned>((n + kThreads - 1) / kThreads);
    DISPATCH(sv,
             (scatter_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data),
                                                                 reinterpret_cast<const int64_t *>(buf + vb),
                                                                 reinterpret_cast<const double2 *>(buf), n)),
             (scatter_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data),
                                                                 reinterpret_cast<const int64_t *>(buf + vb),
                                                                 reinterpret_cast<const double2 *>(buf), n)));
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
}

void set_state_on_wires(StateVec &sv, const double *vals, const std::vector<int> &tbits) {
    sv.set_device();
    WiresArgs p;
    p.k = static_cast<int>(tbits.size());
    PLB_CHECK(p.k <= 40, "too many wires");
    for (int j = 0; j < p.k; j++) p.tbits[j] = tbits[j];
    const uint64_t nvals = uint64_t{1} << p.k;
    PLB_CUDA(cudaMemsetAsync(sv.data, 0, sv.bytes(), sv.stream));
    double2 *dv;
    PLB_CUDA(cudaMalloc(&dv, nvals * sizeof(double2)));
    cudaError_t e = cudaMemcpyAsync(dv, vals, nvals * sizeof(double2), cudaMemcpyHostToDevice, sv.stream);
    const unsigned nb = static_cast<unsigned>((nvals + kThreads - 1) / kThreads);
    if (e == cudaSuccess) {
        DISPATCH(sv, (set_on_wires_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), dv, nvals, p)),
                 (set_on_wires_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), dv, nvals, p)));
        sv.launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv.stream);
    cudaFree(dv);
    PLB_CUDA(e);
}

void collapse_zero(StateVec &sv, int bit, int keep_value) {
    sv.set_device();
    const uint64_t half = sv.length() >> 1;
    const int nb = reduce_blocks(sv, half);
    const BitInsert ins = single_insert(bit);
    const uint64_t zb = keep_value ? 0 : (uint64_t{1} << bit);
    DISPATCH(sv, (collapse_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), half, ins, zb)),
             (collapse_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), half, ins, zb)));
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
}

void pack_bit(const StateVec &sv, int bit, int keep, void *buf) {
    sv.set_device();
    const uint64_t half = sv.length() >> 1;
    const int nb = reduce_blocks(sv, half);
    const BitInsert ins = single_insert(bit);
    const uint64_t sb = keep ? 0 : (uint64_t{1} << bit);
    DISPATCH(sv,
             (pack_kernel<T2, true><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), static_cast<T2 *>(buf),
                                                                    half, ins, sb)),
             (pack_kernel<T2, true><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), static_cast<T2 *>(buf),
                                                                    half, ins, sb)));
    PLB_CUDA(cudaGetLastError());
}
void unpack_bit(StateVec &sv, int bit, int keep, const void *buf) {
    sv.set_device();
    const uint64_t half = sv.length() >> 1;
    const int nb = reduce_blocks(sv, half);
    const BitInsert ins = single_insert(bit);
    const uint64_t sb = keep ? 0 : (uint64_t{1} << bit);
    DISPATCH(sv,
             (pack_kernel<T2, false><<<nb, kThreads, 0, sv.stream>>>(
                 static_cast<T2 *>(sv.data), const_cast<T2 *>(static_cast<const T2 *>(buf)), half, ins, sb)),
             (pack_kernel<T2, false><<<nb, kThreads, 0, sv.stream>>>(
                 static_cast<T2 *>(sv.data), const_cast<T2 *>(static_cast<const T2 *>(buf)), half, ins, sb)));
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
}

void swap_bit_peer(StateVec &sv, int bit, int keep, void *peer, int do_half) {
    // GPU with keep==0 keeps bit==0 amplitudes and gives away bit==1; its peer (keep==1) gives away
    // bit==0.  Each side processes one half of the pair range so the link is used in both directions.
    sv.set_device();
    const uint64_t half = sv.length() >> 1;
    const uint64_t lo = do_half == 0 ? 0 : (do_half == 1 ? 0 : half / 2);
    const uint64_t hi = do_half == 0 ? half : (do_half == 1 ? half / 2 : half);
    const int nb = static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>((hi - lo + kThreads * 4 - 1) / (kThreads * 4),
                                                                             uint64_t(sv.sm_count) * 32)));
    const BitInsert ins = single_insert(bit);
    const uint64_t my_bit = keep ? 0 : (uint64_t{1} << bit);
    const uint64_t peer_bit = keep ? (uint64_t{1} << bit) : 0;
    DISPATCH(sv,
             (swap_peer_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), static_cast<T2 *>(peer),
                                                                   lo, hi, ins, my_bit, peer_bit)),
             (swap_peer_kernel<T2><<<nb, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), static_cast<T2 *>(peer),
                                                                   lo, hi, ins, my_bit, peer_bit)));
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
}

void swap_bits_peer(StateVec &sv, const int *bits, int k, int my_value, void *const *peers) {
    PLB_CHECK(k >= 1 && k <= 3, "swap_bits_peer: 1..3 bits");
    sv.set_device();
    MultiSwapArgs a;
    uint64_t mask = 0;
    for (int i = 0; i < k; i++) mask |= uint64_t{1} << bits[i];
    PLB_CHECK(__builtin_popcountll(mask) == k, "swap_bits_peer: bits must be distinct");
    a.ins.n = 0;
    for (int b = 0; b < 64; b++)
        if (mask >> b & 1) a.ins.lowmask[a.ins.n++] = (uint64_t{1} << b) - 1;
    a.nrest = sv.length() >> k;
    auto deposit = [&](int v) {
        uint64_t o = 0;
        for (int i = 0; i < k; i++)
            if (v >> i & 1) o |= uint64_t{1} << bits[i];
        return o;
    };
    a.peer_off = deposit(my_value);
    a.npartners = 0;
    for (int p = 0; p < (1 << k); p++) {
        if (p == my_value) continue;
        PLB_CHECK(peers[p] != nullptr, "swap_bits_peer: missing peer mapping");
        a.peer[a.npartners] = peers[p];
        a.my_off[a.npartners] = deposit(p);
        a.lower[a.npartners] = my_value < p ? 1 : 0;
        a.npartners++;
    }
    const uint64_t half = a.nrest >> 1;
    const unsigned nbx = static_cast<unsigned>(std::max<uint64_t>(
        1, std::min<uint64_t>((half + kThreads * 4 - 1) / (kThreads * 4), uint64_t(sv.sm_count) * 32 / a.npartners)));
    dim3 grid(nbx, static_cast<unsigned>(a.npartners));
    DISPATCH(sv, (swap_multi_peer_kernel<T2><<<grid, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), a)),
             (swap_multi_peer_kernel<T2><<<grid, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), a)));
    sv.launches++;
    PLB_CUDA(cudaGetLastError());
}

void csr_apply(StateVec &out, const StateVec &in, const int64_t *d_indptr, const int64_t *d_indices, const void *d_vals,
               int64_t nnz) {
    out.set_device();
    const uint64_t nrows = out.length();
    const double avg = static_cast<double>(nnz) / static_cast<double>(nrows);
    const double2 *vals = static_cast<const double2 *>(d_vals);
#define PLB_CSR(G)                                                                                       \
    do {                                                                                                 \
        const uint64_t nb = (nrows * (G) + kThreads - 1) / kThreads;                                     \
        DISPATCH(out,                                                                                    \
                 (csr_apply_kernel<T2, G><<<static_cast<unsigned>(nb), kThreads, 0, out.stream>>>(       \
                     static_cast<T2 *>(out.data), static_cast<const T2 *>(in.data), d_indptr, d_indices, vals, nrows)), \
                 (csr_apply_kernel<T2, G><<<static_cast<unsigned>(nb), kThreads, 0, out.stream>>>(       \
                     static_cast<T2 *>(out.data), static_cast<const T2 *>(in.data), d_indptr, d_indices, vals, nrows))); \
    } while (0)
    if (avg <= 2.0) PLB_CSR(1);
    else if (avg <= 8.0) PLB_CSR(4);
    else if (avg <= 64.0) PLB_CSR(8);
    else PLB_CSR(32);
#undef PLB_CSR
    out.launches++;
    PLB_CUDA(cudaGetLastError());
}

void peer_copy(StateVec &sv, void *dst, const void *src, uint64_t n16, int unroll) {
    sv.set_device();
    const unsigned nb = static_cast<unsigned>(std::min<uint64_t>((n16 + kThreads * 4 - 1) / (kThreads * 4), uint64_t(sv.sm_count) * 32));
    if (unroll >= 8)
        peer_copy_kernel<8><<<nb, kThreads, 0, sv.stream>>>(static_cast<double2 *>(dst), static_cast<const double2 *>(src), n16);
    else
        peer_copy_kernel<4><<<nb, kThreads, 0, sv.stream>>>(static_cast<double2 *>(dst), static_cast<const double2 *>(src), n16);
    PLB_CUDA(cudaGetLastError());
}

} // namespace plb200
