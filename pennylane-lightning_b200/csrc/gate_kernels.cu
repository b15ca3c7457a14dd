// Gate kernels of the plb200 engine — hand-written for sm_100a, HBM-bound by design.
//
// One kernel family per canonical op kind (common.hpp):
//   pairs_kernel     2x2 blocks on amplitude pairs   (all 1-qubit gates, CNOT/CRX/Toffoli via
//                    control masks, SWAP/IsingXY/excitations via arbitrary pair offsets)
//   paulirot_kernel  exp(-i theta/2 P) in one pass    (GateImplementationsLM.hpp:575-629)
//   diag_kernel      diagonal / parity-phase gates, touching only the control subspace
//   dense_kernel     dense 2^k matrices (k<=4 in registers, larger staged in shared memory)
//   project_kernel   control projector used by controlled generators
// Each amplitude is read and written exactly once per op with 128-bit accesses (c128: one
// LDG.128 per amplitude; c64: one LDG.64, vectorised to LDG.128 where index bit 0 is free);
// amplitudes outside the control subspace are never touched (SURVEY.md §8d byte counts).
#include "device.cuh"

#include <algorithm>

namespace plb200 {

namespace {

constexpr int kThreads = 256;

BitInsert make_insert(uint64_t mask) {
    BitInsert bi;
    bi.n = 0;
    for (int b = 0; b < 64; b++)
        if (mask >> b & 1) {
            PLB_CHECK(bi.n < 40, "too many involved bits");
            bi.lowmask[bi.n++] = (uint64_t{1} << b) - 1;
        }
    return bi;
}

inline unsigned grid_for(uint64_t items, int per_thread) {
    uint64_t threads = (items + per_thread - 1) / per_thread;
    uint64_t blocks = (threads + kThreads - 1) / kThreads;
    return static_cast<unsigned>(std::max<uint64_t>(blocks, 1));
}

// ------------------------------------------------------------------------------- pairs
template <typename T2> struct PairsArgs {
    BitInsert ins;
    uint64_t cbits;
    uint64_t ngroups;
    int nblocks;
    uint64_t offA[8], offB[8];
    T2 m[8][4];
};

template <typename T2, int U>
__global__ void __launch_bounds__(kThreads)
    pairs_kernel(T2 *__restrict__ sv, const __grid_constant__ PairsArgs<T2> p) {
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t base[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t g = tid + u * stride;
        ok[u] = g < p.ngroups;
        base[u] = insert_bits(g, p.ins) | p.cbits;
    }
    for (int b = 0; b < p.nblocks; b++) {
        T2 va[U], vb[U];
        const uint64_t oa = p.offA[b], ob = p.offB[b];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) {
                va[u] = sv[base[u] + oa];
                vb[u] = sv[base[u] + ob];
            }
        const T2 m0 = p.m[b][0], m1 = p.m[b][1], m2 = p.m[b][2], m3 = p.m[b][3];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) {
                sv[base[u] + oa] = cfma(m1, vb[u], cmul(m0, va[u]));
                sv[base[u] + ob] = cfma(m3, vb[u], cmul(m2, va[u]));
            }
    }
}

// c64 only: two adjacent amplitudes (index bit 0 free) per 128-bit access
template <int U>
__global__ void __launch_bounds__(kThreads)
    pairs_kernel_c64v2(float4 *__restrict__ sv, const __grid_constant__ PairsArgs<float2> p) {
    // p.ins already contains bit 0 as an inserted position; offsets are in amplitudes.
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t base[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t g = tid + u * stride;
        ok[u] = g < p.ngroups;
        base[u] = insert_bits(g, p.ins) | p.cbits;
    }
    for (int b = 0; b < p.nblocks; b++) {
        float4 va[U], vb[U];
        const uint64_t oa = p.offA[b], ob = p.offB[b];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) {
                va[u] = sv[(base[u] + oa) >> 1];
                vb[u] = sv[(base[u] + ob) >> 1];
            }
        const float2 m0 = p.m[b][0], m1 = p.m[b][1], m2 = p.m[b][2], m3 = p.m[b][3];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ok[u]) {
                const float2 a0 = make_float2(va[u].x, va[u].y), a1 = make_float2(va[u].z, va[u].w);
                const float2 b0 = make_float2(vb[u].x, vb[u].y), b1 = make_float2(vb[u].z, vb[u].w);
                const float2 ra0 = cfma(m1, b0, cmul(m0, a0)), ra1 = cfma(m1, b1, cmul(m0, a1));
                const float2 rb0 = cfma(m3, b0, cmul(m2, a0)), rb1 = cfma(m3, b1, cmul(m2, a1));
                sv[(base[u] + oa) >> 1] = make_float4(ra0.x, ra0.y, ra1.x, ra1.y);
                sv[(base[u] + ob) >> 1] = make_float4(rb0.x, rb0.y, rb1.x, rb1.y);
            }
    }
}

template <typename T2> void launch_pairs(StateVec &sv, const COp &op) {
    PairsArgs<T2> a;
    uint64_t tmask = 0;
    for (int b : op.tbits) tmask |= uint64_t{1} << b;
    const uint64_t involved = tmask | op.cmask;
    const int m = __builtin_popcountll(involved);
    a.cbits = op.cval;
    a.nblocks = static_cast<int>(op.blocks.size());
    PLB_CHECK(a.nblocks <= 8, "too many pair blocks");
    auto local_to_off = [&](uint32_t loc) {
        uint64_t o = 0;
        for (size_t j = 0; j < op.tbits.size(); j++)
            if (loc >> j & 1) o |= uint64_t{1} << op.tbits[j];
        return o;
    };
    for (int b = 0; b < a.nblocks; b++) {
        a.offA[b] = local_to_off(op.blocks[b].a);
        a.offB[b] = local_to_off(op.blocks[b].b);
        for (int q = 0; q < 4; q++)
            a.m[b][q] = mk<T2>(op.blocks[b].m[q].real(), op.blocks[b].m[q].imag());
    }
    T2 *d = static_cast<T2 *>(sv.data);
    if constexpr (sizeof(T2) == 8) {
        if (!(involved & 1) && sv.n - m >= 1) {
            a.ins = make_insert(involved | 1);
            a.ngroups = uint64_t{1} << (sv.n - m - 1);
            constexpr int U = 2;
            pairs_kernel_c64v2<U><<<grid_for(a.ngroups, U), kThreads, 0, sv.stream>>>(
                reinterpret_cast<float4 *>(d), a);
            sv.launches++;
            return;
        }
    }
    a.ins = make_insert(involved);
    a.ngroups = uint64_t{1} << (sv.n - m);
    constexpr int U = 2;
    pairs_kernel<T2, U><<<grid_for(a.ngroups, U), kThreads, 0, sv.stream>>>(d, a);
    sv.launches++;
}

// ---------------------------------------------------------------------------- PauliRot
template <typename T2> struct PauliRotArgs {
    BitInsert ins; // pivot bit (lowest flip bit)
    uint64_t ngroups, x, z;
    T2 c, w;
};

template <typename T2, int U>
__global__ void __launch_bounds__(kThreads)
    paulirot_kernel(T2 *__restrict__ sv, const __grid_constant__ PauliRotArgs<T2> p) {
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t j0[U];
    T2 a[U], b[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t g = tid + u * stride;
        ok[u] = g < p.ngroups;
        j0[u] = insert_bits(g, p.ins);
        if (ok[u]) {
            a[u] = sv[j0[u]];
            b[u] = sv[j0[u] ^ p.x];
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
        if (ok[u]) {
            const uint64_t j1 = j0[u] ^ p.x;
            // new[j0] = c a + w sgn(j1) b ; new[j1] = c b + w sgn(j0) a
            const T2 w0 = (__popcll(j1 & p.z) & 1) ? cneg(p.w) : p.w;
            const T2 w1 = (__popcll(j0[u] & p.z) & 1) ? cneg(p.w) : p.w;
            sv[j0[u]] = cfma(w0, b[u], cmul(p.c, a[u]));
            sv[j1] = cfma(w1, a[u], cmul(p.c, b[u]));
        }
}

template <typename T2> void launch_paulirot(StateVec &sv, const COp &op) {
    PauliRotArgs<T2> a;
    a.x = 0;
    for (int b : op.tbits) a.x |= uint64_t{1} << b;
    a.z = op.pmask;
    a.ins = make_insert(a.x & (~a.x + 1));
    a.ngroups = uint64_t{1} << (sv.n - 1);
    a.c = mk<T2>(op.blocks[0].m[0].real(), op.blocks[0].m[0].imag());
    a.w = mk<T2>(op.blocks[0].m[1].real(), op.blocks[0].m[1].imag());
    constexpr int U = 2;
    paulirot_kernel<T2, U><<<grid_for(a.ngroups, U), kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), a);
    sv.launches++;
}

// -------------------------------------------------------------------------------- diag
enum DiagMode { DIAG_SCALAR = 0, DIAG_SMALL = 1, DIAG_TABLE = 2, DIAG_PARITY = 3 };
template <typename T2> struct DiagArgs {
    BitInsert ins; // control bits
    uint64_t cbits, ngroups, pmask;
    int k;
    int tbits[24];
    T2 small[16];
    const T2 *table;
};

template <typename T2, int MODE>
__device__ __forceinline__ T2 diag_factor(const DiagArgs<T2> &p, uint64_t idx) {
    if constexpr (MODE == DIAG_SCALAR) {
        return p.small[0];
    } else if constexpr (MODE == DIAG_PARITY) {
        return p.small[__popcll(idx & p.pmask) & 1];
    } else {
        unsigned t = 0;
        for (int j = 0; j < p.k; j++) t |= static_cast<unsigned>((idx >> p.tbits[j]) & 1) << j;
        if constexpr (MODE == DIAG_SMALL) return p.small[t];
        else return p.table[t];
    }
}

template <typename T2, int MODE, int U>
__global__ void __launch_bounds__(kThreads)
    diag_kernel(T2 *__restrict__ sv, const __grid_constant__ DiagArgs<T2> p) {
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t idx[U];
    T2 v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t g = tid + u * stride;
        ok[u] = g < p.ngroups;
        idx[u] = insert_bits(g, p.ins) | p.cbits;
        if (ok[u]) v[u] = sv[idx[u]];
    }
#pragma unroll
    for (int u = 0; u < U; u++)
        if (ok[u]) sv[idx[u]] = cmul(v[u], diag_factor<T2, MODE>(p, idx[u]));
}

// c64, bit 0 not a control: 2 amplitudes per float4
template <int MODE, int U>
__global__ void __launch_bounds__(kThreads)
    diag_kernel_c64v2(float4 *__restrict__ sv, const __grid_constant__ DiagArgs<float2> p) {
    const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t idx[U];
    float4 v[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t g = tid + u * stride;
        ok[u] = g < p.ngroups;
        idx[u] = insert_bits(g, p.ins) | p.cbits;
        if (ok[u]) v[u] = sv[idx[u] >> 1];
    }
#pragma unroll
    for (int u = 0; u < U; u++)
        if (ok[u]) {
            const float2 r0 = cmul(make_float2(v[u].x, v[u].y), diag_factor<float2, MODE>(p, idx[u]));
            const float2 r1 = cmul(make_float2(v[u].z, v[u].w), diag_factor<float2, MODE>(p, idx[u] | 1));
            sv[idx[u] >> 1] = make_float4(r0.x, r0.y, r1.x, r1.y);
        }
}

template <typename T2, int MODE> void launch_diag_mode(StateVec &sv, DiagArgs<T2> &a, uint64_t cmask) {
    const int c = __builtin_popcountll(cmask);
    constexpr int U = 4;
    if constexpr (sizeof(T2) == 8) {
        if (!(cmask & 1) && sv.n - c >= 1) {
            a.ins = make_insert(cmask | 1);
            a.ngroups = uint64_t{1} << (sv.n - c - 1);
            diag_kernel_c64v2<MODE, U><<<grid_for(a.ngroups, U), kThreads, 0, sv.stream>>>(
                reinterpret_cast<float4 *>(sv.data), a);
            sv.launches++;
            return;
        }
    }
    a.ins = make_insert(cmask);
    a.ngroups = uint64_t{1} << (sv.n - c);
    diag_kernel<T2, MODE, U><<<grid_for(a.ngroups, U), kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), a);
    sv.launches++;
}

template <typename T2> void launch_diag(StateVec &sv, const COp &op) {
    DiagArgs<T2> a;
    a.cbits = op.cval;
    a.pmask = op.pmask;
    a.k = op.k();
    a.table = nullptr;
    if (op.parity) {
        a.small[0] = mk<T2>(op.pd[0].real(), op.pd[0].imag());
        a.small[1] = mk<T2>(op.pd[1].real(), op.pd[1].imag());
        launch_diag_mode<T2, DIAG_PARITY>(sv, a, op.cmask);
        return;
    }
    PLB_CHECK(a.k <= 24, "diagonal table too large");
    for (int j = 0; j < a.k; j++) a.tbits[j] = op.tbits[j];
    if (a.k == 0) {
        a.small[0] = mk<T2>(op.diag[0].real(), op.diag[0].imag());
        launch_diag_mode<T2, DIAG_SCALAR>(sv, a, op.cmask);
    } else if (a.k <= 4) {
        for (size_t i = 0; i < op.diag.size(); i++) a.small[i] = mk<T2>(op.diag[i].real(), op.diag[i].imag());
        launch_diag_mode<T2, DIAG_SMALL>(sv, a, op.cmask);
    } else {
        std::vector<T2> h(op.diag.size());
        for (size_t i = 0; i < h.size(); i++) h[i] = mk<T2>(op.diag[i].real(), op.diag[i].imag());
        T2 *t = static_cast<T2 *>(sv.table_buf(h.size() * sizeof(T2)));
        // pageable source: the call returns once the bytes are staged (no stream sync needed for h's lifetime)
        PLB_CUDA(cudaMemcpyAsync(t, h.data(), h.size() * sizeof(T2), cudaMemcpyHostToDevice, sv.stream));
        a.table = t;
        launch_diag_mode<T2, DIAG_TABLE>(sv, a, op.cmask);
    }
}

// ------------------------------------------------------------------------------- dense
template <typename T2> struct DenseArgs {
    BitInsert ins;
    uint64_t cbits, ngroups;
    int k;
    uint64_t off[16]; // K <= 4: state offsets of the local indices
    int tbits[20];
    const T2 *mat;  // row-major (small K) ; transposed (generic)
};

template <typename T2, int K>
__global__ void __launch_bounds__(128)
    dense_small_kernel(T2 *__restrict__ sv, const __grid_constant__ DenseArgs<T2> p) {
    constexpr int D = 1 << K;
    __shared__ T2 sm[D * D];
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sm[i] = p.mat[i];
    __syncthreads();
    const uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= p.ngroups) return;
    const uint64_t base = insert_bits(g, p.ins) | p.cbits;
    T2 v[D];
#pragma unroll
    for (int c = 0; c < D; c++) v[c] = sv[base + p.off[c]];
#pragma unroll
    for (int r = 0; r < D; r++) {
        T2 acc = cmul(sm[r * D], v[0]);
#pragma unroll
        for (int c = 1; c < D; c++) acc = cfma(sm[r * D + c], v[c], acc);
        sv[base + p.off[r]] = acc;
    }
}

// generic K (5..): one group per CTA iteration, amplitudes staged in shared memory, matrix
// (transposed) streamed from L2.
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    dense_generic_kernel(T2 *__restrict__ sv, const __grid_constant__ DenseArgs<T2> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T2 *sm = reinterpret_cast<T2 *>(smem_raw);
    const int D = 1 << p.k;
    for (uint64_t g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
        const uint64_t base = insert_bits(g, p.ins) | p.cbits;
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            uint64_t o = 0;
            for (int j = 0; j < p.k; j++) o |= static_cast<uint64_t>((c >> j) & 1) << p.tbits[j];
            sm[c] = sv[base + o];
        }
        __syncthreads();
        for (int r = threadIdx.x; r < D; r += blockDim.x) {
            T2 acc = mk<T2>(0.0, 0.0);
            for (int c = 0; c < D; c++) acc = cfma(p.mat[static_cast<size_t>(c) * D + r], sm[c], acc);
            uint64_t o = 0;
            for (int j = 0; j < p.k; j++) o |= static_cast<uint64_t>((r >> j) & 1) << p.tbits[j];
            sv[base + o] = acc;
        }
        __syncthreads();
    }
}

template <typename T2> void launch_dense(StateVec &sv, const COp &op) {
    if (op.k() >= 5 && launch_dense_mma(sv, op)) return; // batched complex GEMM on the tensor cores
    DenseArgs<T2> a;
    const int k = op.k();
    const int D = 1 << k;
    uint64_t tmask = 0;
    for (int b : op.tbits) tmask |= uint64_t{1} << b;
    const uint64_t involved = tmask | op.cmask;
    const int m = __builtin_popcountll(involved);
    a.ins = make_insert(involved);
    a.cbits = op.cval;
    a.ngroups = uint64_t{1} << (sv.n - m);
    a.k = k;
    for (int j = 0; j < k; j++) a.tbits[j] = op.tbits[j];
    std::vector<T2> h(static_cast<size_t>(D) * D);
    const bool small = k <= 4;
    for (int r = 0; r < D; r++)
        for (int c = 0; c < D; c++) {
            const cd e = op.mat[static_cast<size_t>(r) * D + c];
            h[small ? static_cast<size_t>(r) * D + c : static_cast<size_t>(c) * D + r] = mk<T2>(e.real(), e.imag());
        }
    T2 *t = static_cast<T2 *>(sv.table_buf(h.size() * sizeof(T2)));
    // pageable source: cudaMemcpyAsync returns once the bytes are staged, so `h` may die without a stream sync
    PLB_CUDA(cudaMemcpyAsync(t, h.data(), h.size() * sizeof(T2), cudaMemcpyHostToDevice, sv.stream));
    a.mat = t;
    T2 *d = static_cast<T2 *>(sv.data);
    if (small) {
        for (int c = 0; c < D; c++) {
            uint64_t o = 0;
            for (int j = 0; j < k; j++)
                if (c >> j & 1) o |= uint64_t{1} << op.tbits[j];
            a.off[c] = o;
        }
        const unsigned grid = static_cast<unsigned>((a.ngroups + 127) / 128);
        switch (k) {
        case 1:
            dense_small_kernel<T2, 1><<<grid, 128, 0, sv.stream>>>(d, a);
            break;
        case 2:
            dense_small_kernel<T2, 2><<<grid, 128, 0, sv.stream>>>(d, a);
            break;
        case 3:
            dense_small_kernel<T2, 3><<<grid, 128, 0, sv.stream>>>(d, a);
            break;
        default:
            dense_small_kernel<T2, 4><<<grid, 128, 0, sv.stream>>>(d, a);
            break;
        }
    } else {
        PLB_CHECK(k <= 11, "applyMatrix: dense matrices on more than 11 wires are not supported");
        const size_t smem = static_cast<size_t>(D) * sizeof(T2);
        const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(a.ngroups, uint64_t(sv.sm_count) * 8));
        dense_generic_kernel<T2><<<grid, kThreads, smem, sv.stream>>>(d, a);
    }
    sv.launches++;
}

// ----------------------------------------------------------------------------- project
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    project_kernel(T2 *__restrict__ sv, uint64_t len, uint64_t cmask, uint64_t cval) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += stride)
        if ((i & cmask) != cval) sv[i] = mk<T2>(0.0, 0.0);
}

template <typename T2> void launch_project(StateVec &sv, const COp &op) {
    const uint64_t len = sv.length();
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((len + kThreads - 1) / kThreads, uint64_t(sv.sm_count) * 16));
    project_kernel<T2><<<grid, kThreads, 0, sv.stream>>>(static_cast<T2 *>(sv.data), len, op.cmask, op.cval);
    sv.launches++;
}

template <typename T2> void launch_typed(StateVec &sv, const COp &op) {
    switch (op.kind) {
    case OP_PAIRS:
        if (op.parity) launch_paulirot<T2>(sv, op);
        else launch_pairs<T2>(sv, op);
        break;
    case OP_DIAG:
        launch_diag<T2>(sv, op);
        break;
    case OP_DENSE:
        launch_dense<T2>(sv, op);
        break;
    case OP_PROJECT:
        launch_project<T2>(sv, op);
        break;
    }
    PLB_CUDA(cudaGetLastError());
}

} // namespace

void launch_op(StateVec &sv, const COp &op) {
    sv.set_device();
    if (sv.precision == 64) launch_typed<double2>(sv, op);
    else launch_typed<float2>(sv, op);
}

void launch_ops(StateVec &sv, const std::vector<COp> &ops) {
    for (const auto &op : ops) launch_op(sv, op);
}

} // namespace plb200
