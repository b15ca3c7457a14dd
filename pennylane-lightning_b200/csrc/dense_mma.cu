// Dense 2^k x 2^k matrices on k = 5..7 wires as a batched complex GEMM on the tensor cores.
//
// Replaces custatevecApplyMatrix for QubitUnitary-sized operands (StateVectorCudaManaged.hpp:2531-2646;
// LQ: applyNCN's per-group mat-vec, GateImplementationsLM.hpp:407-498).  The state is read as a matrix
// X[t][c] — t = value of the k target bits, c = everything else — and Y = U X is computed tile by tile:
// one warp owns 8 columns (8 consecutive values of the three lowest index bits, i.e. one 128-byte line per
// row for c128, 64 bytes for c64), keeps the whole D x 8 complex output in its accumulator fragments, walks
// the K dimension reading each X row exactly once (so the update is in place), and takes the matrix
// fragments from shared memory (rows padded by 4 elements: conflict-free quarter-warp accesses).
//   c128: DMMA   mma.sync.m8n8k4  f64 — FP64 has no tcgen05 kind; DMMA is the FP64 tensor instruction of sm_100a
//   c64 : 3xTF32 mma.sync.m16n8k8 tf32 with fp32 accumulators: a = a_hi + a_lo, b = b_hi + b_lo,
//         a b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi   (fp32-level accuracy from tf32 operands)
// Complex product from four real ones: Yr = Ur Xr - Ui Xi, Yi = Ur Xi + Ui Xr.
// Arithmetic intensity: 4 D real FMA per amplitude against 2 x sizeof(amplitude) bytes: D = 32 is 8 flop/B in
// c128 — at the FP64 ridge of a B200 (37 TF/s / 6.55 TB/s = 5.6 flop/B), which is where the fusion pass hands
// over from the structured interpreter to this path.
// Applicability: every target and control bit >= 3 (the 8 columns are the three lowest index bits); otherwise
// the caller keeps its shared-memory mat-vec kernel.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device.cuh"

namespace plb200 {

namespace {

constexpr int kWarps = 4;

struct MmaArgs {
    BitInsert ins;    // zeros at target + control bit positions
    uint64_t cbits;   // control values deposited
    uint64_t ntiles;  // column tiles of 8
    uint64_t toff[128]; // state offset of target value t
};

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- c128: DMMA.  Fragment layouts of m8n8k4 (g = lane / 4, q = lane % 4):
//   A (8x4, row): a = A[g][q]      B (4x8, col): b = B[q][g]      C (8x8): c0 = C[g][2q], c1 = C[g][2q + 1]
template <int K>
__global__ void __launch_bounds__(kWarps * 32)
    dense_dmma_kernel(double2 *__restrict__ sv, const double2 *__restrict__ mat, const __grid_constant__ MmaArgs p) {
    constexpr int D = 1 << K, LD = D + 4, MT = D / 8, KS = D / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *us = reinterpret_cast<double2 *>(smem_raw); // D x LD, row-major
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) us[(i / D) * LD + (i % D)] = mat[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const uint64_t warp = static_cast<uint64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    const uint64_t nwarps = static_cast<uint64_t>(gridDim.x) * kWarps;
    for (uint64_t tile = warp; tile < p.ntiles; tile += nwarps) {
        const uint64_t base = insert_bits(tile << 3, p.ins) | p.cbits;
        double cr[MT][2], ci[MT][2];
#pragma unroll
        for (int m = 0; m < MT; m++) cr[m][0] = cr[m][1] = ci[m][0] = ci[m][1] = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < KS; ks++) {
            const double2 x = sv[base + p.toff[4 * ks + q] + g]; // B[q][g] of this k-step, real and imaginary
            const double nxi = -x.y;
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const double2 a = us[(8 * m + g) * LD + 4 * ks + q];
                dmma(cr[m][0], cr[m][1], a.x, x.x);
                dmma(cr[m][0], cr[m][1], a.y, nxi);
                dmma(ci[m][0], ci[m][1], a.x, x.y);
                dmma(ci[m][0], ci[m][1], a.y, x.x);
            }
        }
        __syncwarp(); // every lane has read its X rows before any row is overwritten
#pragma unroll
        for (int m = 0; m < MT; m++) {
            double2 *row = sv + base + p.toff[8 * m + g] + 2 * q;
            row[0] = make_double2(cr[m][0], ci[m][0]);
            row[1] = make_double2(cr[m][1], ci[m][1]);
        }
    }
}

// ---- c64: 3xTF32.  Fragment layouts of m16n8k8 (g = lane / 4, q = lane % 4):
//   A (16x8, row): a0 = A[g][q], a1 = A[g+8][q], a2 = A[g][q+4], a3 = A[g+8][q+4]
//   B (8x8, col):  b0 = B[q][g], b1 = B[q+4][g]
//   C (16x8):      c0 = C[g][2q], c1 = C[g][2q+1], c2 = C[g+8][2q], c3 = C[g+8][2q+1]
template <int K>
__global__ void __launch_bounds__(kWarps * 32)
    dense_tf32_kernel(float2 *__restrict__ sv, const float2 *__restrict__ mat, const __grid_constant__ MmaArgs p) {
    constexpr int D = 1 << K, LD = D + 4, MT = D / 16, KS = D / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *us = reinterpret_cast<float2 *>(smem_raw);
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) us[(i / D) * LD + (i % D)] = mat[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const uint64_t warp = static_cast<uint64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
    const uint64_t nwarps = static_cast<uint64_t>(gridDim.x) * kWarps;
    for (uint64_t tile = warp; tile < p.ntiles; tile += nwarps) {
        const uint64_t base = insert_bits(tile << 3, p.ins) | p.cbits;
        float cr[MT][4], ci[MT][4];
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
            for (int j = 0; j < 4; j++) cr[m][j] = ci[m][j] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < KS; ks++) {
            const float2 x0 = sv[base + p.toff[8 * ks + q] + g], x1 = sv[base + p.toff[8 * ks + q + 4] + g];
            // hi / lo split of the B fragments (real, imaginary, negated imaginary)
            uint32_t brh[2], brl[2], bih[2], bil[2], nih[2], nil[2];
            const float xr[2] = {x0.x, x1.x}, xi[2] = {x0.y, x1.y};
#pragma unroll
            for (int j = 0; j < 2; j++) {
                brh[j] = to_tf32(xr[j]), brl[j] = to_tf32(xr[j] - __uint_as_float(brh[j]));
                bih[j] = to_tf32(xi[j]), bil[j] = to_tf32(xi[j] - __uint_as_float(bih[j]));
                nih[j] = bih[j] ^ 0x80000000u, nil[j] = bil[j] ^ 0x80000000u;
            }
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const float2 e[4] = {us[(16 * m + g) * LD + 8 * ks + q], us[(16 * m + g + 8) * LD + 8 * ks + q],
                                     us[(16 * m + g) * LD + 8 * ks + q + 4], us[(16 * m + g + 8) * LD + 8 * ks + q + 4]};
                uint32_t arh[4], arl[4], aih[4], ail[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    arh[j] = to_tf32(e[j].x), arl[j] = to_tf32(e[j].x - __uint_as_float(arh[j]));
                    aih[j] = to_tf32(e[j].y), ail[j] = to_tf32(e[j].y - __uint_as_float(aih[j]));
                }
                // Yr += Ur Xr - Ui Xi
                mma_tf32(cr[m], arl, brh[0], brh[1]);
                mma_tf32(cr[m], arh, brl[0], brl[1]);
                mma_tf32(cr[m], ail, nih[0], nih[1]);
                mma_tf32(cr[m], aih, nil[0], nil[1]);
                mma_tf32(cr[m], arh, brh[0], brh[1]);
                mma_tf32(cr[m], aih, nih[0], nih[1]);
                // Yi += Ur Xi + Ui Xr
                mma_tf32(ci[m], arl, bih[0], bih[1]);
                mma_tf32(ci[m], arh, bil[0], bil[1]);
                mma_tf32(ci[m], ail, brh[0], brh[1]);
                mma_tf32(ci[m], aih, brl[0], brl[1]);
                mma_tf32(ci[m], arh, bih[0], bih[1]);
                mma_tf32(ci[m], aih, brh[0], brh[1]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MT; m++) {
            float2 *r0 = sv + base + p.toff[16 * m + g] + 2 * q, *r1 = sv + base + p.toff[16 * m + g + 8] + 2 * q;
            // two adjacent columns = one 16-byte store per row
            *reinterpret_cast<float4 *>(r0) = make_float4(cr[m][0], ci[m][0], cr[m][1], ci[m][1]);
            *reinterpret_cast<float4 *>(r1) = make_float4(cr[m][2], ci[m][2], cr[m][3], ci[m][3]);
        }
    }
}

template <typename T2, class Kern> void launch_mma(StateVec &sv, const COp &op, Kern kern, size_t smem) {
    // per device and cheap: DevicePool / batched adjoints drive several GPUs from one process
    PLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int k = op.k(), D = 1 << k;
    MmaArgs a;
    uint64_t tmask = 0;
    for (int b : op.tbits) tmask |= uint64_t{1} << b;
    const uint64_t involved = tmask | op.cmask;
    a.ins.n = 0;
    for (int b = 0; b < 64; b++)
        if (involved >> b & 1) a.ins.lowmask[a.ins.n++] = (uint64_t{1} << b) - 1;
    a.cbits = op.cval;
    a.ntiles = (uint64_t{1} << (sv.n - __builtin_popcountll(involved))) >> 3;
    for (int t = 0; t < D; t++) {
        uint64_t o = 0;
        for (int j = 0; j < k; j++)
            if (t >> j & 1) o |= uint64_t{1} << op.tbits[j];
        a.toff[t] = o;
    }
    std::vector<T2> h(static_cast<size_t>(D) * D);
    for (size_t i = 0; i < h.size(); i++) h[i] = mk<T2>(op.mat[i].real(), op.mat[i].imag());
    T2 *t = static_cast<T2 *>(sv.table_buf(h.size() * sizeof(T2)));
    // pageable source: the call returns once the bytes are staged, `h` may die afterwards
    PLB_CUDA(cudaMemcpyAsync(t, h.data(), h.size() * sizeof(T2), cudaMemcpyHostToDevice, sv.stream));
    const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((a.ntiles + kWarps - 1) / kWarps, uint64_t(sv.sm_count) * 8));
    kern<<<grid, kWarps * 32, smem, sv.stream>>>(static_cast<T2 *>(sv.data), t, a);
    PLB_CUDA(cudaGetLastError());
    sv.launches++;
}

} // namespace

// Applies an OP_DENSE on 5..7 wires through the tensor cores when its bits allow; false = not applicable.
bool launch_dense_mma(StateVec &sv, const COp &op) {
    if (op.kind != OP_DENSE) return false;
    const int k = op.k();
    if (std::getenv("PLB200_DENSE_MMA") && std::getenv("PLB200_DENSE_MMA")[0] == '0') return false;
    uint64_t involved = op.cmask;
    for (int b : op.tbits) involved |= uint64_t{1} << b;
    if (involved & 7u) return false; // the 8 columns of a tile are index bits 0..2
    if (sv.n - __builtin_popcountll(involved) < 3) return false;
    sv.set_device();
    if (sv.precision == 64) {
        if (k == 5) launch_mma<double2>(sv, op, dense_dmma_kernel<5>, sizeof(double2) * 32 * 36);
        else if (k == 6) launch_mma<double2>(sv, op, dense_dmma_kernel<6>, sizeof(double2) * 64 * 68);
        else return false;
    } else {
        if (k == 5) launch_mma<float2>(sv, op, dense_tf32_kernel<5>, sizeof(float2) * 32 * 36);
        else if (k == 6) launch_mma<float2>(sv, op, dense_tf32_kernel<6>, sizeof(float2) * 64 * 68);
        else if (k == 7) launch_mma<float2>(sv, op, dense_tf32_kernel<7>, sizeof(float2) * 128 * 132);
        else return false;
    }
    return true;
}

} // namespace plb200
