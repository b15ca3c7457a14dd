// Computational-basis sampling ON THE DEVICE for states too large for a host-side alias table
// (plb200_generate_samples_device; replaces Measurements::generate_samples, MeasurementsLQubit.hpp:646-679 /
// MeasurementsGPU.hpp generate_samples, for those sizes).
//
// The reference builds an alias table over the 2^k outcomes on the host (DiscreteRandomVariable,
// MeasurementKernels.hpp:308-381): sequential, and at 30 qubits 8 GiB of probabilities cross PCIe first.  Here the
// state never leaves HBM: ONE sweep produces the probability mass of every chunk of 2^12 amplitudes, a scan of those
// (2^18 doubles at 30 qubits, L2-resident) is the coarse CDF, and every shot — one warp each — draws a Philox
// uniform, binary-searches the coarse CDF and walks the one chunk it landed in.  Bits of wires outside the requested
// subset are dropped (sampling the full distribution and discarding wires IS sampling the marginal).
// Same distribution as the reference, not the same random stream: the alias path (engine.cu) stays the default up to
// 24 wires, where bit-exact samples under a shared seed can be checked against lightning.qubit.
#include <curand_kernel.h>

#include "device.cuh"

namespace plb200 {

namespace {
constexpr int kChunkBits = 12;
constexpr int kThreads = 256;

template <typename T2> __device__ __forceinline__ double prob_of(const T2 a) {
    return static_cast<double>(a.x) * static_cast<double>(a.x) + static_cast<double>(a.y) * static_cast<double>(a.y);
}

// mass of every chunk of 2^cb amplitudes: one warp per chunk, fixed summation order (deterministic)
template <typename T2>
__global__ void __launch_bounds__(kThreads) chunk_mass_kernel(const T2 *__restrict__ a, uint64_t nchunks, int cb, double *__restrict__ mass) {
    const uint64_t warp = (blockIdx.x * static_cast<uint64_t>(kThreads) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nchunks) return;
    const T2 *p = a + (warp << cb);
    double s = 0;
    for (uint32_t i = lane; i < (1u << cb); i += 32) s += prob_of(p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) mass[warp] = s;
}

// exclusive scan of the chunk masses by ONE block (<= 2^22 entries: a few microseconds of L2 traffic); cdf[n] = total
__global__ void __launch_bounds__(1024) scan_kernel(const double *__restrict__ mass, uint64_t n, double *__restrict__ cdf) {
    __shared__ double part[1024];
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    double s = 0;
    for (uint64_t i = lo; i < hi; i++) s += mass[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0;
        for (int i = 0; i < 1024; i++) {
            const double t = part[i];
            part[i] = run;
            run += t;
        }
        cdf[n] = run;
    }
    __syncthreads();
    double run = part[threadIdx.x];
    for (uint64_t i = lo; i < hi; i++) {
        cdf[i] = run;
        run += mass[i];
    }
}

struct SampleArgs {
    int k;        // number of reported wires
    int bits[64]; // index bit of reported wire j (wire order of the caller, first wire first)
};

// one warp per shot
template <typename T2>
__global__ void __launch_bounds__(kThreads)
    draw_kernel(const T2 *__restrict__ a, const double *__restrict__ cdf, uint64_t nchunks, int cb, uint64_t seed, int64_t shots,
                const __grid_constant__ SampleArgs sa, uint64_t *__restrict__ out) {
    const int64_t shot = (blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (shot >= shots) return;
    double u = 0;
    if (lane == 0) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, static_cast<unsigned long long>(shot), 0, &st);
        u = curand_uniform_double(&st); // (0, 1]
    }
    u = __shfl_sync(0xffffffffu, u, 0);
    const double total = cdf[nchunks];
    double target = (1.0 - u) * total; // [0, total)
    // last chunk whose exclusive prefix is <= target and whose mass is positive
    uint64_t lo = 0, hi = nchunks;
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (cdf[mid] <= target) lo = mid;
        else hi = mid;
    }
    while (lo > 0 && !(cdf[lo + 1] > cdf[lo])) lo--; // skip empty chunks (ties of the prefix)
    target -= cdf[lo];
    const T2 *p = a + (lo << cb);
    const uint32_t len = 1u << cb, seg = len >> 5 ? len >> 5 : 1; // contiguous segment per lane
    const uint32_t s0 = lane * seg;
    double mine = 0;
    if (s0 < len)
        for (uint32_t i = 0; i < seg; i++) mine += prob_of(p[s0 + i]);
    double incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    // first lane whose inclusive sum exceeds the target; rounding may leave none: the last lane with mass then
    const unsigned over = __ballot_sync(0xffffffffu, incl > target && mine > 0.0);
    const unsigned any = __ballot_sync(0xffffffffu, mine > 0.0);
    const int pick = over ? __ffs(over) - 1 : (any ? 31 - __clz(any) : 0);
    const double before = __shfl_sync(0xffffffffu, incl - mine, pick);
    if (lane != pick) return;
    double run = before;
    uint32_t idx = s0, last = s0;
    bool found = false;
    for (uint32_t i = 0; i < seg && s0 + i < len; i++) {
        const double q = prob_of(p[s0 + i]);
        if (q > 0.0) last = s0 + i;
        run += q;
        if (!found && run > target && q > 0.0) idx = s0 + i, found = true;
    }
    if (!found) idx = last;
    const uint64_t full = (lo << cb) | idx;
    for (int j = 0; j < sa.k; j++) out[shot * sa.k + j] = (full >> sa.bits[j]) & 1ull;
}
} // namespace

// bits[j]: index bit of the j-th reported wire.  host_out: shots x k values in {0, 1}.
void sample_device(StateVec &sv, const std::vector<int> &bits, int64_t shots, uint64_t seed, uint64_t *host_out) {
    sv.set_device();
    if (shots <= 0) return;
    const int k = static_cast<int>(bits.size());
    PLB_CHECK(k >= 1 && k <= 64, "generate_samples: bad number of wires");
    const int cb = sv.n < kChunkBits ? static_cast<int>(sv.n) : kChunkBits;
    const uint64_t nchunks = sv.length() >> cb;
    PLB_CHECK(nchunks <= (uint64_t{1} << 26), "generate_samples: state too large for the coarse CDF");
    double *mass = nullptr, *cdf = nullptr;
    uint64_t *dout = nullptr;
    cudaError_t e = cudaMalloc(&mass, nchunks * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&cdf, (nchunks + 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&dout, static_cast<size_t>(shots) * k * sizeof(uint64_t));
    if (e == cudaSuccess) {
        SampleArgs sa;
        sa.k = k;
        for (int j = 0; j < k; j++) sa.bits[j] = bits[j];
        const unsigned nb1 = static_cast<unsigned>((nchunks * 32 + kThreads - 1) / kThreads);
        const unsigned nb2 = static_cast<unsigned>((static_cast<uint64_t>(shots) * 32 + kThreads - 1) / kThreads);
        if (sv.precision == 64) {
            chunk_mass_kernel<double2><<<nb1, kThreads, 0, sv.stream>>>(static_cast<const double2 *>(sv.data), nchunks, cb, mass);
            scan_kernel<<<1, 1024, 0, sv.stream>>>(mass, nchunks, cdf);
            draw_kernel<double2><<<nb2, kThreads, 0, sv.stream>>>(static_cast<const double2 *>(sv.data), cdf, nchunks, cb, seed, shots, sa, dout);
        } else {
            chunk_mass_kernel<float2><<<nb1, kThreads, 0, sv.stream>>>(static_cast<const float2 *>(sv.data), nchunks, cb, mass);
            scan_kernel<<<1, 1024, 0, sv.stream>>>(mass, nchunks, cdf);
            draw_kernel<float2><<<nb2, kThreads, 0, sv.stream>>>(static_cast<const float2 *>(sv.data), cdf, nchunks, cb, seed, shots, sa, dout);
        }
        sv.launches += 3;
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(host_out, dout, static_cast<size_t>(shots) * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, sv.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sv.stream);
    }
    cudaFree(mass);
    cudaFree(cdf);
    cudaFree(dout);
    PLB_CUDA(e);
}

} // namespace plb200
