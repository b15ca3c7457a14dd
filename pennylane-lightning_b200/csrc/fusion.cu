// Cache-blocked ("tile") execution of a tape of canonical ops — the gate-fusion pass, and its
// two-state variant that runs the adjoint-Jacobian sweep.
//
// The reference sweeps the whole 2^n vector once per gate ("one full HBM
// read+write per gate ... no fusion anywhere").  Here the tape is scheduled into PASSES: a pass
// owns a set T of M index bits (always including the lowest LOW bits so that every global access
// is a full 128-byte line) and executes, in program order, every pending op whose non-diagonal
// target bits lie in T — controls and diagonal factors may sit on any bit, since outside T they
// are uniform per tile.  One CTA loads a tile of 2^M amplitudes into shared memory, runs the
// pass's ROUNDS on it and writes it back: ONE HBM read+write for the whole group of gates.
// Inside the tile a round picks R "register bits"; what a thread does with them is tile_exec.cuh.
//
// Adjoint variant (AdjointJacobianLQubit.hpp:269-314 as ONE sweep per block of ops): the tile of
// lambda and the tile of H·lambda travel together; walking the tape backwards, a trainable op first
// contributes its overlap Im<H lambda| G |lambda> (G = (controlled) X / Y / Z-parity generator)
// from registers — accumulated per warp in shared memory, per CTA in a partials buffer, reduced in a fixed order — and then
// its inverse is applied to both tiles.  No mu copy, no per-op HBM sweep.
//
// Gate commutation used by the scheduler: ops acting on disjoint bit sets commute, nothing else.
#include "fusion.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>

#include "jit_codegen.hpp"
#include "jit_runtime.hpp"
#include "tile_exec.cuh"

#if defined(PLB200_HOST_EMU)
#include <dlfcn.h>
#include <unistd.h>

#include <fstream>
#include <map>
#endif

namespace plb200 {

namespace {

using namespace tile;

#if !defined(PLB200_HOST_EMU)
// Overlap accumulation of the two-state (adjoint) passes, in a FIXED order: the lanes of a warp are summed by a
// shuffle tree, every warp owns a row of the CTA's shared accumulators (no atomics; a warp walks its tiles in
// order), the rows are added in warp order at the end of the kernel into this CTA's row of a global partials
// buffer, and overlap_reduce_kernel adds the CTAs' rows in CTA order.  Same grid -> same bits, run after run.
struct WarpReduce {
    double *acc; // this warp's row
    __host__ __device__ __forceinline__ void operator()(int slot, double s) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) acc[slot] += s;
#else
        (void)slot, (void)s;
#endif
    }
};
// out[i] = sum over CTAs of part[i * nctas + b] (slot-major partials): one block per slot, thread t adds the CTAs
// t, t + 256, ... in that order, then the 256 partial sums are added by a fixed shuffle / shared-memory tree
__global__ void __launch_bounds__(256) overlap_reduce_kernel(const double *__restrict__ part, unsigned nctas, double *__restrict__ out) {
    __shared__ double sm[8];
    const double *col = part + static_cast<size_t>(blockIdx.x) * nctas;
    double s = 0;
    for (unsigned b = threadIdx.x; b < nctas; b += 256) s += col[b];
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

template <typename T2, class Cfg, bool EXT>
__global__ void __launch_bounds__(1 << (Cfg::M - Cfg::R), Cfg::MINB)
    tile_kernel(T2 *__restrict__ sv0, T2 *__restrict__ sv1, double *__restrict__ acc_g,
                const __grid_constant__ PassParams<T2> pp) {
    using E = Exec<T2, Cfg, EXT>;
    constexpr int M = Cfg::M, LOW = Cfg::LOW, NS = Cfg::NS, NT = E::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *tile0 = smem_raw;
    unsigned char *tile1 = smem_raw + (NS == 2 ? (sizeof(T2) << M) : 0);
    uint64_t *goff = reinterpret_cast<uint64_t *>(smem_raw + NS * (sizeof(T2) << M));
    double *acc = reinterpret_cast<double *>(goff + (1 << (M - LOW)));
    static_assert(sizeof(PassParams<T2>) <= 32764, "kernel parameter space");

    // global offset of every 2^LOW-amplitude line of a tile, from the tile's bit positions
    for (int i = threadIdx.x; i < (1 << (M - LOW)); i += NT) goff[i] = tile_line_offset<M, LOW>(pp.hdr, i);
    if constexpr (NS == 2)
        for (int i = threadIdx.x; i < pp.hdr.nslots * (NT / 32); i += NT) acc[i] = 0.0;
    __syncthreads();
    const uint32_t tid = threadIdx.x;
    const int nrounds = pp.hdr.nrounds;
    const WarpReduce red{acc + (threadIdx.x >> 5) * pp.hdr.nslots};

    for (uint64_t t = blockIdx.x; t < pp.hdr.ntiles; t += gridDim.x) {
        const uint64_t base = insert_bits(t, pp.hdr.tile_ins);
        // pull the NEXT tile of this CTA from HBM into L2 while this one is processed (one prefetch
        // per 128-byte line): no registers or shared memory held, the later load hits L2
        if (t + gridDim.x < pp.hdr.ntiles) {
            const uint64_t nbase = insert_bits(t + gridDim.x, pp.hdr.tile_ins);
            for (int l = tid; l < (1 << (M - LOW)); l += NT) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sv0 + (nbase | goff[l])));
                if constexpr (NS == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv1 + (nbase | goff[l])));
            }
        }
        E::load_tile(tid, base, goff, sv0, tile0);
        if constexpr (NS == 2) E::load_tile(tid, base, goff, sv1, tile1);
        __syncthreads();
        for (int r = 0; r < nrounds; r++) {
            E::round(pp, r, tid, base, tile0, tile1, red);
            __syncthreads();
        }
        E::store_tile(tid, base, goff, sv0, tile0);
        if constexpr (NS == 2) E::store_tile(tid, base, goff, sv1, tile1);
        __syncthreads();
    }
    if constexpr (NS == 2) { // acc_g: the partials buffer, slot-major (one row of gridDim.x CTAs per slot)
        for (int i = threadIdx.x; i < pp.hdr.nslots; i += NT) {
            double t = 0.0;
            for (int w = 0; w < NT / 32; w++) t += acc[w * pp.hdr.nslots + i];
            acc_g[static_cast<size_t>(i) * gridDim.x + blockIdx.x] = t;
        }
    }
}
#endif // !PLB200_HOST_EMU

#if defined(PLB200_HOST_EMU)
// Test-only (PLB200_EMU_JIT=1): the SPECIALISED source of the pass (jit_codegen.hpp) compiled with g++ under
// -DPLB_JIT_HOST and run thread by thread on host memory, instead of the interpreter.
int64_t g_emu_jit_passes = 0;
template <typename T2, class Cfg>
bool emulate_pass_jit(T2 *sv0, const PassParams<T2> &pp, const jit::Route &route = jit::Route{},
                      const jit::RouteParams<T2> *rp = nullptr, T2 *sv1 = nullptr, double *acc = nullptr) {
    {
        const std::string src = jit::generate_pass_source<T2, Cfg>(pp, route);
        if (src.empty()) return false;
        static std::map<uint64_t, void (*)(void *, const void *, const void *, void *, double *)> cache;
        const uint64_t key = jit::fnv1a(src) ^ (static_cast<uint64_t>(src.size()) << 40);
        auto it = cache.find(key);
        if (it == cache.end()) {
            const char *dir = std::getenv("PLB200_EMU_JIT_DIR");
            static int serial = 0; // one counter per instantiation: the precision tag keeps the names apart
            const std::string base = std::string(dir ? dir : "/tmp") + "/plb_jit_" + std::to_string(::getpid()) + "_" +
                                     (sizeof(T2) == 16 ? "d" : "f") + (Cfg::NS == 2 ? "a" : "") + std::to_string(serial++);
            {
                std::ofstream o(base + ".cpp");
                o << src;
            }
            const std::string cmd = "g++ -std=c++17 -O1 -w -shared -fPIC -ffp-contract=off -DPLB_JIT_HOST -o " + base + ".so " + base + ".cpp";
            if (std::system(cmd.c_str()) != 0) fail("emu jit: g++ failed on " + base + ".cpp");
            void *h = dlopen((base + ".so").c_str(), RTLD_NOW | RTLD_LOCAL);
            if (!h) fail(std::string("emu jit: dlopen failed: ") + dlerror());
            auto fn = reinterpret_cast<void (*)(void *, const void *, const void *, void *, double *)>(dlsym(h, "plb_pass_host"));
            if (!fn) fail("emu jit: plb_pass_host missing");
            it = cache.emplace(key, fn).first;
            if (!std::getenv("PLB200_EMU_JIT_KEEP")) {
                ::unlink((base + ".cpp").c_str());
                ::unlink((base + ".so").c_str());
            }
        }
        const jit::RouteParams<T2> none{};
        it->second(sv0, &pp, rp ? rp : &none, sv1, acc);
        g_emu_jit_passes++;
        return true;
    }
}
// Thread-by-thread execution of one pass on host memory (test-only emulation, and the reference
// semantics of tile_kernel: threads of a round are independent, rounds are separated by barriers).
struct HostReduce {
    double *acc;
    __host__ __device__ void operator()(int slot, double s) const { acc[slot] += s; }
};
template <typename T2, class Cfg, bool EXT>
void emulate_pass(T2 *sv0, T2 *sv1, double *acc, const PassParams<T2> &pp) {
    using E = Exec<T2, Cfg, EXT>;
    std::vector<uint64_t> goff_v(size_t{1} << (Cfg::M - Cfg::LOW));
    for (size_t i = 0; i < goff_v.size(); i++)
        goff_v[i] = tile_line_offset<Cfg::M, Cfg::LOW>(pp.hdr, static_cast<int>(i));
    const uint64_t *goff = goff_v.data();
    std::vector<unsigned char> t0(sizeof(T2) << Cfg::M), t1(sizeof(T2) << Cfg::M);
    const HostReduce red{acc};
    for (uint64_t t = 0; t < pp.hdr.ntiles; t++) {
        const uint64_t base = insert_bits(t, pp.hdr.tile_ins);
        for (uint32_t tid = 0; tid < static_cast<uint32_t>(E::NT); tid++) {
            E::load_tile(tid, base, goff, sv0, t0.data());
            if (Cfg::NS == 2) E::load_tile(tid, base, goff, sv1, t1.data());
        }
        for (int r = 0; r < pp.hdr.nrounds; r++)
            for (uint32_t tid = 0; tid < static_cast<uint32_t>(E::NT); tid++)
                E::round(pp, r, tid, base, t0.data(), t1.data(), red);
        for (uint32_t tid = 0; tid < static_cast<uint32_t>(E::NT); tid++) {
            E::store_tile(tid, base, goff, sv0, t0.data());
            if (Cfg::NS == 2) E::store_tile(tid, base, goff, sv1, t1.data());
        }
    }
}
#endif // PLB200_HOST_EMU

// --------------------------------------------------------------------------- host scheduler
struct FOp {
    uint64_t nd = 0;  // non-diagonal target bits
    uint64_t all = 0; // every involved bit
    bool fusable = false;
    // parity form of a fusable diagonal
    uint64_t pmask = 0;
    cd d[2];
};

// the in-place LU of the tile interpreter needs an invertible block (every gate is; a user matrix or
// a generator need not be): others run through the stand-alone kernels
bool well_conditioned(const cd *m) {
    const double nrm = std::norm(m[0]) + std::norm(m[1]) + std::norm(m[2]) + std::norm(m[3]);
    return std::abs(m[0] * m[3] - m[1] * m[2]) > 1e-6 * nrm;
}

// SWAP of two index bits: one (01, 10) block [[0, 1], [1, 0]]
bool is_swap2(const COp &op) {
    if (op.kind != OP_PAIRS || op.parity || op.tbits.size() != 2 || op.blocks.size() != 1) return false;
    const Block2 &b = op.blocks[0];
    return ((b.a == 1 && b.b == 2) || (b.a == 2 && b.b == 1)) && b.m[0] == cd(0.0) && b.m[3] == cd(0.0) &&
           b.m[1] == cd(1.0) && b.m[2] == cd(1.0);
}

// Two- / four-bit pair ops inside tile passes (K_PAIR2 / K_PAIR4).  They need a specialised kernel; PLB200_FUSE_PAIR2
// forces the choice.
bool pair2_enabled(int n) {
    if (const char *e = std::getenv("PLB200_FUSE_PAIR2")) return e[0] != '0';
#if defined(PLB200_HOST_EMU)
    (void)n;
    return false;
#else
    // whenever specialised kernels are in play: compiled at first sight (PLB200_JIT=sync), or in the background
    // (default tier) — until a pass's kernel exists the pass is cut at its pair ops (build_schedule)
    return jit::mode() != jit::Mode::Off && jit::available(nullptr) && n >= jit::min_qubits();
#endif
}
bool is_pair2(const COp &op) {
    if (op.kind != OP_PAIRS || op.parity || op.tbits.size() != 2 || op.blocks.empty() || op.blocks.size() > 2) return false;
    for (const Block2 &b : op.blocks)
        if (b.a > 3 || b.b > 3 || b.a == b.b || !well_conditioned(b.m)) return false;
    return true;
}

// DoubleExcitation-like: one or two 2x2 blocks on a 4-bit target space (K_PAIR4, needs 4 register bits)
bool is_pair4(const COp &op) {
    if (op.kind != OP_PAIRS || op.parity || op.tbits.size() != 4 || op.blocks.empty() || op.blocks.size() > 2) return false;
    for (const Block2 &b : op.blocks)
        if (b.a > 15 || b.b > 15 || b.a == b.b || !well_conditioned(b.m)) return false;
    return true;
}

// dense 4x4 on two wires (QubitUnitary, K_DENSE2)
bool is_dense2(const COp &op) { return op.kind == OP_DENSE && op.tbits.size() == 2 && op.mat.size() == 16; }

FOp classify(const COp &op, bool pair2 = false, int reg_bits = 4) {
    FOp f;
    uint64_t t = 0;
    for (int b : op.tbits) t |= uint64_t{1} << b;
    f.all = t | op.cmask | (op.parity ? op.pmask : 0);
    if (op.kind == OP_PROJECT) f.all = ~uint64_t{0};
    if (op.kind == OP_PAIRS && !op.parity && op.tbits.size() == 1 && op.blocks.size() == 1 && op.blocks[0].a == 0 &&
        op.blocks[0].b == 1 && well_conditioned(op.blocks[0].m)) {
        f.fusable = true;
        f.nd = t;
    } else if (is_swap2(op)) {
        f.fusable = true;
        f.nd = t;
    } else if (pair2 && (is_pair2(op) || (reg_bits >= 4 && is_pair4(op)) || is_dense2(op))) {
        f.fusable = true;
        f.nd = t;
    } else if (op.kind == OP_DIAG) {
        if (op.parity) {
            f.fusable = true, f.pmask = op.pmask, f.d[0] = op.pd[0], f.d[1] = op.pd[1];
        } else if (op.k() == 0) {
            f.fusable = true, f.pmask = 0, f.d[0] = f.d[1] = op.diag[0];
        } else if (op.k() == 1) {
            f.fusable = true, f.pmask = t, f.d[0] = op.diag[0], f.d[1] = op.diag[1];
        } else if (op.k() == 2 && op.diag[0] == op.diag[3] && op.diag[1] == op.diag[2]) {
            f.fusable = true, f.pmask = t, f.d[0] = op.diag[0], f.d[1] = op.diag[1];
        }
        if (!op.parity && op.k() == 0 && op.cmask == 0) f.all = 0; // global scalar commutes with everything
    }
    if (!f.fusable) f.nd = t;
    return f;
}

FOp classify(const AdjItem &it, bool pair2 = false, int reg_bits = 4) {
    if (!it.overlap) return classify(it.op, pair2, reg_bits);
    FOp f;
    const PauliWordMask &w = it.pw;
    f.nd = w.x;
    f.all = w.x | w.z | w.cmask;
    const int nx = __builtin_popcountll(w.x);
    // fused overlaps: Z-parity words, or a single X / Y factor (no extra Z factors)
    f.fusable = (nx == 0) || (nx == 1 && (w.z == 0 || w.z == w.x));
    return f;
}

// ops executable (in order) with non-diagonal bits restricted to S; `pending` lists candidates
void simulate(const std::vector<FOp> &f, const std::vector<int> &pending, uint64_t S, uint64_t full,
              std::vector<int> &out) {
    out.clear();
    uint64_t blocked = 0;
    for (int i : pending) {
        const FOp &o = f[i];
        if (o.fusable && (o.nd & ~S) == 0 && (o.all & blocked) == 0) out.push_back(i);
        else {
            blocked |= o.all;
            if ((blocked & full) == full) break;
        }
    }
}

// greedy growth of a bit set (capacity cap) maximising the number of executable ops
uint64_t grow(const std::vector<FOp> &f, const std::vector<int> &pending, uint64_t S, int cap, uint64_t allowed,
              uint64_t full, std::vector<int> &exec) {
    std::vector<int> tmp;
    simulate(f, pending, S, full, exec);
    while (__builtin_popcountll(S) < cap) {
        // candidate bits: non-diagonal bits of pending ops not in S
        uint64_t cand = 0;
        for (int i : pending) cand |= f[i].nd;
        cand &= allowed & ~S;
        if (!cand) break;
        size_t best = exec.size();
        int best_bit = -1;
        for (int b = 0; b < 64; b++) {
            if (!(cand >> b & 1)) continue;
            simulate(f, pending, S | (uint64_t{1} << b), full, tmp);
            if (tmp.size() > best) best = tmp.size(), best_bit = b;
        }
        if (best_bit < 0) {
            // no single bit helps (e.g. the next op needs two new bits): take the first blocked op's bits
            uint64_t need = 0;
            uint64_t blocked = 0;
            for (int i : pending) {
                const FOp &o = f[i];
                if (o.fusable && (o.all & blocked) == 0 && (o.nd & ~S) != 0 && (o.nd & ~allowed) == 0 &&
                    __builtin_popcountll(S | o.nd) <= cap) {
                    need = o.nd & ~S;
                    break;
                }
                if (!(o.fusable && (o.nd & ~S) == 0 && (o.all & blocked) == 0)) blocked |= o.all;
            }
            if (!need) break;
            S |= need;
        } else
            S |= uint64_t{1} << best_bit;
        simulate(f, pending, S, full, exec);
    }
    return S;
}

#if defined(PLB200_HOST_EMU)
int64_t g_kind_hist[32] = {0}; // encoder coverage: ops emitted per kind (test-only)
int64_t g_plan_hits = 0;       // tapes scheduled from a cached plan (test-only)
#endif

struct HostPass {
    std::vector<int> tbits;               // sorted ascending, size M
    std::vector<std::vector<int>> rounds; // item indices per round
    std::vector<uint64_t> round_bits;     // register bits (global bit masks) per round
};

struct Step {
    int op = -1;       // >= 0: stand-alone item; -1: tile pass; -2: multiply the state by `scale`
    cd scale{1.0, 0.0};
    unsigned grid = 0;
    int nrounds = 0, nops = 0;
    bool ext = false; // needs the extended kernel (two-bit SWAPs / tail ladders)
    bool jit_only = false; // holds K_PAIR2 ops: only a specialised kernel can run it
    bool fallback = false;  // the interpreter's encoding of (a piece of) a pass the consumer refused: no kernel to look for
    bool jit_forms = false; // encoded with forms only the specialised kernels have (scaled rotations, in-stream
                            // ladders); a consumer that cannot run it says so and gets the interpreter encoding
    std::vector<int> slots;          // adjoint: global accumulator slot of each pass-local slot
    std::vector<double> slot_scale;  // ... and the |pending scalar|^2 its overlap was taken under
};

// In-place form of a 2x2 block (tile_exec.cuh): kind, packed parameters, the scalar s that is left
// for the host to carry (only when fold is allowed, else 1) and whether an X must precede it (pivot).
struct PairForm {
    int kind = K_LU_C;
    bool pre_swap = false;
    cd s{1.0, 0.0};
    cd m[4];
};
inline bool is_real(cd z) { return z.imag() == 0.0; }
inline bool is_imag(cd z) { return z.real() == 0.0; }

// jf ("jit forms", only together with fold): the pass will run as a specialised kernel, which has the scaled
// rotations K_SRO[TK]_[RI] (4 FMAs per pair; the cosine or sine goes to the host-carried scalar).  max_growth:
// largest 1 / |cos| the tangent form may take (the stored amplitudes grow by it).
PairForm pair_form(const cd *min, bool fold, bool jf = false, double max_growth = 1.0) {
    PairForm f;
    cd m[4] = {min[0], min[1], min[2], min[3]};
    if (m[0] == cd(0.0) && m[3] == cd(0.0) && m[1] == cd(1.0) && m[2] == cd(1.0)) {
        f.kind = K_SWAP;
        return f;
    }
    const bool all_real = is_real(m[0]) && is_real(m[1]) && is_real(m[2]) && is_real(m[3]);
    const bool rx_like = is_real(m[0]) && is_real(m[3]) && is_imag(m[1]) && is_imag(m[2]);
    if (fold && jf) {
        // [[c, -s], [s, c]] = c [[1, -t], [t, 1]] = s [[k, -1], [1, k]];  [[c, -is], [-is, c]] likewise with -i s
        const bool ry = all_real && m[0] == m[3] && m[1] == -m[2];
        const bool rx = rx_like && m[0] == m[3] && m[1] == m[2];
        if (ry || rx) {
            const double c = m[0].real(), sn = ry ? m[2].real() : -m[2].imag();
            if (std::abs(c * c + sn * sn - 1.0) <= 8e-16) {
                if (std::abs(c) * max_growth >= 1.0) f.kind = ry ? K_SROT_R : K_SROT_I, f.m[0] = cd(sn / c, 0.0), f.s = c;
                else f.kind = ry ? K_SROK_R : K_SROK_I, f.m[0] = cd(c / sn, 0.0), f.s = ry ? cd(sn) : cd(0.0, -sn);
                return f;
            }
        }
    }
    // rotations [[c, -s], [s, c]] / [[c, -is], [-is, c]] with c^2 + s^2 = 1: three shears
    //   [[1, u], [0, 1]] [[1, 0], [l, 1]] [[1, u], [0, 1]],  l = s, u = -s / (1 + c)   (c >= 0)
    // (c < 0: rotate by the angle - pi instead and carry the sign, if a global scalar may be carried)
    if (all_real && m[0] == m[3] && m[1] == -m[2]) {
        double c = m[0].real(), sn = m[2].real();
        if (std::abs(c * c + sn * sn - 1.0) <= 8e-16 && (c >= 0.0 || fold)) {
            if (c < 0.0) c = -c, sn = -sn, f.s = -1.0;
            f.kind = K_LIFT_R, f.m[0] = cd(-sn / (1.0 + c), sn);
            return f;
        }
    }
    if (rx_like && m[0] == m[3] && m[1] == m[2]) {
        double c = m[0].real(), sn = -m[2].imag();
        if (std::abs(c * c + sn * sn - 1.0) <= 8e-16 && (c >= 0.0 || fold)) {
            if (c < 0.0) c = -c, sn = -sn, f.s = -1.0;
            f.kind = K_LIFT_I, f.m[0] = cd(-sn / (1.0 + c), -sn);
            return f;
        }
    }
    if (fold && all_real && m[0] == m[1] && m[0] == m[2] && m[0] == -m[3] && m[0] != cd(0.0)) {
        f.kind = K_HAD, f.s = m[0];
        return f;
    }
    // LU with column pivoting: M = diag(al, be) [[1, 0], [t2, 1]] [[1, t1], [0, 1]] (after an X if |m01| > |m00|)
    if (std::abs(m[1]) > std::abs(m[0])) {
        f.pre_swap = true;
        std::swap(m[0], m[1]);
        std::swap(m[2], m[3]);
    }
    const cd al = m[0], t1 = m[1] / m[0], be = (m[0] * m[3] - m[1] * m[2]) / m[0], t2 = m[2] / be;
    if (all_real) f.kind = K_LU_R, f.m[0] = cd(t1.real(), t2.real()), f.m[1] = cd(al.real(), be.real());
    else f.kind = K_LU_C, f.m[0] = t1, f.m[1] = t2, f.m[2] = al, f.m[3] = be;
    return f;
}

// Pure host: schedule `items` on an n-qubit state into tile passes / stand-alone items.
// allow_scaled: uncontrolled rotations / diagonals may leave a scalar factor with the host (sigma),
// which is multiplied back into the state by a K_DIAG_T op before it can over/underflow, at the end
// of every adjoint pass, and at the end of the tape.
//
// Steps are handed to on_step(step, pass description or nullptr) AS THEY ARE PRODUCED, so the caller can
// launch pass k while pass k+1 is being scheduled (the description is reused: consume it in the call).
//
// jit_forms: tile passes are first encoded with the forms of the specialised kernels (pair_form's scaled
// rotations; controlled phases sharing a register bit merged into in-stream ladders).  on_step may return
// bool for a tile pass: false = "cannot run this encoding" (no kernel yet in the asynchronous tier), and the
// pass is encoded again in the interpreter's forms and handed over a second time.  The schedule itself (tiles,
// rounds, order) is the same either way; only the records and the host-carried scalar differ.
template <typename T2, class Cfg, class OnStep>
void build_schedule(int n, int sm_count, const std::vector<AdjItem> &items, bool allow_scaled, int forms, OnStep &&on_step) {
    // forms: 0 = no kernel will ever be compiled from these passes (interpreter only: the scalar slot exists only
    // where the scalar is folded back); 1 = kernels in play, interpreter's encoding; 2 = kernels in play, jit forms
    const bool jit_forms = forms >= 2, stable_slot = forms >= 1;
    constexpr int M = Cfg::M, LOW = Cfg::LOW, R = Cfg::R, NTB = M - R, SB = Swz<T2>::B;
    constexpr bool is_double = sizeof(T2) == 16;
    constexpr size_t kMinLadder = 6; // shorter runs are cheaper as ordinary ops in the lean kernel
    const double sig_lo = is_double ? 0x1p-200 : 0x1p-20, sig_hi = is_double ? 0x1p200 : 0x1p20;
    if (n < M + 1 || items.size() < 2) {
        for (size_t i = 0; i < items.size(); i++) {
            Step st;
            st.op = static_cast<int>(i);
            on_step(st, nullptr);
        }
        return;
    }
    auto cur = std::make_unique<PassParams<T2>>();
    // PLB200_SCHED_GREEDY1=1 / PLB200_SCHED_MULTISTART=1 force the choice (tests fuzz both on small states).
    // The multi-start tile choice costs the host ~2 ms per pass (30q tape: 41 ms against 2.5 ms for plain greedy)
    // and saves ~15 % of the passes: it pays once a pass takes the GPU longer than that, i.e. from 4 GiB states
    // (c128: 28 qubits, 2.3 ms per pass); below, the host would be the bottleneck (26q: 21 ms of scheduling for
    // 10 ms of GPU work).
    const bool multistart = std::getenv("PLB200_SCHED_MULTISTART") != nullptr ||
                            ((sizeof(T2) << n) >= (size_t{1} << 32) && std::getenv("PLB200_SCHED_GREEDY1") == nullptr);
    const uint64_t full = (n >= 64) ? ~uint64_t{0} : ((uint64_t{1} << n) - 1);
    std::vector<FOp> f(items.size());
    const bool pair2 = Cfg::NS == 1 && pair2_enabled(n);
    for (size_t i = 0; i < items.size(); i++) {
        f[i] = classify(items[i], pair2, R);
        f[i].all &= full;
    }
    std::vector<char> done(items.size(), 0);
    size_t first = 0;
    const size_t window = std::getenv("PLB200_SCHED_WINDOW") ? static_cast<size_t>(std::atoi(std::getenv("PLB200_SCHED_WINDOW"))) : 512;
    const uint64_t lowbits = (uint64_t{1} << LOW) - 1;
    const size_t max_pass_ops = kMaxPassOps - kMaxPassRounds - 2 - 24; // emitted ops; room for scalar ops + ladder headers
    std::vector<int> pending, exec;
    cd sigma{1.0, 0.0};

    auto scale_op = [](cd s) {
        TileOp<T2> t;
        std::memset(&t, 0, sizeof(t));
        t.code = make_code(K_DIAG_T, 0, 0);
        t.m[0] = t.m[1] = mk<T2>(s.real(), s.imag());
        return t;
    };

    // records an item may emit (caps the ops of a pass; a pivoted 2x2 block emits two, a dense 4x4 four, ...)
    std::vector<uint8_t> wt(items.size(), 1);
    for (size_t i = 0; i < items.size(); i++) {
        const AdjItem &it = items[i];
        size_t w = 1;
        if (!it.overlap && it.op.kind == OP_DENSE) w = 4; // K_DENSE2: one record per matrix row
        else if (!it.overlap && it.op.kind == OP_PAIRS && it.op.tbits.size() >= 2 && !is_swap2(it.op))
            w = 2 * it.op.blocks.size(); // K_PAIR2 / K_PAIR4: one record per block (+ a pivot each at most)
        else if (jit_forms && !it.overlap && it.op.kind == OP_DIAG && it.op.cmask != 0 && __builtin_popcountll(f[i].pmask) == 1 &&
                 f[i].d[0] != cd(1.0))
            w = 2; // a controlled two-valued diagonal may split into two controlled phases
        else if (!it.overlap && it.op.kind == OP_PAIRS && f[i].fusable && !it.op.blocks.empty()) {
            // a 2x2 block may need a pivot (an X first): two records — except the forms that never do, whatever
            // their angle: X itself, and the uncontrolled rotations / Hadamard whose scalar goes to the host.  The
            // weight must not depend on the angle (it is part of the plan's signature).
            const cd *m = it.op.blocks[0].m;
            const bool is_x = m[0] == cd(0.0) && m[3] == cd(0.0) && m[1] == cd(1.0) && m[2] == cd(1.0);
            const bool all_real = is_real(m[0]) && is_real(m[1]) && is_real(m[2]) && is_real(m[3]);
            const bool rx_like = is_real(m[0]) && is_real(m[3]) && is_imag(m[1]) && is_imag(m[2]);
            const bool ry = all_real && m[0] == m[3] && m[1] == -m[2];
            const bool rx = rx_like && m[0] == m[3] && m[1] == m[2];
            const double c = m[0].real(), sn = ry ? m[2].real() : -m[2].imag();
            const bool rot = (ry || rx) && std::abs(c * c + sn * sn - 1.0) <= 8e-16;
            const bool had = all_real && m[0] == m[1] && m[0] == m[2] && m[0] == -m[3] && m[0] != cd(0.0);
            w = (is_x || (allow_scaled && it.op.cmask == 0 && (rot || had))) ? 1 : 2;
        }
        wt[i] = static_cast<uint8_t>(std::min<size_t>(w, 255));
    }
    // ---- plan cache.  The plan (which items form which pass, its tile bits, its rounds) is a function of the items'
    // scheduling signature only — bit masks, fusability, record weights — not of their angles: a variational loop or
    // an adjoint sweep per optimiser step re-encodes the cached plan with fresh values and skips the search.
    struct PlanStep {
        int op = -1;
        HostPass hp;
        bool tape_ends = false;
    };
    struct Plan {
        std::vector<uint64_t> sig;
        std::vector<PlanStep> steps;
    };
    static std::mutex cache_mutex;
    static std::deque<std::shared_ptr<const Plan>> cache;
    const bool use_cache = !(std::getenv("PLB200_SCHED_CACHE") && std::getenv("PLB200_SCHED_CACHE")[0] == '0');
    auto plan = std::make_shared<Plan>();
    plan->sig = {static_cast<uint64_t>(n), static_cast<uint64_t>(M), static_cast<uint64_t>(R), static_cast<uint64_t>(Cfg::NS),
                 static_cast<uint64_t>(sizeof(T2)), static_cast<uint64_t>(jit_forms), static_cast<uint64_t>(multistart),
                 static_cast<uint64_t>(window), static_cast<uint64_t>(allow_scaled), static_cast<uint64_t>(items.size())};
    for (size_t i = 0; i < items.size(); i++) {
        plan->sig.push_back(f[i].nd);
        plan->sig.push_back(f[i].all);
        plan->sig.push_back(static_cast<uint64_t>(f[i].fusable) | static_cast<uint64_t>(wt[i]) << 1);
    }
    std::shared_ptr<const Plan> hit;
    if (use_cache) {
        std::lock_guard<std::mutex> lock(cache_mutex);
        for (const auto &c : cache)
            if (c->sig == plan->sig) hit = c;
    }
#if defined(PLB200_HOST_EMU)
    if (hit) g_plan_hits++;
#endif
    // Encode one planned pass (tile bits + rounds) at the current carried scalar and hand it to the consumer; when the
    // consumer refuses it, encode it again in the interpreter's forms / cut it at its pair ops.
    auto run_pass = [&](const HostPass &hp, const bool tape_ends) {
        uint64_t T = 0;
        for (int b : hp.tbits) T |= uint64_t{1} << b;
        // ---- encode the plan (jf: in the forms of the specialised kernels)
        auto encode = [&](const bool jf, const HostPass &hq, const bool tape_done) -> Step {
        std::memset(static_cast<void *>(cur.get()), 0, sizeof(PassParams<T2>));
        PassHdr *hdr = &cur->hdr;
        RoundHdr *rh = cur->rounds;
        TileOp<T2> *top = cur->ops;
        hdr->nrounds = static_cast<int>(hq.rounds.size());
        hdr->ntiles = uint64_t{1} << (n - M);
        hdr->tile_ins.n = 0;
        for (int b : hq.tbits) hdr->tile_ins.lowmask[hdr->tile_ins.n++] = (uint64_t{1} << b) - 1;
        int local_of[64];
        for (int i = 0; i < 64; i++) local_of[i] = -1;
        for (int i = 0; i < M; i++) local_of[hq.tbits[i]] = i;
        auto to_local = [&](uint64_t mask) {
            uint32_t l = 0;
            for (int i = 0; i < M; i++)
                if (mask >> hq.tbits[i] & 1) l |= 1u << i;
            return l;
        };
        Step st;
        int op_cursor = 0;
        double pass_growth = 1.0; // growth of the stored amplitudes by this pass's scaled rotations
        for (size_t r = 0; r < hq.rounds.size(); r++) {
            std::vector<int> rl; // local positions of the register bits, ascending
            for (int i = 0; i < M; i++)
                if (hq.round_bits[r] >> hq.tbits[i] & 1) rl.push_back(i);
            uint32_t rmask_l = 0;
            for (int i = 0; i < R; i++) rmask_l |= 1u << rl[i];
            // ---- thread-bit assignment.  Low SB lane bits: tile bits with independent swizzle
            // columns (conflict-free shared accesses), taken first from the bits no op of this round
            // uses as a control / parity bit; the most used ones end up in the warp-index bits, so
            // conditional ops diverge as little as possible.
            int use[32] = {0};
            for (int idx : hq.rounds[r]) {
                const AdjItem &it = items[idx];
                const uint64_t cm = it.overlap ? it.pw.cmask : it.op.cmask;
                const uint64_t pm = it.overlap ? (it.pw.x ? 0 : it.pw.z) : (it.op.kind == OP_PAIRS ? 0 : f[idx].pmask);
                const uint32_t l = to_local((cm | pm) & T) & ~rmask_l;
                for (int i = 0; i < M; i++)
                    if (l >> i & 1) use[i]++;
            }
            std::vector<int> nr;
            for (int i = 0; i < M; i++)
                if (!(rmask_l >> i & 1)) nr.push_back(i);
            std::stable_sort(nr.begin(), nr.end(), [&](int a, int b) { return use[a] < use[b]; });
            auto col_of = [&](int bit) { return bit < SB ? (1u << bit) : swz_col<T2>(bit); };
            std::vector<int> tpos;
            {
                uint32_t basis[4] = {0, 0, 0, 0}; // reduced basis by leading bit
                std::vector<char> taken(M, 0);
                // First and last round of a pass: when no register bit sits among the tile's LOW contiguous
                // bits, those become the lowest lane bits (their swizzle columns are the unit vectors, so the
                // shared accesses stay conflict-free): 2^LOW consecutive lanes then cover one 128-byte line and
                // a specialised kernel (jit_codegen.hpp) moves that round's registers straight from / to
                // global memory.
                static_assert(SB == LOW, "the contiguous tile bits are the bank-group bits");
                const bool edge_round = (r == 0 || r + 1 == hq.rounds.size()) && Cfg::NS == 1;
                if (edge_round && (rmask_l & ((1u << LOW) - 1u)) == 0) {
                    for (int i = 0; i < LOW; i++) {
                        tpos.push_back(i), taken[i] = 1;
                        basis[i] = 1u << i;
                    }
                }
                for (int cand : nr) {
                    if (taken[cand]) continue;
                    if (static_cast<int>(tpos.size()) == SB) break;
                    uint32_t c = col_of(cand);
                    for (int b = SB - 1; b >= 0 && c; b--)
                        if ((c >> b & 1) && basis[b]) c ^= basis[b];
                    if (!c) continue;
                    basis[31 - __builtin_clz(c)] = c;
                    tpos.push_back(cand), taken[cand] = 1;
                }
                for (int cand : nr)
                    if (!taken[cand]) tpos.push_back(cand);
            }
            rh[r].first_op = op_cursor;
            for (int i = 0; i < NTB; i++) rh[r].w[i] = swz<T2>(1u << tpos[i]) * static_cast<uint32_t>(sizeof(T2));
            for (int u = 0; u < (1 << R); u++) {
                uint32_t o = 0;
                for (int i = 0; i < R; i++)
                    if (u >> i & 1) o |= 1u << rl[i];
                rh[r].sroff[u] = swz<T2>(o) * static_cast<uint32_t>(sizeof(T2));
            }
            auto to_tid = [&](uint32_t local_mask) {
                uint32_t m = 0;
                for (int i = 0; i < NTB; i++)
                    if (local_mask >> tpos[i] & 1) m |= 1u << i;
                return m;
            };
            auto reg_of_local = [&](int lp) { return static_cast<int>(std::find(rl.begin(), rl.end(), lp) - rl.begin()); };
            auto reg_pos = [&](int global_bit) { return reg_of_local(local_of[global_bit]); };
            auto roff_of = [&](int u) {
                uint32_t o = 0;
                for (int i = 0; i < R; i++)
                    if (u >> i & 1) o |= 1u << rl[i];
                return o;
            };

            // phase_S / phase: instead of the item itself, emit "multiply by `phase` where every bit of phase_S
            // is 1" (one factor of a diagonal item that was parked as controlled phases)
            auto emit_item = [&](int idx, uint64_t phase_S = 0, cd phase = cd(1.0)) {
                const AdjItem &it = items[idx];
                const FOp &fo = f[idx];
                TileOp<T2> t;
                std::memset(&t, 0, sizeof(t));
                uint64_t cmask = it.overlap ? it.pw.cmask : it.op.cmask;
                uint64_t cval = it.overlap ? it.pw.cval : it.op.cval;
                uint64_t pmask = it.overlap ? (it.pw.x ? 0 : it.pw.z) : (it.op.kind == OP_PAIRS ? 0 : fo.pmask);
                cd d0 = fo.d[0], d1 = fo.d[1];
                if (phase_S) cmask = cval = phase_S, pmask = 0, d0 = d1 = phase;
                const bool is_diag = !it.overlap && it.op.kind == OP_DIAG;
                if (is_diag) {
                    if (pmask == 0) { // one scalar d0 == d1 on the control subspace
                        if (cmask == 0 && allow_scaled) {
                            sigma *= d0; // global phase: nothing to do on the device
                            return;
                        }
                        // turn one value-1 control into the parity bit: diag(1, d) on it
                        const uint64_t ones = cmask & cval;
                        if (ones) {
                            uint64_t pick = 0;
                            for (int i = 0; i < R && !pick; i++)
                                if (ones >> hq.tbits[rl[i]] & 1) pick = uint64_t{1} << hq.tbits[rl[i]];
                            if (!pick && (ones & T)) pick = uint64_t{1} << __builtin_ctzll(ones & T);
                            if (!pick) pick = uint64_t{1} << __builtin_ctzll(ones);
                            pmask = pick, cmask &= ~pick, cval &= ~pick;
                            d1 = d0, d0 = cd(1.0);
                        }
                    }
                    if (d0 != cd(1.0) && cmask == 0 && allow_scaled && pmask != 0) {
                        sigma *= d0;
                        d1 /= d0, d0 = cd(1.0);
                    }
                }
                const uint32_t cm_l = to_local(cmask & T), cv_l = to_local(cval & T), pm_l = to_local(pmask & T);
                const uint32_t cm_reg = cm_l & rmask_l, pm_reg = pm_l & rmask_l;
                t.cm_tid = to_tid(cm_l & ~rmask_l), t.cv_tid = to_tid(cv_l & ~rmask_l);
                t.pm_tid = to_tid(pm_l & ~rmask_l);
                t.cmask_o = cmask & ~T, t.cval_o = cval & ~T, t.pmask_o = pmask & ~T;
                uint32_t umask = 0, upar = 0;
                for (int u = 0; u < (1 << R); u++) {
                    const uint32_t ro = roff_of(u);
                    if ((ro & cm_l) == (cv_l & rmask_l)) umask |= 1u << u;
                    if (__builtin_popcount(ro & pm_l) & 1) upar |= 1u << u;
                }
                t.umask = umask, t.upar = upar;
                const bool one_ctrl_reg = __builtin_popcount(cm_reg) == 1 && (cv_l & cm_reg) == cm_reg;
                const int creg = cm_reg ? reg_of_local(__builtin_ctz(cm_reg)) : 0;
                if (it.overlap) {
                    t.slot = static_cast<uint32_t>(st.slots.size());
                    st.slots.push_back(it.slot);
                    st.slot_scale.push_back(std::norm(sigma));
                    if (it.pw.x) {
                        t.code = make_code((it.pw.z == it.pw.x) ? K_OVL_Y : K_OVL_X, reg_pos(__builtin_ctzll(it.pw.x)), 0);
                    } else {
                        t.code = make_code(K_OVL_D, 0, 0);
                        t.m[0] = mk<T2>(1.0, it.pw.z ? -1.0 : 1.0);
                    }
                } else if (is_swap2(it.op)) {
                    st.ext = true;
                    t.code = make_code(cm_reg ? K_SWAP2_M : K_SWAP2, reg_pos(it.op.tbits[0]), reg_pos(it.op.tbits[1]));
                } else if (it.op.kind == OP_DENSE) {
                    // dense 4x4 on two register bits: header record (row 0) + three continuation records (rows 1-3)
                    st.jit_only = true;
                    t.code = make_code(K_DENSE2, reg_pos(it.op.tbits[0]), reg_pos(it.op.tbits[1]));
                    if (t.cm_tid | t.cmask_o) t.code |= F_COND;
                    for (int row = 0; row < 4; row++) {
                        TileOp<T2> rec;
                        if (row == 0) rec = t;
                        else std::memset(&rec, 0, sizeof(rec));
                        for (int q = 0; q < 4; q++) rec.m[q] = mk<T2>(it.op.mat[row * 4 + q].real(), it.op.mat[row * 4 + q].imag());
                        top[op_cursor++] = rec;
                    }
                    return;
                } else if (it.op.kind == OP_PAIRS && (it.op.tbits.size() == 2 || it.op.tbits.size() == 4)) {
                    // two- / four-bit pair op: one K_PAIR2 / K_PAIR4 record per 2x2 block (a pivot X on the pair first
                    // when needed)
                    st.jit_only = true;
                    const bool four = it.op.tbits.size() == 4;
                    const int p = reg_pos(it.op.tbits[0]), c = reg_pos(it.op.tbits[1]);
                    uint32_t pos4 = 0;
                    if (four)
                        for (int j = 0; j < 4; j++) pos4 |= static_cast<uint32_t>(reg_pos(it.op.tbits[j])) << (12 + 3 * j);
                    t.code = make_code(four ? K_PAIR4 : K_PAIR2, four ? 0 : p, four ? 0 : c);
                    if (t.cm_tid | t.cmask_o) t.code |= F_COND;
                    auto pack = [&](const Block2 &bl, int form) {
                        return four ? (bl.a | (bl.b << 4) | (static_cast<uint32_t>(form) << 8) | pos4)
                                    : (bl.a | (bl.b << 2) | (static_cast<uint32_t>(form) << 4));
                    };
                    for (const Block2 &bl : it.op.blocks) {
                        const PairForm pf = pair_form(bl.m, false);
                        TileOp<T2> rec = t;
                        if (pf.pre_swap) {
                            TileOp<T2> x = t;
                            x.slot = pack(bl, K_SWAP);
                            top[op_cursor++] = x;
                        }
                        rec.slot = pack(bl, pf.kind);
                        for (int q = 0; q < 4; q++) rec.m[q] = mk<T2>(pf.m[q].real(), pf.m[q].imag());
#if defined(PLB200_HOST_EMU)
                        g_kind_hist[four ? K_PAIR4 : K_PAIR2]++;
#endif
                        top[op_cursor++] = rec;
                    }
                    return;
                } else if (it.op.kind == OP_PAIRS) {
                    const int p = reg_pos(it.op.tbits[0]);
                    // tangent-form rotations let the stored amplitudes grow by 1 / |cos|: at most 2^16 per
                    // rotation and 2^40 (c64) / 2^400 (c128) per round on top of the carried scalar's own range
                    // tangent-form rotations let the stored amplitudes grow by 1 / |cos|: at most 2^16 per
                    // rotation and 2^60 (c64) / 2^400 (c128) per pass on top of the carried scalar's own range;
                    // past 2^20 more (cotangent forms grow by < sqrt 2 each) the three shears take over
                    PairForm pf = pair_form(it.op.blocks[0].m, allow_scaled && cmask == 0, jf,
                                            std::min(0x1p16, (is_double ? 0x1p400 : 0x1p60) / pass_growth));
                    if (pf.kind >= K_SROT_R && pf.kind <= K_SROK_I) {
                        if (pass_growth / std::abs(pf.s) > (is_double ? 0x1p420 : 0x1p80))
                            pf = pair_form(it.op.blocks[0].m, allow_scaled && cmask == 0, false);
                        else
                            st.jit_forms = true, pass_growth /= std::abs(pf.s);
                    }
                    const bool masked = cm_reg != 0;
                    auto masked_kind = [](int k) {
                        return k == K_LIFT_R ? K_LIFT_R_M : k == K_LIFT_I ? K_LIFT_I_M : k == K_LU_R ? K_LU_R_M
                               : k == K_LU_C ? K_LU_C_M : K_SWAP_M;
                    };
                    const int swap_kind = !masked ? K_SWAP : one_ctrl_reg ? K_SWAP_CR : K_SWAP_M;
                    if (pf.pre_swap) { // pivot: X first, same controls
                        TileOp<T2> x = t;
                        x.code = make_code(swap_kind, p, swap_kind == K_SWAP_CR ? creg : 0);
                        if (x.cm_tid | x.cmask_o) x.code |= F_COND;
                        if (swap_kind == K_SWAP_M) st.ext = true;
                        top[op_cursor++] = x;
                    }
                    int kind = pf.kind, c = 0;
                    if (kind == K_SWAP) kind = swap_kind, c = (swap_kind == K_SWAP_CR ? creg : 0);
                    else if (masked) kind = masked_kind(kind);
                    sigma *= pf.s;
                    t.code = make_code(kind, p, c);
                    for (int q = 0; q < 4; q++) t.m[q] = mk<T2>(pf.m[q].real(), pf.m[q].imag());
                } else {
                    const bool one = (d0 == cd(1.0));
                    const int npr = __builtin_popcount(pm_reg);
                    const int p0 = npr >= 1 ? reg_of_local(__builtin_ctz(pm_reg)) : 0;
                    const int p1 = npr >= 2 ? reg_of_local(31 - __builtin_clz(pm_reg)) : 0;
                    if (cm_reg == 0 && npr == 0) t.code = make_code(one ? K_DIAG1_T : K_DIAG_T, 0, 0);
                    else if (cm_reg == 0 && npr == 1)
                        t.code = make_code((one && t.pm_tid == 0 && t.pmask_o == 0) ? K_DIAG1_R : K_DIAG_R, p0, 0);
                    else if (cm_reg == 0 && npr == 2) t.code = make_code(K_DIAG_PP, p0, p1);
                    else if (one_ctrl_reg && npr == 0) t.code = make_code(K_DIAG_CT, creg, 0);
                    else if (one_ctrl_reg && npr == 1) t.code = make_code(K_DIAG_CR, p0, creg);
                    else t.code = make_code(K_DIAG_G, 0, 0);
                    t.m[0] = mk<T2>(d0.real(), d0.imag());
                    t.m[1] = mk<T2>(d1.real(), d1.imag());
                }
                if (t.cm_tid | t.cmask_o) t.code |= F_COND;
                if (t.pm_tid | t.pmask_o) t.code |= F_PAR;
                {
                    const int kd = code_kind(t.code);
                    if (kd == K_LU_R || kd == K_LU_C || kd == K_LIFT_R_M || kd == K_LIFT_I_M || kd == K_LU_R_M ||
                        kd == K_LU_C_M || kd == K_SWAP_M || kd == K_DIAG_PP || kd == K_DIAG_G || kd == K_SWAP2 ||
                        kd == K_SWAP2_M)
                        st.ext = true; // kinds only the extended kernel implements
                    if (kd == K_DIAG1_R || kd == K_DIAG1_T) t.m[0] = t.m[1]; // the one phase, in the prefetched slot
                }
#if defined(PLB200_HOST_EMU)
                g_kind_hist[code_kind(t.code)]++;
#endif
                top[op_cursor++] = t;
            };

            // ---- emission.  Controlled phases "multiply by e^{i phi} where every bit of S is 1" whose S
            // holds at most one register bit P (the rest on thread / outside bits) are not emitted in
            // place: they are parked in a bucket per P (P = R: no register bit).  Such an op commutes
            // with everything else in the round except a later non-diagonal op on its register bit P;
            // when one arrives the bucket is flushed in place as ordinary ops.  What is still parked at
            // the end of the round becomes the round's TAIL: buckets of >= 2 entries as one ladder each
            // (header + entries: one scalar per thread, one multiply), single ones as ordinary ops.
            struct LEntry {
                int idx;
                uint64_t S;
                cd phase, fold;
            };
            // factors of a diagonal item as controlled phases (0: the item is emitted as an ordinary op)
            auto ladder_entries = [&](int idx, LEntry (&e)[2]) {
                const AdjItem &it = items[idx];
                const FOp &fo = f[idx];
                if (it.overlap || it.op.kind != OP_DIAG) return 0;
                int ne = 0;
                if (fo.pmask == 0) {
                    if (it.op.cmask == 0 || it.op.cval != it.op.cmask) return 0;
                    e[ne++] = LEntry{idx, it.op.cmask, fo.d[0], cd(1.0)};
                } else if (__builtin_popcountll(fo.pmask) == 1 && it.op.cmask == 0) {
                    if (!allow_scaled) return 0;
                    e[ne++] = LEntry{idx, fo.pmask, fo.d[1] / fo.d[0], fo.d[0]};
                } else if (jf && __builtin_popcountll(fo.pmask) == 1 && it.op.cval == it.op.cmask) {
                    // controlled diag(d0, d1) (CRZ ...): d0 on the control subspace, d1 / d0 where the target is set too
                    if (fo.d[0] != cd(1.0)) e[ne++] = LEntry{idx, it.op.cmask, fo.d[0], cd(1.0)};
                    e[ne++] = LEntry{idx, it.op.cmask | fo.pmask, fo.d[1] / fo.d[0], cd(1.0)};
                } else
                    return 0;
                for (int q = 0; q < ne; q++)
                    if (__builtin_popcount(to_local(e[q].S & T) & rmask_l) > 1) return 0;
                return ne;
            };
            std::vector<LEntry> bucket[kMaxR + 1];
            // header + packed entries of one ladder on register bit bk (bk = R: none) at the cursor
            auto emit_ladder = [&](int bk) {
                TileOp<T2> hd;
                std::memset(&hd, 0, sizeof(hd));
                hd.code = make_code(K_LADDER, bk, 0);
                hd.slot = static_cast<uint32_t>(bucket[bk].size());
#if defined(PLB200_HOST_EMU)
                g_kind_hist[K_LADDER]++;
#endif
                const int nrec = ladder_records<T2>(static_cast<int>(bucket[bk].size()));
                if (op_cursor + nrec > kMaxPassOps) fail("fusion: pass description overflow");
                top[op_cursor] = hd;
                std::memset(static_cast<void *>(top + op_cursor + 1), 0, sizeof(TileOp<T2>) * static_cast<size_t>(nrec - 1));
                LadderEntry<T2> *ens = reinterpret_cast<LadderEntry<T2> *>(top + op_cursor + 1);
                for (const LEntry &e : bucket[bk]) {
                    ens->cmask_o = e.S & ~T;
                    ens->cm_tid = to_tid(to_local(e.S & T) & ~rmask_l);
                    ens->ph = mk<T2>(e.phase.real(), e.phase.imag());
                    ens++;
                    sigma *= e.fold;
                }
                op_cursor += nrec;
                bucket[bk].clear();
            };
            // parked phases of bucket bk as ops at the cursor: the interpreter's forms emit the items themselves;
            // the specialised kernels take two or more as ONE in-stream ladder (one scalar per thread, one multiply)
            auto flush_bucket = [&](int bk) {
                if (bucket[bk].empty()) return;
                if (jf && bucket[bk].size() >= 2) {
                    st.jit_forms = true;
                    emit_ladder(bk);
                    return;
                }
                for (const LEntry &e : bucket[bk]) {
                    if (jf) {
                        sigma *= e.fold;
                        if (e.S != 0) emit_item(e.idx, e.S, e.phase);
                    } else
                        emit_item(e.idx);
                }
                bucket[bk].clear();
            };
            for (int idx : hq.rounds[r]) {
                const AdjItem &it = items[idx];
                LEntry e[2];
                if (const int ne = ladder_entries(idx, e)) {
                    for (int q = 0; q < ne; q++) {
                        const uint32_t sreg = to_local(e[q].S & T) & rmask_l;
                        bucket[sreg ? reg_of_local(__builtin_ctz(sreg)) : R].push_back(e[q]);
                    }
                    continue;
                }
                // non-diagonal action on register bits: parked phases on those bits must come first
                uint64_t ndbits = 0;
                if (it.overlap) ndbits = it.pw.x;
                else if (it.op.kind == OP_PAIRS || it.op.kind == OP_DENSE)
                    for (int b : it.op.tbits) ndbits |= uint64_t{1} << b;
                for (int i = 0; i < R; i++)
                    if (ndbits >> hq.tbits[rl[i]] & 1) flush_bucket(i);
                emit_item(idx);
            }
            for (int bk = 0; bk <= R; bk++)
                if (jf || bucket[bk].size() < kMinLadder) flush_bucket(bk);
            // The carried scalar has ONE slot per pass, at the end of its last round, whether it is folded back
            // there (adjoint passes, the end of the tape, |sigma| about to leave its range) or not (the slot then
            // holds 1 and the specialised kernels skip it): the structure of a pass — the key of its compiled
            // kernel — must not depend on the angles of the passes before it.
            if (r + 1 == hq.rounds.size()) {
                const double mag = std::abs(sigma);
                const bool fold = sigma != cd(1.0) && (Cfg::NS == 2 || tape_done || mag < sig_lo || mag > sig_hi);
                if (fold || stable_slot) top[op_cursor++] = scale_op(fold ? sigma : cd(1.0));
                if (fold) sigma = cd(1.0);
            }
            rh[r].nops = op_cursor - rh[r].first_op;
            for (int bk = 0; bk <= R; bk++) {
                if (bucket[bk].empty()) continue;
                st.ext = true;
                emit_ladder(bk);
            }
            rh[r].nlad = op_cursor - rh[r].first_op - rh[r].nops;
            if (op_cursor > kMaxPassOps) fail("fusion: pass description overflow");
        }
        hdr->nops_total = op_cursor;
        hdr->nslots = static_cast<int>(st.slots.size());
        st.grid = static_cast<unsigned>(std::min<uint64_t>(hdr->ntiles, uint64_t(sm_count) * 3 * 64));
        st.nrounds = hdr->nrounds, st.nops = 0;
        for (const auto &rr : hq.rounds) st.nops += static_cast<int>(rr.size());
        return st;
        }; // encode
        auto deliver = [&](const Step &st) {
            if constexpr (std::is_same_v<decltype(on_step(st, cur.get())), bool>) return on_step(st, cur.get());
            else {
                on_step(st, cur.get());
                return true;
            }
        };
        const cd sigma_in = sigma;
        Step st = encode(jit_forms, hp, tape_ends);
        if (!deliver(st)) {
            if (!st.jit_forms && !st.jit_only) fail("fusion: a pass in the interpreter's forms was refused");
            sigma = sigma_in;
            if (!st.jit_only) {
                st = encode(false, hp, tape_ends);
                st.fallback = true;
                if (!deliver(st)) fail("fusion: a pass in the interpreter's forms was refused");
            } else {
                // The pass holds two- / four-bit pair ops, which only a specialised kernel can run inside a tile, and
                // that kernel does not exist yet: the pass is cut at those ops — what lies between them runs as
                // interpreter passes over the same tile (same rounds, same order), the pair ops stand-alone.
                HostPass piece;
                piece.tbits = hp.tbits;
                auto flush_piece = [&](bool last) {
                    if (piece.rounds.empty()) return;
                    Step sp = encode(false, piece, last && tape_ends);
                    sp.fallback = true;
                    if (sp.jit_only || !deliver(sp)) fail("fusion: a pass in the interpreter's forms was refused");
                    piece.rounds.clear(), piece.round_bits.clear();
                };
                auto needs_kernel = [&](int idx) {
                    const AdjItem &it = items[idx];
                    return !it.overlap && ((it.op.kind == OP_PAIRS && it.op.tbits.size() >= 2 && !is_swap2(it.op)) || it.op.kind == OP_DENSE);
                };
                for (size_t r = 0; r < hp.rounds.size(); r++) {
                    std::vector<int> run;
                    for (int idx : hp.rounds[r]) {
                        if (!needs_kernel(idx)) {
                            run.push_back(idx);
                            continue;
                        }
                        if (!run.empty()) piece.rounds.push_back(run), piece.round_bits.push_back(hp.round_bits[r]), run.clear();
                        flush_piece(false);
                        Step so;
                        so.op = idx;
                        on_step(so, nullptr);
                    }
                    if (!run.empty()) piece.rounds.push_back(run), piece.round_bits.push_back(hp.round_bits[r]);
                }
                flush_piece(true); // (a scalar left when the pass ends with a pair op is folded by a later slot / the final sweep)
            }
        }
    };

    if (hit) {
        for (const PlanStep &ps : hit->steps) {
            if (ps.op >= 0) {
                Step st;
                st.op = ps.op;
                on_step(st, nullptr);
            } else
                run_pass(ps.hp, ps.tape_ends);
        }
    } else {
    while (true) {
        while (first < items.size() && done[first]) first++;
        if (first >= items.size()) break;
        pending.clear();
        for (size_t i = first; i < items.size() && pending.size() < window; i++)
            if (!done[i]) pending.push_back(static_cast<int>(i));
        // ---- choose the tile bits
        uint64_t T = grow(f, pending, lowbits, M, full, full, exec);
        if (multistart) {
            // greedy growth restarted from every bit some pending op needs: keeps the best tile.
            // (30q benchmark tape: 24 -> 20 passes for ~2 ms more host time per pass, hidden behind
            // the previous pass running on the GPU; not worth it for states a pass sweeps in microseconds)
            uint64_t cand = 0;
            for (int i : pending) cand |= f[i].nd;
            cand &= full & ~lowbits;
            std::vector<int> ex2;
            for (int b = 0; b < n; b++) {
                if (!(cand >> b & 1)) continue;
                const uint64_t Tc = grow(f, pending, lowbits | (uint64_t{1} << b), M, full, full, ex2);
                if (ex2.size() > exec.size()) exec = ex2, T = Tc;
            }
        }
        if (exec.size() < 2) {
            // nothing worth a tile pass: run the first pending item on its own
            Step st;
            st.op = static_cast<int>(first);
            plan->steps.push_back(PlanStep{st.op, HostPass{}, false});
            on_step(st, nullptr);
            done[first] = 1;
            continue;
        }
        // pad T to exactly M bits with the lowest free bits
        for (int b = 0; b < n && __builtin_popcountll(T) < M; b++) T |= uint64_t{1} << b;
        simulate(f, pending, T, full, exec);
        {
            // cap the EMITTED ops (a pivoted 2x2 block emits two); a prefix stays valid
            size_t emitted = 0, keep = 0;
            for (int i : exec) {
                emitted += wt[i];
                if (emitted > max_pass_ops) break;
                keep++;
            }
            exec.resize(keep);
        }
        std::vector<int> pass_ops = exec;

        // ---- rounds
        HostPass hp;
        for (int b = 0; b < n; b++)
            if (T >> b & 1) hp.tbits.push_back(b);
        std::vector<int> rem = pass_ops, rexec;
        while (!rem.empty()) {
            uint64_t seed = f[rem[0]].nd; // guarantees progress
            uint64_t Rb = grow(f, rem, seed, R, T, full, rexec);
            // pad to R bits with tile bits (prefer high local bits)
            for (int i = M - 1; i >= 0 && __builtin_popcountll(Rb) < R; i--) Rb |= uint64_t{1} << hp.tbits[i];
            simulate(f, rem, Rb, full, rexec);
            if (rexec.empty()) fail("fusion scheduler made no progress");
            hp.rounds.push_back(rexec);
            hp.round_bits.push_back(Rb);
            std::vector<int> next;
            size_t k = 0;
            for (int i : rem) {
                if (k < rexec.size() && rexec[k] == i) k++;
                else next.push_back(i);
            }
            rem.swap(next);
            if (hp.rounds.size() == static_cast<size_t>(kMaxPassRounds)) break; // rest waits for the next pass
        }
        pass_ops.clear();
        for (const auto &r : hp.rounds) pass_ops.insert(pass_ops.end(), r.begin(), r.end());
        for (int i : pass_ops) done[i] = 1;
        // the last step of the tape?  (then the carried scalar is multiplied back here)
        bool tape_ends = true;
        for (size_t i = first; i < items.size() && tape_ends; i++) tape_ends = done[i] != 0;
        plan->steps.push_back(PlanStep{-1, hp, tape_ends});
        run_pass(hp, tape_ends);
    }
        if (use_cache) {
            std::lock_guard<std::mutex> lock(cache_mutex);
            cache.push_back(plan);
            if (cache.size() > 32) cache.pop_front();
        }
    }
    if (sigma != cd(1.0)) {
        // stand-alone ops followed the last pass: the carried scalar needs a sweep of its own
        Step st;
        st.op = -2, st.scale = sigma;
        on_step(st, nullptr);
    }
}

template <class Cfg, typename T2> size_t smem_bytes_for() {
    return Cfg::NS * (sizeof(T2) << Cfg::M) + (sizeof(uint64_t) << (Cfg::M - Cfg::LOW)) +
           (Cfg::NS == 2 ? sizeof(double) * kMaxPassOps * ((1u << (Cfg::M - Cfg::R)) / 32) : 0); // one accumulator row per warp
}

// tiles: the state is large enough for tile passes (multi-op expansions pay only then)
std::vector<AdjItem> as_items(const std::vector<COp> &ops, bool tiles = true) {
    std::vector<AdjItem> items;
    items.reserve(ops.size());
    std::vector<COp> pieces;
    for (const COp &o : ops) {
        pieces.clear();
        expand_for_fusion(o, pieces, tiles);
        for (COp &q : pieces) {
            AdjItem it;
            it.op = std::move(q);
            items.push_back(std::move(it));
        }
    }
    return items;
}

// A pass without ops over the lowest M index bits (one round, register bits = the highest tile bits).
template <typename T2, class Cfg> void identity_pass(int n, int sm_count, PassParams<T2> &pp, Step &st) {
    constexpr int M = Cfg::M, R = Cfg::R, NTB = M - R;
    std::memset(static_cast<void *>(&pp), 0, sizeof(PassParams<T2>));
    pp.hdr.nrounds = 1;
    pp.hdr.ntiles = uint64_t{1} << (n - M);
    pp.hdr.tile_ins.n = M;
    for (int b = 0; b < M; b++) pp.hdr.tile_ins.lowmask[b] = (uint64_t{1} << b) - 1;
    RoundHdr &rh = pp.rounds[0];
    for (int i = 0; i < NTB; i++) rh.w[i] = swz<T2>(1u << i) * static_cast<uint32_t>(sizeof(T2));
    for (int u = 0; u < (1 << R); u++) rh.sroff[u] = swz<T2>(static_cast<uint32_t>(u) << NTB) * static_cast<uint32_t>(sizeof(T2));
    st = Step{};
    st.grid = static_cast<unsigned>(std::min<uint64_t>(pp.hdr.ntiles, uint64_t(sm_count) * 3 * 64));
    st.nrounds = 1;
}

// Route description of the caller (engine.cu / the sharded driver): swap k local index bits `lbits` against k
// global (rank) bits while the LAST pass of the tape stores its tile.
template <typename T2> jit::Route tile_route(const RouteSpec &rs, const PassParams<T2> &pp, jit::RouteParams<T2> &rp) {
    jit::Route r;
    r.k = rs.k;
    std::memset(&rp, 0, sizeof(rp));
    for (int i = 0; i < rs.k; i++) {
        r.loc[i] = -1;
        for (int b = 0; b < pp.hdr.tile_ins.n; b++)
            if (pp.hdr.tile_ins.lowmask[b] + 1 == (uint64_t{1} << rs.lbits[i])) r.loc[i] = b;
        rp.lmask |= uint64_t{1} << rs.lbits[i];
        if ((rs.my_value >> i) & 1) rp.rdep |= uint64_t{1} << rs.lbits[i];
        rp.opos[i] = static_cast<uint32_t>(rs.lbits[i]);
    }
    for (int p = 0; p < (1 << rs.k); p++) rp.dst[p] = static_cast<T2 *>(rs.dst[p]);
    // tile-index positions of the swapped bits outside the tile (index bit minus the tile bits below it), ascending
    std::vector<uint32_t> tq;
    for (int i = 0; i < rs.k; i++) {
        if (r.loc[i] >= 0) continue;
        uint32_t below = 0;
        for (int b = 0; b < pp.hdr.tile_ins.n; b++)
            if (pp.hdr.tile_ins.lowmask[b] + 1 < (uint64_t{1} << rs.lbits[i])) below++;
        tq.push_back(static_cast<uint32_t>(rs.lbits[i]) - below);
    }
    std::sort(tq.begin(), tq.end());
    rp.ntq = static_cast<uint32_t>(tq.size());
    for (size_t i = 0; i < tq.size(); i++) rp.tq[i] = tq[i];
    return r;
}

#if !defined(PLB200_HOST_EMU)
bool scaled_forms_enabled() {
    const char *e = std::getenv("PLB200_FUSE_SCALED");
    return !(e && e[0] == '0');
}
// encode passes in the forms of the specialised kernels first (PLB200_JIT_FORMS=0: the interpreter's forms only)
bool jit_forms_enabled() {
    const char *e = std::getenv("PLB200_JIT_FORMS");
    return !(e && e[0] == '0');
}

// Opt the kernels into their dynamic shared memory size, once per device (the attribute is per device:
// DevicePool / batched adjoint runs drive several GPUs from one process).
template <typename T2, class Cfg> void prepare_kernel(int device) {
    static bool done[64] = {false};
    if (device >= 0 && device < 64 && done[device]) return;
    PLB_CUDA(cudaFuncSetAttribute(tile_kernel<T2, Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes_for<Cfg, T2>())));
    PLB_CUDA(cudaFuncSetAttribute(tile_kernel<T2, Cfg, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PLB_CUDA(cudaFuncSetAttribute(tile_kernel<T2, Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes_for<Cfg, T2>())));
    PLB_CUDA(cudaFuncSetAttribute(tile_kernel<T2, Cfg, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (device >= 0 && device < 64) done[device] = true;
}
template <typename T2, class Cfg>
void launch_pass(const Step &st, cudaStream_t stream, T2 *sv0, T2 *sv1, double *acc, const PassParams<T2> &pp) {
    const unsigned nt = 1u << (Cfg::M - Cfg::R);
    if (st.ext) tile_kernel<T2, Cfg, true><<<st.grid, nt, smem_bytes_for<Cfg, T2>(), stream>>>(sv0, sv1, acc, pp);
    else tile_kernel<T2, Cfg, false><<<st.grid, nt, smem_bytes_for<Cfg, T2>(), stream>>>(sv0, sv1, acc, pp);
    PLB_CUDA(cudaGetLastError());
}

// Forward tape: every step is launched the moment the scheduler has produced it (the pass description
// is a by-value kernel parameter), so the host schedules pass k+1 while the GPU runs pass k.
template <typename T2> void run_fused_typed(StateVec &sv, const std::vector<COp> &ops) {
    using Cfg = FwdCfg<T2>;
    const auto items = as_items(ops, sv.n >= Cfg::M + 1);
    prepare_kernel<T2, Cfg>(sv.device);
    // PLB200_FUSE_TRACE=1: per-step device time on stderr (profiling aid; serialises the steps)
    const bool trace = std::getenv("PLB200_FUSE_TRACE") != nullptr;
    const bool use_jit = jit::mode() != jit::Mode::Off && sv.n >= jit::min_qubits();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (trace) {
        PLB_CUDA(cudaEventCreate(&ev0));
        PLB_CUDA(cudaEventCreate(&ev1));
    }
    build_schedule<T2, Cfg>(static_cast<int>(sv.n), sv.sm_count, items, scaled_forms_enabled(), use_jit ? (jit_forms_enabled() ? 2 : 1) : 0,
                            [&](const Step &st, const PassParams<T2> *pp) -> bool {
        if (trace) PLB_CUDA(cudaEventRecord(ev0, sv.stream));
        if (st.op >= 0) launch_op(sv, items[st.op].op);
        else if (st.op == -2) scale(sv, st.scale);
        else {
            // the pass's specialised kernel when the cache has it (jit_runtime.cpp), else the interpreter
            jit::Kernel k;
            if ((use_jit && !st.fallback) || st.jit_only)
                k = jit::lookup(jit::generate_pass_source<T2, Cfg>(*pp), sv.device, smem_bytes_for<Cfg, T2>(), false);
            if (k) jit::launch(k, st.grid, 1u << (Cfg::M - Cfg::R), smem_bytes_for<Cfg, T2>(), sv.stream, sv.data, pp);
            else if (st.jit_forms || st.jit_only) return false; // no kernel (yet): the interpreter's encoding, please
            else launch_pass<T2, Cfg>(st, sv.stream, static_cast<T2 *>(sv.data), nullptr, nullptr, *pp);
            sv.launches++;
        }
        if (trace) {
            PLB_CUDA(cudaEventRecord(ev1, sv.stream));
            PLB_CUDA(cudaEventSynchronize(ev1));
            float ms = 0;
            PLB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
            if (st.op != -1) std::fprintf(stderr, "[plb200 trace] stand-alone op %d: %.3f ms\n", st.op, ms);
            else {
                std::fprintf(stderr, "[plb200 trace] tile pass (%s): %d items, %d records, %d rounds, tile bits",
                             st.ext ? "ext" : "lean", st.nops, pp->hdr.nops_total, pp->hdr.nrounds);
                for (int i = 0; i < pp->hdr.tile_ins.n; i++)
                    std::fprintf(stderr, " %d", 63 - __builtin_clzll(pp->hdr.tile_ins.lowmask[i] + 1));
                std::fprintf(stderr, ": %.3f ms\n", ms);
            }
        }
        return true;
    });
    if (trace) {
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
}

// Forward tape whose last pass stores through a route.  Returns true when the final state went out through the
// route (it then lives in the destination slabs), false when it stayed in place (no tile pass at the end of the
// schedule, last round not line-coalesced, NVRTC missing): the caller then swaps with the stand-alone kernel.
template <typename T2> bool run_fused_routed_typed(StateVec &sv, const std::vector<COp> &ops, const RouteSpec &rs) {
    using Cfg = FwdCfg<T2>;
    const auto items = as_items(ops, sv.n >= Cfg::M + 1);
    prepare_kernel<T2, Cfg>(sv.device);
    const bool use_jit = jit::mode() != jit::Mode::Off && sv.n >= jit::min_qubits();
    const unsigned nt = 1u << (Cfg::M - Cfg::R);
    const size_t smem = smem_bytes_for<Cfg, T2>();
    auto held = std::make_unique<PassParams<T2>>();
    Step held_st;
    bool have = false;
    auto launch_plain = [&](const Step &st, const PassParams<T2> &pp) {
        jit::Kernel k;
        if ((use_jit && !st.fallback) || st.jit_only)
            k = jit::lookup(jit::generate_pass_source<T2, Cfg>(pp), sv.device, smem, st.jit_only || st.jit_forms);
        if (k) jit::launch(k, st.grid, nt, smem, sv.stream, sv.data, &pp);
        else if (st.jit_only) fail("a pass with two-bit pair ops needs its specialised kernel (NVRTC compile failed)");
        else launch_pass<T2, Cfg>(st, sv.stream, static_cast<T2 *>(sv.data), nullptr, nullptr, pp);
        sv.launches++;
    };
    build_schedule<T2, Cfg>(static_cast<int>(sv.n), sv.sm_count, items, scaled_forms_enabled(), use_jit && jit_forms_enabled() ? 2 : 1,
                            [&](const Step &st, const PassParams<T2> *pp) -> bool {
        if (have) launch_plain(held_st, *held), have = false; // the held pass was not the last step
        if (st.op >= 0) launch_op(sv, items[st.op].op);
        else if (st.op == -2) scale(sv, st.scale);
        else {
            // A pass in the specialised forms is held only when its plain kernel exists already (it may turn
            // out not to be the last one); otherwise it comes back in the interpreter's forms.
            if ((st.jit_forms || st.jit_only) && jit::mode() != jit::Mode::Sync &&
                !jit::lookup(jit::generate_pass_source<T2, Cfg>(*pp), sv.device, smem, false))
                return false;
            std::memcpy(static_cast<void *>(held.get()), pp, sizeof(PassParams<T2>));
            held_st = st, have = true;
        }
        return true;
    });
    if (!have) {
        // the schedule did not end with a tile pass (stand-alone kernels at the end, or a tiny batch): an op-less
        // pass carries the state through the route, so that EVERY rank always routes (the ranks' batches
        // differ — controls on global bits — and they must not disagree on how the exchange happens)
        identity_pass<T2, Cfg>(static_cast<int>(sv.n), sv.sm_count, *held, held_st);
    }
    jit::RouteParams<T2> rp;
    const jit::Route route = tile_route<T2>(rs, *held, rp);
    jit::Kernel k = jit::lookup(jit::generate_pass_source<T2, Cfg>(*held, route), sv.device, smem, /*force_sync=*/true);
    if (!k) fail("routed pass: the specialised kernel could not be built (NVRTC)");
    jit::launch(k, held_st.grid, nt, smem, sv.stream, sv.data, held.get(), &rp);
    sv.launches++;
    return true;
}

template <typename T2>
void run_adjoint_typed(StateVec &lambda, StateVec &hl, const std::vector<AdjItem> &items, int n_slots,
                       double *acc_host, int64_t stats[3], StateVec &scratch) {
    using Cfg = AdjCfg<T2>;
    for (int i = 0; i < n_slots; i++) acc_host[i] = 0.0;
    prepare_kernel<T2, Cfg>(lambda.device);
    const bool use_jit = jit::mode() != jit::Mode::Off && lambda.n >= jit::min_qubits();
    // device accumulators: one slab of kMaxPassOps doubles per tile pass; a pass executes >= 2 items
    const size_t max_pass = items.size() / 2 + 1;
    const size_t acc_bytes = max_pass * kMaxPassOps * sizeof(double);
    // + the per-CTA partials of the pass in flight (reduced in CTA order right after it: deterministic)
    // (the buffers live with `scratch` — the caller's persistent state — so that a Jacobian per optimiser step
    // does not pay a cudaMalloc / cudaFree of tens of megabytes each time)
    const size_t max_grid = static_cast<size_t>(std::min<uint64_t>(uint64_t{1} << (lambda.n > Cfg::M ? lambda.n - Cfg::M : 0),
                                                                   uint64_t(lambda.sm_count) * 3 * 64));
    const size_t part_bytes = max_grid * static_cast<size_t>(std::min<int>(kMaxPassOps, std::max(1, n_slots))) * sizeof(double);
    double *dacc = static_cast<double *>(scratch.plan_buf(acc_bytes + part_bytes));
    double *dpart = dacc + max_pass * kMaxPassOps;
    PLB_CUDA(cudaMemsetAsync(dacc, 0, acc_bytes, lambda.stream));
    struct PassSlots {
        std::vector<int> slots;
        std::vector<double> scale;
    };
    std::vector<PassSlots> passes;
    build_schedule<T2, Cfg>(static_cast<int>(lambda.n), lambda.sm_count, items, scaled_forms_enabled(), use_jit ? (jit_forms_enabled() ? 2 : 1) : 0,
                            [&](const Step &st, const PassParams<T2> *pp) -> bool {
        if (st.op >= 0) {
            const AdjItem &it = items[st.op];
            if (it.overlap) {
                double r[2];
                pauli_inner(hl, lambda, &it.pw, 1, r);
                acc_host[it.slot] += r[1];
            } else {
                launch_op(lambda, it.op);
                launch_op(hl, it.op);
            }
            stats[1]++;
            return true;
        }
        if (st.op == -2) fail("fusion: adjoint passes carry no scalar across passes");
        double *pacc = dacc + passes.size() * kMaxPassOps;
        jit::Kernel k;
        if (use_jit && !st.fallback) k = jit::lookup(jit::generate_pass_source<T2, Cfg>(*pp), lambda.device, smem_bytes_for<Cfg, T2>());
        if (!k && st.jit_forms) return false;
        PLB_CHECK(st.grid <= max_grid, "fusion: adjoint pass grid exceeds the partials buffer");
        if (k) {
            void *a0 = lambda.data, *a1 = hl.data;
            void *args[4] = {&a0, &a1, &dpart, const_cast<PassParams<T2> *>(pp)};
            jit::launch_args(k, st.grid, 1u << (Cfg::M - Cfg::R), smem_bytes_for<Cfg, T2>(), lambda.stream, args);
        } else
            launch_pass<T2, Cfg>(st, lambda.stream, static_cast<T2 *>(lambda.data), static_cast<T2 *>(hl.data), dpart, *pp);
        if (pp->hdr.nslots > 0) {
            overlap_reduce_kernel<<<pp->hdr.nslots, 256, 0, lambda.stream>>>(dpart, st.grid, pacc);
            PLB_CUDA(cudaGetLastError());
        }
        lambda.launches++;
        passes.push_back({st.slots, st.slot_scale});
        stats[0]++;
        stats[2] += st.nops;
        return true;
    });
    if (!passes.empty()) {
        std::vector<double> h(passes.size() * kMaxPassOps);
        PLB_CUDA(cudaMemcpyAsync(h.data(), dacc, h.size() * sizeof(double), cudaMemcpyDeviceToHost, lambda.stream));
        PLB_CUDA(cudaStreamSynchronize(lambda.stream));
        for (size_t pi = 0; pi < passes.size(); pi++)
            for (size_t s = 0; s < passes[pi].slots.size(); s++)
                acc_host[passes[pi].slots[s]] += passes[pi].scale[s] * h[pi * kMaxPassOps + s];
    } else
        lambda.sync();
}
#endif // !PLB200_HOST_EMU

} // namespace

// A diagonal the tile encoder has no form for (a table over 2-4 bits that is not a parity form: the residual phases
// of DoubleExcitationPlus / Minus, a diagonal QubitUnitary ...) is, when all but a few of its entries are equal,
// a scalar on its control subspace times one controlled phase per exceptional entry — forms every pass can hold.
// Everything else passes through unchanged.  Only the fused path sees the expansion.
void expand_for_fusion(const COp &op, std::vector<COp> &out, bool tiles) {
    if (tiles && op.kind == OP_PAIRS && op.parity && op.cmask == 0 && op.blocks.size() == 1 && !op.tbits.empty()) {
        // exp(-i theta/2 P) for a Pauli word with X / Y letters (lower_pauli_rot's parity-paired form):
        // = U exp(-i theta/2 Z...Z) U^+ with U = H on the X wires and S H on the Y wires — single-bit pair ops and
        // a parity diagonal, which tile passes hold, instead of a sweep of its own.
        uint64_t x = 0;
        for (int b : op.tbits) x |= uint64_t{1} << b;
        const uint64_t z = op.pmask;
        const int ny = __builtin_popcountll(x & z);
        cd iy = 1.0;
        for (int q = 0; q < (ny & 3); q++) iy *= cd(0.0, 1.0);
        const double c = op.blocks[0].m[0].real();
        const double sn = (op.blocks[0].m[1] / (cd(0.0, -1.0) * iy)).real(); // m[1] = -i s i^ny
        const double r = 0.70710678118654752440;
        auto had = [&](int bit) {
            COp h;
            h.kind = OP_PAIRS;
            h.tbits = {bit};
            Block2 bl;
            bl.a = 0, bl.b = 1;
            bl.m[0] = r, bl.m[1] = r, bl.m[2] = r, bl.m[3] = -r;
            h.blocks.push_back(bl);
            return h;
        };
        auto sgate = [&](int bit, bool dagger) {
            COp d;
            d.kind = OP_DIAG;
            d.tbits = {bit};
            d.diag = {cd(1.0), cd(0.0, dagger ? -1.0 : 1.0)};
            return d;
        };
        for (int b : op.tbits) {
            if (z >> b & 1) out.push_back(sgate(b, true));
            out.push_back(had(b));
        }
        COp d;
        d.kind = OP_DIAG;
        d.parity = true;
        d.pmask = x | z;
        d.pd[0] = cd(c, -sn), d.pd[1] = cd(c, sn);
        out.push_back(d);
        for (int b : op.tbits) {
            out.push_back(had(b));
            if (z >> b & 1) out.push_back(sgate(b, false));
        }
        return;
    }
    const int k = op.k();
    if (op.kind != OP_DIAG || op.parity || k < 2 || k > 4 || classify(op).fusable) {
        out.push_back(op);
        return;
    }
    const int D = 1 << k;
    int best = 0, best_count = 0;
    for (int i = 0; i < D; i++) {
        int c = 0;
        for (int j = 0; j < D; j++) c += op.diag[j] == op.diag[i];
        if (c > best_count) best_count = c, best = i;
    }
    const cd v0 = op.diag[best];
    if (D - best_count > 4 || v0 == cd(0.0)) {
        out.push_back(op);
        return;
    }
    uint64_t tmask = 0;
    for (int b : op.tbits) tmask |= uint64_t{1} << b;
    auto phase_op = [&](cd ph, uint64_t cmask, uint64_t cval) {
        COp d;
        d.kind = OP_DIAG;
        d.cmask = cmask, d.cval = cval;
        d.diag = {ph};
        return d;
    };
    if (v0 != cd(1.0)) out.push_back(phase_op(v0, op.cmask, op.cval));
    for (int i = 0; i < D; i++) {
        if (op.diag[i] == v0) continue;
        uint64_t pat = 0;
        for (int j = 0; j < k; j++)
            if ((i >> j) & 1) pat |= uint64_t{1} << op.tbits[j];
        out.push_back(phase_op(op.diag[i] / v0, op.cmask | tmask, op.cval | pat));
    }
}

void schedule_stats(int n, int precision, const std::vector<COp> &ops, int64_t out[4]) {
    const auto items = as_items(ops, n >= (precision == 64 ? FwdCfg<double2>::M : FwdCfg<float2>::M) + 1);
    out[0] = out[1] = out[2] = out[3] = 0;
    auto count = [&](const Step &s, const void *) {
        if (s.op >= 0) out[1]++;
        else if (s.op == -1) out[0]++, out[2] += s.nrounds, out[3] += s.nops;
    };
    if (precision == 64) build_schedule<double2, FwdCfg<double2>>(n, 148, items, true, 0, count);
    else build_schedule<float2, FwdCfg<float2>>(n, 148, items, true, 0, count);
}

// Host-only: the specialised source of every tile pass of the tape (tools, tests, compile-time checks).
void pass_sources(int n, int precision, const std::vector<COp> &ops, std::vector<std::string> &out) {
    const auto items = as_items(ops, n >= (precision == 64 ? FwdCfg<double2>::M : FwdCfg<float2>::M) + 1);
    const char *jfe = std::getenv("PLB200_JIT_FORMS");
    const bool jit_forms = !(jfe && jfe[0] == '0'); // what the specialised tier compiles
    // PLB200_DUMP_REFUSE=1 (tests): every pass in the specialised forms is refused after its source was taken, as
    // the asynchronous tier does on a first sighting — the carried scalar then follows the interpreter's forms;
    // the sources must not depend on which way it went
    const bool refuse = std::getenv("PLB200_DUMP_REFUSE") != nullptr;
    out.clear();
    auto run = [&](auto t2, auto cfg) {
        using T2 = decltype(t2);
        using Cfg = decltype(cfg);
        build_schedule<T2, Cfg>(n, 148, items, true, jit_forms ? 2 : 1, [&](const Step &s, const PassParams<T2> *pp) -> bool {
            if (s.op != -1) return true;
            if (jit_forms && !s.jit_forms && refuse && !out.empty() && out.back() == "\x01") { // the interpreter encoding of a refused pass
                out.pop_back();
                return true;
            }
            out.push_back(jit::generate_pass_source<T2, Cfg>(*pp));
            if (refuse && s.jit_forms) {
                out.push_back("\x01");
                return false;
            }
            return true;
        });
    };
    if (precision == 64) run(double2{}, FwdCfg<double2>{});
    else run(float2{}, FwdCfg<float2>{});
}

bool build_adjoint_items(int64_t n, const std::vector<GateCall> &calls, const std::vector<int64_t> &tp,
                         int64_t num_param_ops, std::vector<AdjItem> &items, std::vector<double> &sfs) {
    const int64_t n_tp = static_cast<int64_t>(tp.size());
    items.clear();
    sfs.assign(n_tp, 0.0);
    int64_t tpi = n_tp - 1, cur = num_param_ops - 1;
    for (int64_t op_idx = static_cast<int64_t>(calls.size()) - 1; op_idx >= 0; op_idx--) {
        const GateCall &c = calls[op_idx];
        PLB_CHECK(c.params.size() <= 1, "The operation is not supported using the adjoint differentiation method");
        if (c.name == "StatePrep" || c.name == "BasisState") continue;
        if (tpi < 0) break;
        if (!c.params.empty()) {
            if (cur == tp[tpi]) {
                AdjItem it;
                it.overlap = true;
                double gscale = 0;
                if (!generator_as_pauli(n, c, &it.pw, &gscale)) return false;
                it.slot = static_cast<int>(tpi);
                sfs[tpi] = gscale * (c.inverse ? -1.0 : 1.0);
                items.push_back(std::move(it));
                tpi--;
            }
            cur--;
        }
        if (tpi < 0) break;
        GateCall inv = c;
        inv.inverse = !c.inverse;
        std::vector<COp> pieces;
        for (auto &lo : lower_gate(n, inv)) expand_for_fusion(lo, pieces, n >= 13);
        for (auto &lo : pieces) {
            AdjItem it;
            it.op = std::move(lo);
            items.push_back(std::move(it));
        }
    }
    return true;
}

#if !defined(PLB200_HOST_EMU)
void run_fused(StateVec &sv, const std::vector<COp> &ops) {
    sv.set_device();
    if (sv.precision == 64) run_fused_typed<double2>(sv, ops);
    else run_fused_typed<float2>(sv, ops);
}

bool run_fused_routed(StateVec &sv, const std::vector<COp> &ops, const RouteSpec &rs) {
    sv.set_device();
    // depends only on the state size and on whether NVRTC exists: the same answer on every rank
    if (sv.n < (sv.precision == 64 ? FwdCfg<double2>::M : FwdCfg<float2>::M) + 1 || !jit::available(nullptr)) {
        run_fused(sv, ops);
        return false;
    }
    return sv.precision == 64 ? run_fused_routed_typed<double2>(sv, ops, rs) : run_fused_routed_typed<float2>(sv, ops, rs);
}

void run_adjoint_fused(StateVec &lambda, StateVec &hl, const std::vector<AdjItem> &items, int n_slots,
                       double *acc_host, int64_t stats[3], StateVec *scratch) {
    lambda.set_device();
    stats[0] = stats[1] = stats[2] = 0;
    StateVec &owner = scratch ? *scratch : lambda;
    if (lambda.precision == 64) run_adjoint_typed<double2>(lambda, hl, items, n_slots, acc_host, stats, owner);
    else run_adjoint_typed<float2>(lambda, hl, items, n_slots, acc_host, stats, owner);
}
#else
// ------------------------------------------------------------------------------------------
// Test-only host emulation (libplb200_emu.so): the same schedule + encoder + per-thread code,
// executed on host memory.  Stand-alone steps are reported back to the caller (the Python test
// applies them with the numpy oracle), tile passes run through emulate_pass().
template <typename T2, class Cfg>
int emulate_typed(int n, const std::vector<AdjItem> &items, bool scaled, T2 *sv0, T2 *sv1, double *acc_host,
                  int n_slots, int (*standalone)(void *, int), void *ctx, int64_t stats[4]) {
    for (int i = 0; i < n_slots; i++) acc_host[i] = 0.0;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    int rc = 0;
    const char *jfe = std::getenv("PLB200_JIT_FORMS");
    const int jit_forms = std::getenv("PLB200_EMU_JIT") == nullptr ? 0 : (jfe && jfe[0] == '0') ? 1 : 2;
    // PLB200_EMU_REFUSE=1: refuse every pass that needs a specialised kernel, as a consumer without that kernel does
    const bool refuse = std::getenv("PLB200_EMU_REFUSE") != nullptr;
    build_schedule<T2, Cfg>(n, 148, items, scaled, jit_forms, [&](const Step &st, const PassParams<T2> *pp) -> bool {
        if (rc) return true;
        if (st.op >= 0) {
            stats[1]++;
            rc = standalone(ctx, st.op);
            return true;
        }
        if (refuse && (st.jit_forms || st.jit_only)) return false;
        if (st.op == -2) {
            for (T2 *sv : {sv0, sv1}) {
                if (!sv) continue;
                for (uint64_t i = 0; i < (uint64_t{1} << n); i++) {
                    const cd v = cd(sv[i].x, sv[i].y) * st.scale;
                    sv[i] = mk<T2>(v.real(), v.imag());
                }
            }
            return true;
        }
        std::vector<double> acc(kMaxPassOps, 0.0);
        if ((st.jit_only || st.jit_forms || (std::getenv("PLB200_EMU_JIT") && !st.fallback)) &&
            emulate_pass_jit<T2, Cfg>(sv0, *pp, jit::Route{}, nullptr, sv1, acc.data())) {
        } else if (st.jit_only) fail("emu: a pass with two-bit pair ops has no specialised source");
        else if (st.ext) emulate_pass<T2, Cfg, true>(sv0, sv1, acc.data(), *pp);
        else emulate_pass<T2, Cfg, false>(sv0, sv1, acc.data(), *pp);
        for (size_t s = 0; s < st.slots.size(); s++) acc_host[st.slots[s]] += st.slot_scale[s] * acc[s];
        stats[0]++, stats[2] += st.nrounds, stats[3] += st.nops;
        return true;
    });
    return rc;
}
// Test-only: a forward tape whose LAST pass stores through a route (the generated routed code on host memory);
// returns 1 when routed, 0 when the tape was applied in place, < 0 on a stand-alone callback error.
template <typename T2, class Cfg>
int emulate_routed_typed(int n, const std::vector<AdjItem> &items, bool scaled, T2 *sv0, const RouteSpec &rs,
                         int (*standalone)(void *, int), void *ctx) {
    auto held = std::make_unique<PassParams<T2>>();
    Step held_st;
    bool have = false;
    int rc = 0;
    auto plain = [&](const Step &st, const PassParams<T2> &pp) {
        std::vector<double> acc(kMaxPassOps, 0.0);
        if (emulate_pass_jit<T2, Cfg>(sv0, pp)) return;
        if (st.jit_only) fail("emu: a pass with two-bit pair ops has no specialised source");
        if (st.ext) emulate_pass<T2, Cfg, true>(sv0, nullptr, acc.data(), pp);
        else emulate_pass<T2, Cfg, false>(sv0, nullptr, acc.data(), pp);
    };
    const char *jfe = std::getenv("PLB200_JIT_FORMS");
    build_schedule<T2, Cfg>(n, 148, items, scaled, (jfe && jfe[0] == '0') ? 1 : 2, [&](const Step &st, const PassParams<T2> *pp) {
        if (rc) return;
        if (have) plain(held_st, *held), have = false;
        if (st.op >= 0) rc = standalone(ctx, st.op);
        else if (st.op == -2) {
            for (uint64_t i = 0; i < (uint64_t{1} << n); i++) {
                const cd v = cd(sv0[i].x, sv0[i].y) * st.scale;
                sv0[i] = mk<T2>(v.real(), v.imag());
            }
        } else {
            std::memcpy(static_cast<void *>(held.get()), pp, sizeof(PassParams<T2>));
            held_st = st, have = true;
        }
    });
    if (rc) return -1;
    if (!have) identity_pass<T2, Cfg>(n, 148, *held, held_st);
    jit::RouteParams<T2> rp;
    const jit::Route route = tile_route<T2>(rs, *held, rp);
    if (emulate_pass_jit<T2, Cfg>(sv0, *held, route, &rp)) return 1;
    fail("routed pass: no specialised source");
    return 0;
}
int emulate_routed(int n, int precision, const std::vector<AdjItem> &items, bool scaled, void *sv0, const RouteSpec &rs,
                   int (*standalone)(void *, int), void *ctx) {
    if (precision == 64)
        return emulate_routed_typed<double2, FwdCfg<double2>>(n, items, scaled, static_cast<double2 *>(sv0), rs, standalone, ctx);
    return emulate_routed_typed<float2, FwdCfg<float2>>(n, items, scaled, static_cast<float2 *>(sv0), rs, standalone, ctx);
}
// schedule-only statistics of the adjoint sweep: {tile passes, stand-alone items, rounds, fused items}
void emu_adjoint_schedule_stats(int n, int precision, const std::vector<AdjItem> &items, int64_t out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    auto count = [&](const Step &s, const void *) {
        if (s.op >= 0) out[1]++;
        else if (s.op == -1) out[0]++, out[2] += s.nrounds, out[3] += s.nops;
    };
    if (precision == 64) build_schedule<double2, AdjCfg<double2>>(n, 148, items, true, 0, count);
    else build_schedule<float2, AdjCfg<float2>>(n, 148, items, true, 0, count);
}
int64_t emu_jit_passes() { return g_emu_jit_passes; }
int64_t emu_plan_hits() { return g_plan_hits; }
void emu_kind_hist(int64_t out[32], bool reset) {
    for (int i = 0; i < 32; i++) {
        out[i] = g_kind_hist[i];
        if (reset) g_kind_hist[i] = 0;
    }
}
int emulate_fused(int n, int precision, const std::vector<AdjItem> &items, bool adjoint, bool scaled, void *sv0,
                  void *sv1, double *acc_host, int n_slots, int (*standalone)(void *, int), void *ctx,
                  int64_t stats[4]) {
    if (precision == 64) {
        if (adjoint)
            return emulate_typed<double2, AdjCfg<double2>>(n, items, scaled, static_cast<double2 *>(sv0),
                                                           static_cast<double2 *>(sv1), acc_host, n_slots, standalone,
                                                           ctx, stats);
        return emulate_typed<double2, FwdCfg<double2>>(n, items, scaled, static_cast<double2 *>(sv0), nullptr,
                                                       acc_host, n_slots, standalone, ctx, stats);
    }
    if (adjoint)
        return emulate_typed<float2, AdjCfg<float2>>(n, items, scaled, static_cast<float2 *>(sv0),
                                                     static_cast<float2 *>(sv1), acc_host, n_slots, standalone, ctx,
                                                     stats);
    return emulate_typed<float2, FwdCfg<float2>>(n, items, scaled, static_cast<float2 *>(sv0), nullptr, acc_host,
                                                 n_slots, standalone, ctx, stats);
}
#endif

} // namespace plb200
