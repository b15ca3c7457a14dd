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
// Inside the tile a round picks R "register bits": each thread pulls the 2^R amplitudes that
// differ in those bits into registers, applies all of the round's ops there (2x2 updates on
// register pairs, phases on single registers), and stores them back.  Per-gate arithmetic is the
// same 2x2 complex update as the un-fused kernels, so results agree to rounding.
//
// Adjoint variant (AdjointJacobianLQubit.hpp:269-314 as ONE sweep per block of ops): the tile of
// lambda and the tile of H·lambda travel together; walking the tape backwards, a trainable op first
// contributes its overlap Im<H lambda| G |lambda> (G = (controlled) X / Y / Z-parity generator)
// from registers — accumulated per CTA in shared memory, flushed once with fp64 atomics — and then
// its inverse is applied to both tiles.  No mu copy, no per-op HBM sweep.
//
// Gate commutation used by the scheduler: ops acting on disjoint bit sets commute, nothing else.
#include "fusion.hpp"

#include <algorithm>
#include <cstring>

namespace plb200 {

namespace {

constexpr int kMaxR = 4;
constexpr int kMaxPassOps = 224;
constexpr int kMaxPassRounds = 22;

// forward pass: one state, 2^12 (c128) / 2^13 (c64) amplitudes = 64 KiB per tile, 16 per thread
template <typename T2> struct FwdCfg;
template <> struct FwdCfg<double2> {
    static constexpr int M = 12, LOW = 3, R = 4, NS = 1, MINB = 2;
};
template <> struct FwdCfg<float2> {
    static constexpr int M = 13, LOW = 4, R = 4, NS = 1, MINB = 2;
};
// adjoint pass: two states, 2 x 32 KiB tiles, 8 + 8 amplitudes per thread
template <typename T2> struct AdjCfg;
template <> struct AdjCfg<double2> {
    static constexpr int M = 11, LOW = 3, R = 3, NS = 2, MINB = 2;
};
template <> struct AdjCfg<float2> {
    static constexpr int M = 12, LOW = 4, R = 3, NS = 2, MINB = 2;
};

// op kinds inside a tile pass
enum : int {
    TK_GENERAL = 0, // complex 2x2 on register bit p
    TK_SWAP = 1,    // [[0,1],[1,0]] (PauliX / CNOT / Toffoli): pure register exchange, no flops
    TK_REAL = 2,    // real 2x2 (RY, Hadamard)
    TK_RXLIKE = 3,  // real diagonal, imaginary off-diagonal (RX)
    TK_DIAG_T = 4,  // phase m[parity], parity bits on thread / outside bits only
    TK_DIAG_R = 5,  // ... plus exactly one register bit p
    TK_DIAG_G = 6,  // ... any register bits (per-amplitude select)
    TK_DIAG1_T = 7, // as DIAG_T with m[0] == 1: only the parity-1 amplitudes are multiplied
    TK_DIAG1_R = 8, // as DIAG_R with m[0] == 1
    TK_OVL_X = 16,  // adjoint: accumulate Im<h| X_p |l> (controlled by the active mask)
    TK_OVL_Y = 17,  // adjoint: Im<h| Y_p |l>
    TK_OVL_D = 18,  // adjoint: Im<h| D |l>, D = diag(g[parity]), g = (m[0].x, m[0].y)
};

// Controls / parity masks are split on the host into a per-thread part (tile-local index bits that
// are thread bits in this round) and a per-register part (bit patterns over the register index u),
// so that the per-amplitude predicate is a single bit test on a compile-time position.
template <typename T2> struct alignas(16) TileOp {
    int kind;
    int p;
    uint32_t cm_thr, cv_thr; // controls on thread bits (tile-local index space)
    uint32_t pm_thr;         // parity mask on thread bits
    uint32_t umask;          // bit u: register index u satisfies the register-bit controls
    uint32_t upar;           // bit u: parity of u's register bits under the parity mask
    uint32_t slot;           // adjoint: accumulator slot of this overlap inside the pass
    uint64_t cmask_o, cval_o, pmask_o; // bits outside the tile: uniform per tile
    T2 m[4];
};
struct alignas(16) RoundHdr {
    int first_op, nops;
    uint32_t lowmask[kMaxR]; // insertion masks (ascending local positions)
    uint32_t roff[1 << kMaxR];
};
struct alignas(16) PassHdr {
    int nrounds, nops_total;
    uint64_t ntiles;
    int nslots, pad;
    BitInsert tile_ins; // zeros at the M tile bits
};
// The whole pass description travels as a __grid_constant__ kernel parameter (constant bank,
// uniform loads): nothing about the ops is fetched through the LSU/L1 data path.
template <typename T2> struct alignas(16) PassParams {
    PassHdr hdr;
    RoundHdr rounds[kMaxPassRounds];
    TileOp<T2> ops[kMaxPassOps];
};

__device__ __forceinline__ uint32_t swz(uint32_t j) { return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9)) & 7u); }

template <typename T2, int R, int P, int KIND>
__device__ __forceinline__ void apply_pair(T2 (&v)[1 << R], const TileOp<T2> &op, uint32_t active) {
    const T2 m0 = op.m[0], m1 = op.m[1], m2 = op.m[2], m3 = op.m[3];
#pragma unroll
    for (int q = 0; q < (1 << (R - 1)); q++) {
        const int u0 = ((q >> P) << (P + 1)) | (q & ((1 << P) - 1));
        const int u1 = u0 | (1 << P);
        if (active & (1u << u0)) {
            const T2 a = v[u0], b = v[u1];
            if constexpr (KIND == TK_SWAP) {
                v[u0] = b, v[u1] = a;
            } else if constexpr (KIND == TK_REAL) {
                v[u0].x = fma(m1.x, b.x, m0.x * a.x), v[u0].y = fma(m1.x, b.y, m0.x * a.y);
                v[u1].x = fma(m3.x, b.x, m2.x * a.x), v[u1].y = fma(m3.x, b.y, m2.x * a.y);
            } else if constexpr (KIND == TK_RXLIKE) {
                // (m0.x) a + (i m1.y) b ; (i m2.y) a + (m3.x) b
                v[u0].x = fma(-m1.y, b.y, m0.x * a.x), v[u0].y = fma(m1.y, b.x, m0.x * a.y);
                v[u1].x = fma(-m2.y, a.y, m3.x * b.x), v[u1].y = fma(m2.y, a.x, m3.x * b.y);
            } else {
                v[u0] = cfma(m1, b, cmul(m0, a));
                v[u1] = cfma(m3, b, cmul(m2, a));
            }
        }
    }
}

template <typename T2, int R, int P, bool ONE>
__device__ __forceinline__ void apply_diag_r(T2 (&v)[1 << R], const TileOp<T2> &op, uint32_t active, bool pt) {
    if constexpr (ONE) {
        const T2 d1 = op.m[1];
#pragma unroll
        for (int u = 0; u < (1 << R); u++) {
            const bool par = ((u >> P) & 1) ? !pt : pt;
            if ((active & (1u << u)) && par) v[u] = cmul(v[u], d1);
        }
    } else {
        const T2 da = pt ? op.m[1] : op.m[0], db = pt ? op.m[0] : op.m[1];
#pragma unroll
        for (int u = 0; u < (1 << R); u++)
            if (active & (1u << u)) v[u] = cmul(v[u], ((u >> P) & 1) ? db : da);
    }
}

#define PLB_SWITCH_P(R, CALL)                                                                            \
    switch (op.p) {                                                                                      \
    case 0: { constexpr int P = 0; CALL; } break;                                                        \
    case 1: { constexpr int P = (R > 1 ? 1 : 0); CALL; } break;                                          \
    case 2: { constexpr int P = (R > 2 ? 2 : R - 1); CALL; } break;                                      \
    default: { constexpr int P = (R > 3 ? 3 : R - 1); CALL; } break;                                     \
    }

// one gate op on one register set
template <typename T2, int R>
__device__ __forceinline__ void apply_gate(T2 (&v)[1 << R], const TileOp<T2> &op, uint32_t active, bool pt) {
    switch (op.kind) {
    case TK_GENERAL:
        PLB_SWITCH_P(R, (apply_pair<T2, R, P, TK_GENERAL>(v, op, active)));
        break;
    case TK_SWAP:
        PLB_SWITCH_P(R, (apply_pair<T2, R, P, TK_SWAP>(v, op, active)));
        break;
    case TK_REAL:
        PLB_SWITCH_P(R, (apply_pair<T2, R, P, TK_REAL>(v, op, active)));
        break;
    case TK_RXLIKE:
        PLB_SWITCH_P(R, (apply_pair<T2, R, P, TK_RXLIKE>(v, op, active)));
        break;
    case TK_DIAG_T: {
        const T2 d = pt ? op.m[1] : op.m[0];
#pragma unroll
        for (int u = 0; u < (1 << R); u++)
            if (active & (1u << u)) v[u] = cmul(v[u], d);
    } break;
    case TK_DIAG1_T:
        if (pt) {
            const T2 d = op.m[1];
#pragma unroll
            for (int u = 0; u < (1 << R); u++)
                if (active & (1u << u)) v[u] = cmul(v[u], d);
        }
        break;
    case TK_DIAG_R:
        PLB_SWITCH_P(R, (apply_diag_r<T2, R, P, false>(v, op, active, pt)));
        break;
    case TK_DIAG1_R:
        PLB_SWITCH_P(R, (apply_diag_r<T2, R, P, true>(v, op, active, pt)));
        break;
    default: {
        const uint32_t pb = pt ? ~op.upar : op.upar;
        const T2 d0 = op.m[0], d1 = op.m[1];
#pragma unroll
        for (int u = 0; u < (1 << R); u++)
            if (active & (1u << u)) v[u] = cmul(v[u], (pb >> u & 1) ? d1 : d0);
    } break;
    }
}

// Im(conj(a) b), Re(conj(a) b)
template <typename T2> __device__ __forceinline__ double im_cb(T2 a, T2 b) {
    return static_cast<double>(a.x) * b.y - static_cast<double>(a.y) * b.x;
}
template <typename T2> __device__ __forceinline__ double re_cb(T2 a, T2 b) {
    return static_cast<double>(a.x) * b.x + static_cast<double>(a.y) * b.y;
}

template <typename T2, int R, int P, bool ISY>
__device__ __forceinline__ double overlap_pair(const T2 (&l)[1 << R], const T2 (&h)[1 << R], uint32_t active) {
    double s = 0;
#pragma unroll
    for (int q = 0; q < (1 << (R - 1)); q++) {
        const int u0 = ((q >> P) << (P + 1)) | (q & ((1 << P) - 1));
        const int u1 = u0 | (1 << P);
        if (active & (1u << u0)) {
            // (Y l)[u0] = -i l[u1], (Y l)[u1] = i l[u0];  (X l)[u0] = l[u1], (X l)[u1] = l[u0]
            if constexpr (ISY) s += re_cb(h[u1], l[u0]) - re_cb(h[u0], l[u1]);
            else s += im_cb(h[u0], l[u1]) + im_cb(h[u1], l[u0]);
        }
    }
    return s;
}

template <typename T2, class Cfg>
__global__ void __launch_bounds__(1 << (Cfg::M - Cfg::R), Cfg::MINB)
    tile_kernel(T2 *__restrict__ sv0, T2 *__restrict__ sv1, const uint64_t *__restrict__ goff_g,
                double *__restrict__ acc_g, const __grid_constant__ PassParams<T2> pp) {
    constexpr int M = Cfg::M, LOW = Cfg::LOW, R = Cfg::R, NS = Cfg::NS;
    constexpr int NT = 1 << (M - R);
    constexpr int NV = 1 << R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T2 *tile0 = reinterpret_cast<T2 *>(smem_raw);
    T2 *tile1 = tile0 + (NS == 2 ? (1 << M) : 0);
    uint64_t *goff = reinterpret_cast<uint64_t *>(smem_raw + NS * (sizeof(T2) << M));
    double *acc = reinterpret_cast<double *>(goff + (1 << (M - LOW)));
    static_assert(sizeof(PassParams<T2>) <= 32764, "kernel parameter space");

    for (int i = threadIdx.x; i < (1 << (M - LOW)); i += NT) goff[i] = goff_g[i];
    if constexpr (NS == 2)
        for (int i = threadIdx.x; i < pp.hdr.nslots; i += NT) acc[i] = 0.0;
    __syncthreads();
    const uint32_t tid = threadIdx.x;
    const int nrounds = pp.hdr.nrounds;

    for (uint64_t t = blockIdx.x; t < pp.hdr.ntiles; t += gridDim.x) {
        const uint64_t base = insert_bits(t, pp.hdr.tile_ins);
        // ---- load the tile(s) (coalesced 128-byte lines)
        {
            T2 v[NV];
#pragma unroll
            for (int u = 0; u < NV; u++) {
                const uint32_t j = tid + u * NT;
                v[u] = sv0[base | goff[j >> LOW] | (j & ((1u << LOW) - 1))];
            }
#pragma unroll
            for (int u = 0; u < NV; u++) tile0[swz(tid + u * NT)] = v[u];
            if constexpr (NS == 2) {
#pragma unroll
                for (int u = 0; u < NV; u++) {
                    const uint32_t j = tid + u * NT;
                    v[u] = sv1[base | goff[j >> LOW] | (j & ((1u << LOW) - 1))];
                }
#pragma unroll
                for (int u = 0; u < NV; u++) tile1[swz(tid + u * NT)] = v[u];
            }
        }
        __syncthreads();
        // ---- rounds
        for (int r = 0; r < nrounds; r++) {
            const RoundHdr &rh = pp.rounds[r];
            uint32_t jbase = tid;
#pragma unroll
            for (int i = 0; i < R; i++) {
                const uint32_t lm = rh.lowmask[i];
                jbase = ((jbase & ~lm) << 1) | (jbase & lm);
            }
            T2 v[NV];
            T2 h[NS == 2 ? NV : 1];
#pragma unroll
            for (int u = 0; u < NV; u++) v[u] = tile0[swz(jbase | rh.roff[u])];
            if constexpr (NS == 2) {
#pragma unroll
                for (int u = 0; u < NV; u++) h[u] = tile1[swz(jbase | rh.roff[u])];
            }
            const int k_end = rh.first_op + rh.nops;
            for (int k = rh.first_op; k < k_end; k++) {
                const TileOp<T2> &op = pp.ops[k];
                if ((base & op.cmask_o) != op.cval_o) continue; // uniform per tile
                const uint32_t active = ((jbase & op.cm_thr) == op.cv_thr) ? op.umask : 0u;
                const bool pt = ((__popc(jbase & op.pm_thr) + __popcll(base & op.pmask_o)) & 1) != 0;
                if constexpr (NS == 2) {
                    if (op.kind >= TK_OVL_X) {
                        double s = 0;
                        if (op.kind == TK_OVL_X) {
                            PLB_SWITCH_P(R, (s = overlap_pair<T2, R, P, false>(v, h, active)));
                        } else if (op.kind == TK_OVL_Y) {
                            PLB_SWITCH_P(R, (s = overlap_pair<T2, R, P, true>(v, h, active)));
                        } else {
                            const uint32_t pb = pt ? ~op.upar : op.upar;
                            const double g0 = op.m[0].x, g1 = op.m[0].y;
#pragma unroll
                            for (int u = 0; u < NV; u++)
                                if (active & (1u << u)) s += ((pb >> u & 1) ? g1 : g0) * im_cb(h[u], v[u]);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if ((tid & 31) == 0) atomicAdd(&acc[op.slot], s);
                        continue;
                    }
                    apply_gate<T2, R>(h, op, active, pt);
                }
                apply_gate<T2, R>(v, op, active, pt);
            }
#pragma unroll
            for (int u = 0; u < NV; u++) tile0[swz(jbase | rh.roff[u])] = v[u];
            if constexpr (NS == 2) {
#pragma unroll
                for (int u = 0; u < NV; u++) tile1[swz(jbase | rh.roff[u])] = h[u];
            }
            __syncthreads();
        }
        // ---- store the tile(s)
        {
            T2 v[NV];
#pragma unroll
            for (int u = 0; u < NV; u++) v[u] = tile0[swz(tid + u * NT)];
#pragma unroll
            for (int u = 0; u < NV; u++) {
                const uint32_t j = tid + u * NT;
                sv0[base | goff[j >> LOW] | (j & ((1u << LOW) - 1))] = v[u];
            }
            if constexpr (NS == 2) {
#pragma unroll
                for (int u = 0; u < NV; u++) v[u] = tile1[swz(tid + u * NT)];
#pragma unroll
                for (int u = 0; u < NV; u++) {
                    const uint32_t j = tid + u * NT;
                    sv1[base | goff[j >> LOW] | (j & ((1u << LOW) - 1))] = v[u];
                }
            }
        }
        __syncthreads();
    }
    if constexpr (NS == 2) {
        for (int i = threadIdx.x; i < pp.hdr.nslots; i += NT)
            if (acc[i] != 0.0) atomicAdd(&acc_g[i], acc[i]);
    }
}

// --------------------------------------------------------------------------- host scheduler
struct FOp {
    uint64_t nd = 0;  // non-diagonal target bits
    uint64_t all = 0; // every involved bit
    bool fusable = false;
    // parity form of a fusable diagonal
    uint64_t pmask = 0;
    cd d[2];
};

FOp classify(const COp &op) {
    FOp f;
    uint64_t t = 0;
    for (int b : op.tbits) t |= uint64_t{1} << b;
    f.all = t | op.cmask | (op.parity ? op.pmask : 0);
    if (op.kind == OP_PROJECT) f.all = ~uint64_t{0};
    if (op.kind == OP_PAIRS && !op.parity && op.tbits.size() == 1 && op.blocks.size() == 1 && op.blocks[0].a == 0 &&
        op.blocks[0].b == 1) {
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

FOp classify(const AdjItem &it) {
    if (!it.overlap) return classify(it.op);
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

struct HostPass {
    std::vector<int> tbits;               // sorted ascending, size M
    std::vector<std::vector<int>> rounds; // item indices per round
    std::vector<uint64_t> round_bits;     // register bits (global bit masks) per round
};

struct Step {
    int op = -1;       // >= 0: stand-alone item
    size_t off = 0;    // else: offset of the pass's goff table in the arena (bytes)
    size_t params = 0; // index into the params vector
    unsigned grid = 0;
    int nrounds = 0, nops = 0;
    std::vector<int> slots; // adjoint: global accumulator slot of each pass-local slot
};

// Pure host: schedule `items` on an n-qubit state into tile passes / stand-alone items.
template <typename T2, class Cfg>
void build_schedule(int n, int sm_count, const std::vector<AdjItem> &items, std::vector<Step> &steps,
                    std::vector<unsigned char> &arena, std::vector<PassParams<T2>> &params) {
    constexpr int M = Cfg::M, LOW = Cfg::LOW, R = Cfg::R;
    steps.clear();
    arena.clear();
    params.clear();
    if (n < M + 1 || items.size() < 2) {
        for (size_t i = 0; i < items.size(); i++) {
            Step st;
            st.op = static_cast<int>(i);
            steps.push_back(st);
        }
        return;
    }
    const uint64_t full = (n >= 64) ? ~uint64_t{0} : ((uint64_t{1} << n) - 1);
    std::vector<FOp> f(items.size());
    for (size_t i = 0; i < items.size(); i++) {
        f[i] = classify(items[i]);
        f[i].all &= full;
    }
    std::vector<char> done(items.size(), 0);
    size_t first = 0;
    const size_t window = 512;
    const uint64_t lowbits = (uint64_t{1} << LOW) - 1;
    std::vector<int> pending, exec;

    while (true) {
        while (first < items.size() && done[first]) first++;
        if (first >= items.size()) break;
        pending.clear();
        for (size_t i = first; i < items.size() && pending.size() < window; i++)
            if (!done[i]) pending.push_back(static_cast<int>(i));
        // ---- choose the tile bits
        uint64_t T = grow(f, pending, lowbits, M, full, full, exec);
        if (exec.size() < 2) {
            // nothing worth a tile pass: run the first pending item on its own
            Step st;
            st.op = static_cast<int>(first);
            steps.push_back(st);
            done[first] = 1;
            continue;
        }
        // pad T to exactly M bits with the lowest free bits
        for (int b = 0; b < n && __builtin_popcountll(T) < M; b++) T |= uint64_t{1} << b;
        simulate(f, pending, T, full, exec);
        if (exec.size() > static_cast<size_t>(kMaxPassOps)) exec.resize(kMaxPassOps); // a prefix stays valid
        std::vector<int> pass_ops = exec;

        // ---- rounds
        HostPass hp;
        for (int b = 0; b < n; b++)
            if (T >> b & 1) hp.tbits.push_back(b);
        std::vector<int> rem = pass_ops, rexec;
        while (!rem.empty()) {
            uint64_t seed = f[rem[0]].nd; // guarantees progress
            uint64_t Rb = grow(f, rem, seed, R, T, full, rexec);
            // pad to R bits with tile bits (prefer high local bits: conflict-free shared accesses)
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

        // ---- encode the plan
        const size_t goff_n = size_t{1} << (M - LOW);
        const size_t bytes = goff_n * sizeof(uint64_t);
        const size_t off = (arena.size() + 255) & ~size_t{255};
        arena.resize(off + bytes, 0);
        params.emplace_back();
        std::memset(&params.back(), 0, sizeof(PassParams<T2>));
        PassHdr *hdr = &params.back().hdr;
        uint64_t *goff = reinterpret_cast<uint64_t *>(arena.data() + off);
        RoundHdr *rh = params.back().rounds;
        TileOp<T2> *top = params.back().ops;
        hdr->nrounds = static_cast<int>(hp.rounds.size());
        hdr->nops_total = static_cast<int>(pass_ops.size());
        hdr->ntiles = uint64_t{1} << (n - M);
        hdr->tile_ins.n = 0;
        for (int b : hp.tbits) hdr->tile_ins.lowmask[hdr->tile_ins.n++] = (uint64_t{1} << b) - 1;
        int local_of[64];
        for (int i = 0; i < 64; i++) local_of[i] = -1;
        for (int i = 0; i < M; i++) local_of[hp.tbits[i]] = i;
        for (size_t j = 0; j < goff_n; j++) {
            uint64_t o = 0;
            for (int i = LOW; i < M; i++)
                if ((j >> (i - LOW)) & 1) o |= uint64_t{1} << hp.tbits[i];
            goff[j] = o;
        }
        auto to_local = [&](uint64_t mask) {
            uint32_t l = 0;
            for (int i = 0; i < M; i++)
                if (mask >> hp.tbits[i] & 1) l |= 1u << i;
            return l;
        };
        Step st;
        int op_cursor = 0;
        for (size_t r = 0; r < hp.rounds.size(); r++) {
            std::vector<int> rl; // local positions of the register bits, ascending
            for (int i = 0; i < M; i++)
                if (hp.round_bits[r] >> hp.tbits[i] & 1) rl.push_back(i);
            rh[r].first_op = op_cursor;
            rh[r].nops = static_cast<int>(hp.rounds[r].size());
            for (int i = 0; i < R; i++) rh[r].lowmask[i] = (1u << rl[i]) - 1;
            for (int u = 0; u < (1 << R); u++) {
                uint32_t o = 0;
                for (int i = 0; i < R; i++)
                    if (u >> i & 1) o |= 1u << rl[i];
                rh[r].roff[u] = o;
            }
            uint32_t rmask_l = 0;
            for (int i = 0; i < R; i++) rmask_l |= 1u << rl[i];
            for (int idx : hp.rounds[r]) {
                const AdjItem &it = items[idx];
                const FOp &fo = f[idx];
                TileOp<T2> &t = top[op_cursor++];
                const uint64_t cmask = it.overlap ? it.pw.cmask : it.op.cmask;
                const uint64_t cval = it.overlap ? it.pw.cval : it.op.cval;
                const uint64_t pmask = it.overlap ? (it.pw.x ? 0 : it.pw.z) : (it.op.kind == OP_PAIRS ? 0 : fo.pmask);
                const uint32_t cm_l = to_local(cmask & T), cv_l = to_local(cval & T), pm_l = to_local(pmask & T);
                t.cm_thr = cm_l & ~rmask_l, t.cv_thr = cv_l & ~rmask_l;
                t.cmask_o = cmask & ~T, t.cval_o = cval & ~T;
                t.pm_thr = pm_l & ~rmask_l;
                t.pmask_o = pmask & ~T;
                t.umask = 0, t.upar = 0;
                for (int u = 0; u < (1 << R); u++) {
                    const uint32_t ro = rh[r].roff[u];
                    if ((ro & cm_l) == (cv_l & rmask_l)) t.umask |= 1u << u;
                    if (__builtin_popcount(ro & pm_l) & 1) t.upar |= 1u << u;
                }
                auto reg_pos = [&](int global_bit) {
                    const int lp = local_of[global_bit];
                    return static_cast<int>(std::find(rl.begin(), rl.end(), lp) - rl.begin());
                };
                if (it.overlap) {
                    t.slot = static_cast<uint32_t>(st.slots.size());
                    st.slots.push_back(it.slot);
                    if (it.pw.x) {
                        t.kind = (it.pw.z == it.pw.x) ? TK_OVL_Y : TK_OVL_X;
                        t.p = reg_pos(__builtin_ctzll(it.pw.x));
                    } else {
                        t.kind = TK_OVL_D;
                        t.m[0] = mk<T2>(1.0, it.pw.z ? -1.0 : 1.0);
                    }
                } else if (it.op.kind == OP_PAIRS) {
                    const cd *m = it.op.blocks[0].m;
                    t.kind = TK_GENERAL;
                    if (m[0] == cd(0.0) && m[3] == cd(0.0) && m[1] == cd(1.0) && m[2] == cd(1.0)) t.kind = TK_SWAP;
                    else if (m[0].imag() == 0 && m[1].imag() == 0 && m[2].imag() == 0 && m[3].imag() == 0)
                        t.kind = TK_REAL;
                    else if (m[0].imag() == 0 && m[3].imag() == 0 && m[1].real() == 0 && m[2].real() == 0)
                        t.kind = TK_RXLIKE;
                    t.p = reg_pos(it.op.tbits[0]);
                    for (int q = 0; q < 4; q++) t.m[q] = mk<T2>(m[q].real(), m[q].imag());
                } else {
                    const bool one = (fo.d[0] == cd(1.0));
                    const uint32_t pr = pm_l & rmask_l;
                    if (pr == 0) t.kind = one ? TK_DIAG1_T : TK_DIAG_T;
                    else if (__builtin_popcount(pr) == 1) {
                        t.kind = one ? TK_DIAG1_R : TK_DIAG_R;
                        const int lp = __builtin_ctz(pr);
                        t.p = static_cast<int>(std::find(rl.begin(), rl.end(), lp) - rl.begin());
                    } else
                        t.kind = TK_DIAG_G;
                    t.m[0] = mk<T2>(fo.d[0].real(), fo.d[0].imag());
                    t.m[1] = mk<T2>(fo.d[1].real(), fo.d[1].imag());
                }
            }
        }
        hdr->nslots = static_cast<int>(st.slots.size());
        st.off = off;
        st.params = params.size() - 1;
        st.grid = static_cast<unsigned>(std::min<uint64_t>(hdr->ntiles, uint64_t(sm_count) * 3 * 64));
        st.nrounds = hdr->nrounds, st.nops = hdr->nops_total;
        steps.push_back(std::move(st));
    }
}

template <class Cfg, typename T2> size_t smem_bytes_for() {
    return Cfg::NS * (sizeof(T2) << Cfg::M) + (sizeof(uint64_t) << (Cfg::M - Cfg::LOW)) +
           (Cfg::NS == 2 ? sizeof(double) * kMaxPassOps : 0);
}

template <typename T2, class Cfg> void prepare_kernel() {
    static bool done = false;
    if (done) return;
    PLB_CUDA(cudaFuncSetAttribute(tile_kernel<T2, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes_for<Cfg, T2>())));
    PLB_CUDA(cudaFuncSetAttribute(tile_kernel<T2, Cfg>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    done = true;
}

std::vector<AdjItem> as_items(const std::vector<COp> &ops) {
    std::vector<AdjItem> items(ops.size());
    for (size_t i = 0; i < ops.size(); i++) items[i].op = ops[i];
    return items;
}

template <typename T2> void run_fused_typed(StateVec &sv, const std::vector<COp> &ops) {
    using Cfg = FwdCfg<T2>;
    std::vector<Step> steps;
    std::vector<unsigned char> arena;
    std::vector<PassParams<T2>> params;
    const auto items = as_items(ops);
    build_schedule<T2, Cfg>(static_cast<int>(sv.n), sv.sm_count, items, steps, arena, params);
    // ---- upload every pass's offset table once, then launch the whole schedule back to back
    unsigned char *dplan = nullptr;
    if (!arena.empty()) {
        prepare_kernel<T2, Cfg>();
        dplan = static_cast<unsigned char *>(sv.plan_buf(arena.size()));
        PLB_CUDA(cudaMemcpyAsync(dplan, arena.data(), arena.size(), cudaMemcpyHostToDevice, sv.stream));
        PLB_CUDA(cudaStreamSynchronize(sv.stream)); // arena is a local
    }
    for (const Step &st : steps) {
        if (st.op >= 0) {
            launch_op(sv, ops[st.op]);
            continue;
        }
        tile_kernel<T2, Cfg><<<st.grid, 1 << (Cfg::M - Cfg::R), smem_bytes_for<Cfg, T2>(), sv.stream>>>(
            static_cast<T2 *>(sv.data), nullptr, reinterpret_cast<const uint64_t *>(dplan + st.off), nullptr,
            params[st.params]);
        PLB_CUDA(cudaGetLastError());
        sv.launches++;
    }
}

template <typename T2>
void run_adjoint_typed(StateVec &lambda, StateVec &hl, const std::vector<AdjItem> &items, int n_slots,
                       double *acc_host, int64_t stats[3]) {
    using Cfg = AdjCfg<T2>;
    std::vector<Step> steps;
    std::vector<unsigned char> arena;
    std::vector<PassParams<T2>> params;
    build_schedule<T2, Cfg>(static_cast<int>(lambda.n), lambda.sm_count, items, steps, arena, params);
    for (int i = 0; i < n_slots; i++) acc_host[i] = 0.0;
    // device accumulators: one slab of kMaxPassOps doubles per tile pass
    size_t n_pass = 0;
    for (const Step &st : steps)
        if (st.op < 0) n_pass++;
    unsigned char *dplan = nullptr;
    double *dacc = nullptr;
    const size_t acc_off = (arena.size() + 255) & ~size_t{255};
    const size_t acc_bytes = n_pass * kMaxPassOps * sizeof(double);
    if (n_pass) {
        prepare_kernel<T2, Cfg>();
        dplan = static_cast<unsigned char *>(lambda.plan_buf(acc_off + acc_bytes));
        PLB_CUDA(cudaMemcpyAsync(dplan, arena.data(), arena.size(), cudaMemcpyHostToDevice, lambda.stream));
        dacc = reinterpret_cast<double *>(dplan + acc_off);
        PLB_CUDA(cudaMemsetAsync(dacc, 0, acc_bytes, lambda.stream));
        PLB_CUDA(cudaStreamSynchronize(lambda.stream));
    }
    size_t pass_idx = 0;
    for (const Step &st : steps) {
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
            continue;
        }
        tile_kernel<T2, Cfg><<<st.grid, 1 << (Cfg::M - Cfg::R), smem_bytes_for<Cfg, T2>(), lambda.stream>>>(
            static_cast<T2 *>(lambda.data), static_cast<T2 *>(hl.data),
            reinterpret_cast<const uint64_t *>(dplan + st.off), dacc + pass_idx * kMaxPassOps, params[st.params]);
        PLB_CUDA(cudaGetLastError());
        lambda.launches++;
        pass_idx++;
        stats[0]++;
        stats[2] += st.nops;
    }
    if (n_pass) {
        std::vector<double> h(n_pass * kMaxPassOps);
        PLB_CUDA(cudaMemcpyAsync(h.data(), dacc, acc_bytes, cudaMemcpyDeviceToHost, lambda.stream));
        PLB_CUDA(cudaStreamSynchronize(lambda.stream));
        pass_idx = 0;
        for (const Step &st : steps) {
            if (st.op >= 0) continue;
            for (size_t s = 0; s < st.slots.size(); s++) acc_host[st.slots[s]] += h[pass_idx * kMaxPassOps + s];
            pass_idx++;
        }
    } else
        lambda.sync();
}

} // namespace

void schedule_stats(int n, int precision, const std::vector<COp> &ops, int64_t out[4]) {
    std::vector<Step> steps;
    std::vector<unsigned char> arena;
    const auto items = as_items(ops);
    if (precision == 64) {
        std::vector<PassParams<double2>> params;
        build_schedule<double2, FwdCfg<double2>>(n, 148, items, steps, arena, params);
    } else {
        std::vector<PassParams<float2>> params;
        build_schedule<float2, FwdCfg<float2>>(n, 148, items, steps, arena, params);
    }
    out[0] = out[1] = out[2] = out[3] = 0;
    for (const auto &s : steps) {
        if (s.op >= 0) out[1]++;
        else out[0]++, out[2] += s.nrounds, out[3] += s.nops;
    }
}

void run_fused(StateVec &sv, const std::vector<COp> &ops) {
    sv.set_device();
    if (sv.precision == 64) run_fused_typed<double2>(sv, ops);
    else run_fused_typed<float2>(sv, ops);
}

void run_adjoint_fused(StateVec &lambda, StateVec &hl, const std::vector<AdjItem> &items, int n_slots,
                       double *acc_host, int64_t stats[3]) {
    lambda.set_device();
    stats[0] = stats[1] = stats[2] = 0;
    if (lambda.precision == 64) run_adjoint_typed<double2>(lambda, hl, items, n_slots, acc_host, stats);
    else run_adjoint_typed<float2>(lambda, hl, items, n_slots, acc_host, stats);
}

} // namespace plb200
