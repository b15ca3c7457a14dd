// Pass specialisation: CUDA source of ONE fused pass, generated from its encoded description.
//
// The tile interpreter (tile_exec.cuh) spends ~40 decode/dispatch instructions per op and thread against
// 32-48 FP64 instructions of work, plus 12 LOP3/MOV per register pair for every X / CNOT.  A pass kernel
// generated for the pass's STRUCTURE removes all of that: the op sequence is straight-line code, register
// permutations (X, CNOT and SWAP on register bits) are renamings done here at generation time (zero
// instructions), controls / parities on thread bits are compare-with-immediate, shared-memory offsets are
// immediates, and diagonal factors that do not depend on a register bit are multiplied into ONE scalar per
// thread and round.  Everything numeric (angles, phases, the positions of the tile bits, masks over bits
// outside the tile) stays a RUN-TIME argument read from the same PassParams block the interpreter takes
// (constant bank operands), so one compiled kernel serves every parameter set of a variational circuit
// and every tile placement.  The structure key is the generated text.
//
// The text compiles (a) with NVRTC for sm_100a (jit_runtime.cpp) and (b) with g++ under -DPLB_JIT_HOST,
// which tests use to run the generated per-thread code on host memory against the oracle.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tile_exec.cuh"

namespace plb200 {
namespace jit {

using namespace tile;

inline const char *prelude() {
    return R"PLB(
#if defined(PLB_JIT_HOST)
#include <cmath>
#include <cstdint>
#include <cstddef>
typedef uint32_t u32;
typedef uint64_t u64;
#define DEV static inline
#define POPC(x) __builtin_popcount(x)
#define POPCLL(x) __builtin_popcountll(x)
#define PLB_ALIGN(n) alignas(n)
using std::fma;
#if PLB_DOUBLE
typedef double real;
struct alignas(16) T2 { real x, y; };
#else
typedef float real;
struct alignas(8) T2 { real x, y; };
#endif
#else
typedef unsigned int u32;
typedef unsigned long long u64;
#define DEV __device__ __forceinline__
#define POPC(x) __popc(x)
#define POPCLL(x) __popcll(x)
#define PLB_ALIGN(n) __align__(n)
#if PLB_DOUBLE
typedef double real;
typedef double2 T2;
#else
typedef float real;
typedef float2 T2;
#endif
#endif

struct BitInsert { int n; u64 lowmask[40]; };
struct PLB_ALIGN(16) TileOp {
    u32 code, cm_tid, cv_tid, pm_tid, umask, upar, slot;
    u64 cmask_o, cval_o, pmask_o;
    T2 m[4];
};
struct PLB_ALIGN(16) LadderEntry { u64 cmask_o; u32 cm_tid, pad; T2 ph; };
struct PLB_ALIGN(16) RoundHdr { int first_op, nops, nlad, pad; u32 w[9]; u32 sroff[32]; };
struct PLB_ALIGN(16) PassHdr { int nrounds, nops_total; u64 ntiles; int nslots, pad; BitInsert tile_ins; };
struct PLB_ALIGN(16) PassParams { PassHdr hdr; RoundHdr rounds[PLB_MAXROUNDS]; TileOp ops[PLB_MAXOPS + 1]; };
struct PLB_ALIGN(16) RouteParams { T2 *dst[8]; u64 lmask, rdep; u32 opos[4]; u32 tq[3]; u32 ntq; };
// Order in which a routed pass walks its tiles.  A swapped bit that lies OUTSIDE the tile selects the destination
// slab per tile; walking the tiles in index order would send to one peer for a long stretch (first all of the
// pass's stores stay local, then all of them cross NVLink to ONE peer, and every rank picks the same peer at the
// same time).  The loop counter's low bits are therefore placed at those tile-index positions: consecutive CTAs
// — the ones resident together — target all 2^k destinations evenly, for the whole duration of the pass.
DEV u64 route_tile(u64 c, const RouteParams &rp) {
    const u32 k = rp.ntq;
    u64 t = c >> k;
    for (u32 i = 0; i < k; i++) { // tq ascending: insert bit i of the counter at tile-index position tq[i]
        const u64 m = (1ull << rp.tq[i]) - 1ull;
        t = ((t & ~m) << 1) | (t & m) | (((c >> i) & 1ull) << rp.tq[i]);
    }
    return t;
}
static_assert(sizeof(TileOp) == PLB_SIZEOF_TILEOP, "TileOp layout");
static_assert(sizeof(PassParams) == PLB_SIZEOF_PASSPARAMS, "PassParams layout");
static_assert(sizeof(PassHdr) + PLB_MAXROUNDS * sizeof(RoundHdr) == PLB_OFFSETOF_OPS, "PassParams layout");

#if !PLB_DOUBLE && !defined(PLB_JIT_HOST) && !PLB_NO_FFMA2
// c64 on the device, opt-in (PLB200_JIT_FFMA2=1): packed FP32 pairs.  sm_100a's FFMA2 / FMUL2 / FADD2 act on an (x, y) register pair with
// free operand modifiers — broadcast of one scalar to both halves, half swap (LO_HI) and per-half negation —
// which is exactly the shape of complex arithmetic: a shear is ONE instruction per amplitude instead of two,
// a phase multiplication two instead of four.
DEV T2 f2(real x, real y) { T2 r; r.x = x; r.y = y; return r; }
DEV void cmul_ip(T2 &v, const T2 d) {
    const T2 t = __fmul2_rn(f2(-v.y, v.x), f2(d.y, d.y));
    v = __ffma2_rn(v, f2(d.x, d.x), t);
}
DEV void cshear_ip(T2 &a, const T2 t, const T2 &b) {
    a = __ffma2_rn(f2(t.x, t.x), b, a);
    a = __ffma2_rn(f2(-t.y, t.y), f2(b.y, b.x), a);
}
DEV void cswap(T2 &a, T2 &b) { const T2 t = a; a = b; b = t; }
DEV void lift_r(T2 &a, T2 &b, const T2 m) {
    a = __ffma2_rn(f2(m.x, m.x), b, a);
    b = __ffma2_rn(f2(m.y, m.y), a, b);
    a = __ffma2_rn(f2(m.x, m.x), b, a);
}
DEV void lift_i(T2 &a, T2 &b, const T2 m) {
    a = __ffma2_rn(f2(-m.x, m.x), f2(b.y, b.x), a);
    b = __ffma2_rn(f2(-m.y, m.y), f2(a.y, a.x), b);
    a = __ffma2_rn(f2(-m.x, m.x), f2(b.y, b.x), a);
}
DEV void had(T2 &a, T2 &b) {
    a = __fadd2_rn(a, b);
    b = __ffma2_rn(f2((real)-2, (real)-2), b, a);
}
DEV void lu_r(T2 &a, T2 &b, const T2 m0, const T2 m1) {
    a = __ffma2_rn(f2(m0.x, m0.x), b, a);
    b = __ffma2_rn(f2(m0.y, m0.y), a, b);
    a = __fmul2_rn(a, f2(m1.x, m1.x));
    b = __fmul2_rn(b, f2(m1.y, m1.y));
}
// scaled rotations: TWO packed FMAs per pair
DEV void srot_r_t(T2 &a, T2 &b, const real t) {
    const T2 a0 = a, b0 = b;
    a = __ffma2_rn(f2(-t, -t), b0, a0);
    b = __ffma2_rn(f2(t, t), a0, b0);
}
DEV void srot_r_k(T2 &a, T2 &b, const real k) {
    const T2 a0 = a, b0 = b;
    a = __ffma2_rn(f2(k, k), a0, f2(-b0.x, -b0.y));
    b = __ffma2_rn(f2(k, k), b0, a0);
}
DEV void srot_i_t(T2 &a, T2 &b, const real t) {
    const T2 a0 = a, b0 = b;
    a = __ffma2_rn(f2(t, t), f2(b0.y, -b0.x), a0);
    b = __ffma2_rn(f2(t, t), f2(a0.y, -a0.x), b0);
}
DEV void srot_i_k(T2 &a, T2 &b, const real k) {
    const T2 a0 = a, b0 = b;
    a = __ffma2_rn(f2(k, k), f2(-a0.y, a0.x), b0);
    b = __ffma2_rn(f2(k, k), f2(-b0.y, b0.x), a0);
}
#else
DEV void cmul_ip(T2 &v, const T2 d) {
    const real t = v.x * d.y;
    v.x = v.x * d.x;
    v.x = fma(-v.y, d.y, v.x);
    v.y = fma(v.y, d.x, t);
}
DEV void cshear_ip(T2 &a, const T2 t, const T2 &b) {
    a.x = fma(t.x, b.x, a.x);
    a.x = fma(-t.y, b.y, a.x);
    a.y = fma(t.x, b.y, a.y);
    a.y = fma(t.y, b.x, a.y);
}
DEV void cswap(T2 &a, T2 &b) { const T2 t = a; a = b; b = t; }
DEV void lift_r(T2 &a, T2 &b, const T2 m) {
    a.x = fma(m.x, b.x, a.x), a.y = fma(m.x, b.y, a.y);
    b.x = fma(m.y, a.x, b.x), b.y = fma(m.y, a.y, b.y);
    a.x = fma(m.x, b.x, a.x), a.y = fma(m.x, b.y, a.y);
}
DEV void lift_i(T2 &a, T2 &b, const T2 m) {
    a.x = fma(-m.x, b.y, a.x), a.y = fma(m.x, b.x, a.y);
    b.x = fma(-m.y, a.y, b.x), b.y = fma(m.y, a.x, b.y);
    a.x = fma(-m.x, b.y, a.x), a.y = fma(m.x, b.x, a.y);
}
DEV void had(T2 &a, T2 &b) {
    a.x = a.x + b.x, a.y = a.y + b.y;
    b.x = fma((real)-2, b.x, a.x), b.y = fma((real)-2, b.y, a.y);
}
DEV void lu_r(T2 &a, T2 &b, const T2 m0, const T2 m1) {
    a.x = fma(m0.x, b.x, a.x), a.y = fma(m0.x, b.y, a.y);
    b.x = fma(m0.y, a.x, b.x), b.y = fma(m0.y, a.y, b.y);
    a.x = a.x * m1.x, a.y = a.y * m1.x;
    b.x = b.x * m1.y, b.y = b.y * m1.y;
}
// scaled rotations (K_SROT_R / K_SROT_I): 4 FMAs per pair, the cosine / sine is carried by the host
DEV void srot_r_t(T2 &a, T2 &b, const real t) { // a' = a - t b, b' = b + t a
    const T2 a0 = a, b0 = b;
    a.x = fma(-t, b0.x, a0.x), a.y = fma(-t, b0.y, a0.y);
    b.x = fma(t, a0.x, b0.x), b.y = fma(t, a0.y, b0.y);
}
DEV void srot_r_k(T2 &a, T2 &b, const real k) { // a' = k a - b, b' = a + k b
    const T2 a0 = a, b0 = b;
    a.x = fma(k, a0.x, -b0.x), a.y = fma(k, a0.y, -b0.y);
    b.x = fma(k, b0.x, a0.x), b.y = fma(k, b0.y, a0.y);
}
DEV void srot_i_t(T2 &a, T2 &b, const real t) { // a' = a - i t b, b' = b - i t a
    const T2 a0 = a, b0 = b;
    a.x = fma(t, b0.y, a0.x), a.y = fma(-t, b0.x, a0.y);
    b.x = fma(t, a0.y, b0.x), b.y = fma(-t, a0.x, b0.y);
}
DEV void srot_i_k(T2 &a, T2 &b, const real k) { // a' = i k a + b, b' = a + i k b
    const T2 a0 = a, b0 = b;
    a.x = fma(-k, a0.y, b0.x), a.y = fma(k, a0.x, b0.y);
    b.x = fma(-k, b0.y, a0.x), b.y = fma(k, b0.x, a0.y);
}
#endif
DEV void lu_c(T2 &a, T2 &b, const T2 m0, const T2 m1, const T2 m2, const T2 m3) {
    cshear_ip(a, m0, b);
    cshear_ip(b, m1, a);
    cmul_ip(a, m2);
    cmul_ip(b, m3);
}
// adjoint (two-state) passes: Im / Re of conj(a) b in double, and the per-CTA overlap accumulators
DEV double im_cb(const T2 a, const T2 b) { return (double)a.x * (double)b.y - (double)a.y * (double)b.x; }
DEV double re_cb(const T2 a, const T2 b) { return (double)a.x * (double)b.x + (double)a.y * (double)b.y; }
#if defined(PLB_JIT_HOST)
DEV void ovl_reduce(double *acc, int slot, double s) { acc[slot] += s; }
#else
DEV void ovl_reduce(double *acc, int slot, double s) { // acc: this WARP's row of the CTA's accumulators: fixed order, no atomics
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) acc[slot] += s;
}
#endif
DEV u64 insert_bits_m(u64 x, const BitInsert &bi) {
#pragma unroll
    for (int i = 0; i < PLB_M; i++) {
        const u64 lm = bi.lowmask[i];
        x = ((x & ~lm) << 1) | (x & lm);
    }
    return x;
}
DEV u64 tile_line_offset(const PassHdr &h, int i) {
    u64 o = 0;
#pragma unroll
    for (int b = PLB_LOW; b < PLB_M; b++)
        if ((i >> (b - PLB_LOW)) & 1) o |= h.tile_ins.lowmask[b] + 1;
    return o;
}
// shared-memory swizzle of tile_exec.cuh (Swz<T2>): element j lives at j ^ g(j)
DEV constexpr u32 swz(u32 j) {
    u32 g = 0;
    for (int i = PLB_SWZ_B; i < 13; i++)
        if ((j >> i) & 1u) g ^= (u32)((PLB_SWZ_COLS >> (4 * i)) & 15ull);
    return j ^ g;
}
#define PLB_NV (1 << PLB_R)
DEV void load_tile(u32 tid, u64 base, const u64 *goff, const T2 *__restrict__ sv, unsigned char *smem) {
    const u32 st = swz(tid) * (u32)sizeof(T2);
    T2 v[PLB_NV];
#pragma unroll
    for (int u = 0; u < PLB_NV; u++) {
        const u32 j = tid + u * PLB_NT;
        v[u] = sv[base | goff[j >> PLB_LOW] | (j & ((1u << PLB_LOW) - 1))];
    }
#pragma unroll
    for (int u = 0; u < PLB_NV; u++) *(T2 *)(smem + (st ^ (swz(u * PLB_NT) * (u32)sizeof(T2)))) = v[u];
}
DEV void store_tile(u32 tid, u64 base, const u64 *goff, T2 *__restrict__ sv, const unsigned char *smem) {
    const u32 st = swz(tid) * (u32)sizeof(T2);
    T2 v[PLB_NV];
#pragma unroll
    for (int u = 0; u < PLB_NV; u++) v[u] = *(const T2 *)(smem + (st ^ (swz(u * PLB_NT) * (u32)sizeof(T2))));
#pragma unroll
    for (int u = 0; u < PLB_NV; u++) {
        const u32 j = tid + u * PLB_NT;
        sv[base | goff[j >> PLB_LOW] | (j & ((1u << PLB_LOW) - 1))] = v[u];
    }
}
)PLB";
}

// Routed pass ("swap-out"): the LAST round of the pass stores every amplitude to the slab and position it has
// AFTER an exchange of k local index bits with k global (rank) bits — its own ping-pong slab when the local
// bits already equal this rank's global bits, otherwise straight into a peer's slab over NVLink.  The index-bit
// swap of the sharded mode then costs no sweep of its own: the transfer rides on the pass's store phase.
// loc[i]: tile-local position of swapped local bit i, or -1 when the bit lies outside the pass's tile.
struct Route {
    int k = 0;
    int loc[3] = {-1, -1, -1};
};
// run-time half of a route (second kernel argument of a routed pass)
template <typename T2> struct alignas(16) RouteParams {
    T2 *dst[8];        // destination slab by the value of the swapped LOCAL bits (bit i <-> swapped bit i)
    uint64_t lmask;    // the swapped local bits, as a mask over the slab index
    uint64_t rdep;     // this rank's values of the swapped global bits, deposited at those positions
    uint32_t opos[4];  // index-bit position of swapped bit i (used when it lies outside the tile)
    uint32_t tq[3];    // tile-index positions of the swapped bits outside the tile, ascending (route_tile)
    uint32_t ntq;
};

template <typename T2, class Cfg> class Gen {
    static constexpr int M = Cfg::M, LOW = Cfg::LOW, R = Cfg::R, NV = 1 << R, NTB = M - R, NS = Cfg::NS;
    char cur_set = 'v'; // register set statements are emitted for: 'v' (lambda / the state), 'h' (H lambda, NS == 2)
    const PassParams<T2> &pp;
    std::string s;
    int perm[NV]; // logical register u lives in variable v<perm[u]>

    void add(const char *fmt, ...) __attribute__((format(printf, 2, 3))) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        s += buf;
    }
    static std::string hex(uint32_t v) {
        char b[16];
        snprintf(b, sizeof(b), "0x%xu", v);
        return b;
    }
    std::string V(int u) const { return std::string(1, cur_set) + std::to_string(perm[u]); }
    std::string Vs(char set, int u) const { return std::string(1, set) + std::to_string(perm[u]); }

    // condition of op K: tile-uniform part (outside controls) and thread part
    std::string cond_of(const TileOp<T2> &op, int K) const {
        std::string c;
        if (op.cmask_o) c += "((base & pp.ops[" + std::to_string(K) + "].cmask_o) == pp.ops[" + std::to_string(K) + "].cval_o)";
        if (op.cm_tid) {
            if (!c.empty()) c += " && ";
            c += "((tid & " + hex(op.cm_tid) + ") == " + hex(op.cv_tid) + ")";
        }
        return c;
    }
    // parity expression of op K, or "" when it has none
    std::string par_of(const TileOp<T2> &op, int K) const {
        std::string p;
        if (op.pm_tid) p += "POPC(tid & " + hex(op.pm_tid) + ")";
        if (op.pmask_o) {
            if (!p.empty()) p += " + ";
            p += "POPCLL(base & pp.ops[" + std::to_string(K) + "].pmask_o)";
        }
        if (p.empty()) return p;
        return "(((" + p + ") & 1) != 0)";
    }

    // local tile-bit positions of the thread bits / register bits of a round, recovered from the encoded
    // shared-memory offsets (the swizzle only touches the bank-group bits, so the top set bit is the position)
    static int top_bit(uint32_t v) { return 31 - __builtin_clz(v); }
    void round_bits(const RoundHdr &rh, int tpos[NTB], int rl[R]) const {
        for (int i = 0; i < NTB; i++) tpos[i] = top_bit(rh.w[i] / static_cast<uint32_t>(sizeof(T2)));
        for (int i = 0; i < R; i++) rl[i] = top_bit(rh.sroff[1 << i] / static_cast<uint32_t>(sizeof(T2)));
    }
    // A round can exchange its registers with GLOBAL memory directly (no staging through shared memory, no
    // barrier) when the lowest LOW lane bits are the tile's LOW contiguous bits: every 2^LOW lanes then cover one
    // full 128-byte line per access.  The scheduler arranges that for the first and the last round of a pass
    // whenever none of those bits is a register bit (fusion.cu, thread-bit assignment).
    bool direct_ok(int r) const {
        if (NS == 2 || std::getenv("PLB200_JIT_NO_DIRECT")) return false;
        int tpos[NTB], rl[R];
        round_bits(round_hdr(r), tpos, rl);
        for (int i = 0; i < LOW; i++)
            if (tpos[i] != i) return false;
        return true;
    }

    // records that follow op K and belong to it (ladder entries, rows 1-3 of a dense 4x4)
    int extra_records(int K) const {
        const TileOp<T2> &o = pp.ops[K];
        const int kind = code_kind(o.code);
        if (kind == K_LADDER) return ladder_records<T2>(static_cast<int>(o.slot)) - 1;
        if (kind == K_DENSE2) return 3;
        return 0;
    }

    // from_global / to_global: the round's registers come from / go to the state vector instead of the tile
    // round r of the pass; r == nrounds: the op-less TRANSFER round a routed pass appends when its last round is
    // not line-coalesced (register bits = the highest tile bits; it replaces the store phase at the same cost)
    RoundHdr transfer_round;
    const RoundHdr &round_hdr(int r) const { return r < pp.hdr.nrounds ? pp.rounds[r] : transfer_round; }
    void make_transfer_round() {
        std::memset(&transfer_round, 0, sizeof(transfer_round));
        int rl[R], tpos[NTB];
        for (int i = 0; i < R; i++) rl[i] = M - R + i;
        for (int i = 0; i < NTB; i++) tpos[i] = i;
        for (int i = 0; i < NTB; i++) transfer_round.w[i] = swz<T2>(1u << tpos[i]) * static_cast<uint32_t>(sizeof(T2));
        for (int u = 0; u < NV; u++) {
            uint32_t o = 0;
            for (int i = 0; i < R; i++)
                if (u >> i & 1) o |= 1u << rl[i];
            transfer_round.sroff[u] = swz<T2>(o) * static_cast<uint32_t>(sizeof(T2));
        }
    }

    void gen_round(int r, bool from_global, bool to_global) {
        const RoundHdr &rh = round_hdr(r);
        int tpos[NTB], rl[R];
        round_bits(rh, tpos, rl);
        const bool routed = to_global && route.k > 0;
        if (NS == 2)
            add("DEV void round_%d(const PassParams &pp, const u32 tid, const u64 base, unsigned char *smem, unsigned char *smem1, double *acc) {\n", r);
        else
            add("DEV void round_%d(const PassParams &pp, const u32 tid, const u64 base, unsigned char *smem, T2 *__restrict__ sv%s) {\n", r,
                routed ? ", const RouteParams &rp" : "");
        const uint32_t lowmask = ((1u << Swz<T2>::B) - 1u) * static_cast<uint32_t>(sizeof(T2));
        std::vector<uint32_t> lows;
        auto low_id = [&](uint32_t low) {
            for (size_t i = 0; i < lows.size(); i++)
                if (lows[i] == low) return static_cast<int>(i);
            lows.push_back(low);
            return static_cast<int>(lows.size() - 1);
        };
        if (!from_global || !to_global) {
            // thread part of the swizzled offset; the swizzle only mixes into the bank-group bits, everything
            // above is a plain (immediate) offset
            s += "    const u32 sb = 0u";
            for (int i = 0; i < NTB; i++) add(" ^ ((tid >> %d & 1u) * %s)", i, hex(rh.w[i]).c_str());
            s += ";\n";
            for (int u = 0; u < NV; u++) low_id(rh.sroff[u] & lowmask);
            for (size_t i = 0; i < lows.size(); i++) {
                add("    unsigned char *const b%zu = smem + (sb ^ %s);\n", i, hex(lows[i]).c_str());
                if (NS == 2) add("    unsigned char *const c%zu = smem1 + (sb ^ %s);\n", i, hex(lows[i]).c_str());
            }
        }
        if (from_global || to_global) {
            // element offset of this thread inside the tile (global index bits) and of each register
            s += "    const u64 goff_t = 0ull";
            for (int i = 0; i < NTB; i++) add(" | ((tid >> %d & 1u) ? pp.hdr.tile_ins.lowmask[%d] + 1ull : 0ull)", i, tpos[i]);
            s += ";\n    T2 *__restrict__ const gp = sv + (base | goff_t);\n";
            for (int i = 0; i < R; i++) add("    const u64 ro%d = pp.hdr.tile_ins.lowmask[%d] + 1ull;\n", i, rl[i]);
        }
        auto greg = [&](int u) {
            std::string e;
            for (int i = 0; i < R; i++)
                if ((u >> i) & 1) e += (e.empty() ? "ro" : " | ro") + std::to_string(i);
            return e.empty() ? std::string("0ull") : "(" + e + ")";
        };
        for (int u = 0; u < NV; u++) {
            perm[u] = u;
            if (from_global) add("    T2 v%d = gp[%s];\n", u, greg(u).c_str());
            else
                add("    T2 v%d = *(const T2 *)(b%d + %s);\n", u, low_id(rh.sroff[u] & lowmask), hex(rh.sroff[u] & ~lowmask).c_str());
            if (NS == 2)
                add("    T2 h%d = *(const T2 *)(c%d + %s);\n", u, low_id(rh.sroff[u] & lowmask), hex(rh.sroff[u] & ~lowmask).c_str());
        }
        // thread-scalar factors of this round (K_DIAG_T / K_DIAG1_T, ladders on no register bit): merged into ONE
        // scalar per thread, applied once at the end of the round (it commutes with every op), when there are
        // >= 2 of them or a ladder among them.  Ladder records may sit among the regular ops (the encoder's
        // "jit forms") or in the tail section; their entry records are skipped here.
        int n_tscalar = 0;
        bool r_ladder = false;
        for (int k = 0; k < rh.nops + rh.nlad; k++) {
            const TileOp<T2> &o = pp.ops[rh.first_op + k];
            const int kind = code_kind(o.code);
            if (kind == K_LADDER && code_p(o.code) >= R) r_ladder = true;
            if (kind == K_DIAG_T || kind == K_DIAG1_T) n_tscalar++;
            k += extra_records(rh.first_op + k);
        }
        const bool merge_ts = n_tscalar >= 2 || r_ladder;
        if (merge_ts) s += "    T2 ts; ts.x = (real)1; ts.y = (real)0; bool ts_any = false;\n";

        for (int k = 0; k < rh.nops + rh.nlad; k++) {
            const int K = rh.first_op + k;
            if (code_kind(pp.ops[K].code) == K_LADDER) {
                k += gen_ladder(K) - 1;
                continue;
            }
            if (pp.ops[K].code & F_OVL) {
                gen_overlap(K);
                continue;
            }
            if (NS == 2) { // H lambda first, as the interpreter does; renamings / the thread scalar happen once
                cur_set = 'h';
                gen_op(K, merge_ts, false);
                cur_set = 'v';
            }
            gen_op(K, merge_ts, true);
            k += extra_records(K);
        }

        if (merge_ts) {
            s += "    if (ts_any) {\n";
            for (int u = 0; u < NV; u++) {
                add("        cmul_ip(%s, ts);\n", Vs('v', u).c_str());
                if (NS == 2) add("        cmul_ip(%s, ts);\n", Vs('h', u).c_str());
            }
            s += "    }\n";
        }
        // routed stores: swapped bit i of this amplitude selects the destination slab; inside the slab the
        // swapped positions take this rank's global-bit values
        int reg_of_swapped[3] = {-1, -1, -1};
        if (routed) {
            s += "    const u32 sel_t = 0u";
            for (int i = 0; i < route.k; i++) {
                const int lp = route.loc[i];
                if (lp < 0) add(" | ((u32)((base >> rp.opos[%d]) & 1ull) << %d)", i, i);
                else {
                    int ti = -1;
                    for (int j = 0; j < NTB; j++)
                        if (tpos[j] == lp) ti = j;
                    for (int j = 0; j < R; j++)
                        if (rl[j] == lp) reg_of_swapped[i] = j;
                    if (ti >= 0) add(" | ((tid >> %d & 1u) << %d)", ti, i);
                }
            }
            s += ";\n    const u64 gb = ((base | goff_t) & ~rp.lmask) | rp.rdep;\n";
            unsigned done = 0;
            for (int u = 0; u < NV; u++) {
                unsigned x = 0;
                for (int i = 0; i < route.k; i++)
                    if (reg_of_swapped[i] >= 0 && ((u >> reg_of_swapped[i]) & 1)) x |= 1u << i;
                if (done >> x & 1u) continue;
                done |= 1u << x;
                add("    T2 *__restrict__ const d%u = rp.dst[sel_t | %uu] + gb;\n", x, x);
            }
        }
        auto greg_routed = [&](int u) { // register part of the index without the swapped positions
            std::string e;
            for (int i = 0; i < R; i++) {
                bool swapped = false;
                for (int q = 0; q < route.k; q++) swapped = swapped || reg_of_swapped[q] == i;
                if (((u >> i) & 1) && !swapped) e += (e.empty() ? "ro" : " | ro") + std::to_string(i);
            }
            return e.empty() ? std::string("0ull") : "(" + e + ")";
        };
        for (int u = 0; u < NV; u++) {
            if (routed) {
                unsigned x = 0;
                for (int i = 0; i < route.k; i++)
                    if (reg_of_swapped[i] >= 0 && ((u >> reg_of_swapped[i]) & 1)) x |= 1u << i;
                add("    d%u[%s] = %s;\n", x, greg_routed(u).c_str(), V(u).c_str());
            } else if (to_global) add("    gp[%s] = %s;\n", greg(u).c_str(), V(u).c_str());
            else
                add("    *(T2 *)(b%d + %s) = %s;\n", low_id(rh.sroff[u] & lowmask), hex(rh.sroff[u] & ~lowmask).c_str(), V(u).c_str());
            if (NS == 2)
                add("    *(T2 *)(c%d + %s) = %s;\n", low_id(rh.sroff[u] & lowmask), hex(rh.sroff[u] & ~lowmask).c_str(), Vs('h', u).c_str());
        }
        s += "}\n";
    }

    // adjoint: thread-local contribution to Im<H lambda| G |lambda> of one overlap op, reduced per warp into the
    // CTA's accumulator (tile_exec.cuh overlap_op); l = the 'v' registers, h = the 'h' registers
    void gen_overlap(int K) {
        const TileOp<T2> &op = pp.ops[K];
        const int kind = code_kind(op.code), p = code_p(op.code);
        const std::string O = "pp.ops[" + std::to_string(K) + "]";
        add("    { // overlap op %d kind %d P %d slot %u\n", K, kind, p, op.slot);
        const bool out_cond = (op.code & F_COND) && op.cmask_o;
        if (out_cond) s += "        if ((base & " + O + ".cmask_o) == " + O + ".cval_o) {\n";
        s += "        double ovl = 0.0;\n";
        const bool thr_cond = (op.code & F_COND) && op.cm_tid;
        if (thr_cond) s += "        if ((tid & " + hex(op.cm_tid) + ") == " + hex(op.cv_tid) + ") {\n";
        if (kind == K_OVL_D) {
            const std::string par = (op.code & F_PAR) ? par_of(op, K) : std::string();
            add("        const double g0 = (double)%s.m[0].x, g1 = (double)%s.m[0].y;\n", O.c_str(), O.c_str());
            if (par.empty()) s += "        const double gE = g0, gO = g1;\n";
            else s += "        const bool pt = " + par + ";\n        const double gE = pt ? g1 : g0, gO = pt ? g0 : g1;\n";
            for (int u = 0; u < NV; u++)
                if ((op.umask >> u) & 1u)
                    add("        ovl += %s * im_cb(%s, %s);\n", ((op.upar >> u) & 1u) ? "gO" : "gE", Vs('h', u).c_str(), Vs('v', u).c_str());
        } else {
            std::vector<std::pair<int, int>> pr;
            pairs_on(p, pr);
            for (auto [u0, u1] : pr) {
                if (!((op.umask >> u0) & 1u)) continue;
                if (kind == K_OVL_Y)
                    add("        ovl += re_cb(%s, %s) - re_cb(%s, %s);\n", Vs('h', u1).c_str(), Vs('v', u0).c_str(), Vs('h', u0).c_str(), Vs('v', u1).c_str());
                else
                    add("        ovl += im_cb(%s, %s) + im_cb(%s, %s);\n", Vs('h', u0).c_str(), Vs('v', u1).c_str(), Vs('h', u1).c_str(), Vs('v', u0).c_str());
            }
        }
        if (thr_cond) s += "        }\n";
        add("        ovl_reduce(acc, %u, ovl);\n", op.slot);
        if (out_cond) s += "        }\n";
        s += "    }\n";
    }

    // pairs (u0, u1) on register bit P
    static void pairs_on(int P, std::vector<std::pair<int, int>> &out) {
        out.clear();
        for (int q = 0; q < (1 << (R - 1)); q++) {
            const int u0 = ((q >> P) << (P + 1)) | (q & ((1 << P) - 1));
            out.emplace_back(u0, u0 | (1 << P));
        }
    }

    void gen_op(int K, bool merge_ts, bool last) {
        const TileOp<T2> &op = pp.ops[K];
        const uint32_t code = op.code;
        const int kind = code_kind(code), P = code_p(code), C = static_cast<int>((code >> 16) & 7u);
        const std::string cond = (code & F_COND) ? cond_of(op, K) : std::string();
        const std::string par = (code & F_PAR) ? par_of(op, K) : std::string();
        const std::string O = "pp.ops[" + std::to_string(K) + "]";
        std::vector<std::pair<int, int>> pr;
        add("    { // op %d kind %d P %d C %d\n", K, kind, P, C);
        auto open_cond = [&]() {
            if (!cond.empty()) s += "        if (" + cond + ") {\n";
        };
        auto close_cond = [&]() {
            if (!cond.empty()) s += "        }\n";
        };
        auto decl_m = [&](int n) {
            for (int q = 0; q < n; q++) add("        const T2 m%d = %s.m[%d];\n", q, O.c_str(), q);
        };
        // phases of a two-valued diagonal: dA (parity 0) / dB (parity 1) after the thread / outside parity
        auto decl_dab = [&]() {
            decl_m(2);
            if (par.empty()) s += "        const T2 dA = m0, dB = m1;\n";
            else {
                s += "        const bool pt = " + par + ";\n";
                s += "        const T2 dA = pt ? m1 : m0, dB = pt ? m0 : m1;\n";
            }
        };
        switch (kind) {
        case K_LIFT_R: case K_LIFT_I: case K_HAD: case K_LU_R: case K_LU_C:
        case K_LIFT_R_M: case K_LIFT_I_M: case K_LU_R_M: case K_LU_C_M: {
            const bool masked = kind >= K_LIFT_R_M;
            const int base_kind = kind == K_LIFT_R_M ? K_LIFT_R : kind == K_LIFT_I_M ? K_LIFT_I
                                  : kind == K_LU_R_M ? K_LU_R : kind == K_LU_C_M ? K_LU_C : kind;
            decl_m(base_kind == K_LU_C ? 4 : base_kind == K_LU_R ? 2 : base_kind == K_HAD ? 0 : 1);
            open_cond();
            pairs_on(P, pr);
            for (auto [u0, u1] : pr) {
                if (masked && !((op.umask >> u0) & 1u)) continue;
                const std::string a = V(u0), b = V(u1);
                switch (base_kind) {
                case K_LIFT_R: add("            lift_r(%s, %s, m0);\n", a.c_str(), b.c_str()); break;
                case K_LIFT_I: add("            lift_i(%s, %s, m0);\n", a.c_str(), b.c_str()); break;
                case K_HAD: add("            had(%s, %s);\n", a.c_str(), b.c_str()); break;
                case K_LU_R: add("            lu_r(%s, %s, m0, m1);\n", a.c_str(), b.c_str()); break;
                default: add("            lu_c(%s, %s, m0, m1, m2, m3);\n", a.c_str(), b.c_str()); break;
                }
            }
            close_cond();
        } break;
        case K_SROT_R: case K_SROT_I: case K_SROK_R: case K_SROK_I: { // scaled rotations: tangent / cotangent form
            decl_m(1);
            open_cond();
            pairs_on(P, pr);
            const char *fn = kind == K_SROT_R ? "srot_r_t" : kind == K_SROT_I ? "srot_i_t" : kind == K_SROK_R ? "srot_r_k" : "srot_i_k";
            for (auto [u0, u1] : pr) add("            %s(%s, %s, m0.x);\n", fn, V(u0).c_str(), V(u1).c_str());
            close_cond();
        } break;
        case K_DENSE2: { // out[j] = sum_c M[j][c] in[c] on the register quadruples over bits P (matrix bit 0), C (bit 1)
            open_cond();
            for (int u = 0; u < NV; u++) {
                if (((u >> P) & 1) || ((u >> C) & 1) || !((op.umask >> u) & 1u)) continue;
                const int q[4] = {u, u | (1 << P), u | (1 << C), u | (1 << P) | (1 << C)};
                s += "            {\n";
                for (int c = 0; c < 4; c++) add("                const T2 o%d = %s;\n", c, V(q[c]).c_str());
                for (int j = 0; j < 4; j++) {
                    add("                T2 n%d = o0; cmul_ip(n%d, pp.ops[%d].m[0]);\n", j, j, K + j);
                    for (int c = 1; c < 4; c++) add("                cshear_ip(n%d, pp.ops[%d].m[%d], o%d);\n", j, K + j, c, c);
                }
                for (int j = 0; j < 4; j++) add("                %s = n%d;\n", V(q[j]).c_str(), j);
                s += "            }\n";
            }
            close_cond();
        } break;
        case K_SWAP: case K_SWAP_M: case K_SWAP_CR: case K_SWAP2: case K_SWAP2_M: {
            pr.clear();
            if (kind == K_SWAP2 || kind == K_SWAP2_M) {
                for (int u = 0; u < NV; u++) {
                    if (((u >> P) & 1) || ((u >> C) & 1)) continue;
                    const int ua = u | (1 << P), ub = u | (1 << C);
                    if (kind == K_SWAP2_M && !((op.umask >> ua) & 1u)) continue;
                    pr.emplace_back(ua, ub);
                }
            } else {
                std::vector<std::pair<int, int>> all;
                pairs_on(P, all);
                for (auto [u0, u1] : all) {
                    if (kind == K_SWAP_M && !((op.umask >> u0) & 1u)) continue;
                    if (kind == K_SWAP_CR && !((u0 >> C) & 1)) continue;
                    pr.emplace_back(u0, u1);
                }
            }
            if (cond.empty()) { // a renaming: no instructions (once for both register sets)
                if (last)
                    for (auto [a, b] : pr) std::swap(perm[a], perm[b]);
                s += "        // register renaming\n";
            } else {
                open_cond();
                for (auto [a, b] : pr) add("            cswap(%s, %s);\n", V(a).c_str(), V(b).c_str());
                close_cond();
            }
        } break;
        case K_PAIR2: case K_PAIR4: { // one 2x2 block of a two- / four-bit pair op on the registers whose target bits = a and = b
            const bool four = kind == K_PAIR4;
            const unsigned a2 = four ? (op.slot & 15u) : (op.slot & 3u), b2 = four ? ((op.slot >> 4) & 15u) : ((op.slot >> 2) & 3u);
            const int form = static_cast<int>((op.slot >> (four ? 8 : 4)) & 15u);
            int pos[4] = {P, C, 0, 0};
            if (four)
                for (int j = 0; j < 4; j++) pos[j] = static_cast<int>((op.slot >> (12 + 3 * j)) & 7u);
            const int nb = four ? 4 : 2;
            int tmask = 0;
            for (int j = 0; j < nb; j++) tmask |= 1 << pos[j];
            auto dep = [&](unsigned x) {
                int r = 0;
                for (int j = 0; j < nb; j++) r |= static_cast<int>((x >> j) & 1u) << pos[j];
                return r;
            };
            pr.clear();
            for (int u = 0; u < NV; u++) {
                if (u & tmask) continue;
                const int ua = u | dep(a2), ub = u | dep(b2);
                if (!((op.umask >> ua) & 1u)) continue;
                pr.emplace_back(ua, ub);
            }
            if (form == K_SWAP) {
                if (cond.empty()) {
                    if (last)
                        for (auto [x, y] : pr) std::swap(perm[x], perm[y]);
                    s += "        // register renaming\n";
                } else {
                    open_cond();
                    for (auto [x, y] : pr) add("            cswap(%s, %s);\n", V(x).c_str(), V(y).c_str());
                    close_cond();
                }
                break;
            }
            decl_m(form == K_LU_C ? 4 : form == K_LU_R ? 2 : 1);
            open_cond();
            for (auto [x, y] : pr) {
                const std::string va = V(x), vb = V(y);
                switch (form) {
                case K_LIFT_R: add("            lift_r(%s, %s, m0);\n", va.c_str(), vb.c_str()); break;
                case K_LIFT_I: add("            lift_i(%s, %s, m0);\n", va.c_str(), vb.c_str()); break;
                case K_LU_R: add("            lu_r(%s, %s, m0, m1);\n", va.c_str(), vb.c_str()); break;
                case K_LU_C: add("            lu_c(%s, %s, m0, m1, m2, m3);\n", va.c_str(), vb.c_str()); break;
                default: ok = false; break;
                }
            }
            close_cond();
        } break;
        case K_DIAG_R: case K_DIAG_PP: case K_DIAG_CR: {
            decl_dab();
            open_cond();
            for (int u = 0; u < NV; u++) {
                if (kind == K_DIAG_CR && !((u >> C) & 1)) continue;
                const bool odd = kind == K_DIAG_PP ? ((((u >> P) ^ (u >> C)) & 1) != 0) : (((u >> P) & 1) != 0);
                add("            cmul_ip(%s, %s);\n", V(u).c_str(), odd ? "dB" : "dA");
            }
            close_cond();
        } break;
        case K_DIAG1_R: {
            decl_m(1);
            open_cond();
            for (int u = 0; u < NV; u++)
                if ((u >> P) & 1) add("            cmul_ip(%s, m0);\n", V(u).c_str());
            close_cond();
        } break;
        case K_DIAG_T: case K_DIAG1_T: case K_DIAG_CT: {
            // one scalar for all registers (K_DIAG_CT: for the registers whose bit P is set)
            if (kind == K_DIAG1_T) {
                decl_m(1);
                if (par.empty()) break; // parity 0 everywhere: identity
                s += "        const bool pt = " + par + ";\n        const T2 d = m0;\n";
            } else {
                decl_m(2);
                if (par.empty()) s += "        const T2 d = m0;\n";
                else s += "        const bool pt = " + par + ";\n        const T2 d = pt ? m1 : m0;\n";
            }
            std::string c2 = cond;
            if (kind == K_DIAG1_T) c2 = c2.empty() ? "pt" : "(" + c2 + ") && pt";
            // the pass's slot for the host-carried scalar: holds 1 unless the scalar is folded back here
            if (kind == K_DIAG_T && cond.empty() && par.empty()) c2 = "m0.x != (real)1 || m0.y != (real)0";
            if (!c2.empty()) s += "        if (" + c2 + ") {\n";
            if (kind != K_DIAG_CT && merge_ts) {
                if (last) s += "            cmul_ip(ts, d); ts_any = true;\n";
            }
            else
                for (int u = 0; u < NV; u++) {
                    if (kind == K_DIAG_CT && !((u >> P) & 1)) continue;
                    add("            cmul_ip(%s, d);\n", V(u).c_str());
                }
            if (!c2.empty()) s += "        }\n";
        } break;
        case K_DIAG_G: {
            decl_dab(); // dA: parity bit 0 after the thread parity, dB: 1
            open_cond();
            for (int u = 0; u < NV; u++) {
                if (!((op.umask >> u) & 1u)) continue;
                add("            cmul_ip(%s, %s);\n", V(u).c_str(), ((op.upar >> u) & 1u) ? "dB" : "dA");
            }
            close_cond();
        } break;
        default: ok = false; break;
        }
        s += "    }\n";
    }

    // One ladder (header record K + packed entries): the thread multiplies the phases of its active entries into
    // one scalar and applies it to the registers whose bit p is set; p == R (no register bit): the scalar joins
    // the round's thread scalar `ts`.  Returns the number of records.
    int gen_ladder(int K) {
        const TileOp<T2> &hd = pp.ops[K];
        const int n = static_cast<int>(hd.slot), p = code_p(hd.code);
        const LadderEntry<T2> *en = reinterpret_cast<const LadderEntry<T2> *>(&pp.ops[K + 1]);
        const bool to_ts = p >= R;
        add("    { // ladder of %d controlled phases on register bit %d\n", n, p);
        add("        const LadderEntry *en = (const LadderEntry *)&pp.ops[%d];\n", K + 1);
        auto cond_of_entry = [&](int e) {
            std::string c;
            if (en[e].cmask_o) c = "((base & en[" + std::to_string(e) + "].cmask_o) == en[" + std::to_string(e) + "].cmask_o)";
            if (en[e].cm_tid) {
                if (!c.empty()) c += " && ";
                c += "((tid & " + hex(en[e].cm_tid) + ") == " + hex(en[e].cm_tid) + ")";
            }
            return c;
        };
        if (to_ts) {
            for (int e = 0; e < n; e++) {
                const std::string c = cond_of_entry(e);
                if (c.empty()) add("        cmul_ip(ts, en[%d].ph); ts_any = true;\n", e);
                else add("        if (%s) { cmul_ip(ts, en[%d].ph); ts_any = true; }\n", c.c_str(), e);
            }
            s += "    }\n";
            return ladder_records<T2>(n);
        }
        // unconditional entries first: the scalar starts as their product and is always applied
        bool have_t = false;
        for (int e = 0; e < n; e++) {
            if (!cond_of_entry(e).empty()) continue;
            if (!have_t) add("        T2 t = en[%d].ph;\n", e);
            else add("        cmul_ip(t, en[%d].ph);\n", e);
            have_t = true;
        }
        const bool always = have_t;
        if (!have_t) s += "        T2 t; t.x = (real)1; t.y = (real)0; bool any = false;\n";
        for (int e = 0; e < n; e++) {
            const std::string c = cond_of_entry(e);
            if (c.empty()) continue;
            add("        if (%s) { cmul_ip(t, en[%d].ph);%s }\n", c.c_str(), e, always ? "" : " any = true;");
        }
        s += always ? "        {\n" : "        if (any) {\n";
        for (int u = 0; u < NV; u++)
            if ((u >> p) & 1) {
                add("            cmul_ip(%s, t);\n", Vs('v', u).c_str());
                if (NS == 2) add("            cmul_ip(%s, t);\n", Vs('h', u).c_str());
            }
        s += "        }\n    }\n";
        return ladder_records<T2>(n);
    }

    // resident CTAs per SM the kernel is compiled for (register cap); PLB200_JIT_MINB overrides (tuning)
    static int minb() {
        const char *e = std::getenv(NS == 2 ? "PLB200_JIT_ADJ_MINB" : "PLB200_JIT_MINB");
        const int v = e ? std::atoi(e) : 0;
        // measured on the 30-qubit benchmark tape: the specialised code wants 128 registers (no spills; ptxas
        // takes 202 uncapped): c128 4 CTAs x 128 threads 211 ms (5: 232, 6: 282 with 0.5 KB of spills per
        // thread), c64 2 CTAs x 256 threads 143 ms (3: 155, 4: 264)
        return v > 0 ? v : std::max(1, 512 / (1 << NTB));
    }

    // drivers of an adjoint pass: two tiles travel together, overlaps accumulate per warp in shared memory and leave
    // as one row of per-CTA partials (as tile_kernel does; the host reduces the rows in CTA order)
    void gen_two_state_drivers(int nr) {
        std::string dev, host;
        for (int r = 0; r < nr; r++) {
            const std::string call = "round_" + std::to_string(r) + "(pp, tid, base, smem, smem1, ";
            dev += "        " + call + "accw); __syncthreads();\n";
            host += "        for (u32 tid = 0; tid < PLB_NT; tid++) { " + call + "acc); }\n";
        }
        s += "#if defined(PLB_JIT_HOST)\n"
             "extern \"C\" void plb_pass_host(T2 *sv, const PassParams *ppp, const RouteParams *, T2 *sv1, double *acc) {\n"
             "    const PassParams &pp = *ppp;\n"
             "    alignas(16) static unsigned char smem[(sizeof(T2) << PLB_M)], smem1[(sizeof(T2) << PLB_M)];\n"
             "    static u64 goff[1 << (PLB_M - PLB_LOW)];\n"
             "    for (int i = 0; i < (1 << (PLB_M - PLB_LOW)); i++) goff[i] = tile_line_offset(pp.hdr, i);\n"
             "    for (u64 t = 0; t < pp.hdr.ntiles; t++) {\n"
             "        const u64 base = insert_bits_m(t, pp.hdr.tile_ins);\n"
             "        for (u32 tid = 0; tid < PLB_NT; tid++) { load_tile(tid, base, goff, sv, smem); load_tile(tid, base, goff, sv1, smem1); }\n" +
             host +
             "        for (u32 tid = 0; tid < PLB_NT; tid++) { store_tile(tid, base, goff, sv, smem); store_tile(tid, base, goff, sv1, smem1); }\n"
             "    }\n}\n"
             "#else\n"
             "extern \"C\" __global__ void __launch_bounds__(PLB_NT, PLB_MINB)\n"
             "    plb_pass(T2 *__restrict__ sv, T2 *__restrict__ sv1, double *__restrict__ acc_g, const __grid_constant__ PassParams pp) {\n"
             "    extern __shared__ __align__(16) unsigned char smem[];\n"
             "    unsigned char *smem1 = smem + (sizeof(T2) << PLB_M);\n"
             "    u64 *goff = (u64 *)(smem + 2 * (sizeof(T2) << PLB_M));\n"
             "    double *acc = (double *)(goff + (1 << (PLB_M - PLB_LOW)));\n"
             "    for (int i = threadIdx.x; i < (1 << (PLB_M - PLB_LOW)); i += PLB_NT) goff[i] = tile_line_offset(pp.hdr, i);\n"
             "    for (int i = threadIdx.x; i < pp.hdr.nslots * (PLB_NT / 32); i += PLB_NT) acc[i] = 0.0;\n"
             "    __syncthreads();\n"
             "    const u32 tid = threadIdx.x;\n"
             "    double *accw = acc + (threadIdx.x >> 5) * pp.hdr.nslots; // one accumulator row per warp\n"
             "    for (u64 t = blockIdx.x; t < pp.hdr.ntiles; t += gridDim.x) {\n"
             "        const u64 base = insert_bits_m(t, pp.hdr.tile_ins);\n"
             "        if (t + gridDim.x < pp.hdr.ntiles) {\n"
             "            const u64 nbase = insert_bits_m(t + gridDim.x, pp.hdr.tile_ins);\n"
             "            for (int l = tid; l < (1 << (PLB_M - PLB_LOW)); l += PLB_NT) {\n"
             "                asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(sv + (nbase | goff[l])));\n"
             "                asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(sv1 + (nbase | goff[l])));\n"
             "            }\n"
             "        }\n"
             "        load_tile(tid, base, goff, sv, smem);\n"
             "        load_tile(tid, base, goff, sv1, smem1);\n"
             "        __syncthreads();\n" +
             dev +
             "        store_tile(tid, base, goff, sv, smem);\n"
             "        store_tile(tid, base, goff, sv1, smem1);\n"
             "        __syncthreads();\n"
             "    }\n"
             "    for (int i = threadIdx.x; i < pp.hdr.nslots; i += PLB_NT) { // rows in warp order -> this CTA's column of the slot-major partials\n"
             "        double t = 0.0;\n"
             "        for (int w = 0; w < PLB_NT / 32; w++) t += acc[w * pp.hdr.nslots + i];\n"
             "        acc_g[(size_t)i * gridDim.x + blockIdx.x] = t;\n"
             "    }\n"
             "}\n"
             "#endif\n";
    }

  public:
    bool ok = true;
    Route route;
    explicit Gen(const PassParams<T2> &p, const Route &rt = Route{}) : pp(p), route(rt) {}

    std::string run() {
        constexpr bool dbl = sizeof(T2) == 16;
        add("#define PLB_DOUBLE %d\n#define PLB_M %d\n#define PLB_LOW %d\n#define PLB_R %d\n#define PLB_NT %d\n#define PLB_MINB %d\n",
            dbl ? 1 : 0, M, LOW, R, 1 << NTB, minb());
        // packed FP32 (FFMA2) is opt-in: measured SLOWER on the 30-qubit c64 tape (182 vs 124 ms): the 64-bit
        // register-pair alignment of FFMA2 operands costs MOVs and spills that outweigh the halved FMA count
        add("#define PLB_NO_FFMA2 %d\n", std::getenv("PLB200_JIT_FFMA2") ? 0 : 1);
        // L2 prefetch distance in tiles of this CTA (tuning knob; 1 = the next tile, 0 = none)
        add("#define PLB_PREFETCH %dull\n", std::getenv("PLB200_JIT_PREFETCH") ? std::max(0, std::atoi(std::getenv("PLB200_JIT_PREFETCH"))) : 1);
        add("#define PLB_MAXROUNDS %d\n#define PLB_MAXOPS %d\n#define PLB_SWZ_B %d\n#define PLB_SWZ_COLS 0x%llxull\n", kMaxPassRounds,
            kMaxPassOps, Swz<T2>::B, static_cast<unsigned long long>(Swz<T2>::cols));
        add("#define PLB_SIZEOF_TILEOP %zu\n#define PLB_SIZEOF_PASSPARAMS %zu\n#define PLB_OFFSETOF_OPS %zu\n", sizeof(TileOp<T2>),
            sizeof(PassParams<T2>), offsetof(PassParams<T2>, ops));
        s += prelude();
        int nr = pp.hdr.nrounds;
        const bool first_direct = direct_ok(0);
        bool last_direct = direct_ok(nr - 1);
        if (route.k > 0 && !last_direct) {
            // the routed store phase needs a line-coalesced last round: append the transfer round
            make_transfer_round();
            nr++;
            last_direct = direct_ok(nr - 1);
            if (!last_direct) {
                ok = false;
                return s;
            }
        }
        for (int r = 0; r < nr; r++) gen_round(r, r == 0 && first_direct, r == nr - 1 && last_direct);
        if (NS == 2) {
            gen_two_state_drivers(nr);
            return s;
        }
        // ---- the two drivers: the device kernel and the test-only host loop (phases separated by barriers on
        // the device are completed for every thread before the next phase starts on the host)
        std::string dev, host;
        auto phase = [&](const std::string &call, bool barrier) {
            dev += "        " + call + ";" + (barrier ? " __syncthreads();" : "") + "\n";
            host += "        for (u32 tid = 0; tid < PLB_NT; tid++) { " + call + "; }\n";
        };
        // smem of the previous tile is still being read by the slower warps' last phase
        if ((first_direct || last_direct) && nr > 1) dev += "        __syncthreads();\n";
        if (!first_direct) phase("load_tile(tid, base, goff, sv, smem)", true);
        for (int r = 0; r < nr; r++) {
            const bool to_g = r == nr - 1 && last_direct;
            phase("round_" + std::to_string(r) + "(pp, tid, base, smem, sv" + (to_g && route.k > 0 ? ", rp)" : ")"), !to_g);
        }
        if (!last_direct) phase("store_tile(tid, base, goff, sv, smem)", !first_direct || nr == 1);
        s += "#if defined(PLB_JIT_HOST)\n"
             "extern \"C\" void plb_pass_host(T2 *sv, const PassParams *ppp, const RouteParams *rpp, T2 *, double *) {\n"
             "    const PassParams &pp = *ppp;\n"
             "    const RouteParams &rp = *rpp; (void)rp;\n"
             "    alignas(16) static unsigned char smem[(sizeof(T2) << PLB_M)];\n"
             "    static u64 goff[1 << (PLB_M - PLB_LOW)];\n"
             "    for (int i = 0; i < (1 << (PLB_M - PLB_LOW)); i++) goff[i] = tile_line_offset(pp.hdr, i);\n"
             "    for (u64 t = 0; t < pp.hdr.ntiles; t++) {\n"
             "        const u64 base = insert_bits_m(" + std::string(route.k > 0 ? "route_tile(t, rp)" : "t") + ", pp.hdr.tile_ins);\n" +
             host +
             "    }\n}\n"
             "#else\n"
             "extern \"C\" __global__ void __launch_bounds__(PLB_NT, PLB_MINB)\n"
             "    plb_pass(T2 *__restrict__ sv, const __grid_constant__ PassParams pp" +
             std::string(route.k > 0 ? ", const __grid_constant__ RouteParams rp" : "") + ") {\n"
             "    extern __shared__ __align__(16) unsigned char smem[];\n"
             "    u64 *goff = (u64 *)(smem + (sizeof(T2) << PLB_M));\n"
             "    for (int i = threadIdx.x; i < (1 << (PLB_M - PLB_LOW)); i += PLB_NT) goff[i] = tile_line_offset(pp.hdr, i);\n"
             "    __syncthreads();\n"
             "    const u32 tid = threadIdx.x;\n"
             "    for (u64 t = blockIdx.x; t < pp.hdr.ntiles; t += gridDim.x) {\n"
             "        const u64 base = insert_bits_m(" + std::string(route.k > 0 ? "route_tile(t, rp)" : "t") + ", pp.hdr.tile_ins);\n"
             "        if (PLB_PREFETCH > 0 && t + PLB_PREFETCH * gridDim.x < pp.hdr.ntiles) {\n"
             "            const u64 nbase = insert_bits_m(" + std::string(route.k > 0 ? "route_tile(t + PLB_PREFETCH * gridDim.x, rp)" : "t + PLB_PREFETCH * gridDim.x") + ", pp.hdr.tile_ins);\n"
             "            for (int l = tid; l < (1 << (PLB_M - PLB_LOW)); l += PLB_NT)\n"
             "                asm volatile(\"prefetch.global.L2 [%0];\" ::\"l\"(sv + (nbase | goff[l])));\n"
             "        }\n" +
             dev +
             "    }\n}\n"
             "#endif\n";
        return s;
    }
};

// Source of the specialised kernel of a forward pass, or "" when the pass holds an op kind the generator
// does not cover (the interpreter kernel then runs it).
template <typename T2, class Cfg>
std::string generate_pass_source(const PassParams<T2> &pp, const Route &route = Route{}) {
    Gen<T2, Cfg> g(pp, route);
    std::string src = g.run();
    return g.ok ? src : std::string();
}

} // namespace jit
} // namespace plb200
