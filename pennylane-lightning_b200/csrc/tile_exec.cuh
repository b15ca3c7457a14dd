// The tile interpreter: what ONE thread of a fused pass does (fusion.cu holds the scheduler, the
// encoder and the kernel that calls this).  Everything here is __host__ __device__ so the same code
// runs (a) inside tile_kernel on the GPU and (b) thread-by-thread on the CPU in the test-only host
// emulation (tests build it with -DPLB200_HOST_EMU; the product library never does), which lets the
// scheduler/encoder/op semantics be checked without a GPU.
//
// A pass stages a tile of 2^M amplitudes in shared memory.  A ROUND picks R register bits: thread t
// pulls the 2^R amplitudes that differ in those bits, applies the round's ops in registers and
// writes them back.  Ops are dispatched through ONE switch on a host-built code
// (kind, register-bit P, second register-bit C), every case being straight-line, predicate-free
// arithmetic on compile-time register indices:
//   * rotations (RX / RY and their controlled forms) run as three shears (lifting), 6 FMAs per pair
//     instead of 4 mul + 4 FMA, every one of them in place; other 2x2 blocks run as an in-place LU
//     (two shears + two scalings); a global scalar left over (sign of a rotation beyond pi/2,
//     1/sqrt2 of a Hadamard) is accumulated on the host and applied once (it commutes with every op);
//   * RZ-like diagonals have their first phase folded into the same scalar, so only the amplitudes
//     with parity 1 are multiplied;
//   * controls on thread bits are one compare per thread and op, controls outside the tile one
//     uniform compare per tile and op; a single control on a register bit selects compile-time
//     register subsets (SWAP_CR / DIAG_CR / DIAG_CT); anything else takes a generic masked path.
#pragma once
#include <cmath>
#include <type_traits>

#include "device.cuh"

namespace plb200 {
namespace tile {

#define PLB_HD __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define PLB_POPC(x) __popc(x)
#define PLB_POPCLL(x) __popcll(x)
#else
#define PLB_POPC(x) __builtin_popcount(x)
#define PLB_POPCLL(x) __builtin_popcountll(x)
#endif

constexpr int kMaxR = 5;
constexpr int kMaxPassOps = 208;
constexpr int kMaxPassRounds = 22;
constexpr int kMaxThreadBits = 9;

// forward pass: one state; 16 (c128) / 32 (c64) amplitudes per thread = 64 data registers either way
// (measured: c128 with R = 5 needs 255 registers, 8 warps/SM, and loses to R = 4 at 16 warps/SM although
// it executes 25 % fewer instructions).  c128 tiles are 2^11 amplitudes = 32 KiB on 128 threads, four
// CTAs per SM: measured 4 % faster than 2^12 / 256 threads / two CTAs although it needs 24 instead of
// 21 passes for the 30-qubit benchmark tape (smaller barriers, memory phases of four CTAs overlap).
template <typename T2> struct FwdCfg;
#ifndef PLB200_FWD128_M
#define PLB200_FWD128_M 11
#endif
#ifndef PLB200_FWD128_MINB
#define PLB200_FWD128_MINB 6 /* 80 registers, 24 warps/SM: measured 283 ms vs 314 (5) vs 319 (4) on the 30q tape */
#endif
#ifndef PLB200_FWD128_R
#define PLB200_FWD128_R 4
#endif
template <> struct FwdCfg<double2> {
    static constexpr int M = PLB200_FWD128_M, LOW = 3, R = PLB200_FWD128_R, NS = 1, MINB = PLB200_FWD128_MINB;
};
#ifndef PLB200_FWD64_M
#define PLB200_FWD64_M 13
#endif
#ifndef PLB200_FWD64_MINB
#define PLB200_FWD64_MINB 3 /* 80 registers, 24 warps/SM: 176 ms vs 207 (2) on the 30q tape */
#endif
#ifndef PLB200_FWD64_R
#define PLB200_FWD64_R 5
#endif
template <> struct FwdCfg<float2> {
    static constexpr int M = PLB200_FWD64_M, LOW = 4, R = PLB200_FWD64_R, NS = 1, MINB = PLB200_FWD64_MINB;
};
// adjoint pass: two states, 2 x 32 KiB tiles, 8 + 8 amplitudes per thread
template <typename T2> struct AdjCfg;
#ifndef PLB200_ADJ_MINB
#define PLB200_ADJ_MINB 2
#endif
template <> struct AdjCfg<double2> {
    static constexpr int M = 11, LOW = 3, R = 3, NS = 2, MINB = PLB200_ADJ_MINB;
};
template <> struct AdjCfg<float2> {
    static constexpr int M = 12, LOW = 4, R = 3, NS = 2, MINB = 2; // 512 threads: 64 registers
};

// op kinds; dispatch code = kind << 4 | P << 2 | C.  Every form updates its registers IN PLACE (each
// statement overwrites one operand with an FMA of the current values), so the register file carries
// the tile state through the op loop without copies.
enum : int {
    // pair ops on register bit P (a = bit clear, b = bit set)
    K_LIFT_R = 0, // real rotation as three shears: a += u b; b += l a; a += u b       m[0] = (u, l)
    K_LIFT_I = 1, // same with imaginary shears:    a += iu b; b += il a; a += iu b    m[0] = (u, l)
    K_HAD = 2,    // s [[1, 1], [1, -1]]:           a += b; b = a - 2 b
    K_LU_R = 3,   // real 2x2 = diag(al, be) L(t2) U(t1): a += t1 b; b += t2 a; a *= al; b *= be
                  //                                m[0] = (t1, t2), m[1] = (al, be)
    K_LU_C = 4,   // complex 2x2, same factorisation   m[0..3] = t1, t2, al, be
    K_SWAP = 5,   // [[0, 1], [1, 0]]
    // ... the same under a register-level predicate (umask)
    K_LIFT_R_M = 6,
    K_LIFT_I_M = 7,
    K_LU_R_M = 8,
    K_LU_C_M = 9,
    K_SWAP_M = 10,
    // diagonal ops, phases m[0] (parity 0) / m[1] (parity 1)
    K_DIAG_R = 11,  // parity = thread part ^ register bit P
    K_DIAG1_R = 12, // parity = register bit P only, m[0] == 1
    K_DIAG_T = 13,  // parity on thread / outside bits only
    K_DIAG1_T = 14, // ... with m[0] == 1
    K_DIAG_PP = 15, // parity = thread part ^ register bits P ^ C
    // exactly one control (value 1) on register bit C
    K_SWAP_CR = 16, // X on register bit P
    K_DIAG_CR = 17, // phase by (thread parity ^ register bit P)
    K_DIAG_CT = 18, // phase by thread parity                 (control bit stored in the P field)
    K_DIAG_G = 19,  // generic: register-level predicate umask / parity upar
    // SWAP of the two register bits P, C (exchange the registers with (P, C) = (1, 0) and (0, 1))
    K_SWAP2 = 20,
    K_SWAP2_M = 21, // ... under the register-level predicate umask
    // LADDER (tail section of a round, not dispatched through the switch): header of a run of
    // controlled phases sharing one register bit P (P = R: none); the next `slot` records are entries
    // (thread / outside controls + phase m[0]).  The thread multiplies the phases of its active
    // entries into one scalar and applies it once.
    K_LADDER = 22,
    // adjoint: accumulate Im<h| G |l>
    K_OVL_X = 23,
    K_OVL_Y = 24,
    K_OVL_D = 25, // G = diag(g[parity]), g = (m[0].x, m[0].y)
    // One 2x2 block of a TWO-bit pair op (IsingXX / XY / YY, SingleExcitation(+-), PSWAP, ...): register bits P, C
    // carry the op's two target bits; the block acts on the amplitudes with (C, P) = a and = b (2-bit values in
    // slot bits 0-1 / 2-3), in the in-place form slot bits 4-7 (K_LIFT_R / K_LIFT_I / K_LU_R / K_LU_C / K_SWAP),
    // for the registers with umask set.  Only the SPECIALISED kernels implement it (jit_codegen.hpp): a pass
    // holding one is never handed to the interpreter.
    K_PAIR2 = 26,
    // SCALED rotations (specialised kernels only, "jit forms" of the encoder): an uncontrolled rotation leaves its
    // cosine (or sine) with the host-carried scalar, so a pair costs 4 FMAs instead of the 6 of three shears:
    //   K_SROT_R: c [[1, -t], [t, 1]]   a' = a - t b, b' = b + t a      t = m[0].x = tan
    //   K_SROK_R: s [[k, -1], [1, k]]   a' = k a - b, b' = a + k b      k = m[0].x = cot
    // K_SROT_I / K_SROK_I are the RX-like twins [[1, -it], [-it, 1]] / [[ik, 1], [1, ik]].  The tangent form is
    // used for every angle whose tangent stays below 2^16 (the stored amplitudes grow by 1 / cos, floating point
    // keeps their relative precision, the encoder bounds the growth inside a round), so the structure of a pass
    // does not depend on its angles except within 2e-5 of a half turn, where the cotangent form takes over.
    K_SROT_R = 27,
    K_SROT_I = 28,
    K_SROK_R = 29,
    K_SROK_I = 30,
    // One 2x2 block of a FOUR-bit pair op (DoubleExcitation: a Givens rotation on |0011>, |1100>): all four target
    // bits are register bits of the round; slot = a | b << 4 | form << 8 | pos0 << 12 | pos1 << 15 | pos2 << 18 |
    // pos3 << 21 (a, b: 4-bit values over the target bits, pos_j: register bit carrying target bit j).  Specialised
    // kernels only, like K_PAIR2.
    K_PAIR4 = 31,
    // Dense 4x4 matrix on TWO register bits P (matrix bit 0) and C (matrix bit 1) — QubitUnitary on two wires and its
    // controlled forms.  Four records: record j carries row j of the matrix in m[0..3]; only the first is an op.
    // Specialised kernels only.
    K_DENSE2 = 32,
    K_LAST_OVL = K_OVL_D,
    K_FIRST_DIAG = K_DIAG_R,
    K_FIRST_OVL = K_OVL_X,
};
// Code word of an op: bits 0-7 = dense case index of apply_gate's switch (one jump table), bits 8-12 =
// kind (bit 19: its sixth bit, kinds >= 32 exist in the specialised kernels only), bits 13-15 = P, bits 16-18 = C,
// then flags.
constexpr uint32_t F_COND = 1u << 20; // has controls (outside / thread bits): needs the predicate
constexpr uint32_t F_PAR = 1u << 21;  // phase depends on a thread / outside parity
constexpr uint32_t F_OVL = 1u << 22;  // adjoint overlap op
PLB_HD constexpr int kind_cases(int kind) {
    return (kind == K_DIAG_PP || kind == K_SWAP_CR || kind == K_DIAG_CR || kind == K_SWAP2 || kind == K_SWAP2_M)
               ? kMaxR * (kMaxR - 1)
           : (kind == K_LADDER || kind >= K_PAIR2) ? 0
           : (kind == K_DIAG_T || kind == K_DIAG1_T || kind == K_DIAG_G) ? 1
                                                                        : kMaxR;
}
PLB_HD constexpr int kind_base(int kind) {
    int b = 0;
    for (int k = 0; k < kind; k++) b += kind_cases(k);
    return b;
}
PLB_HD constexpr int pc_index(int p, int c) { return p * (kMaxR - 1) + (c < p ? c : c - 1); }
PLB_HD constexpr uint32_t make_code(int kind, int p, int c) {
    const int n = kind_cases(kind);
    const int sub = n == kMaxR * (kMaxR - 1) ? pc_index(p, c) : n == kMaxR ? p : 0;
    return static_cast<uint32_t>((kind_base(kind) + sub) & 255) | static_cast<uint32_t>(kind & 31) << 8 |
           static_cast<uint32_t>(kind >> 5) << 19 | static_cast<uint32_t>(p) << 13 | static_cast<uint32_t>(c) << 16 |
           ((kind >= K_FIRST_OVL && kind <= K_LAST_OVL) ? F_OVL : 0u);
}
PLB_HD constexpr int code_kind(uint32_t code) { return static_cast<int>(((code >> 8) & 31u) | (((code >> 19) & 1u) << 5)); }
PLB_HD constexpr int code_p(uint32_t code) { return static_cast<int>((code >> 13) & 7u); }

template <typename T2> struct alignas(16) TileOp {
    uint32_t code;
    uint32_t cm_tid, cv_tid; // controls on thread bits, as masks over threadIdx bits
    uint32_t pm_tid;         // parity mask over threadIdx bits
    uint32_t umask, upar;    // generic masked forms: bit u = register u is active / has odd parity
    uint32_t slot;           // adjoint: accumulator slot inside the pass
    uint64_t cmask_o, cval_o, pmask_o; // bits outside the tile: uniform per tile
    T2 m[4];
};
// Entry of a tail ladder, packed 4 (c128) / 3 (c64) to a TileOp-sized record: the phase applies where
// every bit of cmask_o (outside the tile) and of cm_tid (thread bits) is 1.
template <typename T2> struct alignas(16) LadderEntry {
    uint64_t cmask_o;
    uint32_t cm_tid, pad;
    T2 ph;
};
template <typename T2> PLB_HD constexpr int ladder_entries_per_record() {
    return static_cast<int>(sizeof(TileOp<T2>) / sizeof(LadderEntry<T2>));
}
template <typename T2> PLB_HD constexpr int ladder_records(int n) {
    return 1 + (n + ladder_entries_per_record<T2>() - 1) / ladder_entries_per_record<T2>();
}
struct alignas(16) RoundHdr {
    int first_op, nops; // regular ops [first_op, first_op + nops), then nlad records of tail ladders
    int nlad, pad;
    uint32_t w[kMaxThreadBits]; // swizzled byte offset contributed by thread bit i
    uint32_t sroff[1 << kMaxR]; // swizzled byte offset of register u
};
struct alignas(16) PassHdr {
    int nrounds, nops_total;
    uint64_t ntiles;
    int nslots, pad;
    BitInsert tile_ins; // zeros at the M tile bits
};
// Global offset of line i (2^LOW amplitudes) of a tile: bit b of i -> tile bit LOW + b.
template <int M, int LOW> PLB_HD uint64_t tile_line_offset(const PassHdr &h, int i) {
    uint64_t o = 0;
#pragma unroll
    for (int b = LOW; b < M; b++)
        if ((i >> (b - LOW)) & 1) o |= h.tile_ins.lowmask[b] + 1; // lowmask = (1 << position) - 1
    return o;
}
// The whole pass description travels as a __grid_constant__ kernel parameter (constant bank,
// uniform loads): nothing about the ops is fetched through the LSU/L1 data path.
template <typename T2> struct alignas(16) PassParams {
    PassHdr hdr;
    RoundHdr rounds[kMaxPassRounds];
    TileOp<T2> ops[kMaxPassOps + 1]; // +1: the interpreter prefetches op k+1
};

// ---- shared-memory swizzle.  Amplitude j of the tile lives at element j ^ g(j), where g is a
// GF(2)-linear map of the bits >= B into the B "bank-group" bits (B = 3: 8 x 16 B for c128,
// B = 4: 16 x 8 B for c64).  Each tile bit has a column; no nonzero column repeats more than twice
// (c128) / at all (c64), so after removing any 4 register bits the remaining bits still hold B
// independent columns: the host maps those to the low lane bits and every quarter/half-warp access
// of a round touches all banks exactly once.  Linearity makes the address of
// (thread part | register part) the XOR of two precomputed halves.
template <typename T2> struct Swz;
template <> struct Swz<double2> {
    static constexpr int B = 3;
    // columns of bits 0..12: 1,2,4, 3,5,6, 7,1,2, 4,3,5, 6 (4 bits each, bit 0 first)
    static constexpr uint64_t cols = 0x6534217653421ull;
};
template <> struct Swz<float2> {
    static constexpr int B = 4;
    // 1,2,4,8, 3,5,6,9, 10,12,7,11, 13
    static constexpr uint64_t cols = 0xDB7CA96538421ull;
};
template <typename T2> PLB_HD constexpr uint32_t swz_col(int bit) {
    return static_cast<uint32_t>((Swz<T2>::cols >> (4 * bit)) & 15u);
}
template <typename T2> PLB_HD constexpr uint32_t swz(uint32_t j) {
    uint32_t g = 0;
    for (int i = Swz<T2>::B; i < 13; i++)
        if ((j >> i) & 1u) g ^= swz_col<T2>(i);
    return j ^ g;
}

// ---- arithmetic (host + device), all in place
template <typename T2> PLB_HD void cmul_ip(T2 &v, const T2 d) { // v *= d
    const auto t = v.x * d.y;
    v.x = v.x * d.x;
    v.x = fma(-v.y, d.y, v.x);
    v.y = fma(v.y, d.x, t);
}
template <typename T2> PLB_HD void cshear_ip(T2 &a, const T2 t, const T2 &b) { // a += t * b
    a.x = fma(t.x, b.x, a.x);
    a.x = fma(-t.y, b.y, a.x);
    a.y = fma(t.x, b.y, a.y);
    a.y = fma(t.y, b.x, a.y);
}
// exchange two scalars without a temporary the compiler could hoist (three XORs on the bit patterns)
PLB_HD void xswap(double &a, double &b) {
#if defined(__CUDA_ARCH__)
    asm volatile("xor.b64 %0, %0, %1;\n\txor.b64 %1, %1, %0;\n\txor.b64 %0, %0, %1;" : "+d"(a), "+d"(b));
#else
    const double t = a;
    a = b, b = t;
#endif
}
PLB_HD void xswap(float &a, float &b) {
#if defined(__CUDA_ARCH__)
    asm volatile("xor.b32 %0, %0, %1;\n\txor.b32 %1, %1, %0;\n\txor.b32 %0, %0, %1;" : "+f"(a), "+f"(b));
#else
    const float t = a;
    a = b, b = t;
#endif
}

template <typename T2, int R, int P, int KIND, bool MASKED>
PLB_HD void pair_op(T2 (&v)[1 << R], const TileOp<T2> &op, const T2 m0, uint32_t active) {
    T2 m1 = m0, m2 = m0, m3 = m0;
    if constexpr (KIND == K_LU_R || KIND == K_LU_C) m1 = op.m[1];
    if constexpr (KIND == K_LU_C) m2 = op.m[2], m3 = op.m[3];
#pragma unroll
    for (int q = 0; q < (1 << (R - 1)); q++) {
        const int u0 = ((q >> P) << (P + 1)) | (q & ((1 << P) - 1));
        const int u1 = u0 | (1 << P);
        if (MASKED && !((active >> u0) & 1u)) continue;
        T2 &a = v[u0], &b = v[u1];
        if constexpr (KIND == K_SWAP) {
            xswap(a.x, b.x), xswap(a.y, b.y);
        } else if constexpr (KIND == K_LIFT_R) {
            a.x = fma(m0.x, b.x, a.x), a.y = fma(m0.x, b.y, a.y);
            b.x = fma(m0.y, a.x, b.x), b.y = fma(m0.y, a.y, b.y);
            a.x = fma(m0.x, b.x, a.x), a.y = fma(m0.x, b.y, a.y);
        } else if constexpr (KIND == K_LIFT_I) {
            a.x = fma(-m0.x, b.y, a.x), a.y = fma(m0.x, b.x, a.y);
            b.x = fma(-m0.y, a.y, b.x), b.y = fma(m0.y, a.x, b.y);
            a.x = fma(-m0.x, b.y, a.x), a.y = fma(m0.x, b.x, a.y);
        } else if constexpr (KIND == K_HAD) {
            a.x = a.x + b.x, a.y = a.y + b.y;
            b.x = fma(static_cast<decltype(b.x)>(-2), b.x, a.x), b.y = fma(static_cast<decltype(b.y)>(-2), b.y, a.y);
        } else if constexpr (KIND == K_LU_R) {
            a.x = fma(m0.x, b.x, a.x), a.y = fma(m0.x, b.y, a.y);
            b.x = fma(m0.y, a.x, b.x), b.y = fma(m0.y, a.y, b.y);
            a.x = a.x * m1.x, a.y = a.y * m1.x;
            b.x = b.x * m1.y, b.y = b.y * m1.y;
        } else {
            cshear_ip(a, m0, b);
            cshear_ip(b, m1, a);
            cmul_ip(a, m2);
            cmul_ip(b, m3);
        }
    }
}

// X on register bit P for the registers whose bit C is set
template <typename T2, int R, int P, int C> PLB_HD void swap_cr(T2 (&v)[1 << R]) {
#pragma unroll
    for (int q = 0; q < (1 << (R - 1)); q++) {
        const int u0 = ((q >> P) << (P + 1)) | (q & ((1 << P) - 1));
        const int u1 = u0 | (1 << P);
        if ((u0 >> C) & 1) xswap(v[u0].x, v[u1].x), xswap(v[u0].y, v[u1].y);
    }
}

// SWAP of register bits P and C: exchange v[u | P] and v[u | C] for every u with both bits clear
template <typename T2, int R, int P, int C, bool MASKED>
PLB_HD void swap2(T2 (&v)[1 << R], uint32_t active) {
#pragma unroll
    for (int u = 0; u < (1 << R); u++) {
        if (((u >> P) & 1) || ((u >> C) & 1)) continue;
        const int ua = u | (1 << P), ub = u | (1 << C);
        if (MASKED && !((active >> ua) & 1u)) continue;
        xswap(v[ua].x, v[ub].x), xswap(v[ua].y, v[ub].y);
    }
}

// v[u] *= (bit P of u ? dB : dA); ONE: dA == 1.  CTRL >= 0: only registers whose bit CTRL is set.
template <typename T2, int R, int P, bool ONE, int CTRL>
PLB_HD void diag_bit(T2 (&v)[1 << R], T2 dA, T2 dB) {
#pragma unroll
    for (int u = 0; u < (1 << R); u++) {
        if (CTRL >= 0 && !((u >> (CTRL < 0 ? 0 : CTRL)) & 1)) continue;
        if ((u >> P) & 1) cmul_ip(v[u], dB);
        else if (!ONE) cmul_ip(v[u], dA);
    }
}
// v[u] *= d for every register (CTRL < 0) or those whose bit CTRL is set
template <typename T2, int R, int CTRL> PLB_HD void diag_all(T2 (&v)[1 << R], T2 d) {
#pragma unroll
    for (int u = 0; u < (1 << R); u++) {
        if (CTRL >= 0 && !((u >> (CTRL < 0 ? 0 : CTRL)) & 1)) continue;
        cmul_ip(v[u], d);
    }
}
// v[u] *= (bit P ^ bit C of u) ? dB : dA
template <typename T2, int R, int P, int C> PLB_HD void diag_pp(T2 (&v)[1 << R], T2 dA, T2 dB) {
#pragma unroll
    for (int u = 0; u < (1 << R); u++) cmul_ip(v[u], (((u >> P) ^ (u >> C)) & 1) ? dB : dA);
}

#define PLB_CASE(IDX, PV, CV, STMT)                                                                      \
    case (IDX): {                                                                                        \
        constexpr int P = (PV), C = (CV);                                                                \
        (void)P, (void)C;                                                                                \
        STMT;                                                                                            \
    } break;
// case lists per number of register bits: only the register positions that exist are instantiated
#define PLB_CASES_P_3(KIND, STMT) PLB_CASE(kind_base(KIND) + 0, 0, 0, STMT) PLB_CASE(kind_base(KIND) + 1, 1, 0, STMT) PLB_CASE(kind_base(KIND) + 2, 2, 0, STMT)
#define PLB_CASES_PC_3(KIND, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 1), 0, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 2), 0, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 0), 1, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 2), 1, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 0), 2, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 1), 2, 1, STMT)
#define PLB_CASES_P_4(KIND, STMT) PLB_CASE(kind_base(KIND) + 0, 0, 0, STMT) PLB_CASE(kind_base(KIND) + 1, 1, 0, STMT) PLB_CASE(kind_base(KIND) + 2, 2, 0, STMT) PLB_CASE(kind_base(KIND) + 3, 3, 0, STMT)
#define PLB_CASES_PC_4(KIND, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 1), 0, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 2), 0, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 3), 0, 3, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 0), 1, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 2), 1, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 3), 1, 3, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 0), 2, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 1), 2, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 3), 2, 3, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 0), 3, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 1), 3, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 2), 3, 2, STMT)
#define PLB_CASES_P_5(KIND, STMT) PLB_CASE(kind_base(KIND) + 0, 0, 0, STMT) PLB_CASE(kind_base(KIND) + 1, 1, 0, STMT) PLB_CASE(kind_base(KIND) + 2, 2, 0, STMT) PLB_CASE(kind_base(KIND) + 3, 3, 0, STMT) PLB_CASE(kind_base(KIND) + 4, 4, 0, STMT)
#define PLB_CASES_PC_5(KIND, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 1), 0, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 2), 0, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 3), 0, 3, STMT) PLB_CASE(kind_base(KIND) + pc_index(0, 4), 0, 4, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 0), 1, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 2), 1, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 3), 1, 3, STMT) PLB_CASE(kind_base(KIND) + pc_index(1, 4), 1, 4, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 0), 2, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 1), 2, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 3), 2, 3, STMT) PLB_CASE(kind_base(KIND) + pc_index(2, 4), 2, 4, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 0), 3, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 1), 3, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 2), 3, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(3, 4), 3, 4, STMT) PLB_CASE(kind_base(KIND) + pc_index(4, 0), 4, 0, STMT) PLB_CASE(kind_base(KIND) + pc_index(4, 1), 4, 1, STMT) PLB_CASE(kind_base(KIND) + pc_index(4, 2), 4, 2, STMT) PLB_CASE(kind_base(KIND) + pc_index(4, 3), 4, 3, STMT)

// One gate op on one register set.  m0 = op.m[0] (prefetched by the caller).  pt: parity of the
// thread + outside bits under the op's parity mask.
// EXT = false: the lean interpreter = the op kinds of the common gate sets only (rotations, Hadamard,
// X / CNOT, phase gates and their singly-controlled forms); EXT = true adds the LU forms of arbitrary
// 2x2 blocks, the register-masked forms, two-register-bit parities and two-bit SWAPs (and round() adds
// the tail ladders).  A pass runs the lean kernel unless it holds one of those: its dispatch tree and
// code are half the size, and its op loop decodes on the uniform datapath.
#define PLB_NO_CASES(KIND, STMT)
#define PLB_DEFINE_APPLY_GATE(RV, EXTV, CASES_P, CASES_PC, CASES_EXT, CASES_P_EXT)                       \
    template <typename T2>                                                                               \
    PLB_HD void apply_gate(T2 (&v)[1 << RV], const TileOp<T2> &op, uint32_t code, const T2 m0, bool pt,  \
                           std::integral_constant<bool, EXTV>) {                                         \
        constexpr int R = RV;                                                                            \
        switch (code & 255u) {                                                                           \
            CASES_P(K_LIFT_R, (pair_op<T2, R, P, K_LIFT_R, false>(v, op, m0, 0u)))                       \
            CASES_P(K_LIFT_I, (pair_op<T2, R, P, K_LIFT_I, false>(v, op, m0, 0u)))                       \
            CASES_P(K_HAD, (pair_op<T2, R, P, K_HAD, false>(v, op, m0, 0u)))                             \
            CASES_P_EXT(K_LU_R, (pair_op<T2, R, P, K_LU_R, false>(v, op, m0, 0u)))                           \
            CASES_P_EXT(K_LU_C, (pair_op<T2, R, P, K_LU_C, false>(v, op, m0, 0u)))                           \
            CASES_P(K_SWAP, (pair_op<T2, R, P, K_SWAP, false>(v, op, m0, 0u)))                           \
            CASES_P_EXT(K_LIFT_R_M, (pair_op<T2, R, P, K_LIFT_R, true>(v, op, m0, op.umask)))                \
            CASES_P_EXT(K_LIFT_I_M, (pair_op<T2, R, P, K_LIFT_I, true>(v, op, m0, op.umask)))                \
            CASES_P_EXT(K_LU_R_M, (pair_op<T2, R, P, K_LU_R, true>(v, op, m0, op.umask)))                    \
            CASES_P_EXT(K_LU_C_M, (pair_op<T2, R, P, K_LU_C, true>(v, op, m0, op.umask)))                    \
            CASES_P_EXT(K_SWAP_M, (pair_op<T2, R, P, K_SWAP, true>(v, op, m0, op.umask)))                    \
            CASES_P(K_DIAG_R, (diag_bit<T2, R, P, false, -1>(v, pt ? op.m[1] : m0, pt ? m0 : op.m[1])))  \
            CASES_P(K_DIAG1_R, (diag_bit<T2, R, P, true, -1>(v, m0, m0)))                                \
            PLB_CASE(kind_base(K_DIAG_T), 0, 0, (diag_all<T2, R, -1>(v, pt ? op.m[1] : m0)))             \
            PLB_CASE(kind_base(K_DIAG1_T), 0, 0, (pt ? diag_all<T2, R, -1>(v, m0) : (void)0))            \
            CASES_EXT(K_DIAG_PP, (diag_pp<T2, R, P, C>(v, pt ? op.m[1] : m0, pt ? m0 : op.m[1])))         \
            CASES_PC(K_SWAP_CR, (swap_cr<T2, R, P, C>(v)))                                               \
            CASES_EXT(K_SWAP2, (swap2<T2, R, P, C, false>(v, 0u)))                                       \
            CASES_EXT(K_SWAP2_M, (swap2<T2, R, P, C, true>(v, op.umask)))                                \
            CASES_PC(K_DIAG_CR, (diag_bit<T2, R, P, false, C>(v, pt ? op.m[1] : m0, pt ? m0 : op.m[1]))) \
            CASES_P(K_DIAG_CT, (diag_all<T2, R, P>(v, pt ? op.m[1] : m0)))                               \
        default: { /* K_DIAG_G */                                                                        \
            const uint32_t pb = pt ? ~op.upar : op.upar;                                                 \
            const uint32_t active = op.umask;                                                            \
            const T2 d1 = op.m[1];                                                                       \
            _Pragma("unroll") for (int u = 0; u < (1 << R); u++) if (active & (1u << u))                 \
                cmul_ip(v[u], (pb >> u & 1) ? d1 : m0);                                                  \
        } break;                                                                                         \
        }                                                                                                \
    }
PLB_DEFINE_APPLY_GATE(3, false, PLB_CASES_P_3, PLB_CASES_PC_3, PLB_NO_CASES, PLB_NO_CASES)
PLB_DEFINE_APPLY_GATE(4, false, PLB_CASES_P_4, PLB_CASES_PC_4, PLB_NO_CASES, PLB_NO_CASES)
PLB_DEFINE_APPLY_GATE(5, false, PLB_CASES_P_5, PLB_CASES_PC_5, PLB_NO_CASES, PLB_NO_CASES)
PLB_DEFINE_APPLY_GATE(3, true, PLB_CASES_P_3, PLB_CASES_PC_3, PLB_CASES_PC_3, PLB_CASES_P_3)
PLB_DEFINE_APPLY_GATE(4, true, PLB_CASES_P_4, PLB_CASES_PC_4, PLB_CASES_PC_4, PLB_CASES_P_4)
PLB_DEFINE_APPLY_GATE(5, true, PLB_CASES_P_5, PLB_CASES_PC_5, PLB_CASES_PC_5, PLB_CASES_P_5)

// Im(conj(a) b), Re(conj(a) b)
template <typename T2> PLB_HD double im_cb(T2 a, T2 b) {
    return static_cast<double>(a.x) * b.y - static_cast<double>(a.y) * b.x;
}
template <typename T2> PLB_HD double re_cb(T2 a, T2 b) {
    return static_cast<double>(a.x) * b.x + static_cast<double>(a.y) * b.y;
}
template <typename T2, int R, int P, bool ISY>
PLB_HD double overlap_pair(const T2 (&l)[1 << R], const T2 (&h)[1 << R], uint32_t active) {
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
// thread-local overlap contribution of one K_OVL_* op
template <typename T2, int R>
PLB_HD double overlap_op(const T2 (&l)[1 << R], const T2 (&h)[1 << R], const TileOp<T2> &op, uint32_t code,
                         bool thr_ok, bool pt) {
    const uint32_t active = thr_ok ? static_cast<uint32_t>(op.umask) : 0u;
    const int kind = code_kind(code), p = code_p(code);
    double s = 0;
    if (kind == K_OVL_D) {
        const uint32_t pb = pt ? ~static_cast<uint32_t>(op.upar) : static_cast<uint32_t>(op.upar);
        const double g0 = op.m[0].x, g1 = op.m[0].y;
#pragma unroll
        for (int u = 0; u < (1 << R); u++)
            if (active & (1u << u)) s += ((pb >> u & 1) ? g1 : g0) * im_cb(h[u], l[u]);
        return s;
    }
    const bool isy = kind == K_OVL_Y;
    switch (p) {
    case 0: s = isy ? overlap_pair<T2, R, 0, true>(l, h, active) : overlap_pair<T2, R, 0, false>(l, h, active); break;
    case 1: {
        constexpr int P = R > 1 ? 1 : 0;
        s = isy ? overlap_pair<T2, R, P, true>(l, h, active) : overlap_pair<T2, R, P, false>(l, h, active);
    } break;
    case 2: {
        constexpr int P = R > 2 ? 2 : R - 1;
        s = isy ? overlap_pair<T2, R, P, true>(l, h, active) : overlap_pair<T2, R, P, false>(l, h, active);
    } break;
    case 3: {
        constexpr int P = R > 3 ? 3 : R - 1;
        s = isy ? overlap_pair<T2, R, P, true>(l, h, active) : overlap_pair<T2, R, P, false>(l, h, active);
    } break;
    default: {
        constexpr int P = R > 4 ? 4 : R - 1;
        s = isy ? overlap_pair<T2, R, P, true>(l, h, active) : overlap_pair<T2, R, P, false>(l, h, active);
    } break;
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// Per-thread pieces of a pass.  smem0/smem1: the staged tile(s); goff: global offset of each
// 2^LOW-amplitude line of the tile; acc: per-CTA overlap accumulators (adjoint).
template <typename T2, class Cfg, bool EXT> struct Exec {
    static constexpr int M = Cfg::M, LOW = Cfg::LOW, R = Cfg::R, NS = Cfg::NS;
    static constexpr int NT = 1 << (M - R), NV = 1 << R;

    static PLB_HD uint32_t tid_swz_bytes(uint32_t tid) { return swz<T2>(tid) * static_cast<uint32_t>(sizeof(T2)); }

    static PLB_HD void load_tile(uint32_t tid, uint64_t base, const uint64_t *goff, const T2 *__restrict__ sv,
                                 unsigned char *smem) {
        const uint32_t st = tid_swz_bytes(tid);
        T2 v[NV];
#pragma unroll
        for (int u = 0; u < NV; u++) {
            const uint32_t j = tid + u * NT;
            v[u] = sv[base | goff[j >> LOW] | (j & ((1u << LOW) - 1))];
        }
#pragma unroll
        for (int u = 0; u < NV; u++)
            *reinterpret_cast<T2 *>(smem + (st ^ (swz<T2>(u * NT) * static_cast<uint32_t>(sizeof(T2))))) = v[u];
    }
    static PLB_HD void store_tile(uint32_t tid, uint64_t base, const uint64_t *goff, T2 *__restrict__ sv,
                                  const unsigned char *smem) {
        const uint32_t st = tid_swz_bytes(tid);
        T2 v[NV];
#pragma unroll
        for (int u = 0; u < NV; u++)
            v[u] = *reinterpret_cast<const T2 *>(smem + (st ^ (swz<T2>(u * NT) * static_cast<uint32_t>(sizeof(T2)))));
#pragma unroll
        for (int u = 0; u < NV; u++) {
            const uint32_t j = tid + u * NT;
            sv[base | goff[j >> LOW] | (j & ((1u << LOW) - 1))] = v[u];
        }
    }

    // v[u] *= t for the registers whose bit p is set (p == R: all registers)
    static PLB_HD void ladder_apply(T2 (&v)[NV], int p, T2 t) {
        switch (p) {
        case 0: diag_bit<T2, R, 0, true, -1>(v, t, t); break;
        case 1: diag_bit<T2, R, (R > 1 ? 1 : 0), true, -1>(v, t, t); break;
        case 2: diag_bit<T2, R, (R > 2 ? 2 : 0), true, -1>(v, t, t); break;
        case 3: if constexpr (R > 3) { diag_bit<T2, R, (R > 3 ? 3 : 0), true, -1>(v, t, t); break; }
        case 4: if constexpr (R > 4) { diag_bit<T2, R, (R > 4 ? 4 : 0), true, -1>(v, t, t); break; }
        default: diag_all<T2, R, -1>(v, t); break;
        }
    }

    // One round of one thread.  `reduce(slot, s)` receives this thread's overlap contributions.
    template <class Reduce>
    static PLB_HD void round(const PassParams<T2> &pp, int r, uint32_t tid, uint64_t base, unsigned char *smem0,
                             unsigned char *smem1, Reduce &&reduce) {
        const RoundHdr &rh = pp.rounds[r];
        uint32_t sb = 0;
#pragma unroll
        for (int i = 0; i < M - R; i++)
            if ((tid >> i) & 1u) sb ^= rh.w[i];
        T2 v[NV];
        T2 h[NS == 2 ? NV : 1];
#pragma unroll
        for (int u = 0; u < NV; u++) v[u] = *reinterpret_cast<const T2 *>(smem0 + (sb ^ rh.sroff[u]));
        if constexpr (NS == 2) {
#pragma unroll
            for (int u = 0; u < NV; u++) h[u] = *reinterpret_cast<const T2 *>(smem1 + (sb ^ rh.sroff[u]));
        }
        // software-pipelined decode: the code word and first parameter pair of op k+1 are fetched
        // from the constant bank while op k executes (ops[] has one slack entry)
        const TileOp<T2> *ops = pp.ops + rh.first_op;
        const int nops = rh.nops;
        uint32_t code = ops[0].code;
        T2 m0 = ops[0].m[0];
        for (int k = 0; k < nops; k++) {
            const TileOp<T2> &op = ops[k];
            const uint32_t ccode = code;
            const T2 cm0 = m0;
            code = ops[k + 1].code;
            m0 = ops[k + 1].m[0];
            bool thr_ok = true, pt = false;
            if (ccode & F_COND) {
                if ((base & op.cmask_o) != op.cval_o) continue; // uniform per tile
                thr_ok = (tid & op.cm_tid) == op.cv_tid;
            }
            if (ccode & F_PAR) pt = ((PLB_POPC(tid & op.pm_tid) + PLB_POPCLL(base & op.pmask_o)) & 1) != 0;
            if constexpr (NS == 2) {
                if (ccode & F_OVL) {
                    reduce(static_cast<int>(op.slot), overlap_op<T2, R>(v, h, op, ccode, thr_ok, pt));
                    continue;
                }
                if (thr_ok) apply_gate<T2>(h, op, ccode, cm0, pt, std::integral_constant<bool, EXT>{});
            }
            if (thr_ok) apply_gate<T2>(v, op, ccode, cm0, pt, std::integral_constant<bool, EXT>{});
        }
        // ---- tail ladders (EXT kernels only): runs of controlled phases that commute to the end of
        // the round.  Each is a header (the register bit P they share, or P = R: none) + `slot` entries
        // (thread / outside controls + phase); the thread multiplies the phases of its active entries
        // and applies the product once.
        if constexpr (EXT) {
            const TileOp<T2> *lad = ops + nops;
            const int nl = rh.nlad;
            for (int q = 0; q < nl;) {
                const int n = static_cast<int>(lad[q].slot);
                const int p = code_p(lad[q].code);
                T2 t;
                t.x = 1, t.y = 0;
                bool any = false;
                const LadderEntry<T2> *en = reinterpret_cast<const LadderEntry<T2> *>(lad + q + 1);
#pragma unroll 4
                for (int e = 0; e < n; e++) {
                    const uint64_t co = en[e].cmask_o;
                    if ((base & co) != co) continue; // uniform per tile
                    const uint32_t ct = en[e].cm_tid;
                    if ((tid & ct) == ct) {
                        cmul_ip(t, en[e].ph);
                        any = true;
                    }
                }
                q += ladder_records<T2>(n);
                if (!any) continue;
                if constexpr (NS == 2) ladder_apply(h, p, t);
                ladder_apply(v, p, t);
            }
        }
#pragma unroll
        for (int u = 0; u < NV; u++) *reinterpret_cast<T2 *>(smem0 + (sb ^ rh.sroff[u])) = v[u];
        if constexpr (NS == 2) {
#pragma unroll
            for (int u = 0; u < NV; u++) *reinterpret_cast<T2 *>(smem1 + (sb ^ rh.sroff[u])) = h[u];
        }
    }

};

} // namespace tile
} // namespace plb200
