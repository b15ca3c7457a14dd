// TEST INFRASTRUCTURE — not part of the product.  tests/_emu/libplb200_emu.so = lowering.cpp + fusion.cu built
// with -DPLB200_HOST_EMU + this file: the fusion scheduler, the pass encoder and the per-thread tile
// interpreter (tile_exec.cuh) executed thread-by-thread on HOST memory, so that `-m "not gpu"` tests
// can check them against the numpy oracle without a GPU.  libplb200.so contains none of this and has
// no CPU execution path.
#include <complex>
#include <cstring>

#include "../../include/plb200.h"
#include "fusion.hpp"

namespace plb200 {
int emulate_fused(int n, int precision, const std::vector<AdjItem> &items, bool adjoint, bool scaled, void *sv0,
                  void *sv1, double *acc_host, int n_slots, int (*standalone)(void *, int), void *ctx,
                  int64_t stats[4]);
}

namespace plb200 {
int emulate_routed(int n, int precision, const std::vector<AdjItem> &items, bool scaled, void *sv0, const RouteSpec &rs,
                   int (*standalone)(void *, int), void *ctx);
void emu_kind_hist(int64_t out[32], bool reset);
int64_t emu_jit_passes();
int64_t emu_plan_hits();
void emu_adjoint_schedule_stats(int n, int precision, const std::vector<AdjItem> &items, int64_t out[4]);
}
using namespace plb200;

namespace {
thread_local std::string g_err;

template <class T> std::vector<T> vec(const T *p, int64_t n) { return (p && n > 0) ? std::vector<T>(p, p + n) : std::vector<T>(); }

GateCall call_from_blob(const plb200_ops_t &b, int64_t i) {
    GateCall g;
    g.name = b.names[i];
    g.wires = vec(b.wires + b.wires_off[i], b.wires_off[i + 1] - b.wires_off[i]);
    g.ctrl_wires = vec(b.ctrl_wires + b.ctrl_off[i], b.ctrl_off[i + 1] - b.ctrl_off[i]);
    g.ctrl_values = vec(b.ctrl_values + b.ctrl_off[i], b.ctrl_off[i + 1] - b.ctrl_off[i]);
    g.params = vec(b.params + b.params_off[i], b.params_off[i + 1] - b.params_off[i]);
    g.inverse = b.inverses[i] != 0;
    if (b.mats && b.mats_off) {
        const int64_t m0 = b.mats_off[i], m1 = b.mats_off[i + 1];
        for (int64_t k = m0; k < m1; k++) g.matrix.emplace_back(b.mats[2 * k], b.mats[2 * k + 1]);
    }
    return g;
}

// Naive host application of one canonical op (only for the ops the scheduler leaves stand-alone).
template <typename C> int host_apply(const COp &op, int n, C *psi) {
    const uint64_t len = uint64_t{1} << n;
    uint64_t tmask = 0;
    for (int b : op.tbits) tmask |= uint64_t{1} << b;
    auto off = [&](uint32_t loc) {
        uint64_t o = 0;
        for (size_t j = 0; j < op.tbits.size(); j++)
            if (loc >> j & 1) o |= uint64_t{1} << op.tbits[j];
        return o;
    };
    using R = typename C::value_type;
    auto mul = [](cd m, C v) { return cd(v.real(), v.imag()) * m; };
    auto cast = [](cd v) { return C(static_cast<R>(v.real()), static_cast<R>(v.imag())); };
    switch (op.kind) {
    case OP_PAIRS: {
        if (op.parity) { // exp(-i theta/2 P): new[j] = c psi[j] + w sgn(j ^ x) psi[j ^ x]
            const uint64_t x = tmask, z = op.pmask, pivot = x & (~x + 1);
            const cd c = op.blocks[0].m[0], w = op.blocks[0].m[1];
            for (uint64_t j0 = 0; j0 < len; j0++) {
                if (j0 & pivot) continue;
                const uint64_t j1 = j0 ^ x;
                const cd a = cd(psi[j0].real(), psi[j0].imag()), b = cd(psi[j1].real(), psi[j1].imag());
                const cd w0 = (__builtin_popcountll(j1 & z) & 1) ? -w : w, w1 = (__builtin_popcountll(j0 & z) & 1) ? -w : w;
                psi[j0] = cast(c * a + w0 * b), psi[j1] = cast(c * b + w1 * a);
            }
            return 0;
        }
        for (uint64_t g = 0; g < len; g++) {
            if ((g & tmask) != 0 || (g & op.cmask) != op.cval) continue;
            for (const Block2 &bl : op.blocks) {
                const uint64_t ia = g | off(bl.a), ib = g | off(bl.b);
                const C a = psi[ia], b = psi[ib];
                psi[ia] = cast(mul(bl.m[0], a) + mul(bl.m[1], b));
                psi[ib] = cast(mul(bl.m[2], a) + mul(bl.m[3], b));
            }
        }
        return 0;
    }
    case OP_DIAG:
        for (uint64_t i = 0; i < len; i++) {
            if ((i & op.cmask) != op.cval) continue;
            cd d;
            if (op.parity) d = op.pd[__builtin_popcountll(i & op.pmask) & 1];
            else {
                unsigned t = 0;
                for (int j = 0; j < op.k(); j++) t |= static_cast<unsigned>((i >> op.tbits[j]) & 1) << j;
                d = op.diag[t];
            }
            psi[i] = cast(mul(d, psi[i]));
        }
        return 0;
    case OP_DENSE: {
        const int k = op.k(), D = 1 << k;
        std::vector<cd> in(D);
        for (uint64_t g = 0; g < len; g++) {
            if ((g & tmask) != 0 || (g & op.cmask) != op.cval) continue;
            for (int c = 0; c < D; c++) in[c] = cd(psi[g | off(c)].real(), psi[g | off(c)].imag());
            for (int r = 0; r < D; r++) {
                cd s = 0;
                for (int c = 0; c < D; c++) s += op.mat[static_cast<size_t>(r) * D + c] * in[c];
                psi[g | off(r)] = cast(s);
            }
        }
        return 0;
    }
    default: // OP_PROJECT
        for (uint64_t i = 0; i < len; i++)
            if ((i & op.cmask) != op.cval) psi[i] = C(0, 0);
        return 0;
    }
}

template <typename C> void host_pauli_inner(const C *a, const C *b, int n, const PauliWordMask &w, double out[2]) {
    const uint64_t len = uint64_t{1} << n;
    cd s = 0;
    for (uint64_t i = 0; i < len; i++) {
        if ((i & w.cmask) != w.cval) continue;
        const uint64_t j = i ^ w.x;
        cd t = std::conj(cd(a[i].real(), a[i].imag())) * cd(b[j].real(), b[j].imag());
        if (__builtin_popcountll(j & w.z) & 1) t = -t;
        s += t;
    }
    static const cd ipow[4] = {cd(1, 0), cd(0, 1), cd(-1, 0), cd(0, -1)};
    s *= ipow[w.ny & 3];
    out[0] = s.real(), out[1] = s.imag();
}

struct Ctx {
    int n, precision;
    const std::vector<AdjItem> *items;
    void *sv0, *sv1;
    double *acc;
};
int standalone_cb(void *p, int idx) {
    Ctx &c = *static_cast<Ctx *>(p);
    const AdjItem &it = (*c.items)[idx];
    if (it.overlap) {
        double r[2];
        if (c.precision == 64)
            host_pauli_inner(static_cast<std::complex<double> *>(c.sv1), static_cast<std::complex<double> *>(c.sv0), c.n,
                             it.pw, r);
        else
            host_pauli_inner(static_cast<std::complex<float> *>(c.sv1), static_cast<std::complex<float> *>(c.sv0), c.n,
                             it.pw, r);
        c.acc[it.slot] += r[1];
        return 0;
    }
    for (void *sv : {c.sv0, c.sv1}) {
        if (!sv) continue;
        if (c.precision == 64) host_apply(it.op, c.n, static_cast<std::complex<double> *>(sv));
        else host_apply(it.op, c.n, static_cast<std::complex<float> *>(sv));
    }
    return 0;
}
} // namespace

extern "C" {
const char *plb200_emu_last_error(void) { return g_err.c_str(); }
// passes executed through the g++-compiled specialised source (PLB200_EMU_JIT=1) so far
int64_t plb200_emu_jit_passes(void) { return emu_jit_passes(); }
// tapes whose schedule came from the plan cache so far
int64_t plb200_emu_plan_hits(void) { return emu_plan_hits(); }
// ops emitted by the pass encoder per interpreter kind (tile_exec.cuh enum) since the last reset
void plb200_emu_kind_histogram(int64_t *out32, int reset) { emu_kind_hist(out32, reset != 0); }

// schedule only (no state): fills the kind histogram and stats4 like plb200_schedule_stats
int plb200_emu_schedule(int64_t n, int precision, const plb200_ops_t *ops, int64_t *stats4) {
    try {
        std::vector<COp> all;
        for (int64_t i = 0; i < ops->n_ops; i++)
            for (auto &lo : lower_gate(n, call_from_blob(*ops, i))) all.push_back(std::move(lo));
        schedule_stats(static_cast<int>(n), precision, all, stats4);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// schedule of the backward adjoint sweep only (no states): stats4 as above, -1 in stats4[0] when a
// trainable generator is not a Pauli word (un-fused route)
int plb200_emu_adjoint_schedule(int64_t n, int precision, const plb200_ops_t *ops, const int64_t *trainable,
                                int64_t n_tp, int64_t *stats4) {
    try {
        std::vector<GateCall> calls(ops->n_ops);
        int64_t num_param_ops = 0;
        for (int64_t i = 0; i < ops->n_ops; i++) {
            calls[i] = call_from_blob(*ops, i);
            if (!calls[i].params.empty()) num_param_ops++;
        }
        std::vector<AdjItem> items;
        std::vector<double> sfs;
        if (!build_adjoint_items(n, calls, std::vector<int64_t>(trainable, trainable + n_tp), num_param_ops, items, sfs)) {
            stats4[0] = -1;
            return 0;
        }
        emu_adjoint_schedule_stats(static_cast<int>(n), precision, items, stats4);
        stats4[3] = static_cast<int64_t>(items.size()) * 1000000 + stats4[3] % 1000000; // items packed for reporting
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// state: 2^n interleaved complex (precision 64: double, 32: float), updated in place.
// stats4 = {tile passes, stand-alone ops, rounds, ops inside tile passes}
int plb200_emu_apply_ops(int64_t n, int precision, const plb200_ops_t *ops, void *state, int scaled, int64_t *stats4) {
    try {
        std::vector<AdjItem> items;
        for (int64_t i = 0; i < ops->n_ops; i++) {
            std::vector<COp> pieces;
            for (auto &lo : lower_gate(n, call_from_blob(*ops, i))) expand_for_fusion(lo, pieces, n >= (precision == 64 ? 12 : 14));
            for (auto &lo : pieces) {
                AdjItem it;
                it.op = std::move(lo);
                items.push_back(std::move(it));
            }
        }
        double dummy = 0;
        Ctx ctx{static_cast<int>(n), precision, &items, state, nullptr, &dummy};
        // the stand-alone accumulators are added after emulate_fused zeroes its own
        return emulate_fused(static_cast<int>(n), precision, items, false, scaled != 0, state, nullptr, &dummy, 0,
                             standalone_cb, &ctx, stats4);
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// Routed tape on host memory: like plb200_emu_apply_ops, but the last pass stores through the swap-out route
// (dst[p]: host array receiving the amplitudes whose swapped local bits read p).  *routed as in
// plb200_sv_apply_ops_route; when 0 the tape was applied to `state` in place.
int plb200_emu_apply_ops_route(int64_t n, int precision, const plb200_ops_t *ops, void *state, int64_t k,
                               const int64_t *lbits, int64_t my_value, void *const *dst, int *routed) {
    try {
        std::vector<AdjItem> items;
        for (int64_t i = 0; i < ops->n_ops; i++) {
            std::vector<COp> pieces;
            for (auto &lo : lower_gate(n, call_from_blob(*ops, i))) expand_for_fusion(lo, pieces, n >= (precision == 64 ? 12 : 14));
            for (auto &lo : pieces) {
                AdjItem it;
                it.op = std::move(lo);
                items.push_back(std::move(it));
            }
        }
        RouteSpec rs;
        rs.k = static_cast<int>(k), rs.my_value = static_cast<int>(my_value);
        for (int64_t i = 0; i < k; i++) rs.lbits[i] = static_cast<int>(lbits[i]);
        for (int p = 0; p < (1 << k); p++) rs.dst[p] = dst[p];
        double dummy = 0;
        Ctx ctx{static_cast<int>(n), precision, &items, state, nullptr, &dummy};
        const int r = emulate_routed(static_cast<int>(n), precision, items, true, state, rs, standalone_cb, &ctx);
        if (r < 0) return 1;
        *routed = r;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// Backward adjoint sweep on prepared lambda (= U psi) and hl (= H lambda); jac[n_tp] = -2 sf Im<hl|G|lambda>.
// Returns 2 when some trainable generator is not a (controlled) Pauli word (the engine then takes its
// un-fused route).
int plb200_emu_adjoint_sweep(int64_t n, int precision, const plb200_ops_t *ops, const int64_t *trainable, int64_t n_tp,
                             void *lambda, void *hl, int scaled, double *jac, int64_t *stats4) {
    try {
        std::vector<GateCall> calls(ops->n_ops);
        int64_t num_param_ops = 0;
        for (int64_t i = 0; i < ops->n_ops; i++) {
            calls[i] = call_from_blob(*ops, i);
            if (!calls[i].params.empty()) num_param_ops++;
        }
        std::vector<AdjItem> items;
        std::vector<double> sfs;
        if (!build_adjoint_items(n, calls, std::vector<int64_t>(trainable, trainable + n_tp), num_param_ops, items, sfs))
            return 2;
        std::vector<double> acc(n_tp, 0.0), acc_alone(n_tp, 0.0);
        Ctx ctx{static_cast<int>(n), precision, &items, lambda, hl, acc_alone.data()};
        if (int rc = emulate_fused(static_cast<int>(n), precision, items, true, scaled != 0, lambda, hl, acc.data(),
                                   static_cast<int>(n_tp), standalone_cb, &ctx, stats4))
            return rc;
        for (int64_t p = 0; p < n_tp; p++) jac[p] = -2.0 * sfs[p] * (acc[p] + acc_alone[p]);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
}
