// Tape scheduling: cache-blocked ("tile") execution of a canonical-op list (fusion.cu).
#pragma once
#include <string>
#include <vector>

#include "device.cuh"

namespace plb200 {
// Applies `ops` in order with as few HBM sweeps as the scheduler finds; arithmetic per gate is
// identical to launch_ops().
void run_fused(StateVec &sv, const std::vector<COp> &ops);
// What the fused path schedules instead of `op`: the op itself, or — for a diagonal the tile encoder has no form
// for whose entries are all equal but a few — a scalar on its control subspace and one controlled phase per
// exceptional entry; for a Pauli rotation with X / Y letters — when `tiles` says the state is large enough for tile
// passes — basis changes (H, S) around a parity diagonal (appended to `out`).
void expand_for_fusion(const COp &op, std::vector<COp> &out, bool tiles = true);
// Host-only: out = {tile passes, stand-alone kernels, rounds, ops executed inside tile passes}
void schedule_stats(int n, int precision, const std::vector<COp> &ops, int64_t out[4]);

// Swap-out route of the sharded mode: exchange the k local index bits `lbits` with k global (rank) bits while
// the last pass of the tape stores its tiles.  dst[p] = slab that receives the amplitudes whose swapped local
// bits read p (this rank's own ping-pong slab for p == my_value, a peer-mapped slab otherwise); my_value = this
// rank's values of the swapped global bits.
struct RouteSpec {
    int k = 0;
    int lbits[3] = {0, 0, 0};
    int my_value = 0;
    void *dst[8] = {nullptr};
};
// Applies the tape like run_fused(); returns true when the final state left through the route.
bool run_fused_routed(StateVec &sv, const std::vector<COp> &ops, const RouteSpec &rs);

// Host-only: CUDA source of the specialised kernel of every tile pass of the tape (jit_codegen.hpp)
void pass_sources(int n, int precision, const std::vector<COp> &ops, std::vector<std::string> &out);

// One step of the backward adjoint sweep: either "apply this (already inverted) op to lambda and
// H lambda" or "accumulate Im<H lambda| P |lambda> into slot" for a (controlled) Pauli word P.
struct AdjItem {
    bool overlap = false;
    COp op;
    PauliWordMask pw;
    int slot = -1;
};
// Backward-sweep item list of the adjoint method (AdjointJacobianLQubit.hpp:269-314): for every op,
// last to first, [overlap with its generator if it is the next trainable one] then [its inverse].
// sfs[p] = generator scale factor x (-1 for inverse ops).  Returns false when a trainable generator is
// not a (controlled) Pauli word (the caller then uses the copy + generator + dot route).
bool build_adjoint_items(int64_t n, const std::vector<GateCall> &calls, const std::vector<int64_t> &tp,
                         int64_t num_param_ops, std::vector<AdjItem> &items, std::vector<double> &sfs);
// Runs the items in order on the pair (lambda, H lambda) with tile passes over both states;
// acc_host[slot] receives Im<H lambda|P|lambda>.  stats = {tile passes, stand-alone items, fused items}.
// scratch: the state whose plan buffer holds the accumulators (a persistent one saves an allocation per call).
void run_adjoint_fused(StateVec &lambda, StateVec &hl, const std::vector<AdjItem> &items, int n_slots,
                       double *acc_host, int64_t stats[3], StateVec *scratch = nullptr);
} // namespace plb200
