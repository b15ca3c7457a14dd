// Tape scheduling: cache-blocked ("tile") execution of a canonical-op list (fusion.cu).
#pragma once
#include <vector>

#include "device.cuh"

namespace plb200 {
// Applies `ops` in order with as few HBM sweeps as the scheduler finds; arithmetic per gate is
// identical to launch_ops().
void run_fused(StateVec &sv, const std::vector<COp> &ops);
// Host-only: out = {tile passes, stand-alone kernels, rounds, ops executed inside tile passes}
void schedule_stats(int n, int precision, const std::vector<COp> &ops, int64_t out[4]);
} // namespace plb200
