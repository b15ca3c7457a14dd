#include "fusion.hpp"

namespace plb200 {
void run_fused(StateVec &sv, const std::vector<COp> &ops) { launch_ops(sv, ops); }
} // namespace plb200
