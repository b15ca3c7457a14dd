// Run-time half of the pass specialiser: compile cache + launch (jit_runtime.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "common.hpp"

namespace plb200 {
namespace jit {

enum class Mode : int { Off = 0, Async = 1, Sync = 2 };
Mode mode();            // PLB200_JIT = 0 | async (default) | sync, or the value of set_mode()
void set_mode(int m);   // -1: back to the environment's choice
int min_qubits();       // PLB200_JIT_MIN_QUBITS (default 20): smaller states always run the interpreter
bool available(std::string *why = nullptr); // NVRTC and the driver API could be loaded

struct Kernel {
    void *fn = nullptr;
    explicit operator bool() const { return fn != nullptr; }
};
// The compiled kernel for this source on `device` (whose context must be current), or an empty handle when
// it is not (yet) available — the caller then runs the interpreter kernel.
// force_sync: compile now even in the background mode (routed passes have no interpreter equivalent)
Kernel lookup(const std::string &src, int device, size_t smem_bytes, bool force_sync = false);
void launch(const Kernel &k, unsigned grid, unsigned block, size_t smem_bytes, void *stream, void *sv,
            const void *pass_params, const void *route_params = nullptr);
// generic form: args[] as cuLaunchKernel takes them (adjoint passes: two states + accumulators + description)
void launch_args(const Kernel &k, unsigned grid, unsigned block, size_t smem_bytes, void *stream, void **args);
void wait_idle();
// {compiled, loaded from disk, launches of compiled kernels, launches left to the interpreter, failed,
//  compile microseconds, queued + in flight, structures seen}
void stats(int64_t out[8]);
// compile without caching / loading (tests, tools): log receives the NVRTC log on failure
bool compile_only(const std::string &src, std::string &log, size_t *cubin_bytes);

inline uint64_t fnv1a(const std::string &s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    return h;
}

} // namespace jit
} // namespace plb200
