// Run-time half of the pass specialiser (jit_codegen.hpp): NVRTC -> cubin for sm_100a -> driver-API module,
// a structure-keyed kernel cache, background compilation and an optional on-disk cubin cache.
//
// libnvrtc and libcuda are dlopen'ed on first use: libplb200.so itself links neither, so it loads (and its
// symbol table can be checked) on a machine without a driver.  When NVRTC is not present the fused path
// simply keeps running the interpreter kernel (tile_kernel) — still CUDA, never a CPU path.
//
// Tiering (PLB200_JIT): "async" (default) — a pass structure is compiled in the background the SECOND time it
// is seen (one-shot circuits never pay a compile) and the interpreter runs it until the kernel is ready;
// "sync" — compiled at first sight, blocking; "0" — interpreter only.  plb200_jit_wait() blocks until the
// queue is empty (bench.py calls it after its warm-up steps, outside the timed region).
#include "jit_runtime.hpp"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace plb200 {
namespace jit {

namespace {

// ---- NVRTC, by name
using nvrtcProgram = struct _nvrtcProgram *;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    int (*Version)(int *, int *) = nullptr;
    bool ok = false;
    std::string why;
};
// ---- the driver API, by name
using CUmodule = struct CUmod_st *;
using CUfunction = struct CUfunc_st *;
using CUstream = struct CUstream_st *;
struct Driver {
    int (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    int (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    int (*FuncSetAttribute)(CUfunction, int, int) = nullptr;
    int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **,
                        void **) = nullptr;
    int (*GetErrorString)(int, const char **) = nullptr;
    bool ok = false;
    std::string why;
};

template <class F> bool sym(void *h, const char *name, F &f) {
    f = reinterpret_cast<F>(dlsym(h, name));
    return f != nullptr;
}

const Nvrtc &nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        void *h = nullptr;
        if (const char *e = std::getenv("PLB200_NVRTC_LIB")) h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
        for (const char *nm : names)
            if (!h) h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            r.why = "libnvrtc not found";
            return r;
        }
        r.ok = sym(h, "nvrtcCreateProgram", r.CreateProgram) && sym(h, "nvrtcCompileProgram", r.CompileProgram) &&
               sym(h, "nvrtcGetCUBINSize", r.GetCUBINSize) && sym(h, "nvrtcGetCUBIN", r.GetCUBIN) &&
               sym(h, "nvrtcGetProgramLogSize", r.GetProgramLogSize) && sym(h, "nvrtcGetProgramLog", r.GetProgramLog) &&
               sym(h, "nvrtcDestroyProgram", r.DestroyProgram) && sym(h, "nvrtcVersion", r.Version);
        if (!r.ok) r.why = "libnvrtc lacks a required entry point";
        return r;
    }();
    return n;
}

const Driver &driver() {
    static Driver d = [] {
        Driver r;
        void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            r.why = "libcuda.so.1 not found";
            return r;
        }
        r.ok = sym(h, "cuModuleLoadData", r.ModuleLoadData) && sym(h, "cuModuleGetFunction", r.ModuleGetFunction) &&
               sym(h, "cuFuncSetAttribute", r.FuncSetAttribute) && sym(h, "cuLaunchKernel", r.LaunchKernel) &&
               sym(h, "cuGetErrorString", r.GetErrorString);
        if (!r.ok) r.why = "libcuda lacks a required entry point";
        return r;
    }();
    return d;
}

std::string cu_err(int rc) {
    const char *s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(rc, &s);
    return s ? s : ("CUDA driver error " + std::to_string(rc));
}

// ---- cache
enum class St { Seen, Queued, Ready, Failed };
struct Entry {
    St st = St::Seen;
    int seen = 0;
    std::string src;          // kept while queued
    std::vector<char> cubin;  // once ready
    CUfunction fn[16] = {nullptr};
    int smem_set[16] = {0};
};

// never destroyed: worker threads may still be inside NVRTC when the process exits
std::mutex &g_mu = *new std::mutex;
std::condition_variable &g_cv_work = *new std::condition_variable, &g_cv_idle = *new std::condition_variable;
auto &g_cache = *new std::unordered_map<uint64_t, std::unique_ptr<Entry>>;
auto &g_queue = *new std::deque<uint64_t>;
int g_inflight = 0;
auto &g_workers = *new std::vector<std::thread>;
bool g_stop = false;
std::atomic<int64_t> g_stat_compiled{0}, g_stat_hits{0}, g_stat_interp{0}, g_stat_failed{0}, g_stat_disk{0};
std::atomic<int64_t> g_stat_compile_us{0};
std::atomic<int> g_mode_override{-1};

std::string cache_dir() {
    if (const char *e = std::getenv("PLB200_JIT_CACHE_DIR")) return e;
    const char *home = std::getenv("HOME");
    return std::string(home ? home : "/tmp") + "/.cache/plb200_jit";
}
bool disk_cache_enabled() {
    const char *e = std::getenv("PLB200_JIT_DISK_CACHE");
    return !(e && e[0] == '0');
}
std::string cache_path(uint64_t key, size_t len) {
    int maj = 0, min = 0;
    if (nvrtc().ok) nvrtc().Version(&maj, &min);
    char b[96];
    snprintf(b, sizeof(b), "/%016llx_%zu_nvrtc%d.%d.cubin", static_cast<unsigned long long>(key), len, maj, min);
    return cache_dir() + b;
}

bool compile_source(const std::string &src, std::vector<char> &cubin, std::string &log) {
    const Nvrtc &n = nvrtc();
    if (!n.ok) {
        log = n.why;
        return false;
    }
    nvrtcProgram prog = nullptr;
    if (n.CreateProgram(&prog, src.c_str(), "plb_pass.cu", 0, nullptr, nullptr) != 0) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=true"};
    const int rc = n.CompileProgram(prog, 4, opts);
    if (rc != 0) {
        size_t ls = 0;
        n.GetProgramLogSize(prog, &ls);
        log.resize(ls);
        if (ls) n.GetProgramLog(prog, log.data());
        n.DestroyProgram(&prog);
        return false;
    }
    size_t sz = 0;
    n.GetCUBINSize(prog, &sz);
    cubin.resize(sz);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    return sz > 0;
}

void compile_entry(uint64_t key) {
    std::string src;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        src = g_cache[key]->src;
    }
    std::vector<char> cubin;
    std::string log;
    bool ok = false, from_disk = false;
    const std::string path = cache_path(key, src.size());
    if (disk_cache_enabled()) {
        std::ifstream f(path, std::ios::binary);
        if (f) {
            cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            ok = from_disk = !cubin.empty();
        }
    }
    if (!ok) {
        const auto t0 = std::chrono::steady_clock::now();
        ok = compile_source(src, cubin, log);
        g_stat_compile_us += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (ok && disk_cache_enabled()) {
            const std::string dir = cache_dir();
            ::mkdir((dir.substr(0, dir.rfind('/'))).c_str(), 0755);
            ::mkdir(dir.c_str(), 0755);
            const std::string tmp = path + ".tmp" + std::to_string(::getpid());
            std::ofstream o(tmp, std::ios::binary);
            if (o) {
                o.write(cubin.data(), static_cast<std::streamsize>(cubin.size()));
                o.close();
                ::rename(tmp.c_str(), path.c_str());
            }
        }
    }
    if (!ok && std::getenv("PLB200_JIT_VERBOSE")) std::fprintf(stderr, "[plb200 jit] compile failed:\n%s\n", log.c_str());
    std::lock_guard<std::mutex> lk(g_mu);
    Entry &e = *g_cache[key];
    e.src.clear();
    e.src.shrink_to_fit();
    if (ok) {
        e.cubin = std::move(cubin);
        e.st = St::Ready;
        (from_disk ? g_stat_disk : g_stat_compiled)++;
    } else {
        e.st = St::Failed;
        g_stat_failed++;
    }
}

void worker() {
    for (;;) {
        uint64_t key;
        {
            std::unique_lock<std::mutex> lk(g_mu);
            g_cv_work.wait(lk, [] { return g_stop || !g_queue.empty(); });
            if (g_stop) return;
            key = g_queue.front();
            g_queue.pop_front();
            g_inflight++;
        }
        compile_entry(key);
        {
            std::lock_guard<std::mutex> lk(g_mu);
            g_inflight--;
        }
        g_cv_idle.notify_all();
    }
}

void ensure_workers() { // g_mu held
    if (!g_workers.empty()) return;
    unsigned n = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("PLB200_JIT_THREADS")) n = static_cast<unsigned>(std::atoi(e));
    n = std::max(1u, std::min(n, 32u));
    for (unsigned i = 0; i < n; i++) g_workers.emplace_back(worker);
    // the threads are detached at exit: a process that ends while a compile runs must not block on it
    std::atexit([] {
        {
            std::lock_guard<std::mutex> lk(g_mu);
            g_stop = true;
        }
        g_cv_work.notify_all();
        for (auto &t : g_workers) t.detach();
    });
}

} // namespace

Mode mode() {
    const int o = g_mode_override.load();
    if (o >= 0) return static_cast<Mode>(o);
    static const Mode m = [] {
        const char *e = std::getenv("PLB200_JIT");
        if (!e) return Mode::Async;
        if (!std::strcmp(e, "0") || !std::strcmp(e, "off")) return Mode::Off;
        if (!std::strcmp(e, "sync") || !std::strcmp(e, "2")) return Mode::Sync;
        return Mode::Async;
    }();
    return m;
}
void set_mode(int m) { g_mode_override.store(m); }

int min_qubits() {
    const char *e = std::getenv("PLB200_JIT_MIN_QUBITS");
    return e ? std::atoi(e) : 20;
}

bool available(std::string *why) {
    if (!nvrtc().ok) {
        if (why) *why = nvrtc().why;
        return false;
    }
    if (!driver().ok) {
        if (why) *why = driver().why;
        return false;
    }
    return true;
}

Kernel lookup(const std::string &src, int device, size_t smem_bytes, bool force_sync) {
    Kernel k;
    if (src.empty() || (mode() == Mode::Off && !force_sync) || !available(nullptr) || device < 0 || device >= 16) {
        g_stat_interp++;
        return k;
    }
    const uint64_t key = fnv1a(src) ^ (static_cast<uint64_t>(src.size()) << 40);
    std::unique_lock<std::mutex> lk(g_mu);
    auto &slot = g_cache[key];
    if (!slot) slot = std::make_unique<Entry>();
    Entry &e = *slot;
    e.seen++;
    const bool sync = force_sync || mode() == Mode::Sync;
    if (e.st == St::Seen && (sync || e.seen >= 2)) {
        e.src = src;
        e.st = St::Queued;
        if (sync) {
            lk.unlock();
            compile_entry(key);
            g_cv_idle.notify_all();
            lk.lock();
        } else {
            ensure_workers();
            g_queue.push_back(key);
            g_cv_work.notify_one();
        }
    }
    if (sync && e.st == St::Queued) // another thread (or a background worker) is compiling it
        g_cv_idle.wait(lk, [&] { return e.st != St::Queued; });
    if (e.st != St::Ready) {
        g_stat_interp++;
        return k;
    }
    if (!e.fn[device]) { // load into the calling thread's current context (the caller has set the device)
        CUmodule mod = nullptr;
        int rc = driver().ModuleLoadData(&mod, e.cubin.data());
        if (rc == 0) rc = driver().ModuleGetFunction(&e.fn[device], mod, "plb_pass");
        if (rc == 0) rc = driver().FuncSetAttribute(e.fn[device], 8 /* MAX_DYNAMIC_SHARED_SIZE_BYTES */, static_cast<int>(smem_bytes));
        if (rc == 0) rc = driver().FuncSetAttribute(e.fn[device], 9 /* PREFERRED_SHARED_MEMORY_CARVEOUT */, 100);
        if (rc != 0) {
            if (std::getenv("PLB200_JIT_VERBOSE")) std::fprintf(stderr, "[plb200 jit] module load failed: %s\n", cu_err(rc).c_str());
            e.fn[device] = nullptr;
            e.st = St::Failed;
            g_stat_failed++;
            g_stat_interp++;
            return k;
        }
    }
    g_stat_hits++;
    k.fn = e.fn[device];
    return k;
}

void launch(const Kernel &k, unsigned grid, unsigned block, size_t smem_bytes, void *stream, void *sv, const void *pass_params,
            const void *route_params) {
    void *args[3] = {&sv, const_cast<void *>(pass_params), const_cast<void *>(route_params)};
    const int rc = driver().LaunchKernel(static_cast<CUfunction>(k.fn), grid, 1, 1, block, 1, 1, static_cast<unsigned>(smem_bytes),
                                         static_cast<CUstream>(stream), args, nullptr);
    if (rc != 0) fail("JIT pass kernel launch failed: " + cu_err(rc));
}

void launch_args(const Kernel &k, unsigned grid, unsigned block, size_t smem_bytes, void *stream, void **args) {
    const int rc = driver().LaunchKernel(static_cast<CUfunction>(k.fn), grid, 1, 1, block, 1, 1, static_cast<unsigned>(smem_bytes),
                                         static_cast<CUstream>(stream), args, nullptr);
    if (rc != 0) fail("JIT pass kernel launch failed: " + cu_err(rc));
}

void wait_idle() {
    std::unique_lock<std::mutex> lk(g_mu);
    g_cv_idle.wait(lk, [] { return g_queue.empty() && g_inflight == 0; });
}

void stats(int64_t out[8]) {
    out[0] = g_stat_compiled, out[1] = g_stat_disk, out[2] = g_stat_hits, out[3] = g_stat_interp, out[4] = g_stat_failed;
    out[5] = g_stat_compile_us;
    std::lock_guard<std::mutex> lk(g_mu);
    out[6] = static_cast<int64_t>(g_queue.size()) + g_inflight;
    out[7] = static_cast<int64_t>(g_cache.size());
}

bool compile_only(const std::string &src, std::string &log, size_t *cubin_bytes) {
    std::vector<char> cubin;
    const bool ok = compile_source(src, cubin, log);
    if (cubin_bytes) *cubin_bytes = cubin.size();
    return ok;
}

} // namespace jit
} // namespace plb200
