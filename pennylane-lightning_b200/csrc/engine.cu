// C ABI of the plb200 engine (include/plb200.h): state-vector lifecycle, gate application,
// observables, measurements, sampling and the adjoint-Jacobian sweep.  Host orchestration only;
// the arithmetic lives in gate_kernels.cu / measure_kernels.cu / fused_kernels.cu.
#include "../../include/plb200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <stack>

#include "device.cuh"
#include "fusion.hpp"
#include "jit_runtime.hpp"

using namespace plb200;

struct plb200_sv {
    StateVec s;
};

struct plb200_obs {
    enum Kind { NAMED, HERMITIAN, TENSOR, HAMILTONIAN, SPARSE } kind = NAMED;
    // SPARSE: CSR over the full 2^n index space (SparseHamiltonianBase, Observables.hpp:592-699)
    std::vector<int64_t> indptr, indices;
    std::vector<cd> values;
    std::string name;
    std::vector<int64_t> wires;
    std::vector<double> params;
    std::vector<cd> matrix;
    std::vector<std::shared_ptr<plb200_obs>> terms;
    std::vector<double> coeffs;
};

namespace {

thread_local std::string g_err;

#define ABI_TRY try {
#define ABI_CATCH                                                                                        \
    }                                                                                                    \
    catch (const std::exception &e) {                                                                    \
        g_err = e.what();                                                                                \
        return 1;                                                                                        \
    }                                                                                                    \
    catch (...) {                                                                                        \
        g_err = "unknown error";                                                                         \
        return 1;                                                                                        \
    }                                                                                                    \
    return 0;

std::vector<int64_t> vi(const int64_t *p, int64_t n) { return n > 0 ? std::vector<int64_t>(p, p + n) : std::vector<int64_t>{}; }
std::vector<uint8_t> vb(const uint8_t *p, int64_t n) { return n > 0 ? std::vector<uint8_t>(p, p + n) : std::vector<uint8_t>{}; }
std::vector<double> vd(const double *p, int64_t n) { return n > 0 ? std::vector<double>(p, p + n) : std::vector<double>{}; }
std::vector<cd> vc(const double *p, int64_t n) {
    std::vector<cd> v(n > 0 ? n : 0);
    for (int64_t i = 0; i < n; i++) v[i] = cd(p[2 * i], p[2 * i + 1]);
    return v;
}

void init_sv(StateVec &s, int64_t n, int precision, int device, void *stream, void *ext) {
    PLB_CHECK(n >= 0 && n <= 40, "Invalid number of qubits");
    PLB_CHECK(precision == 32 || precision == 64, "precision must be 32 (c64) or 64 (c128)");
    int count = 0;
    PLB_CUDA(cudaGetDeviceCount(&count));
    PLB_CHECK(device >= 0 && device < count, "Invalid CUDA device");
    s.n = n;
    s.precision = precision;
    s.device = device;
    s.stream = static_cast<cudaStream_t>(stream);
    s.set_device();
    cudaDeviceProp prop;
    PLB_CUDA(cudaGetDeviceProperties(&prop, device));
    PLB_CHECK(prop.major >= 10, "plb200 requires an sm_100 (Blackwell) device; no fallback path exists");
    s.sm_count = prop.multiProcessorCount;
    if (ext) {
        s.data = ext;
        s.owned = false;
    } else {
        PLB_CUDA(cudaMalloc(&s.data, s.bytes()));
        s.owned = true;
        PLB_CUDA(cudaMemsetAsync(s.data, 0, s.bytes(), s.stream));
        const double one64[2] = {1.0, 0.0};
        const float one32[2] = {1.0f, 0.0f};
        PLB_CUDA(cudaMemcpyAsync(s.data, precision == 64 ? static_cast<const void *>(one64) : one32, s.elem_bytes(),
                                 cudaMemcpyHostToDevice, s.stream));
        PLB_CUDA(cudaStreamSynchronize(s.stream));
    }
}

void free_sv(StateVec &s) {
    cudaSetDevice(s.device);
    if (s.owned && s.data) cudaFree(s.data);
    if (s.alt) cudaFree(s.alt);
    s.alt = nullptr;
    if (s.red) cudaFree(s.red);
    if (s.tbl) cudaFree(s.tbl);
    if (s.plan) cudaFree(s.plan);
    s.data = nullptr, s.red = nullptr, s.tbl = nullptr, s.plan = nullptr;
}

// temporary state sharing device/stream/precision with `like`
struct TempState {
    StateVec s;
    explicit TempState(const StateVec &like, bool copy) {
        s.n = like.n, s.precision = like.precision, s.device = like.device, s.stream = like.stream;
        s.sm_count = like.sm_count;
        s.set_device();
        PLB_CUDA(cudaMalloc(&s.data, s.bytes()));
        s.owned = true;
        if (copy) PLB_CUDA(cudaMemcpyAsync(s.data, like.data, s.bytes(), cudaMemcpyDeviceToDevice, s.stream));
    }
    ~TempState() { free_sv(s); }
    TempState(const TempState &) = delete;
    TempState &operator=(const TempState &) = delete;
};

void copy_state(StateVec &dst, const StateVec &src) {
    PLB_CHECK(dst.n == src.n && dst.precision == src.precision, "state vectors are incompatible");
    dst.set_device();
    PLB_CUDA(cudaMemcpyAsync(dst.data, src.data, dst.bytes(), cudaMemcpyDeviceToDevice, dst.stream));
}

GateCall make_call(const char *name, const int64_t *cw, const uint8_t *cv, int64_t nc, const int64_t *w, int64_t nw,
                   int inverse, const double *params, int64_t np) {
    GateCall g;
    g.name = name ? name : "";
    g.wires = vi(w, nw);
    g.ctrl_wires = vi(cw, nc);
    g.ctrl_values = vb(cv, nc);
    g.params = vd(params, np);
    g.inverse = inverse != 0;
    return g;
}

GateCall call_from_blob(const plb200_ops_t &b, int64_t i) {
    GateCall g = make_call(b.names[i], b.ctrl_wires + b.ctrl_off[i], b.ctrl_values + b.ctrl_off[i],
                           b.ctrl_off[i + 1] - b.ctrl_off[i], b.wires + b.wires_off[i],
                           b.wires_off[i + 1] - b.wires_off[i], b.inverses[i], b.params + b.params_off[i],
                           b.params_off[i + 1] - b.params_off[i]);
    if (b.mats && b.mats_off) g.matrix = vc(b.mats + 2 * b.mats_off[i], b.mats_off[i + 1] - b.mats_off[i]);
    return g;
}

void apply_call(StateVec &s, const GateCall &g) { launch_ops(s, lower_gate(s.n, g)); }

// ------------------------------------------------------------------------- observables
struct PauliTerm {
    double coeff;
    PauliWordMask w;
};

// Expand an observable into sum_k coeff_k P_k if it is built from Pauli/Identity/Hadamard named
// observables; returns false for Hermitian factors, other named gates or > max_terms terms.
bool expand_pauli(int64_t n, const plb200_obs &o, std::vector<PauliTerm> &out, size_t max_terms) {
    auto bit = [&](int64_t w) {
        PLB_CHECK(w >= 0 && w < n, "Invalid wire index");
        return uint64_t{1} << (n - 1 - w);
    };
    switch (o.kind) {
    case plb200_obs::NAMED: {
        if (o.wires.size() != 1) return false;
        PauliTerm t{1.0, {}};
        const uint64_t b = bit(o.wires[0]);
        if (o.name == "Identity") {
        } else if (o.name == "PauliX") t.w.x = b;
        else if (o.name == "PauliY") t.w.x = b, t.w.z = b, t.w.ny = 1;
        else if (o.name == "PauliZ") t.w.z = b;
        else if (o.name == "Hadamard") {
            PauliTerm tx{M_SQRT1_2, {}}, tz{M_SQRT1_2, {}};
            tx.w.x = b, tz.w.z = b;
            out.push_back(tx), out.push_back(tz);
            return true;
        } else
            return false;
        out.push_back(t);
        return true;
    }
    case plb200_obs::HERMITIAN:
        return false;
    case plb200_obs::TENSOR: {
        std::vector<PauliTerm> acc{{1.0, {}}};
        for (const auto &t : o.terms) {
            std::vector<PauliTerm> f;
            if (!expand_pauli(n, *t, f, max_terms)) return false;
            std::vector<PauliTerm> next;
            for (const auto &a : acc)
                for (const auto &b : f) {
                    if ((a.w.x | a.w.z) & (b.w.x | b.w.z)) return false; // overlapping wires: not a plain word
                    PauliTerm p{a.coeff * b.coeff, {}};
                    p.w.x = a.w.x | b.w.x, p.w.z = a.w.z | b.w.z, p.w.ny = a.w.ny + b.w.ny;
                    next.push_back(p);
                }
            if (next.size() > max_terms) return false;
            acc.swap(next);
        }
        out.insert(out.end(), acc.begin(), acc.end());
        return true;
    }
    case plb200_obs::HAMILTONIAN: {
        for (size_t k = 0; k < o.terms.size(); k++) {
            std::vector<PauliTerm> f;
            if (!expand_pauli(n, *o.terms[k], f, max_terms)) return false;
            for (auto &t : f) t.coeff *= o.coeffs[k];
            out.insert(out.end(), f.begin(), f.end());
            if (out.size() > max_terms) return false;
        }
        return true;
    }
    }
    return false;
}

void obs_apply(const plb200_obs &o, StateVec &s);

// device copy of a CSR matrix, alive for one call
struct DeviceCsr {
    int64_t *indptr = nullptr, *indices = nullptr;
    void *vals = nullptr;
    int64_t nnz = 0;
    DeviceCsr(const StateVec &s, const int64_t *ip, const int64_t *ix, const cd *v, int64_t nrows) {
        PLB_CHECK(static_cast<uint64_t>(nrows) == s.length(), "sparse matrix dimension must equal the state-vector length");
        nnz = ip[nrows];
        s.set_device();
        PLB_CUDA(cudaMalloc(&indptr, sizeof(int64_t) * (nrows + 1)));
        PLB_CUDA(cudaMalloc(&indices, sizeof(int64_t) * std::max<int64_t>(nnz, 1)));
        PLB_CUDA(cudaMalloc(&vals, sizeof(cd) * std::max<int64_t>(nnz, 1)));
        PLB_CUDA(cudaMemcpyAsync(indptr, ip, sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice, s.stream));
        PLB_CUDA(cudaMemcpyAsync(indices, ix, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, s.stream));
        PLB_CUDA(cudaMemcpyAsync(vals, v, sizeof(cd) * nnz, cudaMemcpyHostToDevice, s.stream));
    }
    ~DeviceCsr() {
        cudaFree(indptr), cudaFree(indices), cudaFree(vals);
    }
    DeviceCsr(const DeviceCsr &) = delete;
    DeviceCsr &operator=(const DeviceCsr &) = delete;
};
// dst = A src
void sparse_apply_to(const plb200_obs &o, StateVec &dst, const StateVec &src) {
    DeviceCsr d(src, o.indptr.data(), o.indices.data(), o.values.data(), static_cast<int64_t>(o.indptr.size()) - 1);
    csr_apply(dst, src, d.indptr, d.indices, d.vals, d.nnz);
    dst.sync(); // the device copy of the matrix dies with this scope
}

void hamiltonian_apply_generic(const plb200_obs &o, StateVec &s) {
    // sum_k c_k O_k |s>  (ObservablesLQubit.hpp:156-199): accumulator + one scratch copy
    TempState orig(s, true), tmp(s, false);
    PLB_CUDA(cudaMemsetAsync(s.data, 0, s.bytes(), s.stream));
    for (size_t k = 0; k < o.terms.size(); k++) {
        copy_state(tmp.s, orig.s);
        obs_apply(*o.terms[k], tmp.s);
        s.launches += tmp.s.launches, tmp.s.launches = 0;
        axpy(s, cd(o.coeffs[k], 0.0), tmp.s);
    }
    s.sync();
}

void obs_apply(const plb200_obs &o, StateVec &s) {
    switch (o.kind) {
    case plb200_obs::NAMED: {
        GateCall g;
        g.name = o.name, g.wires = o.wires, g.params = o.params;
        apply_call(s, g);
        break;
    }
    case plb200_obs::HERMITIAN:
        launch_ops(s, lower_matrix(s.n, o.matrix, o.wires, {}, {}, false, true));
        break;
    case plb200_obs::TENSOR:
        for (const auto &t : o.terms) obs_apply(*t, s);
        break;
    case plb200_obs::SPARSE: {
        TempState in(s, true);
        sparse_apply_to(o, s, in.s);
        s.launches += in.s.launches;
        break;
    }
    case plb200_obs::HAMILTONIAN: {
        std::vector<PauliTerm> terms;
        if (expand_pauli(s.n, o, terms, 1 << 16)) {
            TempState in(s, true);
            std::vector<PauliWordMask> w(terms.size());
            std::vector<double> c(terms.size());
            for (size_t i = 0; i < terms.size(); i++) w[i] = terms[i].w, c[i] = terms[i].coeff;
            pauli_sum_apply(s, in.s, w.data(), c.data(), static_cast<int64_t>(w.size()));
            s.sync();
        } else
            hamiltonian_apply_generic(o, s);
        break;
    }
    }
}

// out-of-place: dst = O src, without touching src
void obs_apply_to(const plb200_obs &o, StateVec &dst, const StateVec &src) {
    if (o.kind == plb200_obs::SPARSE) {
        sparse_apply_to(o, dst, src);
        return;
    }
    std::vector<PauliTerm> terms;
    if (expand_pauli(src.n, o, terms, 1 << 16) && terms.size() > 1) {
        std::vector<PauliWordMask> w(terms.size());
        std::vector<double> c(terms.size());
        for (size_t i = 0; i < terms.size(); i++) w[i] = terms[i].w, c[i] = terms[i].coeff;
        pauli_sum_apply(dst, src, w.data(), c.data(), static_cast<int64_t>(w.size()));
        return;
    }
    copy_state(dst, src);
    obs_apply(o, dst);
}

double expval_terms(StateVec &s, const std::vector<PauliTerm> &terms) {
    std::vector<PauliWordMask> w(terms.size());
    for (size_t i = 0; i < terms.size(); i++) w[i] = terms[i].w;
    std::vector<double> r(2 * terms.size());
    pauli_inner(s, s, w.data(), static_cast<int64_t>(w.size()), r.data());
    double e = 0;
    for (size_t i = 0; i < terms.size(); i++) e += terms[i].coeff * r[2 * i];
    return e;
}

double expval_obs(StateVec &s, const plb200_obs &o) {
    std::vector<PauliTerm> terms;
    if (expand_pauli(s.n, o, terms, 1 << 16)) return expval_terms(s, terms);
    // generic: copy, apply, Re<O psi|psi>  (MeasurementsLQubit.hpp:372-394,698-704)
    TempState t(s, true);
    obs_apply(o, t.s);
    s.launches += t.s.launches;
    double d[2];
    dot(t.s, s, s, d);
    return d[0];
}

double var_obs(StateVec &s, const plb200_obs &o) {
    // <O psi|O psi> - <psi|O psi>^2 (MeasurementsLQubit.hpp:431-457,716-727)
    TempState t(s, false);
    obs_apply_to(o, t.s, s);
    s.launches += t.s.launches;
    const double ms = norm2(t.s);
    double d[2];
    dot(s, t.s, s, d);
    return ms - d[0] * d[0];
}

std::vector<int> wires_to_bits_msb_first(int64_t n, const std::vector<int64_t> &wires) {
    std::vector<int> b(wires.size());
    for (size_t i = 0; i < wires.size(); i++) {
        PLB_CHECK(wires[i] >= 0 && wires[i] < n, "Invalid wire index");
        b[i] = static_cast<int>(n - 1 - wires[i]);
    }
    return b;
}

void probs_impl(StateVec &s, const int64_t *wires, int64_t nw, double *out) {
    if (nw < 0) {
        probs_all(s, out);
        return;
    }
    auto w = vi(wires, nw);
    // full register in natural order -> elementwise
    bool natural = (nw == s.n);
    for (int64_t i = 0; natural && i < nw; i++) natural = (w[i] == i);
    if (natural) {
        probs_all(s, out);
        return;
    }
    {
        std::vector<int64_t> sorted(w);
        std::sort(sorted.begin(), sorted.end());
        PLB_CHECK(std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end(),
                  "Wires must be unique");
    }
    if (nw == 0) {
        out[0] = norm2(s);
        return;
    }
    probs_wires(s, wires_to_bits_msb_first(s.n, w), out);
}

// Alias-method sampler, reproducing DiscreteRandomVariable (MeasurementKernels.hpp:308-381)
// and Measurements::generate_samples (MeasurementsLQubit.hpp:662-679) step for step, including
// the arithmetic type of every intermediate, so that a shared seed gives identical samples for identical
// probabilities (marginals over <= 11 wires come from an atomically accumulated histogram whose last bits can vary
// run to run; a bucket comparison can then flip in the rare case of a tie at the last bit).
template <typename P>
void alias_samples(const std::vector<double> &probs_d, int64_t n_wires, int64_t shots, std::mt19937 &gen,
                   uint64_t *out) {
    const size_t n = probs_d.size();
    constexpr size_t none = std::numeric_limits<size_t>::max();
    std::vector<std::pair<double, size_t>> bucket(n, {0.0, none});
    std::stack<size_t> under, over;
    for (size_t i = 0; i < n; i++) {
        bucket[i].first = n * static_cast<P>(probs_d[i]);
        if (bucket[i].first < 1.0) under.push(i);
        else over.push(i);
    }
    while (!under.empty() && !over.empty()) {
        const size_t i = over.top();
        over.pop();
        const size_t j = under.top();
        under.pop();
        bucket[j].second = i;
        bucket[i].first += bucket[j].first - 1.0;
        if (bucket[i].first < 1.0) under.push(i);
        else over.push(i);
    }
    std::uniform_real_distribution<P> dist{0.0, 1.0};
    for (int64_t s = 0; s < shots; s++) {
        size_t idx = static_cast<size_t>(dist(gen) * n);
        if (dist(gen) >= bucket[idx].first && bucket[idx].second != none) idx = bucket[idx].second;
        for (int64_t j = 0; j < n_wires; j++) out[s * n_wires + (n_wires - 1 - j)] = (idx >> j) & 1U;
    }
}

} // namespace

// =========================================================================== C ABI
extern "C" {

const char *plb200_last_error(void) { return g_err.c_str(); }
const char *plb200_version(void) { return "plb200 0.1.0 (sm_100a)"; }

int plb200_device_count(int *count) {
    ABI_TRY
    PLB_CUDA(cudaGetDeviceCount(count));
    ABI_CATCH
}
int plb200_device_arch(int device, int *arch) {
    ABI_TRY
    cudaDeviceProp p;
    PLB_CUDA(cudaGetDeviceProperties(&p, device));
    *arch = p.major * 10 + p.minor;
    ABI_CATCH
}

int plb200_sv_create(plb200_sv **out, int64_t n, int precision, int device, void *stream) {
    ABI_TRY
    auto sv = std::make_unique<plb200_sv>();
    init_sv(sv->s, n, precision, device, stream, nullptr);
    *out = sv.release();
    ABI_CATCH
}
int plb200_sv_create_external(plb200_sv **out, int64_t n, int precision, int device, void *stream, void *ptr) {
    ABI_TRY
    PLB_CHECK(ptr != nullptr, "device_ptr must not be null");
    auto sv = std::make_unique<plb200_sv>();
    init_sv(sv->s, n, precision, device, stream, ptr);
    *out = sv.release();
    ABI_CATCH
}
int plb200_sv_destroy(plb200_sv *sv) {
    if (sv) {
        free_sv(sv->s);
        delete sv;
    }
    return 0;
}
int64_t plb200_sv_num_qubits(const plb200_sv *sv) { return sv->s.n; }
int64_t plb200_sv_length(const plb200_sv *sv) { return static_cast<int64_t>(sv->s.length()); }
int plb200_sv_precision(const plb200_sv *sv) { return sv->s.precision; }
int plb200_sv_device(const plb200_sv *sv) { return sv->s.device; }
void *plb200_sv_device_ptr(const plb200_sv *sv) { return sv->s.data; }
int64_t plb200_sv_kernel_launches(const plb200_sv *sv) { return sv->s.launches; }
int plb200_sv_sync(plb200_sv *sv) {
    ABI_TRY
    sv->s.set_device();
    sv->s.sync();
    ABI_CATCH
}

int plb200_sv_h2d(plb200_sv *sv, const void *host, int64_t n_elems, int async) {
    ABI_TRY
    PLB_CHECK(n_elems >= 0 && static_cast<size_t>(n_elems) <= sv->s.length(), "Invalid size of the host buffer");
    sv->s.set_device();
    PLB_CUDA(cudaMemcpyAsync(sv->s.data, host, n_elems * sv->s.elem_bytes(), cudaMemcpyHostToDevice, sv->s.stream));
    if (!async) sv->s.sync();
    ABI_CATCH
}
int plb200_sv_d2h(plb200_sv *sv, void *host, int64_t n_elems, int async) {
    ABI_TRY
    PLB_CHECK(n_elems >= 0 && static_cast<size_t>(n_elems) <= sv->s.length(), "Invalid size of the host buffer");
    sv->s.set_device();
    PLB_CUDA(cudaMemcpyAsync(host, sv->s.data, n_elems * sv->s.elem_bytes(), cudaMemcpyDeviceToHost, sv->s.stream));
    if (!async) sv->s.sync();
    ABI_CATCH
}
int plb200_sv_d2d(plb200_sv *dst, const plb200_sv *src) {
    ABI_TRY
    copy_state(dst->s, src->s);
    dst->s.sync();
    ABI_CATCH
}

int plb200_sv_set_basis_state_index(plb200_sv *sv, int64_t index) {
    ABI_TRY
    StateVec &s = sv->s;
    PLB_CHECK(index >= 0 && static_cast<size_t>(index) < s.length(), "Invalid index");
    s.set_device();
    PLB_CUDA(cudaMemsetAsync(s.data, 0, s.bytes(), s.stream));
    const double one64[2] = {1.0, 0.0};
    const float one32[2] = {1.0f, 0.0f};
    PLB_CUDA(cudaMemcpyAsync(static_cast<char *>(s.data) + index * s.elem_bytes(),
                             s.precision == 64 ? static_cast<const void *>(one64) : one32, s.elem_bytes(),
                             cudaMemcpyHostToDevice, s.stream));
    s.sync();
    ABI_CATCH
}
int plb200_sv_reset(plb200_sv *sv) { return plb200_sv_set_basis_state_index(sv, 0); }
int plb200_sv_set_basis_state(plb200_sv *sv, const int64_t *state, const int64_t *wires, int64_t nw) {
    int64_t index = 0;
    try {
        for (int64_t k = 0; k < nw; k++) {
            PLB_CHECK(wires[k] >= 0 && wires[k] < sv->s.n, "Invalid wire index");
            index |= (state[k] ? int64_t{1} : 0) << (sv->s.n - 1 - wires[k]);
        }
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
    return plb200_sv_set_basis_state_index(sv, index);
}
int plb200_sv_set_state_vector(plb200_sv *sv, const double *values, const int64_t *wires, int64_t nw) {
    ABI_TRY
    auto w = vi(wires, nw);
    std::vector<int> tbits(nw);
    for (int64_t j = 0; j < nw; j++) {
        PLB_CHECK(w[nw - 1 - j] >= 0 && w[nw - 1 - j] < sv->s.n, "Invalid wire index");
        tbits[j] = static_cast<int>(sv->s.n - 1 - w[nw - 1 - j]);
    }
    set_state_on_wires(sv->s, values, tbits);
    ABI_CATCH
}
int plb200_sv_set_state_indices(plb200_sv *sv, const int64_t *indices, const double *values, int64_t n) {
    ABI_TRY
    for (int64_t i = 0; i < n; i++)
        PLB_CHECK(indices[i] >= 0 && static_cast<size_t>(indices[i]) < sv->s.length(), "Invalid index");
    scatter_values(sv->s, indices, values, n);
    sv->s.sync();
    ABI_CATCH
}
int plb200_sv_normalize(plb200_sv *sv) {
    ABI_TRY
    StateVec &s = sv->s;
    const double nrm = std::sqrt(norm2(s));
    const double eps = s.precision == 64 ? std::numeric_limits<double>::epsilon()
                                         : static_cast<double>(std::numeric_limits<float>::epsilon());
    PLB_CHECK(!(nrm < eps * 1e2), "Vector has norm close to zero and cannot be normalized");
    scale(s, cd(1.0 / nrm, 0.0));
    s.sync();
    ABI_CATCH
}
int plb200_sv_collapse(plb200_sv *sv, int64_t wire, int branch) {
    ABI_TRY
    StateVec &s = sv->s;
    PLB_CHECK(wire >= 0 && wire < s.n, "Invalid wire index");
    collapse_zero(s, static_cast<int>(s.n - 1 - wire), branch ? 1 : 0);
    // normalize(), StateVectorLQubit.hpp:908-921
    const double nrm = std::sqrt(norm2(s));
    const double eps = s.precision == 64 ? std::numeric_limits<double>::epsilon()
                                         : static_cast<double>(std::numeric_limits<float>::epsilon());
    PLB_CHECK(!(nrm < eps * 1e2), "Vector has norm close to zero and cannot be normalized");
    scale(s, cd(1.0 / nrm, 0.0));
    s.sync();
    ABI_CATCH
}

int plb200_sv_apply(plb200_sv *sv, const char *name, const int64_t *cw, const uint8_t *cv, int64_t nc,
                    const int64_t *w, int64_t nw, int inverse, const double *params, int64_t np) {
    ABI_TRY
    apply_call(sv->s, make_call(name, cw, cv, nc, w, nw, inverse, params, np));
    ABI_CATCH
}
int plb200_validate_op(int64_t n, const char *name, const int64_t *cw, const uint8_t *cv, int64_t nc, const int64_t *w,
                       int64_t nw, int inverse, const double *params, int64_t np) {
    ABI_TRY
    (void)lower_gate(n, make_call(name, cw, cv, nc, w, nw, inverse, params, np));
    ABI_CATCH
}
int plb200_validate_ops(int64_t n, const plb200_ops_t *ops) {
    ABI_TRY
    for (int64_t i = 0; i < ops->n_ops; i++) (void)lower_gate(n, call_from_blob(*ops, i));
    ABI_CATCH
}
int plb200_sv_apply_matrix(plb200_sv *sv, const double *matrix, const int64_t *cw, const uint8_t *cv, int64_t nc,
                           const int64_t *w, int64_t nw, int inverse) {
    ABI_TRY
    PLB_CHECK(nw > 0, "Number of wires must be larger than 0");
    PLB_CHECK(nw <= 20, "applyMatrix supports at most 20 target wires");
    launch_ops(sv->s, lower_matrix(sv->s.n, vc(matrix, int64_t{1} << (2 * nw)), vi(w, nw), vi(cw, nc), vb(cv, nc),
                                   inverse != 0));
    ABI_CATCH
}
int plb200_sv_apply_pauli_rot(plb200_sv *sv, const int64_t *w, int64_t nw, int inverse, double theta,
                              const char *word) {
    ABI_TRY
    launch_ops(sv->s, lower_pauli_rot(sv->s.n, vi(w, nw), inverse != 0, theta, word));
    ABI_CATCH
}
int plb200_sv_apply_generator(plb200_sv *sv, const char *name, const int64_t *cw, const uint8_t *cv, int64_t nc,
                              const int64_t *w, int64_t nw, int adj, double *scale_out) {
    ABI_TRY
    (void)adj; // every generator of the reference is Hermitian: adj is ignored there too
    launch_ops(sv->s, lower_generator(sv->s.n, make_call(name, cw, cv, nc, w, nw, 0, nullptr, 0), scale_out));
    ABI_CATCH
}
int plb200_sv_apply_ops(plb200_sv *sv, const plb200_ops_t *ops, int fuse) {
    ABI_TRY
    StateVec &s = sv->s;
    std::vector<COp> all;
    for (int64_t i = 0; i < ops->n_ops; i++) {
        auto l = lower_gate(s.n, call_from_blob(*ops, i));
        all.insert(all.end(), std::make_move_iterator(l.begin()), std::make_move_iterator(l.end()));
    }
    const int64_t before = s.launches;
    if (fuse) run_fused(s, all);
    else launch_ops(s, all);
    s.last_stats[0] = ops->n_ops;
    s.last_stats[1] = s.launches - before;
    ABI_CATCH
}
// ---- routed tape (sharded mode): the last pass stores through an index-bit swap into ping-pong slabs
int plb200_sv_alloc_alt(plb200_sv *sv) {
    ABI_TRY
    StateVec &s = sv->s;
    PLB_CHECK(s.owned, "a state vector on caller-owned memory cannot allocate a ping-pong slab");
    s.set_device();
    if (!s.alt) PLB_CUDA(cudaMalloc(&s.alt, s.bytes()));
    ABI_CATCH
}
void *plb200_sv_alt_ptr(const plb200_sv *sv) { return sv->s.alt; }
int plb200_sv_apply_ops_route(plb200_sv *sv, const plb200_ops_t *ops, int64_t k, const int64_t *lbits, int64_t my_value,
                              void *const *dst, int *routed) {
    ABI_TRY
    StateVec &s = sv->s;
    PLB_CHECK(k >= 1 && k <= 3, "Invalid number of bits");
    PLB_CHECK(s.alt != nullptr, "plb200_sv_alloc_alt must be called first");
    PLB_CHECK(my_value >= 0 && my_value < (int64_t{1} << k), "Invalid rank value");
    RouteSpec rs;
    rs.k = static_cast<int>(k);
    rs.my_value = static_cast<int>(my_value);
    uint64_t seen = 0;
    for (int64_t i = 0; i < k; i++) {
        PLB_CHECK(lbits[i] >= 0 && lbits[i] < s.n && !((seen >> lbits[i]) & 1), "Invalid bit");
        seen |= uint64_t{1} << lbits[i];
        rs.lbits[i] = static_cast<int>(lbits[i]);
    }
    for (int p = 0; p < (1 << k); p++) {
        rs.dst[p] = p == my_value ? s.alt : dst[p];
        PLB_CHECK(rs.dst[p] != nullptr, "missing destination slab");
    }
    std::vector<COp> all;
    for (int64_t i = 0; i < ops->n_ops; i++) {
        auto l = lower_gate(s.n, call_from_blob(*ops, i));
        all.insert(all.end(), std::make_move_iterator(l.begin()), std::make_move_iterator(l.end()));
    }
    const int64_t before = s.launches;
    const bool r = run_fused_routed(s, all, rs);
    s.last_stats[0] = ops->n_ops;
    s.last_stats[1] = s.launches - before;
    // the routed pass wrote this rank's share into its own alt slab (and the peers write theirs into it):
    // from here on that slab is the state
    if (r) std::swap(s.data, s.alt);
    *routed = r ? 1 : 0;
    ABI_CATCH
}

int plb200_schedule_stats(int64_t n, int precision, const plb200_ops_t *ops, int64_t *out4) {
    ABI_TRY
    std::vector<COp> all;
    for (int64_t i = 0; i < ops->n_ops; i++) {
        auto l = lower_gate(n, call_from_blob(*ops, i));
        all.insert(all.end(), std::make_move_iterator(l.begin()), std::make_move_iterator(l.end()));
    }
    schedule_stats(static_cast<int>(n), precision, all, out4);
    ABI_CATCH
}
// ---- pass specialisation (jit_codegen.hpp / jit_runtime.cpp)
int plb200_jit_available(void) { return jit::available(nullptr) ? 1 : 0; }
int plb200_jit_mode(void) { return static_cast<int>(jit::mode()); }
void plb200_jit_set_mode(int mode) { jit::set_mode(mode); }
int plb200_jit_wait(void) {
    ABI_TRY
    jit::wait_idle();
    ABI_CATCH
}
void plb200_jit_stats(int64_t *out8) { jit::stats(out8); }
int plb200_jit_dump_sources(int64_t n, int precision, const plb200_ops_t *ops, const char *dir, int64_t *n_passes) {
    ABI_TRY
    std::vector<COp> all;
    for (int64_t i = 0; i < ops->n_ops; i++) {
        auto l = lower_gate(n, call_from_blob(*ops, i));
        all.insert(all.end(), std::make_move_iterator(l.begin()), std::make_move_iterator(l.end()));
    }
    std::vector<std::string> srcs;
    pass_sources(static_cast<int>(n), precision, all, srcs);
    for (size_t i = 0; i < srcs.size(); i++) {
        const std::string path = std::string(dir) + "/pass_" + std::to_string(i) + ".cu";
        FILE *f = std::fopen(path.c_str(), "w");
        PLB_CHECK(f != nullptr, "cannot write " + path);
        std::fwrite(srcs[i].data(), 1, srcs[i].size(), f);
        std::fclose(f);
    }
    *n_passes = static_cast<int64_t>(srcs.size());
    ABI_CATCH
}
int plb200_jit_compile_check(int64_t n, int precision, const plb200_ops_t *ops, int64_t *n_passes, int64_t *n_ok) {
    ABI_TRY
    std::vector<COp> all;
    for (int64_t i = 0; i < ops->n_ops; i++) {
        auto l = lower_gate(n, call_from_blob(*ops, i));
        all.insert(all.end(), std::make_move_iterator(l.begin()), std::make_move_iterator(l.end()));
    }
    std::vector<std::string> srcs;
    pass_sources(static_cast<int>(n), precision, all, srcs);
    *n_passes = static_cast<int64_t>(srcs.size());
    *n_ok = 0;
    for (const auto &src : srcs) {
        std::string log;
        size_t bytes = 0;
        PLB_CHECK(!src.empty(), "a pass has no specialised source");
        PLB_CHECK(jit::compile_only(src, log, &bytes), "NVRTC: " + log);
        (*n_ok)++;
    }
    ABI_CATCH
}

int plb200_sv_last_apply_stats(const plb200_sv *sv, int64_t *stats2) {
    stats2[0] = sv->s.last_stats[0];
    stats2[1] = sv->s.last_stats[1];
    return 0;
}

int plb200_sv_dot(const plb200_sv *a, const plb200_sv *b, double *out) {
    ABI_TRY
    dot(a->s, b->s, const_cast<StateVec &>(a->s), out);
    ABI_CATCH
}
int plb200_sv_axpy(plb200_sv *y, const double *alpha, const plb200_sv *x) {
    ABI_TRY
    axpy(y->s, cd(alpha[0], alpha[1]), x->s);
    ABI_CATCH
}
int plb200_sv_scale(plb200_sv *sv, const double *alpha) {
    ABI_TRY
    scale(sv->s, cd(alpha[0], alpha[1]));
    ABI_CATCH
}
int plb200_sv_norm2(const plb200_sv *sv, double *out) {
    ABI_TRY
    *out = norm2(const_cast<StateVec &>(sv->s));
    ABI_CATCH
}

// ------------------------------------------------------------------------- observables
int plb200_obs_named(plb200_obs **out, const char *name, const int64_t *wires, int64_t nw, const double *params,
                     int64_t np) {
    ABI_TRY
    PLB_CHECK(gate_known(name), std::string("Gate operation does not exist for ") + name);
    auto o = std::make_unique<plb200_obs>();
    o->kind = plb200_obs::NAMED;
    o->name = name;
    o->wires = vi(wires, nw);
    o->params = vd(params, np);
    *out = o.release();
    ABI_CATCH
}
int plb200_obs_hermitian(plb200_obs **out, const double *matrix, const int64_t *wires, int64_t nw) {
    ABI_TRY
    PLB_CHECK(nw > 0 && nw <= 20, "Invalid number of wires for a Hermitian observable");
    auto o = std::make_unique<plb200_obs>();
    o->kind = plb200_obs::HERMITIAN;
    o->wires = vi(wires, nw);
    o->matrix = vc(matrix, int64_t{1} << (2 * nw));
    *out = o.release();
    ABI_CATCH
}
int plb200_obs_tensor(plb200_obs **out, const plb200_obs *const *terms, int64_t n) {
    ABI_TRY
    auto o = std::make_unique<plb200_obs>();
    o->kind = plb200_obs::TENSOR;
    std::vector<int64_t> all;
    for (int64_t i = 0; i < n; i++) {
        o->terms.push_back(std::make_shared<plb200_obs>(*terms[i]));
        std::vector<int64_t> tw = terms[i]->wires;
        if (terms[i]->kind == plb200_obs::TENSOR || terms[i]->kind == plb200_obs::HAMILTONIAN) tw = terms[i]->wires;
        all.insert(all.end(), tw.begin(), tw.end());
    }
    std::vector<int64_t> sorted(all);
    std::sort(sorted.begin(), sorted.end());
    PLB_CHECK(std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end(),
              "All wires in observables must be disjoint.");
    o->wires = all;
    *out = o.release();
    ABI_CATCH
}
int plb200_obs_hamiltonian(plb200_obs **out, const double *coeffs, const plb200_obs *const *terms, int64_t n) {
    ABI_TRY
    auto o = std::make_unique<plb200_obs>();
    o->kind = plb200_obs::HAMILTONIAN;
    o->coeffs = vd(coeffs, n);
    std::vector<int64_t> all;
    for (int64_t i = 0; i < n; i++) {
        o->terms.push_back(std::make_shared<plb200_obs>(*terms[i]));
        for (auto w : terms[i]->wires)
            if (std::find(all.begin(), all.end(), w) == all.end()) all.push_back(w);
    }
    std::sort(all.begin(), all.end());
    o->wires = all;
    *out = o.release();
    ABI_CATCH
}
int plb200_obs_destroy(plb200_obs *o) {
    delete o;
    return 0;
}
int plb200_obs_apply(const plb200_obs *o, plb200_sv *sv) {
    ABI_TRY
    obs_apply(*o, sv->s);
    ABI_CATCH
}

// ------------------------------------------------------------------------ measurements
int plb200_probs(plb200_sv *sv, const int64_t *wires, int64_t nw, double *out) {
    ABI_TRY
    probs_impl(sv->s, wires, nw, out);
    ABI_CATCH
}

static plb200_obs named_obs(const char *name, const int64_t *wires, int64_t nw) {
    plb200_obs o;
    o.kind = plb200_obs::NAMED;
    o.name = name;
    o.wires = vi(wires, nw);
    return o;
}

int plb200_expval_named(plb200_sv *sv, const char *name, const int64_t *wires, int64_t nw, double *out) {
    ABI_TRY
    const std::string nm(name);
    PLB_CHECK(nm == "Identity" || nm == "PauliX" || nm == "PauliY" || nm == "PauliZ" || nm == "Hadamard",
              "Expval does not exist for named observable " + nm);
    *out = expval_obs(sv->s, named_obs(name, wires, nw));
    ABI_CATCH
}
int plb200_var_named(plb200_sv *sv, const char *name, const int64_t *wires, int64_t nw, double *out) {
    ABI_TRY
    *out = var_obs(sv->s, named_obs(name, wires, nw));
    ABI_CATCH
}
int plb200_expval_matrix(plb200_sv *sv, const double *matrix, const int64_t *wires, int64_t nw, double *out) {
    ABI_TRY
    StateVec &s = sv->s;
    PLB_CHECK(nw > 0 && nw <= 20, "The size of matrix does not match with the given number of wires");
    auto w = vi(wires, nw);
    auto m = vc(matrix, int64_t{1} << (2 * nw));
    if (nw <= 4) {
        std::vector<int> tbits(nw);
        uint64_t seen = 0;
        for (int64_t j = 0; j < nw; j++) {
            PLB_CHECK(w[nw - 1 - j] >= 0 && w[nw - 1 - j] < s.n, "Invalid wire index");
            tbits[j] = static_cast<int>(s.n - 1 - w[nw - 1 - j]);
            PLB_CHECK(!(seen >> tbits[j] & 1), "Wires must be unique");
            seen |= uint64_t{1} << tbits[j];
        }
        *out = expval_matrix_small(s, m, tbits);
    } else {
        plb200_obs o;
        o.kind = plb200_obs::HERMITIAN;
        o.wires = w;
        o.matrix = m;
        *out = expval_obs(s, o);
    }
    ABI_CATCH
}
int plb200_var_matrix(plb200_sv *sv, const double *matrix, const int64_t *wires, int64_t nw, double *out) {
    ABI_TRY
    PLB_CHECK(nw > 0 && nw <= 20, "The size of matrix does not match with the given number of wires");
    plb200_obs o;
    o.kind = plb200_obs::HERMITIAN;
    o.wires = vi(wires, nw);
    o.matrix = vc(matrix, int64_t{1} << (2 * nw));
    *out = var_obs(sv->s, o);
    ABI_CATCH
}
int plb200_expval_pauli_words_each(plb200_sv *sv, const char *const *words, const int64_t *wires,
                                   const int64_t *wires_off, int64_t n_words, double *out) {
    ABI_TRY
    StateVec &s = sv->s;
    std::vector<PauliWordMask> w(n_words);
    for (int64_t k = 0; k < n_words; k++)
        w[k] = pauli_word_mask(s.n, words[k], vi(wires + wires_off[k], wires_off[k + 1] - wires_off[k]));
    std::vector<double> r(2 * n_words);
    pauli_inner(s, s, w.data(), n_words, r.data());
    for (int64_t k = 0; k < n_words; k++) out[k] = r[2 * k];
    ABI_CATCH
}
int plb200_expval_pauli_words(plb200_sv *sv, const char *const *words, const int64_t *wires,
                              const int64_t *wires_off, const double *coeffs, int64_t n_words, double *out) {
    std::vector<double> each(n_words > 0 ? n_words : 0);
    if (int rc = plb200_expval_pauli_words_each(sv, words, wires, wires_off, n_words, each.data())) return rc;
    double e = 0;
    for (int64_t k = 0; k < n_words; k++) e += coeffs[k] * each[k];
    *out = e;
    return 0;
}
int plb200_expval_obs(plb200_sv *sv, const plb200_obs *o, double *out) {
    ABI_TRY
    *out = expval_obs(sv->s, *o);
    ABI_CATCH
}
int plb200_var_obs(plb200_sv *sv, const plb200_obs *o, double *out) {
    ABI_TRY
    *out = var_obs(sv->s, *o);
    ABI_CATCH
}
int plb200_generate_samples_device(plb200_sv *sv, const int64_t *wires, int64_t nw, int64_t shots, int64_t seed,
                                   uint64_t *out) {
    ABI_TRY
    StateVec &s = sv->s;
    std::vector<int64_t> w;
    if (nw < 0) {
        w.resize(s.n);
        for (int64_t i = 0; i < s.n; i++) w[i] = i;
    } else
        w.assign(wires, wires + nw);
    PLB_CHECK(!w.empty(), "generate_samples: no wires");
    uint64_t seen = 0;
    std::vector<int> bits;
    for (int64_t x : w) {
        PLB_CHECK(x >= 0 && x < s.n, "Invalid wire index");
        PLB_CHECK(!(seen >> x & 1), "Wires must be unique");
        seen |= uint64_t{1} << x;
        bits.push_back(static_cast<int>(s.n - 1 - x));
    }
    uint64_t sd = static_cast<uint64_t>(seed);
    if (seed < 0) {
        std::random_device rd;
        sd = (static_cast<uint64_t>(rd()) << 32) | rd();
    }
    sample_device(s, bits, shots, sd, out);
    ABI_CATCH
}
int plb200_generate_samples(plb200_sv *sv, const int64_t *wires, int64_t nw, int64_t shots, int64_t seed,
                            uint64_t *out) {
    {
        // above 24 wires the table no longer belongs on the host (2^k doubles over PCIe, a sequential build):
        // the device sampler takes over (same distribution, its own random stream); PLB200_SAMPLES=device|alias forces
        const char *e = std::getenv("PLB200_SAMPLES");
        const int64_t kk = nw < 0 ? sv->s.n : nw;
        if ((e && !std::strcmp(e, "device")) || (kk > 24 && !(e && !std::strcmp(e, "alias"))))
            return plb200_generate_samples_device(sv, wires, nw, shots, seed, out);
    }
    ABI_TRY
    StateVec &s = sv->s;
    const int64_t k = nw < 0 ? s.n : nw;
    PLB_CHECK(k <= 34, "generate_samples: too many wires for a host-side alias table");
    std::vector<double> p(size_t{1} << k);
    probs_impl(s, wires, nw, p.data());
    std::mt19937 gen;
    if (seed >= 0) gen.seed(static_cast<std::size_t>(seed));
    else {
        std::random_device rd;
        gen.seed(rd());
    }
    if (s.precision == 64) alias_samples<double>(p, k, shots, gen, out);
    else alias_samples<float>(p, k, shots, gen, out);
    ABI_CATCH
}

// --------------------------------------------------------------------- adjoint Jacobian
int plb200_adjoint_jacobian(const plb200_sv *sv, const plb200_obs *const *obs, int64_t n_obs,
                            const plb200_ops_t *ops, const int64_t *trainable, int64_t n_tp, int apply_ops,
                            double *jac) {
    ABI_TRY
    const StateVec &ref = sv->s;
    if (n_tp == 0) return 0;
    const int64_t n_ops = ops->n_ops;
    std::vector<GateCall> calls(n_ops);
    int64_t num_param_ops = 0;
    for (int64_t i = 0; i < n_ops; i++) {
        calls[i] = call_from_blob(*ops, i);
        if (!calls[i].params.empty()) num_param_ops++;
    }
    std::vector<int64_t> tp(trainable, trainable + n_tp);
    for (int64_t i = 0; i < n_obs * n_tp; i++) jac[i] = 0.0;

    // lambda = U psi (or psi), H_lambda_i = O_i lambda     (AdjointJacobianLQubit.hpp:378-425)
    TempState lambda(ref, true);
    if (apply_ops)
        for (const auto &c : calls) apply_call(lambda.s, c);
    std::vector<std::unique_ptr<TempState>> hl;
    for (int64_t i = 0; i < n_obs; i++) {
        hl.push_back(std::make_unique<TempState>(ref, false));
        obs_apply_to(*obs[i], hl.back()->s, lambda.s);
    }
    // ---- single observable, every trainable generator a (controlled) Pauli word: the whole backward
    //      sweep runs as fused two-state tile passes (fusion.cu)
    bool fuse_ok = (n_obs == 1) && std::getenv("PLB200_ADJOINT_UNFUSED") == nullptr;
    if (fuse_ok) {
        std::vector<AdjItem> items;
        std::vector<double> sfs;
        fuse_ok = build_adjoint_items(ref.n, calls, tp, num_param_ops, items, sfs);
        if (fuse_ok) {
            std::vector<double> acc(n_tp, 0.0);
            int64_t st[3];
            run_adjoint_fused(lambda.s, hl[0]->s, items, static_cast<int>(n_tp), acc.data(), st, &const_cast<StateVec &>(ref));
            for (int64_t p = 0; p < n_tp; p++) jac[p] = -2.0 * sfs[p] * acc[p];
            const_cast<StateVec &>(ref).launches += lambda.s.launches + hl[0]->s.launches;
            const_cast<StateVec &>(ref).last_stats[0] = st[0], const_cast<StateVec &>(ref).last_stats[1] = st[1];
            return 0;
        }
    }
    std::unique_ptr<TempState> mu; // only for generators that are not (controlled) Pauli words

    int64_t tp_idx = n_tp - 1;             // trainableParamNumber
    int64_t current_param_idx = num_param_ops - 1;
    for (int64_t op_idx = n_ops - 1; op_idx >= 0; op_idx--) {
        const GateCall &c = calls[op_idx];
        PLB_CHECK(c.params.size() <= 1,
                  "The operation is not supported using the adjoint differentiation method");
        if (c.name == "StatePrep" || c.name == "BasisState") continue;
        if (tp_idx < 0) break;
        if (!c.params.empty()) {
            if (current_param_idx == tp[tp_idx]) {
                PauliWordMask pw;
                double gscale = 0;
                std::vector<double> r(2);
                if (generator_as_pauli(ref.n, c, &pw, &gscale)) {
                    const double sf = gscale * (c.inverse ? -1.0 : 1.0);
                    for (int64_t o = 0; o < n_obs; o++) {
                        pauli_inner(hl[o]->s, lambda.s, &pw, 1, r.data());
                        jac[o * n_tp + tp_idx] = -2.0 * sf * r[1];
                    }
                } else {
                    if (!mu) mu = std::make_unique<TempState>(ref, false);
                    copy_state(mu->s, lambda.s);
                    GateCall gc = c;
                    gc.params.clear();
                    launch_ops(mu->s, lower_generator(ref.n, gc, &gscale));
                    const double sf = gscale * (c.inverse ? -1.0 : 1.0);
                    for (int64_t o = 0; o < n_obs; o++) {
                        dot(hl[o]->s, mu->s, mu->s, r.data());
                        jac[o * n_tp + tp_idx] = -2.0 * sf * r[1];
                    }
                }
                tp_idx--;
            }
            current_param_idx--;
        }
        if (tp_idx < 0) break;
        GateCall inv = c;
        inv.inverse = !c.inverse;
        const auto lowered = lower_gate(ref.n, inv);
        launch_ops(lambda.s, lowered);
        for (int64_t o = 0; o < n_obs; o++) launch_ops(hl[o]->s, lowered);
    }
    lambda.s.sync();
    int64_t total = lambda.s.launches + (mu ? mu->s.launches : 0);
    for (auto &h : hl) total += h->s.launches;
    const_cast<StateVec &>(ref).launches += total;
    ABI_CATCH
}

// ------------------------------------------------------------------ sparse observables / CSR overloads
int plb200_obs_sparse(plb200_obs **out, const int64_t *indptr, const int64_t *indices, const double *data, int64_t n_rows) {
    ABI_TRY
    PLB_CHECK(n_rows >= 1 && (n_rows & (n_rows - 1)) == 0, "sparse matrix dimension must be a power of two");
    auto o = std::make_unique<plb200_obs>();
    o->kind = plb200_obs::SPARSE;
    o->indptr.assign(indptr, indptr + n_rows + 1);
    const int64_t nnz = indptr[n_rows];
    PLB_CHECK(indptr[0] == 0 && nnz >= 0, "invalid CSR row pointer");
    o->indices.assign(indices, indices + nnz);
    for (int64_t k = 0; k < nnz; k++) PLB_CHECK(indices[k] >= 0 && indices[k] < n_rows, "invalid CSR column index");
    o->values = vc(data, nnz);
    *out = o.release();
    ABI_CATCH
}
// <psi| A |psi> and its variance for a CSR matrix given directly (Measurements::expval / var CSR overloads,
// lightning_gpu/bindings/LGPUBindings.hpp:65-150, MeasurementsGPU.hpp sparse paths)
int plb200_expval_sparse(plb200_sv *sv, const int64_t *indptr, const int64_t *indices, const double *data, int64_t n_rows,
                         double *out) {
    ABI_TRY
    StateVec &s = sv->s;
    const auto vals = vc(data, indptr[n_rows]);
    DeviceCsr d(s, indptr, indices, vals.data(), n_rows);
    TempState t(s, false);
    csr_apply(t.s, s, d.indptr, d.indices, d.vals, d.nnz);
    double r[2];
    dot(s, t.s, s, r);
    s.launches += t.s.launches;
    *out = r[0];
    ABI_CATCH
}
int plb200_var_sparse(plb200_sv *sv, const int64_t *indptr, const int64_t *indices, const double *data, int64_t n_rows,
                      double *out) {
    ABI_TRY
    StateVec &s = sv->s;
    const auto vals = vc(data, indptr[n_rows]);
    DeviceCsr d(s, indptr, indices, vals.data(), n_rows);
    TempState t(s, false);
    csr_apply(t.s, s, d.indptr, d.indices, d.vals, d.nnz);
    const double ms = norm2(t.s);
    double r[2];
    dot(s, t.s, s, r);
    s.launches += t.s.launches;
    *out = ms - r[0] * r[0];
    ABI_CATCH
}

// ------------------------------------------------------------------ Hermitian eigen-decomposition (host)
// Replaces the LAPACK zheev/cheev call the reference dlopens from scipy-openblas for Hermitian observables
// measured with shots (core/utils/UtilLinearAlg.hpp:59-117, Observables.hpp:236-262): cyclic complex Jacobi.
// eigvals ascending; unitary (row-major) = V^dagger, i.e. row j = conj(eigenvector j), the matrix that rotates
// the state into the observable's eigenbasis.
int plb200_hermitian_eigh(const double *matrix, int64_t dim, double *eigvals, double *unitary) {
    ABI_TRY
    PLB_CHECK(dim >= 1 && dim <= 4096, "Hermitian eigen-decomposition: dimension out of range");
    const size_t n = static_cast<size_t>(dim);
    std::vector<cd> A = vc(matrix, dim * dim), V(n * n, cd(0.0));
    double scale = 0.0;
    for (size_t i = 0; i < n; i++) {
        V[i * n + i] = 1.0;
        for (size_t j = 0; j < n; j++) {
            PLB_CHECK(std::abs(A[i * n + j] - std::conj(A[j * n + i])) <= 1e-10 * (1.0 + std::abs(A[i * n + j])),
                      "The matrix passed to HermitianObs is not a Hermitian matrix.");
            scale = std::max(scale, std::abs(A[i * n + j]));
        }
    }
    const double tiny = std::max(scale, 1e-300) * 1e-17;
    for (int sweep = 0; sweep < 80; sweep++) {
        double off = 0.0;
        for (size_t p = 0; p < n; p++)
            for (size_t q = p + 1; q < n; q++) off = std::max(off, std::abs(A[p * n + q]));
        if (off <= tiny) break;
        for (size_t p = 0; p < n; p++)
            for (size_t q = p + 1; q < n; q++) {
                const cd apq = A[p * n + q];
                const double mag = std::abs(apq);
                if (mag <= tiny) continue;
                const cd ph = apq / mag; // a_pq = |a_pq| ph
                const double tau = (A[q * n + q].real() - A[p * n + p].real()) / (2.0 * mag);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::abs(tau) + std::sqrt(1.0 + tau * tau));
                const double c = 1.0 / std::sqrt(1.0 + t * t), sn = t * c;
                // J = D R: D = diag(1, conj(ph)) on (p, q) makes the pivot real, R = [[c, sn], [-sn, c]]
                const cd jpp = c, jpq = sn, jqp = -sn * std::conj(ph), jqq = c * std::conj(ph);
                for (size_t k = 0; k < n; k++) { // A <- A J (columns p, q)
                    const cd akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = akp * jpp + akq * jqp;
                    A[k * n + q] = akp * jpq + akq * jqq;
                    const cd vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = vkp * jpp + vkq * jqp;
                    V[k * n + q] = vkp * jpq + vkq * jqq;
                }
                for (size_t k = 0; k < n; k++) { // A <- J^dagger A (rows p, q)
                    const cd apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = std::conj(jpp) * apk + std::conj(jqp) * aqk;
                    A[q * n + k] = std::conj(jpq) * apk + std::conj(jqq) * aqk;
                }
            }
    }
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), size_t{0});
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return A[a * n + a].real() < A[b * n + b].real(); });
    for (size_t j = 0; j < n; j++) {
        eigvals[j] = A[order[j] * n + order[j]].real();
        for (size_t k = 0; k < n; k++) {
            const cd v = std::conj(V[k * n + order[j]]); // row j of V^dagger
            unitary[2 * (j * n + k)] = v.real();
            unitary[2 * (j * n + k) + 1] = v.imag();
        }
    }
    ABI_CATCH
}

// ------------------------------------------------------------------ vector-Jacobian product (state cotangent)
// VectorJacobianProduct::operator() (lightning_qubit/algorithms/VectorJacobianProduct.hpp:43-163, bound at
// LQubitBindings.hpp:413-462): vjp[k] = i * sf_k * <G_k mu | lambda> with lambda = (U) psi and mu = dy swept
// backwards together.  dy: 2^n host complex (re, im) pairs; out: n_tp complex pairs.
int plb200_vjp(const plb200_sv *sv, const double *dy, const plb200_ops_t *ops, const int64_t *trainable, int64_t n_tp,
               int apply_ops, double *out) {
    ABI_TRY
    const StateVec &ref = sv->s;
    for (int64_t i = 0; i < 2 * n_tp; i++) out[i] = 0.0;
    if (n_tp == 0) return 0;
    const int64_t n_ops = ops->n_ops;
    std::vector<GateCall> calls(n_ops);
    int64_t num_param_ops = 0;
    for (int64_t i = 0; i < n_ops; i++) {
        calls[i] = call_from_blob(*ops, i);
        if (!calls[i].params.empty()) num_param_ops++;
    }
    std::vector<int64_t> tp(trainable, trainable + n_tp);
    TempState lambda(ref, true), mu(ref, false), mu_d(ref, false);
    if (apply_ops)
        for (const auto &c : calls) apply_call(lambda.s, c);
    {
        const size_t len = ref.length();
        if (ref.precision == 64)
            PLB_CUDA(cudaMemcpyAsync(mu.s.data, dy, len * 16, cudaMemcpyHostToDevice, mu.s.stream));
        else {
            std::vector<float> f(2 * len);
            for (size_t i = 0; i < 2 * len; i++) f[i] = static_cast<float>(dy[i]);
            PLB_CUDA(cudaMemcpyAsync(mu.s.data, f.data(), len * 8, cudaMemcpyHostToDevice, mu.s.stream));
            PLB_CUDA(cudaStreamSynchronize(mu.s.stream));
        }
    }
    int64_t tp_idx = n_tp - 1, current_param_idx = num_param_ops - 1;
    for (int64_t op_idx = n_ops - 1; op_idx >= 0; op_idx--) {
        const GateCall &c = calls[op_idx];
        PLB_CHECK(c.params.size() <= 1, "The operation is not supported using the adjoint differentiation method");
        if (c.name == "StatePrep" || c.name == "BasisState") continue;
        if (tp_idx < 0) break;
        if (!c.params.empty()) {
            if (current_param_idx == tp[tp_idx]) {
                copy_state(mu_d.s, mu.s);
                GateCall gc = c;
                gc.params.clear();
                double gscale = 0;
                launch_ops(mu_d.s, lower_generator(ref.n, gc, &gscale));
                const double sf = gscale * (c.inverse ? -1.0 : 1.0);
                double r[2];
                dot(mu_d.s, lambda.s, mu_d.s, r); // <mu_d | lambda>
                // i * sf * (re + i im) = sf * (-im + i re)
                out[2 * tp_idx] = -sf * r[1];
                out[2 * tp_idx + 1] = sf * r[0];
                tp_idx--;
            }
            current_param_idx--;
        }
        if (tp_idx < 0) break;
        GateCall inv = c;
        inv.inverse = !c.inverse;
        const auto lowered = lower_gate(ref.n, inv);
        launch_ops(lambda.s, lowered);
        launch_ops(mu.s, lowered);
    }
    lambda.s.sync();
    const_cast<StateVec &>(ref).launches += lambda.s.launches + mu.s.launches + mu_d.s.launches;
    ABI_CATCH
}

// ------------------------------------------------------------------------- distributed
int plb200_sv_pack_bit(const plb200_sv *sv, int64_t bit, int keep, void *buf) {
    ABI_TRY
    PLB_CHECK(bit >= 0 && bit < sv->s.n, "Invalid bit");
    pack_bit(sv->s, static_cast<int>(bit), keep, buf);
    ABI_CATCH
}
int plb200_sv_unpack_bit(plb200_sv *sv, int64_t bit, int keep, const void *buf) {
    ABI_TRY
    PLB_CHECK(bit >= 0 && bit < sv->s.n, "Invalid bit");
    unpack_bit(sv->s, static_cast<int>(bit), keep, buf);
    ABI_CATCH
}
int plb200_sv_swap_bit_peer(plb200_sv *sv, int64_t bit, int keep, void *peer, int do_half) {
    ABI_TRY
    PLB_CHECK(bit >= 0 && bit < sv->s.n, "Invalid bit");
    swap_bit_peer(sv->s, static_cast<int>(bit), keep, peer, do_half);
    ABI_CATCH
}

int plb200_sv_swap_bits_peer(plb200_sv *sv, const int64_t *bits, int64_t k, int64_t my_value,
                             void *const *peer_device_ptrs) {
    ABI_TRY
    PLB_CHECK(k >= 1 && k <= 3, "Invalid number of bits");
    int b[3];
    for (int64_t i = 0; i < k; i++) {
        PLB_CHECK(bits[i] >= 0 && bits[i] < sv->s.n, "Invalid bit");
        b[i] = static_cast<int>(bits[i]);
    }
    PLB_CHECK(my_value >= 0 && my_value < (int64_t{1} << k), "Invalid rank value");
    swap_bits_peer(sv->s, b, static_cast<int>(k), static_cast<int>(my_value), peer_device_ptrs);
    ABI_CATCH
}

// bandwidth probe (tools/peer_bw.py): copy n_bytes between device pointers on this state's stream; either pointer
// may be a peer mapping
int plb200_sv_peer_copy(plb200_sv *sv, void *dst, const void *src, int64_t n_bytes, int unroll) {
    ABI_TRY
    PLB_CHECK(n_bytes >= 0 && n_bytes % 16 == 0, "peer_copy: bytes must be a multiple of 16");
    peer_copy(sv->s, dst, src, static_cast<uint64_t>(n_bytes / 16), unroll);
    ABI_CATCH
}
int plb200_sv_ipc_handle_alt(const plb200_sv *sv, unsigned char *handle64) {
    ABI_TRY
    PLB_CHECK(sv->s.alt != nullptr, "no ping-pong slab");
    sv->s.set_device();
    cudaIpcMemHandle_t h;
    PLB_CUDA(cudaIpcGetMemHandle(&h, sv->s.alt));
    std::memcpy(handle64, &h, 64);
    ABI_CATCH
}
int plb200_sv_ipc_handle(const plb200_sv *sv, unsigned char *handle64) {
    ABI_TRY
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    sv->s.set_device();
    cudaIpcMemHandle_t h;
    PLB_CUDA(cudaIpcGetMemHandle(&h, sv->s.data));
    std::memcpy(handle64, &h, 64);
    ABI_CATCH
}
int plb200_ipc_open(const unsigned char *handle64, int device, void **peer_ptr) {
    ABI_TRY
    PLB_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    PLB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ABI_CATCH
}
int plb200_ipc_close(void *peer_ptr, int device) {
    ABI_TRY
    PLB_CUDA(cudaSetDevice(device));
    PLB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    ABI_CATCH
}

} // extern "C"
