// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// extern "C" shim around the *unmodified* lightning.qubit C++ core, compiled from
// the sources where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/liblq_ref.so.  Nothing here re-implements the reference: every entry
// point only marshals plain pointers into the reference's own classes
//   StateVectorLQubitManaged   (simulators/lightning_qubit/StateVectorLQubitManaged.hpp:52-188)
//   Measurements               (simulators/lightning_qubit/measurements/MeasurementsLQubit.hpp)
//   NamedObs/HermitianObs/TensorProdObs/Hamiltonian (…/observables/ObservablesLQubit.hpp)
//   AdjointJacobian            (…/algorithms/AdjointJacobianLQubit.hpp:347-491)
// Used only by tests/, __graft_entry__.smoke(), tests/golden/make_golden.py and
// bench.py's cpu_baseline / --impl reference legs.
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <span>
#include <string>
#include <vector>
#include <chrono>
#include <omp.h>

#include "AdjointJacobianLQubit.hpp"
#include "JacobianData.hpp"
#include "MeasurementsLQubit.hpp"
#include "ObservablesLQubit.hpp"
#include "StateVectorLQubitManaged.hpp"
#include "VectorJacobianProduct.hpp"

using namespace Pennylane::LightningQubit;
using namespace Pennylane::LightningQubit::Measures;
using namespace Pennylane::LightningQubit::Observables;
using namespace Pennylane::LightningQubit::Algorithms;
using Pennylane::Algorithms::JacobianData;
using Pennylane::Algorithms::OpsData;
using Pennylane::Observables::Observable;

static thread_local std::string g_err;
extern "C" const char *lqref_last_error() { return g_err.c_str(); }

#define LQ_TRY try {
#define LQ_CATCH                                                               \
    }                                                                          \
    catch (const std::exception &e) {                                          \
        g_err = e.what();                                                      \
        return 1;                                                              \
    }                                                                          \
    return 0;

namespace {
template <class T> using SV = StateVectorLQubitManaged<T>;
template <class T> using Obs = std::shared_ptr<Observable<SV<T>>>;

std::vector<std::size_t> vec_sz(const int64_t *p, int64_t n) {
    return std::vector<std::size_t>(p, p + n);
}
std::vector<bool> vec_b(const uint8_t *p, int64_t n) {
    std::vector<bool> v(n);
    for (int64_t i = 0; i < n; i++) v[i] = p[i] != 0;
    return v;
}
template <class T> std::vector<T> vec_fp(const double *p, int64_t n) {
    std::vector<T> v(n);
    for (int64_t i = 0; i < n; i++) v[i] = static_cast<T>(p[i]);
    return v;
}
template <class T>
std::vector<std::complex<T>> vec_c(const double *p, int64_t n) {
    std::vector<std::complex<T>> v(n);
    for (int64_t i = 0; i < n; i++)
        v[i] = {static_cast<T>(p[2 * i]), static_cast<T>(p[2 * i + 1])};
    return v;
}

// Flattened tape, shared layout with include/plb200.h (plb200_ops_t).
struct OpsBlob {
    int64_t n_ops;
    const char *const *names;
    const int64_t *wires, *wires_off;
    const int64_t *ctrl_wires, *ctrl_off;
    const uint8_t *ctrl_values;
    const double *params;
    const int64_t *params_off;
    const uint8_t *inverses;
    const double *mats; // interleaved complex128
    const int64_t *mats_off; // offsets in complex elements
};

template <class T> OpsData<SV<T>> make_ops(const OpsBlob &b) {
    std::vector<std::string> names;
    std::vector<std::vector<T>> params;
    std::vector<std::vector<std::size_t>> wires, cw;
    std::vector<bool> inv;
    std::vector<std::vector<std::complex<T>>> mats;
    std::vector<std::vector<bool>> cv;
    for (int64_t i = 0; i < b.n_ops; i++) {
        names.emplace_back(b.names[i]);
        params.push_back(vec_fp<T>(b.params + b.params_off[i],
                                   b.params_off[i + 1] - b.params_off[i]));
        wires.push_back(
            vec_sz(b.wires + b.wires_off[i], b.wires_off[i + 1] - b.wires_off[i]));
        cw.push_back(vec_sz(b.ctrl_wires + b.ctrl_off[i],
                            b.ctrl_off[i + 1] - b.ctrl_off[i]));
        cv.push_back(vec_b(b.ctrl_values + b.ctrl_off[i],
                           b.ctrl_off[i + 1] - b.ctrl_off[i]));
        inv.push_back(b.inverses[i] != 0);
        mats.push_back(vec_c<T>(b.mats + 2 * b.mats_off[i],
                                b.mats_off[i + 1] - b.mats_off[i]));
    }
    return OpsData<SV<T>>(names, params, wires, inv, mats, cw, cv);
}
} // namespace

#define DEFINE_API(SFX, T)                                                     \
    extern "C" {                                                               \
    void *lqref_sv_create_##SFX(int64_t n) {                                   \
        try {                                                                  \
            return new SV<T>(static_cast<std::size_t>(n));                     \
        } catch (const std::exception &e) {                                    \
            g_err = e.what();                                                  \
            return nullptr;                                                    \
        }                                                                      \
    }                                                                          \
    void lqref_sv_destroy_##SFX(void *h) { delete static_cast<SV<T> *>(h); }   \
    void *lqref_sv_data_##SFX(void *h) {                                       \
        return static_cast<SV<T> *>(h)->getData();                             \
    }                                                                          \
    int64_t lqref_sv_length_##SFX(void *h) {                                   \
        return static_cast<int64_t>(static_cast<SV<T> *>(h)->getLength());     \
    }                                                                          \
    int lqref_sv_apply_##SFX(void *h, const char *name, const int64_t *cw,     \
                             const uint8_t *cv, int64_t nc, const int64_t *w,  \
                             int64_t nw, int inverse, const double *params,    \
                             int64_t np) {                                     \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        if (nc == 0)                                                           \
            sv->applyOperation(std::string(name), vec_sz(w, nw), inverse != 0, \
                               vec_fp<T>(params, np));                         \
        else                                                                   \
            sv->applyOperation(std::string(name), vec_sz(cw, nc),              \
                               vec_b(cv, nc), vec_sz(w, nw), inverse != 0,     \
                               vec_fp<T>(params, np));                         \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_apply_matrix_##SFX(void *h, const double *mat,                \
                                    const int64_t *cw, const uint8_t *cv,      \
                                    int64_t nc, const int64_t *w, int64_t nw,  \
                                    int inverse) {                             \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        auto m = vec_c<T>(mat, (int64_t{1} << (2 * nw)));                      \
        if (nc == 0)                                                           \
            sv->applyMatrix(m.data(), vec_sz(w, nw), inverse != 0);            \
        else                                                                   \
            sv->applyControlledMatrix(m.data(), vec_sz(cw, nc), vec_b(cv, nc), \
                                      vec_sz(w, nw), inverse != 0);            \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_apply_pauli_rot_##SFX(void *h, const int64_t *w, int64_t nw,  \
                                       int inverse, double theta,              \
                                       const char *word) {                     \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        sv->applyPauliRot(vec_sz(w, nw), inverse != 0,                         \
                          std::vector<T>{static_cast<T>(theta)},               \
                          std::string(word));                                  \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_apply_generator_##SFX(void *h, const char *name,              \
                                       const int64_t *cw, const uint8_t *cv,   \
                                       int64_t nc, const int64_t *w,           \
                                       int64_t nw, int adj, double *scale) {   \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        if (nc == 0)                                                           \
            *scale = sv->applyGenerator(std::string(name), vec_sz(w, nw),      \
                                        adj != 0);                             \
        else                                                                   \
            *scale = sv->applyGenerator(std::string(name), vec_sz(cw, nc),     \
                                        vec_b(cv, nc), vec_sz(w, nw),          \
                                        adj != 0);                             \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_set_basis_state_##SFX(void *h, const int64_t *state,          \
                                       const int64_t *w, int64_t nw) {         \
        LQ_TRY static_cast<SV<T> *>(h)->setBasisState(vec_sz(state, nw),       \
                                                      vec_sz(w, nw));          \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_set_state_vector_##SFX(void *h, const double *state,          \
                                        const int64_t *w, int64_t nw) {        \
        LQ_TRY auto s = vec_c<T>(state, int64_t{1} << nw);                     \
        static_cast<SV<T> *>(h)->setStateVector(s, vec_sz(w, nw));             \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_reset_##SFX(void *h) {                                        \
        LQ_TRY static_cast<SV<T> *>(h)->resetStateVector();                    \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_collapse_##SFX(void *h, int64_t wire, int branch) {           \
        LQ_TRY static_cast<SV<T> *>(h)->collapse(wire, branch != 0);           \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_sv_normalize_##SFX(void *h) {                                    \
        LQ_TRY static_cast<SV<T> *>(h)->normalize();                           \
        LQ_CATCH                                                               \
    }                                                                          \
    /* ---------------- observables ---------------- */                        \
    void *lqref_obs_named_##SFX(const char *name, const int64_t *w,            \
                                int64_t nw) {                                  \
        try {                                                                  \
            return new Obs<T>(std::make_shared<NamedObs<SV<T>>>(               \
                std::string(name), vec_sz(w, nw)));                            \
        } catch (const std::exception &e) {                                    \
            g_err = e.what();                                                  \
            return nullptr;                                                    \
        }                                                                      \
    }                                                                          \
    void *lqref_obs_hermitian_##SFX(const double *mat, const int64_t *w,       \
                                    int64_t nw) {                              \
        try {                                                                  \
            return new Obs<T>(std::make_shared<HermitianObs<SV<T>>>(           \
                vec_c<T>(mat, int64_t{1} << (2 * nw)), vec_sz(w, nw)));        \
        } catch (const std::exception &e) {                                    \
            g_err = e.what();                                                  \
            return nullptr;                                                    \
        }                                                                      \
    }                                                                          \
    void *lqref_obs_tensor_##SFX(void *const *terms, int64_t n) {              \
        try {                                                                  \
            std::vector<Obs<T>> v;                                             \
            for (int64_t i = 0; i < n; i++)                                    \
                v.push_back(*static_cast<Obs<T> *>(terms[i]));                 \
            return new Obs<T>(std::make_shared<TensorProdObs<SV<T>>>(v));      \
        } catch (const std::exception &e) {                                    \
            g_err = e.what();                                                  \
            return nullptr;                                                    \
        }                                                                      \
    }                                                                          \
    void *lqref_obs_hamiltonian_##SFX(const double *coeffs,                    \
                                      void *const *terms, int64_t n) {         \
        try {                                                                  \
            std::vector<Obs<T>> v;                                             \
            for (int64_t i = 0; i < n; i++)                                    \
                v.push_back(*static_cast<Obs<T> *>(terms[i]));                 \
            return new Obs<T>(std::make_shared<Hamiltonian<SV<T>>>(            \
                vec_fp<T>(coeffs, n), v));                                     \
        } catch (const std::exception &e) {                                    \
            g_err = e.what();                                                  \
            return nullptr;                                                    \
        }                                                                      \
    }                                                                          \
    void *lqref_obs_sparse_##SFX(const int64_t *indptr, const int64_t *indices,\
                                 const double *data, int64_t n_rows,           \
                                 const int64_t *wires, int64_t nw) {           \
        try {                                                                  \
            using SH = SparseHamiltonian<SV<T>>;                               \
            const int64_t nnz = indptr[n_rows];                                \
            std::vector<std::complex<T>> d(nnz);                               \
            for (int64_t k = 0; k < nnz; k++)                                  \
                d[k] = {static_cast<T>(data[2 * k]),                           \
                        static_cast<T>(data[2 * k + 1])};                      \
            std::vector<typename SH::IdxT> ix(indices, indices + nnz),         \
                ip(indptr, indptr + n_rows + 1);                               \
            return new Obs<T>(std::make_shared<SH>(d, ix, ip, vec_sz(wires, nw)));\
        } catch (const std::exception &e) {                                    \
            g_err = e.what();                                                  \
            return nullptr;                                                    \
        }                                                                      \
    }                                                                          \
    void lqref_obs_destroy_##SFX(void *o) { delete static_cast<Obs<T> *>(o); } \
    int lqref_obs_apply_##SFX(void *o, void *h) {                              \
        LQ_TRY(*static_cast<Obs<T> *>(o))                                      \
            ->applyInPlace(*static_cast<SV<T> *>(h));                          \
        LQ_CATCH                                                               \
    }                                                                          \
    /* ---------------- measurements ---------------- */                       \
    int lqref_probs_##SFX(void *h, const int64_t *w, int64_t nw, double *out) {\
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        auto p = (nw < 0) ? m.probs() : m.probs(vec_sz(w, nw));                \
        for (std::size_t i = 0; i < p.size(); i++) out[i] = p[i];              \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_expval_named_##SFX(void *h, const char *name, const int64_t *w,  \
                                 int64_t nw, double *out) {                    \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        *out = m.expval(std::string(name), vec_sz(w, nw));                     \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_var_named_##SFX(void *h, const char *name, const int64_t *w,     \
                              int64_t nw, double *out) {                       \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        *out = m.var(std::string(name), vec_sz(w, nw));                        \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_expval_matrix_##SFX(void *h, const double *mat,                  \
                                  const int64_t *w, int64_t nw, double *out) { \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        *out = m.expval(vec_c<T>(mat, int64_t{1} << (2 * nw)), vec_sz(w, nw)); \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_var_matrix_##SFX(void *h, const double *mat, const int64_t *w,   \
                               int64_t nw, double *out) {                      \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        *out = m.var(vec_c<T>(mat, int64_t{1} << (2 * nw)), vec_sz(w, nw));    \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_expval_obs_##SFX(void *h, void *o, double *out) {                \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        *out = m.expval(**static_cast<Obs<T> *>(o));                           \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_var_obs_##SFX(void *h, void *o, double *out) {                   \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        *out = m.var(**static_cast<Obs<T> *>(o));                              \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_generate_samples_##SFX(void *h, const int64_t *w, int64_t nw,    \
                                     int64_t shots, int64_t seed,              \
                                     uint64_t *out) {                          \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        if (seed >= 0) m.setSeed(static_cast<std::size_t>(seed));              \
        auto s = (nw < 0) ? m.generate_samples(shots)                          \
                          : m.generate_samples(vec_sz(w, nw), shots);          \
        for (std::size_t i = 0; i < s.size(); i++) out[i] = s[i];              \
        LQ_CATCH                                                               \
    }                                                                          \
    /* ---------------- shot-based API (MeasurementsBase.hpp:159-521) ---- */  \
    int lqref_expval_shots_##SFX(void *h, void *o, int64_t shots,              \
                                 int64_t seed, const int64_t *range,           \
                                 int64_t n_range, double *out) {               \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        m.setSeed(static_cast<std::size_t>(seed));                             \
        *out = m.expval(**static_cast<Obs<T> *>(o),                            \
                        static_cast<std::size_t>(shots),                       \
                        vec_sz(range, n_range));                               \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_var_shots_##SFX(void *h, void *o, int64_t shots, int64_t seed,   \
                              double *out) {                                   \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        m.setSeed(static_cast<std::size_t>(seed));                             \
        *out = m.var(**static_cast<Obs<T> *>(o),                               \
                     static_cast<std::size_t>(shots));                         \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_probs_shots_##SFX(void *h, const int64_t *w, int64_t nw,         \
                                int64_t shots, int64_t seed, double *out) {    \
        LQ_TRY Measurements<SV<T>> m(*static_cast<SV<T> *>(h));                \
        m.setSeed(static_cast<std::size_t>(seed));                             \
        auto p = m.probs(vec_sz(w, nw), static_cast<std::size_t>(shots));      \
        for (std::size_t i = 0; i < p.size(); i++) out[i] = p[i];              \
        LQ_CATCH                                                               \
    }                                                                          \
    /* ---------------- adjoint Jacobian ---------------- */                   \
    int lqref_apply_ops_##SFX(void *h, const OpsBlob *blob) {                  \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        for (int64_t i = 0; i < blob->n_ops; i++) {                            \
            int64_t nc = blob->ctrl_off[i + 1] - blob->ctrl_off[i];            \
            int64_t nw = blob->wires_off[i + 1] - blob->wires_off[i];          \
            int64_t np = blob->params_off[i + 1] - blob->params_off[i];        \
            int64_t nm = blob->mats_off[i + 1] - blob->mats_off[i];            \
            if (nm > 0) {                                                      \
                auto m = vec_c<T>(blob->mats + 2 * blob->mats_off[i], nm);     \
                if (nc == 0)                                                   \
                    sv->applyMatrix(m.data(),                                  \
                                    vec_sz(blob->wires + blob->wires_off[i],   \
                                           nw),                                \
                                    blob->inverses[i] != 0);                   \
                else                                                           \
                    sv->applyControlledMatrix(                                 \
                        m.data(),                                              \
                        vec_sz(blob->ctrl_wires + blob->ctrl_off[i], nc),      \
                        vec_b(blob->ctrl_values + blob->ctrl_off[i], nc),      \
                        vec_sz(blob->wires + blob->wires_off[i], nw),          \
                        blob->inverses[i] != 0);                               \
            } else if (nc == 0)                                                \
                sv->applyOperation(                                            \
                    std::string(blob->names[i]),                               \
                    vec_sz(blob->wires + blob->wires_off[i], nw),              \
                    blob->inverses[i] != 0,                                    \
                    vec_fp<T>(blob->params + blob->params_off[i], np));        \
            else                                                               \
                sv->applyOperation(                                            \
                    std::string(blob->names[i]),                               \
                    vec_sz(blob->ctrl_wires + blob->ctrl_off[i], nc),          \
                    vec_b(blob->ctrl_values + blob->ctrl_off[i], nc),          \
                    vec_sz(blob->wires + blob->wires_off[i], nw),              \
                    blob->inverses[i] != 0,                                    \
                    vec_fp<T>(blob->params + blob->params_off[i], np));        \
        }                                                                      \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_adjoint_jacobian_##SFX(void *h, void *const *obs, int64_t n_obs, \
                                     const OpsBlob *blob, const int64_t *tp,   \
                                     int64_t n_tp, int apply_ops,              \
                                     double *jac) {                            \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        std::vector<Obs<T>> ov;                                                \
        for (int64_t i = 0; i < n_obs; i++)                                    \
            ov.push_back(*static_cast<Obs<T> *>(obs[i]));                      \
        auto ops = make_ops<T>(*blob);                                         \
        JacobianData<SV<T>> jd(ops.getTotalNumParams(), sv->getLength(),       \
                               sv->getData(), ov, ops, vec_sz(tp, n_tp));      \
        std::vector<T> j(n_obs * n_tp, 0);                                     \
        AdjointJacobian<SV<T>> adj;                                            \
        adj.adjointJacobian(std::span<T>{j}, jd, *sv, apply_ops != 0);         \
        for (std::size_t i = 0; i < j.size(); i++) jac[i] = j[i];              \
        LQ_CATCH                                                               \
    }                                                                          \
    int lqref_vjp_##SFX(void *h, const double *dy, const OpsBlob *blob,        \
                        const int64_t *tp, int64_t n_tp, int apply_ops,        \
                        double *out) {                                         \
        LQ_TRY auto *sv = static_cast<SV<T> *>(h);                             \
        auto ops = make_ops<T>(*blob);                                         \
        JacobianData<SV<T>> jd(ops.getTotalNumParams(), sv->getLength(),       \
                               sv->getData(), {}, ops, vec_sz(tp, n_tp));      \
        std::vector<std::complex<T>> d(sv->getLength()), j(n_tp);              \
        for (std::size_t i = 0; i < d.size(); i++)                             \
            d[i] = {static_cast<T>(dy[2 * i]), static_cast<T>(dy[2 * i + 1])}; \
        Pennylane::LightningQubit::Algorithms::VectorJacobianProduct<SV<T>> v; \
        v(std::span<std::complex<T>>{j},                                       \
          jd, std::span<const std::complex<T>>{d.data(), d.size()},            \
          apply_ops != 0);                                                     \
        for (std::size_t i = 0; i < j.size(); i++)                             \
            out[2 * i] = j[i].real(), out[2 * i + 1] = j[i].imag();            \
        LQ_CATCH                                                               \
    }                                                                          \
    }

DEFINE_API(c64, float)
DEFINE_API(c128, double)

extern "C" int lqref_num_threads() { return omp_get_max_threads(); }
extern "C" void lqref_set_num_threads(int n) { omp_set_num_threads(n); }
