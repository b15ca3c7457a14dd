// StateVectorB200<PrecisionT> — the B200 backend's state vector, implementing the interface of
// StateVectorBase<PrecisionT, Derived> (core/simulators/base/StateVectorBase.hpp:43-205) and the
// GPU-backend extras of StateVectorCudaManaged (lightning_gpu/StateVectorCudaManaged.hpp) that the
// binding layer calls (core/bindings/Bindings.hpp:152-276,883-921; LGPUBindings.hpp:283-408).
// All arithmetic happens in libplb200.so behind the C ABI (include/plb200.h).
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/plb200.h"
#include "DevTag.hpp"
#include "Error.hpp"

// tag types of core/utils/Memory.hpp, declared here so that this header also builds outside the reference tree
namespace Pennylane::Util::MemoryStorageLocation {
struct Internal;
struct External;
struct Undefined;
} // namespace Pennylane::Util::MemoryStorageLocation

namespace Pennylane::LightningB200 {

namespace detail {
inline std::vector<int64_t> to_i64(const std::vector<std::size_t> &v) { return {v.begin(), v.end()}; }
inline std::vector<uint8_t> to_u8(const std::vector<bool> &v) {
    std::vector<uint8_t> r(v.size());
    for (std::size_t i = 0; i < v.size(); i++) r[i] = v[i] ? 1 : 0;
    return r;
}
template <class T> std::vector<double> to_f64(const std::vector<T> &v) { return {v.begin(), v.end()}; }
template <class T> std::vector<double> to_c128(const std::complex<T> *m, std::size_t n) {
    std::vector<double> r(2 * n);
    for (std::size_t i = 0; i < n; i++) r[2 * i] = m[i].real(), r[2 * i + 1] = m[i].imag();
    return r;
}

// Flattened tape builder for plb200_ops_t
struct OpsBlob {
    std::vector<std::string> names;
    std::vector<const char *> name_ptrs;
    std::vector<int64_t> wires, wires_off{0}, ctrl_wires, ctrl_off{0}, params_off{0}, mats_off{0};
    std::vector<uint8_t> ctrl_values, inverses;
    std::vector<double> params, mats;
    template <class P>
    void add(const std::string &name, const std::vector<std::size_t> &w, bool inverse, const std::vector<P> &p,
             const std::vector<std::size_t> &cw = {}, const std::vector<bool> &cv = {},
             const std::vector<std::complex<P>> &matrix = {}) {
        names.push_back(name);
        wires.insert(wires.end(), w.begin(), w.end());
        wires_off.push_back(static_cast<int64_t>(wires.size()));
        ctrl_wires.insert(ctrl_wires.end(), cw.begin(), cw.end());
        for (bool b : cv) ctrl_values.push_back(b ? 1 : 0);
        ctrl_off.push_back(static_cast<int64_t>(ctrl_wires.size()));
        params.insert(params.end(), p.begin(), p.end());
        params_off.push_back(static_cast<int64_t>(params.size()));
        inverses.push_back(inverse ? 1 : 0);
        for (const auto &c : matrix) mats.push_back(c.real()), mats.push_back(c.imag());
        mats_off.push_back(static_cast<int64_t>(mats.size() / 2));
    }
    plb200_ops_t view() {
        name_ptrs.clear();
        for (auto &s : names) name_ptrs.push_back(s.c_str());
        // keep data() non-null for empty vectors
        if (ctrl_wires.capacity() == 0) ctrl_wires.reserve(1);
        if (ctrl_values.capacity() == 0) ctrl_values.reserve(1);
        if (params.capacity() == 0) params.reserve(1);
        if (mats.capacity() == 0) mats.reserve(1);
        if (wires.capacity() == 0) wires.reserve(1);
        return plb200_ops_t{static_cast<int64_t>(names.size()), name_ptrs.data(), wires.data(), wires_off.data(),
                            ctrl_wires.data(), ctrl_off.data(), ctrl_values.data(), params.data(),
                            params_off.data(), inverses.data(), mats.data(), mats_off.data()};
    }
};
} // namespace detail

template <class Precision = double> class StateVectorB200 {
  public:
    using PrecisionT = Precision;
    // where the amplitudes live, as the reference's templates ask (core/utils/Memory.hpp:191-208;
    // MeasurementsBase.hpp:291,485 branch on it): device memory owned by the engine = "Undefined", like LGPU
    using MemoryStorageT = Pennylane::Util::MemoryStorageLocation::Undefined;
    using ComplexT = std::complex<PrecisionT>;
    using CFP_t = ComplexT; // device elements are layout-compatible interleaved (re, im)

    StateVectorB200() = delete;
    explicit StateVectorB200(std::size_t num_qubits) : StateVectorB200(num_qubits, DevTag<int>{0, nullptr}) {}
    StateVectorB200(std::size_t num_qubits, const DevTag<int> &dev_tag) : num_qubits_{num_qubits}, dev_tag_{dev_tag} {
        PLB200_ABI(plb200_sv_create(&h_, static_cast<int64_t>(num_qubits), prec(), dev_tag.getDeviceID(),
                                    dev_tag.getStreamID()));
    }
    StateVectorB200(const ComplexT *host_data, std::size_t length, const DevTag<int> &dev_tag = DevTag<int>{0, nullptr})
        : StateVectorB200(log2_checked(length), dev_tag) {
        CopyHostDataToGpu(host_data, length, false);
    }
    StateVectorB200(const StateVectorB200 &other) : StateVectorB200(other.num_qubits_, other.dev_tag_) {
        PLB200_ABI(plb200_sv_d2d(h_, other.handle()));
    }
    StateVectorB200 &operator=(const StateVectorB200 &) = delete;
    // the lazily queued gates travel with the handle (a moved-from object owns nothing)
    StateVectorB200(StateVectorB200 &&o) noexcept
        : h_{o.h_}, num_qubits_{o.num_qubits_}, dev_tag_{o.dev_tag_}, lazy_{o.lazy_}, queue_{std::move(o.queue_)} {
        o.h_ = nullptr;
        o.queue_ = detail::OpsBlob{};
    }
    ~StateVectorB200() {
        if (h_) plb200_sv_destroy(h_);
    }

    // ---- StateVectorBase interface -------------------------------------------------------
    [[nodiscard]] auto getNumQubits() const -> std::size_t { return num_qubits_; }
    [[nodiscard]] auto getTotalNumQubits() const -> std::size_t { return num_qubits_; }
    [[nodiscard]] auto getLength() const -> std::size_t { return std::size_t{1} << num_qubits_; }
    [[nodiscard]] auto getData() -> CFP_t * {
        flush();
        return static_cast<CFP_t *>(plb200_sv_device_ptr(h_));
    }
    [[nodiscard]] auto getData() const -> const CFP_t * {
        flush();
        return static_cast<const CFP_t *>(plb200_sv_device_ptr(h_));
    }
    [[nodiscard]] auto getDevTag() const -> const DevTag<int> & { return dev_tag_; }
    // Every consumer of the device state (measurements, observables, adjoint, copies) goes through
    // handle(): pending gates are flushed first.
    [[nodiscard]] plb200_sv *handle() const {
        flush();
        return h_;
    }
    // Lazy gate queue: the per-gate API (one Python->C++ call per gate, Bindings.hpp:223-276) only
    // validates and records the gate; the queue is applied as ONE fused tape on the first read.
    // PLB200_LAZY=0 restores one kernel launch per call.
    void flush() const {
        if (queue_.names.empty()) return;
        auto v = queue_.view();
        detail::OpsBlob done;
        std::swap(done, queue_); // queue_ is empty again even if the call throws
        v = done.view();
        PLB200_ABI(plb200_sv_apply_ops(h_, &v, 1));
    }
    [[nodiscard]] std::size_t pendingOps() const { return queue_.names.size(); }

    void applyOperation(const std::string &opName, const std::vector<std::size_t> &wires, bool inverse = false,
                        const std::vector<PrecisionT> &params = {}) {
        const auto w = detail::to_i64(wires);
        const auto p = detail::to_f64(params);
        if (lazy_) {
            PLB200_ABI(plb200_validate_op(static_cast<int64_t>(num_qubits_), opName.c_str(), nullptr, nullptr, 0,
                                          w.data(), static_cast<int64_t>(w.size()), inverse, p.data(),
                                          static_cast<int64_t>(p.size())));
            queue_.template add<PrecisionT>(opName, wires, inverse, params);
            if (queue_.names.size() >= kMaxQueue) flush();
            return;
        }
        PLB200_ABI(plb200_sv_apply(h_, opName.c_str(), nullptr, nullptr, 0, w.data(), static_cast<int64_t>(w.size()),
                                   inverse, p.data(), static_cast<int64_t>(p.size())));
    }
    void applyOperation(const std::string &opName, const std::vector<std::size_t> &controlled_wires,
                        const std::vector<bool> &controlled_values, const std::vector<std::size_t> &wires,
                        bool inverse = false, const std::vector<PrecisionT> &params = {}) {
        PLB200_ABORT_IF_NOT(controlled_wires.size() == controlled_values.size(),
                            "`controlled_wires` must have the same size as `controlled_values`.");
        const auto w = detail::to_i64(wires), cw = detail::to_i64(controlled_wires);
        const auto cv = detail::to_u8(controlled_values);
        const auto p = detail::to_f64(params);
        if (lazy_) {
            PLB200_ABI(plb200_validate_op(static_cast<int64_t>(num_qubits_), opName.c_str(), cw.data(), cv.data(),
                                          static_cast<int64_t>(cw.size()), w.data(), static_cast<int64_t>(w.size()),
                                          inverse, p.data(), static_cast<int64_t>(p.size())));
            queue_.template add<PrecisionT>(opName, wires, inverse, params, controlled_wires, controlled_values);
            if (queue_.names.size() >= kMaxQueue) flush();
            return;
        }
        PLB200_ABI(plb200_sv_apply(h_, opName.c_str(), cw.data(), cv.data(), static_cast<int64_t>(cw.size()), w.data(),
                                   static_cast<int64_t>(w.size()), inverse, p.data(),
                                   static_cast<int64_t>(p.size())));
    }
    // overloads with an explicit matrix: used when `opName` is not a native gate
    // (StateVectorLQubit.hpp:434-483, AdjointJacobianBase.hpp:78-96)
    void applyOperation(const std::string &opName, const std::vector<std::size_t> &wires, bool inverse,
                        const std::vector<PrecisionT> &params, const std::vector<ComplexT> &matrix) {
        if (isNativeGate(opName) || matrix.empty()) applyOperation(opName, wires, inverse, params);
        else applyMatrix(matrix, wires, inverse);
    }
    void applyOperation(const std::string &opName, const std::vector<std::size_t> &controlled_wires,
                        const std::vector<bool> &controlled_values, const std::vector<std::size_t> &wires,
                        bool inverse, const std::vector<PrecisionT> &params, const std::vector<ComplexT> &matrix) {
        if (isNativeGate(opName) || matrix.empty()) {
            if (controlled_wires.empty()) applyOperation(opName, wires, inverse, params);
            else applyOperation(opName, controlled_wires, controlled_values, wires, inverse, params);
        } else if (controlled_wires.empty())
            applyMatrix(matrix, wires, inverse);
        else
            applyControlledMatrix(matrix.data(), controlled_wires, controlled_values, wires, inverse);
    }
    // applyOperations (StateVectorBase.hpp:116-160): the whole list goes to the engine in one call so
    // that it can be scheduled into cache-blocked passes.
    void applyOperations(const std::vector<std::string> &ops, const std::vector<std::vector<std::size_t>> &ops_wires,
                         const std::vector<bool> &ops_adjoint,
                         const std::vector<std::vector<PrecisionT>> &ops_params) {
        const std::size_t n = ops.size();
        PLB200_ABORT_IF(n != ops_wires.size(), "Invalid arguments: number of operations, wires, inverses, and "
                                               "parameters must all be equal");
        PLB200_ABORT_IF(n != ops_adjoint.size(), "Invalid arguments: number of operations, wires, inverses, and "
                                                 "parameters must all be equal");
        PLB200_ABORT_IF(n != ops_params.size(), "Invalid arguments: number of operations, wires, inverses, and "
                                                "parameters must all be equal");
        flush();
        detail::OpsBlob blob;
        for (std::size_t i = 0; i < n; i++) blob.add<PrecisionT>(ops[i], ops_wires[i], ops_adjoint[i], ops_params[i]);
        const auto v = blob.view();
        PLB200_ABI(plb200_sv_apply_ops(h_, &v, 1));
    }
    void applyOperations(const std::vector<std::string> &ops, const std::vector<std::vector<std::size_t>> &ops_wires,
                         const std::vector<bool> &ops_adjoint) {
        applyOperations(ops, ops_wires, ops_adjoint, std::vector<std::vector<PrecisionT>>(ops.size()));
    }
    void applyOperations(detail::OpsBlob &blob, bool fuse = true) {
        flush();
        const auto v = blob.view();
        PLB200_ABI(plb200_sv_apply_ops(h_, &v, fuse ? 1 : 0));
    }

    // Matrices on up to kLazyMatrixWires wires and Pauli rotations join the lazy queue (validated now, applied
    // with the rest of the tape as fused passes); larger matrices flush and run at once.
    static constexpr std::size_t kLazyMatrixWires = 4;
    bool queue_matrix(const ComplexT *matrix, const std::vector<std::size_t> &controlled_wires,
                      const std::vector<bool> &controlled_values, const std::vector<std::size_t> &wires, bool inverse) {
        if (!lazy_ || wires.empty() || wires.size() > kLazyMatrixWires) return false;
        const std::vector<ComplexT> m(matrix, matrix + (std::size_t{1} << (2 * wires.size())));
        detail::OpsBlob one;
        one.template add<PrecisionT>("Matrix", wires, inverse, {}, controlled_wires, controlled_values, m);
        const auto v = one.view();
        PLB200_ABI(plb200_validate_ops(static_cast<int64_t>(num_qubits_), &v));
        queue_.template add<PrecisionT>("Matrix", wires, inverse, {}, controlled_wires, controlled_values, m);
        if (queue_.names.size() >= kMaxQueue) flush();
        return true;
    }
    void applyMatrix(const ComplexT *matrix, const std::vector<std::size_t> &wires, bool inverse = false) {
        PLB200_ABORT_IF(wires.empty(), "Number of wires must be larger than 0");
        if (queue_matrix(matrix, {}, {}, wires, inverse)) return;
        flush();
        const auto w = detail::to_i64(wires);
        const auto m = detail::to_c128(matrix, std::size_t{1} << (2 * wires.size()));
        PLB200_ABI(plb200_sv_apply_matrix(h_, m.data(), nullptr, nullptr, 0, w.data(), static_cast<int64_t>(w.size()),
                                          inverse));
    }
    void applyMatrix(const std::vector<ComplexT> &matrix, const std::vector<std::size_t> &wires, bool inverse = false) {
        PLB200_ABORT_IF(matrix.size() != (std::size_t{1} << (2 * wires.size())),
                        "The size of matrix does not match with the given number of wires");
        applyMatrix(matrix.data(), wires, inverse);
    }
    void applyControlledMatrix(const ComplexT *matrix, const std::vector<std::size_t> &controlled_wires,
                               const std::vector<bool> &controlled_values, const std::vector<std::size_t> &wires,
                               bool inverse = false) {
        PLB200_ABORT_IF(wires.empty(), "Number of wires must be larger than 0");
        PLB200_ABORT_IF_NOT(controlled_wires.size() == controlled_values.size(),
                            "`controlled_wires` must have the same size as `controlled_values`.");
        if (queue_matrix(matrix, controlled_wires, controlled_values, wires, inverse)) return;
        flush();
        const auto w = detail::to_i64(wires), cw = detail::to_i64(controlled_wires);
        const auto cv = detail::to_u8(controlled_values);
        const auto m = detail::to_c128(matrix, std::size_t{1} << (2 * wires.size()));
        PLB200_ABI(plb200_sv_apply_matrix(h_, m.data(), cw.data(), cv.data(), static_cast<int64_t>(cw.size()),
                                          w.data(), static_cast<int64_t>(w.size()), inverse));
    }
    void applyPauliRot(const std::vector<std::size_t> &wires, bool inverse, const std::vector<PrecisionT> &params,
                       const std::string &word) {
        PLB200_ABORT_IF_NOT(wires.size() == word.size(), "wires and word have incompatible dimensions.");
        PLB200_ABORT_IF(params.empty(), "PauliRot needs one parameter");
        if (lazy_) {
            const std::string name = "PauliRot[" + word + "]";
            detail::OpsBlob one;
            one.template add<PrecisionT>(name, wires, inverse, {params[0]});
            const auto v = one.view();
            PLB200_ABI(plb200_validate_ops(static_cast<int64_t>(num_qubits_), &v));
            queue_.template add<PrecisionT>(name, wires, inverse, {params[0]});
            if (queue_.names.size() >= kMaxQueue) flush();
            return;
        }
        flush();
        const auto w = detail::to_i64(wires);
        PLB200_ABI(plb200_sv_apply_pauli_rot(h_, w.data(), static_cast<int64_t>(w.size()), inverse,
                                             static_cast<double>(params[0]), word.c_str()));
    }
    [[nodiscard]] auto applyGenerator(const std::string &opName, const std::vector<std::size_t> &wires,
                                      bool adj = false) -> PrecisionT {
        flush();
        const auto w = detail::to_i64(wires);
        double scale = 0;
        PLB200_ABI(plb200_sv_apply_generator(h_, opName.c_str(), nullptr, nullptr, 0, w.data(),
                                             static_cast<int64_t>(w.size()), adj, &scale));
        return static_cast<PrecisionT>(scale);
    }
    [[nodiscard]] auto applyGenerator(const std::string &opName, const std::vector<std::size_t> &controlled_wires,
                                      const std::vector<bool> &controlled_values,
                                      const std::vector<std::size_t> &wires, bool adj = false) -> PrecisionT {
        flush();
        const auto w = detail::to_i64(wires), cw = detail::to_i64(controlled_wires);
        const auto cv = detail::to_u8(controlled_values);
        double scale = 0;
        PLB200_ABI(plb200_sv_apply_generator(h_, opName.c_str(), cw.data(), cv.data(), static_cast<int64_t>(cw.size()),
                                             w.data(), static_cast<int64_t>(w.size()), adj, &scale));
        return static_cast<PrecisionT>(scale);
    }

    // ---- state preparation (Bindings.hpp:883-921) -----------------------------------------
    void resetStateVector(bool = false) {
        queue_ = detail::OpsBlob{}; // pending gates are overwritten anyway
        PLB200_ABI(plb200_sv_reset(h_));
    }
    void setBasisState(const std::vector<std::size_t> &state, const std::vector<std::size_t> &wires, bool = false) {
        PLB200_ABORT_IF(state.size() != wires.size(), "state and wires must have equal dimensions.");
        queue_ = detail::OpsBlob{};
        const auto s = detail::to_i64(state), w = detail::to_i64(wires);
        PLB200_ABI(plb200_sv_set_basis_state(h_, s.data(), w.data(), static_cast<int64_t>(w.size())));
    }
    void setBasisState(std::size_t index) {
        queue_ = detail::OpsBlob{};
        PLB200_ABI(plb200_sv_set_basis_state_index(h_, static_cast<int64_t>(index)));
    }
    void setStateVector(const ComplexT *state, std::size_t state_size, const std::vector<std::size_t> &wires,
                        bool = false) {
        PLB200_ABORT_IF_NOT(state_size == (std::size_t{1} << wires.size()), "Inconsistent state and wires dimensions.");
        queue_ = detail::OpsBlob{};
        const auto w = detail::to_i64(wires);
        const auto v = detail::to_c128(state, state_size);
        PLB200_ABI(plb200_sv_set_state_vector(h_, v.data(), w.data(), static_cast<int64_t>(w.size())));
    }
    void setStateVector(const std::vector<ComplexT> &state, const std::vector<std::size_t> &wires) {
        setStateVector(state.data(), state.size(), wires);
    }
    void setStateVector(const std::vector<std::size_t> &indices, const std::vector<ComplexT> &values) {
        PLB200_ABORT_IF(indices.size() != values.size(), "Indices and values length must match");
        flush(); // only the listed amplitudes are overwritten: the others must see the pending gates
        const auto i = detail::to_i64(indices);
        const auto v = detail::to_c128(values.data(), values.size());
        PLB200_ABI(plb200_sv_set_state_indices(h_, i.data(), v.data(), static_cast<int64_t>(i.size())));
    }
    void updateData(const StateVectorB200 &other) {
        queue_ = detail::OpsBlob{};
        PLB200_ABI(plb200_sv_d2d(h_, other.handle()));
    }
    void updateData(const ComplexT *host, std::size_t length) { CopyHostDataToGpu(host, length, false); }
    void updateData(const std::vector<ComplexT> &host) { CopyHostDataToGpu(host.data(), host.size(), false); }
    void collapse(std::size_t wire, bool branch) {
        flush();
        PLB200_ABI(plb200_sv_collapse(h_, static_cast<int64_t>(wire), branch));
    }
    void normalize() {
        flush();
        PLB200_ABI(plb200_sv_normalize(h_));
    }

    // ---- copies (StateVectorCudaBase.hpp) --------------------------------------------------
    void CopyHostDataToGpu(const ComplexT *host, std::size_t length, bool async = false) {
        PLB200_ABORT_IF_NOT(length == getLength(), "Sizes do not match for Host and GPU data");
        queue_ = detail::OpsBlob{};
        PLB200_ABI(plb200_sv_h2d(h_, host, static_cast<int64_t>(length), async));
    }
    void CopyGpuDataToHost(ComplexT *host, std::size_t length, bool async = false) const {
        PLB200_ABORT_IF_NOT(length == getLength(), "Sizes do not match for Host and GPU data");
        flush();
        PLB200_ABI(plb200_sv_d2h(h_, host, static_cast<int64_t>(length), async));
    }
    void CopyGpuDataToGpuIn(const StateVectorB200 &other, bool = false) { updateData(other); }
    [[nodiscard]] auto getDataVector() const -> std::vector<ComplexT> {
        std::vector<ComplexT> v(getLength());
        CopyGpuDataToHost(v.data(), v.size());
        return v;
    }
    [[nodiscard]] std::int64_t kernelLaunches() const {
        flush();
        return plb200_sv_kernel_launches(h_);
    }

    static bool isNativeGate(const std::string &name) {
        static const char *names[] = {"Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T", "PhaseShift",
                                      "RX", "RY", "RZ", "Rot", "CNOT", "CY", "CZ", "SWAP", "IsingXX", "IsingXY",
                                      "IsingYY", "IsingZZ", "ControlledPhaseShift", "CRX", "CRY", "CRZ", "CRot",
                                      "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus", "PSWAP",
                                      "Toffoli", "CSWAP", "DoubleExcitation", "DoubleExcitationMinus",
                                      "DoubleExcitationPlus", "MultiRZ", "GlobalPhase", "PCPhase"};
        for (const char *s : names)
            if (name == s) return true;
        return false;
    }

  private:
    static constexpr int prec() { return sizeof(PrecisionT) == 8 ? PLB200_C128 : PLB200_C64; }
    static std::size_t log2_checked(std::size_t length) {
        PLB200_ABORT_IF(length == 0 || (length & (length - 1)) != 0, "The size of provided data must be a power of 2.");
        std::size_t n = 0;
        while ((std::size_t{1} << n) < length) n++;
        return n;
    }
    static bool lazy_default() {
        const char *e = std::getenv("PLB200_LAZY");
        return !(e && e[0] == '0');
    }
    static constexpr std::size_t kMaxQueue = 1 << 16;
    plb200_sv *h_ = nullptr;
    std::size_t num_qubits_;
    DevTag<int> dev_tag_;
    bool lazy_ = lazy_default();
    mutable detail::OpsBlob queue_;
};

} // namespace Pennylane::LightningB200
