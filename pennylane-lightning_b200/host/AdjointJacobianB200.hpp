// OpsData / JacobianData / AdjointJacobian for the B200 backend, mirroring
// core/algorithms/JacobianData.hpp:39-393 and AdjointJacobianBase.hpp / AdjointJacobianLQubit.hpp:347-491.
// The sweep itself (lambda, H lambda_i, fused Im<H lambda|G|lambda> reductions) runs inside libplb200.
#pragma once
#include <complex>
#include <memory>
#include <exception>
#include <span>
#include <thread>
#include <string>
#include <vector>

#include "DevTag.hpp"
#include "ObservablesB200.hpp"
#include "StateVectorB200.hpp"

namespace Pennylane::LightningB200::Algorithms {

template <class StateVectorT> class OpsData {
    using PrecisionT = typename StateVectorT::PrecisionT;
    using ComplexT = typename StateVectorT::ComplexT;

  public:
    OpsData(std::vector<std::string> ops_name, const std::vector<std::vector<PrecisionT>> &ops_params,
            std::vector<std::vector<std::size_t>> ops_wires, std::vector<bool> ops_inverses,
            std::vector<std::vector<ComplexT>> ops_matrices, std::vector<std::vector<std::size_t>> ops_controlled_wires,
            std::vector<std::vector<bool>> ops_controlled_values)
        : ops_name_{std::move(ops_name)}, ops_params_{ops_params}, ops_wires_{std::move(ops_wires)},
          ops_inverses_{std::move(ops_inverses)}, ops_matrices_{std::move(ops_matrices)},
          ops_controlled_wires_{std::move(ops_controlled_wires)},
          ops_controlled_values_{std::move(ops_controlled_values)} {
        count();
    }
    OpsData(std::vector<std::string> ops_name, const std::vector<std::vector<PrecisionT>> &ops_params,
            std::vector<std::vector<std::size_t>> ops_wires, std::vector<bool> ops_inverses,
            std::vector<std::vector<ComplexT>> ops_matrices)
        : ops_name_{std::move(ops_name)}, ops_params_{ops_params}, ops_wires_{std::move(ops_wires)},
          ops_inverses_{std::move(ops_inverses)}, ops_matrices_{std::move(ops_matrices)},
          ops_controlled_wires_(ops_name_.size()), ops_controlled_values_(ops_name_.size()) {
        count();
    }
    OpsData(const std::vector<std::string> &ops_name, const std::vector<std::vector<PrecisionT>> &ops_params,
            std::vector<std::vector<std::size_t>> ops_wires, std::vector<bool> ops_inverses)
        : ops_name_{ops_name}, ops_params_{ops_params}, ops_wires_{std::move(ops_wires)},
          ops_inverses_{std::move(ops_inverses)}, ops_matrices_(ops_name.size()),
          ops_controlled_wires_(ops_name.size()), ops_controlled_values_(ops_name.size()) {
        count();
    }
    [[nodiscard]] auto getSize() const -> std::size_t { return ops_name_.size(); }
    [[nodiscard]] auto getOpsName() const -> const std::vector<std::string> & { return ops_name_; }
    [[nodiscard]] auto getOpsParams() const -> const std::vector<std::vector<PrecisionT>> & { return ops_params_; }
    [[nodiscard]] auto getOpsWires() const -> const std::vector<std::vector<std::size_t>> & { return ops_wires_; }
    [[nodiscard]] auto getOpsControlledWires() const -> const std::vector<std::vector<std::size_t>> & {
        return ops_controlled_wires_;
    }
    [[nodiscard]] auto getOpsControlledValues() const -> const std::vector<std::vector<bool>> & {
        return ops_controlled_values_;
    }
    [[nodiscard]] auto getOpsInverses() const -> const std::vector<bool> & { return ops_inverses_; }
    [[nodiscard]] auto getOpsMatrices() const -> const std::vector<std::vector<ComplexT>> & { return ops_matrices_; }
    [[nodiscard]] auto hasParams(std::size_t index) const -> bool { return !ops_params_[index].empty(); }
    [[nodiscard]] auto getNumParOps() const -> std::size_t { return num_par_ops_; }
    [[nodiscard]] auto getNumNonParOps() const -> std::size_t { return num_nonpar_ops_; }
    [[nodiscard]] auto getTotalNumParams() const -> std::size_t {
        std::size_t n = 0;
        for (const auto &p : ops_params_) n += p.size();
        return n;
    }
    void fill(detail::OpsBlob &blob) const {
        for (std::size_t i = 0; i < ops_name_.size(); i++)
            blob.add<PrecisionT>(ops_name_[i], ops_wires_[i], ops_inverses_[i], ops_params_[i], ops_controlled_wires_[i],
                                 ops_controlled_values_[i], ops_matrices_[i]);
    }

  private:
    void count() {
        num_par_ops_ = 0;
        for (const auto &p : ops_params_) num_par_ops_ += static_cast<std::size_t>(!p.empty());
        num_nonpar_ops_ = ops_params_.size() - num_par_ops_;
    }
    std::size_t num_par_ops_ = 0, num_nonpar_ops_ = 0;
    std::vector<std::string> ops_name_;
    std::vector<std::vector<PrecisionT>> ops_params_;
    std::vector<std::vector<std::size_t>> ops_wires_;
    std::vector<bool> ops_inverses_;
    std::vector<std::vector<ComplexT>> ops_matrices_;
    std::vector<std::vector<std::size_t>> ops_controlled_wires_;
    std::vector<std::vector<bool>> ops_controlled_values_;
};

template <class StateVectorT> class JacobianData {
    using CFP_t = typename StateVectorT::CFP_t;
    using ObsPtr = std::shared_ptr<Observables::Observable<StateVectorT>>;

  public:
    JacobianData(std::size_t num_params, std::size_t num_elem, const CFP_t *sv_ptr, std::vector<ObsPtr> obs,
                 OpsData<StateVectorT> ops, std::vector<std::size_t> trainP)
        : num_parameters{num_params}, num_elements{num_elem}, psi{sv_ptr}, observables{std::move(obs)},
          operations{std::move(ops)}, trainableParams{std::move(trainP)} {}
    [[nodiscard]] auto getNumParams() const -> std::size_t { return num_parameters; }
    [[nodiscard]] auto getSizeStateVec() const -> std::size_t { return num_elements; }
    [[nodiscard]] auto getPtrStateVec() const -> const CFP_t * { return psi; }
    [[nodiscard]] auto getObservables() const -> const std::vector<ObsPtr> & { return observables; }
    [[nodiscard]] auto getNumObservables() const -> std::size_t { return observables.size(); }
    [[nodiscard]] auto getOperations() const -> const OpsData<StateVectorT> & { return operations; }
    [[nodiscard]] auto getTrainableParams() const -> const std::vector<std::size_t> & { return trainableParams; }
    [[nodiscard]] auto hasTrainableParams() const -> bool { return !trainableParams.empty(); }

  private:
    std::size_t num_parameters, num_elements;
    const CFP_t *psi;
    const std::vector<ObsPtr> observables;
    const OpsData<StateVectorT> operations;
    const std::vector<std::size_t> trainableParams;
};

template <class StateVectorT> class AdjointJacobian {
    using PrecisionT = typename StateVectorT::PrecisionT;

  public:
    AdjointJacobian() = default;
    // jac: observable-major [obs][param] as returned to Python (Bindings.hpp:710-729).
    // `ref_data` must be the state vector jd.getPtrStateVec() points into (device resident).
    void adjointJacobian(std::span<PrecisionT> jac, const JacobianData<StateVectorT> &jd, const StateVectorT &ref_data,
                         bool apply_operations = false) {
        const auto &obs = jd.getObservables();
        const auto &tp = jd.getTrainableParams();
        if (!jd.hasTrainableParams()) return;
        PLB200_ABORT_IF_NOT(jac.size() == tp.size() * obs.size(),
                            "The size of preallocated jacobian must be same as the number of trainable parameters "
                            "times the number of observables provided.");
        detail::OpsBlob blob;
        jd.getOperations().fill(blob);
        const auto v = blob.view();
        std::vector<const plb200_obs *> hs;
        for (const auto &o : obs) hs.push_back(o->handle());
        const auto t = detail::to_i64(tp);
        std::vector<double> out(jac.size());
        PLB200_ABI(plb200_adjoint_jacobian(ref_data.handle(), hs.data(), static_cast<int64_t>(hs.size()), &v, t.data(),
                                           static_cast<int64_t>(t.size()), apply_operations, out.data()));
        for (std::size_t i = 0; i < out.size(); i++) jac[i] = static_cast<PrecisionT>(out[i]);
    }
    // Observable batching over the GPUs of the box (lightning_gpu/algorithms/AdjointJacobianGPU.hpp:123-225): the
    // observables are cut into one contiguous chunk per device, every chunk runs its own backward sweep on its own
    // GPU from a copy of the state (one std::thread and one DevTag per device), rows land in jac in order.
    void batchAdjointJacobian(std::span<PrecisionT> jac, const JacobianData<StateVectorT> &jd,
                              const StateVectorT &ref_data, bool apply_operations = false) {
        const auto &obs = jd.getObservables();
        const auto &tp = jd.getTrainableParams();
        if (!jd.hasTrainableParams()) return;
        const std::size_t n_dev = std::min<std::size_t>(DevicePool<int>::getTotalDevices(), obs.size());
        if (n_dev <= 1) {
            adjointJacobian(jac, jd, ref_data, apply_operations);
            return;
        }
        PLB200_ABORT_IF_NOT(jac.size() == tp.size() * obs.size(),
                            "The size of preallocated jacobian must be same as the number of trainable parameters "
                            "times the number of observables provided.");
        const auto host = ref_data.getDataVector();
        detail::OpsBlob blob;
        jd.getOperations().fill(blob);
        const auto v = blob.view();
        const auto t = detail::to_i64(tp);
        std::vector<std::thread> threads;
        std::vector<std::exception_ptr> errors(n_dev);
        const std::size_t per = (obs.size() + n_dev - 1) / n_dev;
        for (std::size_t d = 0; d < n_dev; d++) {
            const std::size_t first = d * per, last = std::min(obs.size(), first + per);
            if (first >= last) break;
            threads.emplace_back([&, d, first, last]() {
                try {
                    // device d's own copy of the state (device 0 could share ref_data; a copy keeps it uniform)
                    StateVectorT local(host.data(), host.size(), DevTag<int>{static_cast<int>(d), nullptr});
                    std::vector<const plb200_obs *> hs;
                    for (std::size_t o = first; o < last; o++) hs.push_back(obs[o]->handle());
                    std::vector<double> out((last - first) * tp.size());
                    PLB200_ABI(plb200_adjoint_jacobian(local.handle(), hs.data(), static_cast<int64_t>(hs.size()), &v,
                                                       t.data(), static_cast<int64_t>(t.size()), apply_operations,
                                                       out.data()));
                    for (std::size_t i = 0; i < out.size(); i++) jac[first * tp.size() + i] = static_cast<PrecisionT>(out[i]);
                } catch (...) {
                    errors[d] = std::current_exception();
                }
            });
        }
        for (auto &th : threads) th.join();
        for (auto &e : errors)
            if (e) std::rethrow_exception(e);
    }
};

// VectorJacobianProduct (lightning_qubit/algorithms/VectorJacobianProduct.hpp:43-163; bound as
// VectorJacobianProductC64/128, LQubitBindings.hpp:413-462): vjp[k] = sum_i conj(dy_i) d psi_i / d theta_k.
template <class StateVectorT> class VectorJacobianProduct {
    using PrecisionT = typename StateVectorT::PrecisionT;
    using ComplexT = typename StateVectorT::ComplexT;

  public:
    VectorJacobianProduct() = default;
    void operator()(std::span<ComplexT> jac, const JacobianData<StateVectorT> &jd, std::span<const ComplexT> dy,
                    const StateVectorT &ref_data, bool apply_operations = false) {
        PLB200_ABORT_IF_NOT(dy.size() == jd.getSizeStateVec(), "dy must have the size of the state vector");
        if (!jd.hasTrainableParams()) return;
        const auto &tp = jd.getTrainableParams();
        PLB200_ABORT_IF_NOT(jac.size() == tp.size(),
                            "The size of preallocated jacobian must be same as the number of trainable parameters.");
        detail::OpsBlob blob;
        jd.getOperations().fill(blob);
        const auto v = blob.view();
        const auto d = detail::to_c128(dy.data(), dy.size());
        const auto t = detail::to_i64(tp);
        std::vector<double> out(2 * jac.size());
        PLB200_ABI(plb200_vjp(ref_data.handle(), d.data(), &v, t.data(), static_cast<int64_t>(t.size()), apply_operations,
                              out.data()));
        for (std::size_t i = 0; i < jac.size(); i++)
            jac[i] = ComplexT{static_cast<PrecisionT>(out[2 * i]), static_cast<PrecisionT>(out[2 * i + 1])};
    }
};

} // namespace Pennylane::LightningB200::Algorithms
