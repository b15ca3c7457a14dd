// Observable hierarchy of the B200 backend, mirroring core/observables/Observables.hpp:36-584
// (Observable, NamedObsBase, HermitianObsBase, TensorProdObsBase, HamiltonianBase) and the LGPU
// finals in lightning_gpu/observables/ObservablesGPU.hpp.  Each object owns a plb200_obs tree.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "StateVectorB200.hpp"

namespace Pennylane::LightningB200::Observables {

template <class StateVectorT> class Observable {
  public:
    using PrecisionT = typename StateVectorT::PrecisionT;
    virtual ~Observable() {
        if (h_) plb200_obs_destroy(h_);
    }
    Observable(const Observable &) = delete;
    Observable &operator=(const Observable &) = delete;

    // Apply the observable to the given state vector in place (Observables.hpp:63)
    virtual void applyInPlace(StateVectorT &sv) const { PLB200_ABI(plb200_obs_apply(h_, sv.handle())); }
    // Rotate `sv` into the eigenbasis of the observable and report eigenvalues / wires for
    // shot-based measurement (Observables.hpp:63-78,172-202,414-448).
    virtual void applyInPlaceShots(StateVectorT &sv, std::vector<std::vector<PrecisionT>> &eigenValues,
                                   std::vector<std::size_t> &ob_wires) const = 0;
    [[nodiscard]] virtual auto getObsName() const -> std::string = 0;
    [[nodiscard]] virtual auto getWires() const -> std::vector<std::size_t> = 0;
    [[nodiscard]] virtual auto getObs() const -> std::vector<std::shared_ptr<Observable<StateVectorT>>> { return {}; }
    [[nodiscard]] virtual auto getCoeffs() const -> std::vector<PrecisionT> { return {}; }
    [[nodiscard]] bool operator==(const Observable &other) const {
        return typeid(*this) == typeid(other) && isEqual(other);
    }
    [[nodiscard]] bool operator!=(const Observable &other) const { return !(*this == other); }
    [[nodiscard]] const plb200_obs *handle() const { return h_; }

  protected:
    Observable() = default;
    [[nodiscard]] virtual bool isEqual(const Observable &other) const = 0;
    plb200_obs *h_ = nullptr;
};

template <class StateVectorT> class NamedObs final : public Observable<StateVectorT> {
  public:
    using PrecisionT = typename StateVectorT::PrecisionT;
    NamedObs(std::string obs_name, std::vector<std::size_t> wires, std::vector<PrecisionT> params = {})
        : obs_name_{std::move(obs_name)}, wires_{std::move(wires)}, params_{std::move(params)} {
        const auto w = detail::to_i64(wires_);
        const auto p = detail::to_f64(params_);
        PLB200_ABI(plb200_obs_named(&this->h_, obs_name_.c_str(), w.data(), static_cast<int64_t>(w.size()), p.data(),
                                    static_cast<int64_t>(p.size())));
    }
    [[nodiscard]] auto getObsName() const -> std::string override {
        std::ostringstream s;
        s << obs_name_ << "[";
        for (std::size_t i = 0; i < wires_.size(); i++) s << (i ? ", " : "") << wires_[i];
        s << "]";
        return s.str();
    }
    [[nodiscard]] auto getWires() const -> std::vector<std::size_t> override { return wires_; }
    void applyInPlaceShots(StateVectorT &sv, std::vector<std::vector<PrecisionT>> &eigenValues,
                           std::vector<std::size_t> &ob_wires) const override {
        ob_wires.clear();
        eigenValues.clear();
        ob_wires.push_back(wires_[0]);
        if (obs_name_ == "PauliX") {
            sv.applyOperation("Hadamard", wires_, false);
        } else if (obs_name_ == "PauliY") {
            sv.applyOperations({"PauliZ", "S", "Hadamard"}, {wires_, wires_, wires_}, {false, false, false});
        } else if (obs_name_ == "Hadamard") {
            const PrecisionT theta = -M_PI / 4.0;
            sv.applyOperation("RY", wires_, false, {theta});
        } else if (obs_name_ == "PauliZ" || obs_name_ == "Identity") {
        } else {
            PLB200_ABORT("Provided NamedObs does not support shot measurement.");
        }
        if (obs_name_ == "Identity") eigenValues.push_back({1, 1});
        else eigenValues.push_back({1, -1});
    }

  private:
    [[nodiscard]] bool isEqual(const Observable<StateVectorT> &other) const override {
        const auto &o = static_cast<const NamedObs &>(other);
        return obs_name_ == o.obs_name_ && wires_ == o.wires_ && params_ == o.params_;
    }
    std::string obs_name_;
    std::vector<std::size_t> wires_;
    std::vector<PrecisionT> params_;
};

template <class StateVectorT> class HermitianObs final : public Observable<StateVectorT> {
  public:
    using PrecisionT = typename StateVectorT::PrecisionT;
    using ComplexT = typename StateVectorT::ComplexT;
    using MatrixT = std::vector<ComplexT>;
    HermitianObs(MatrixT matrix, std::vector<std::size_t> wires) : matrix_{std::move(matrix)}, wires_{std::move(wires)} {
        PLB200_ABORT_IF(matrix_.size() != (std::size_t{1} << (2 * wires_.size())),
                        "The size of matrix does not match with the given number of wires");
        const auto w = detail::to_i64(wires_);
        const auto m = detail::to_c128(matrix_.data(), matrix_.size());
        PLB200_ABI(plb200_obs_hermitian(&this->h_, m.data(), w.data(), static_cast<int64_t>(w.size())));
    }
    [[nodiscard]] auto getMatrix() const -> const MatrixT & { return matrix_; }
    [[nodiscard]] auto getObsName() const -> std::string override { return "Hermitian"; }
    // Rotate into the observable's eigenbasis (Observables.hpp:236-303): the reference diagonalises with LAPACK
    // zheev loaded at run time from scipy-openblas; here the engine's own Jacobi solver (plb200_hermitian_eigh).
    void applyInPlaceShots(StateVectorT &sv, std::vector<std::vector<PrecisionT>> &eigenValues,
                           std::vector<std::size_t> &ob_wires) const override {
        if (eigenVals_.empty()) {
            const std::size_t dim = std::size_t{1} << wires_.size();
            const auto m = detail::to_c128(matrix_.data(), matrix_.size());
            std::vector<double> ev(dim), u(2 * dim * dim);
            PLB200_ABI(plb200_hermitian_eigh(m.data(), static_cast<int64_t>(dim), ev.data(), u.data()));
            eigenVals_.assign(ev.begin(), ev.end());
            unitary_.resize(dim * dim);
            for (std::size_t i = 0; i < dim * dim; i++)
                unitary_[i] = ComplexT{static_cast<PrecisionT>(u[2 * i]), static_cast<PrecisionT>(u[2 * i + 1])};
        }
        eigenValues.clear();
        ob_wires = wires_;
        sv.applyMatrix(unitary_, wires_);
        eigenValues.push_back(eigenVals_);
    }
    [[nodiscard]] auto getWires() const -> std::vector<std::size_t> override { return wires_; }

  private:
    [[nodiscard]] bool isEqual(const Observable<StateVectorT> &other) const override {
        const auto &o = static_cast<const HermitianObs &>(other);
        return matrix_ == o.matrix_ && wires_ == o.wires_;
    }
    MatrixT matrix_;
    std::vector<std::size_t> wires_;
    mutable std::vector<PrecisionT> eigenVals_;
    mutable MatrixT unitary_;
};

// SparseHamiltonian (Observables.hpp:592-699, lightning_gpu/observables/ObservablesGPU.hpp): CSR over the full
// index space, applied by the engine's CSR kernel.
template <class StateVectorT> class SparseHamiltonian final : public Observable<StateVectorT> {
  public:
    using PrecisionT = typename StateVectorT::PrecisionT;
    using ComplexT = typename StateVectorT::ComplexT;
    using IdxT = int64_t;
    SparseHamiltonian(std::vector<ComplexT> data, std::vector<IdxT> indices, std::vector<IdxT> offsets,
                      std::vector<std::size_t> wires)
        : data_{std::move(data)}, indices_{std::move(indices)}, offsets_{std::move(offsets)}, wires_{std::move(wires)} {
        PLB200_ABORT_IF(data_.size() != indices_.size(), "data and indices must have the same size");
        PLB200_ABORT_IF(offsets_.empty(), "offsets must not be empty");
        const auto d = detail::to_c128(data_.data(), data_.size());
        PLB200_ABI(plb200_obs_sparse(&this->h_, offsets_.data(), indices_.data(), d.data(),
                                     static_cast<int64_t>(offsets_.size()) - 1));
    }
    static auto create(std::initializer_list<ComplexT> data, std::initializer_list<IdxT> indices,
                       std::initializer_list<IdxT> offsets, std::initializer_list<std::size_t> wires)
        -> std::shared_ptr<SparseHamiltonian> {
        return std::make_shared<SparseHamiltonian>(std::vector<ComplexT>(data), std::vector<IdxT>(indices),
                                                   std::vector<IdxT>(offsets), std::vector<std::size_t>(wires));
    }
    void applyInPlaceShots(StateVectorT &, std::vector<std::vector<PrecisionT>> &,
                           std::vector<std::size_t> &) const override {
        PLB200_ABORT("SparseHamiltonian observables do not support shot measurement.");
    }
    [[nodiscard]] auto getObsName() const -> std::string override {
        std::ostringstream s;
        s << "SparseHamiltonian: {\n'data' : \n";
        for (const auto &d : data_) s << "{" << d.real() << ", " << d.imag() << "}, ";
        s << "\n'indices' : \n";
        for (const auto &i : indices_) s << i << ", ";
        s << "\n'offsets' : \n";
        for (const auto &o : offsets_) s << o << ", ";
        s << "\n}";
        return s.str();
    }
    [[nodiscard]] auto getWires() const -> std::vector<std::size_t> override { return wires_; }

  private:
    [[nodiscard]] bool isEqual(const Observable<StateVectorT> &other) const override {
        const auto &o = static_cast<const SparseHamiltonian &>(other);
        return data_ == o.data_ && indices_ == o.indices_ && offsets_ == o.offsets_;
    }
    std::vector<ComplexT> data_;
    std::vector<IdxT> indices_, offsets_;
    std::vector<std::size_t> wires_;
};

template <class StateVectorT> class TensorProdObs final : public Observable<StateVectorT> {
  public:
    using ObsPtr = std::shared_ptr<Observable<StateVectorT>>;
    explicit TensorProdObs(std::vector<ObsPtr> obs) : obs_{std::move(obs)} {
        std::vector<const plb200_obs *> hs;
        for (const auto &o : obs_) {
            hs.push_back(o->handle());
            const auto w = o->getWires();
            all_wires_.insert(all_wires_.end(), w.begin(), w.end());
        }
        PLB200_ABI(plb200_obs_tensor(&this->h_, hs.data(), static_cast<int64_t>(hs.size())));
    }
    static auto create(std::initializer_list<ObsPtr> obs) -> std::shared_ptr<TensorProdObs> {
        return std::make_shared<TensorProdObs>(std::vector<ObsPtr>(obs));
    }
    [[nodiscard]] auto getSize() const -> std::size_t { return obs_.size(); }
    void applyInPlaceShots(StateVectorT &sv, std::vector<std::vector<typename StateVectorT::PrecisionT>> &eigenValues,
                           std::vector<std::size_t> &ob_wires) const override {
        for (const auto &ob : obs_)
            if (ob->getObsName().find("Hamiltonian") != std::string::npos)
                PLB200_ABORT("Hamiltonian observables as a term of an TensorProd observable do not support shot "
                             "measurement.");
        eigenValues.clear();
        ob_wires.clear();
        for (const auto &ob : obs_) {
            std::vector<std::vector<typename StateVectorT::PrecisionT>> ev;
            std::vector<std::size_t> w;
            ob->applyInPlaceShots(sv, ev, w);
            ob_wires.push_back(w[0]);
            eigenValues.push_back(ev[0]);
        }
    }
    [[nodiscard]] auto getWires() const -> std::vector<std::size_t> override { return all_wires_; }
    [[nodiscard]] auto getObs() const -> std::vector<ObsPtr> override { return obs_; }
    [[nodiscard]] auto getObsName() const -> std::string override {
        std::ostringstream s;
        for (std::size_t i = 0; i < obs_.size(); i++) s << (i ? " @ " : "") << obs_[i]->getObsName();
        return s.str();
    }

  private:
    [[nodiscard]] bool isEqual(const Observable<StateVectorT> &other) const override {
        const auto &o = static_cast<const TensorProdObs &>(other);
        if (obs_.size() != o.obs_.size()) return false;
        for (std::size_t i = 0; i < obs_.size(); i++)
            if (*obs_[i] != *o.obs_[i]) return false;
        return true;
    }
    std::vector<ObsPtr> obs_;
    std::vector<std::size_t> all_wires_;
};

template <class StateVectorT> class Hamiltonian final : public Observable<StateVectorT> {
  public:
    using PrecisionT = typename StateVectorT::PrecisionT;
    using ObsPtr = std::shared_ptr<Observable<StateVectorT>>;
    Hamiltonian(std::vector<PrecisionT> coeffs, std::vector<ObsPtr> obs) : coeffs_{std::move(coeffs)}, obs_{std::move(obs)} {
        PLB200_ABORT_IF(coeffs_.size() != obs_.size(), "coeffs and obs must have the same size");
        std::vector<const plb200_obs *> hs;
        for (const auto &o : obs_) hs.push_back(o->handle());
        const auto c = detail::to_f64(coeffs_);
        PLB200_ABI(plb200_obs_hamiltonian(&this->h_, c.data(), hs.data(), static_cast<int64_t>(hs.size())));
    }
    static auto create(std::initializer_list<PrecisionT> coeffs, std::initializer_list<ObsPtr> obs)
        -> std::shared_ptr<Hamiltonian> {
        return std::make_shared<Hamiltonian>(std::vector<PrecisionT>(coeffs), std::vector<ObsPtr>(obs));
    }
    [[nodiscard]] auto getWires() const -> std::vector<std::size_t> override {
        std::vector<std::size_t> all;
        for (const auto &o : obs_)
            for (auto w : o->getWires())
                if (std::find(all.begin(), all.end(), w) == all.end()) all.push_back(w);
        std::sort(all.begin(), all.end());
        return all;
    }
    [[nodiscard]] auto getObs() const -> std::vector<ObsPtr> override { return obs_; }
    [[nodiscard]] auto getCoeffs() const -> std::vector<PrecisionT> override { return coeffs_; }
    void applyInPlaceShots(StateVectorT &, std::vector<std::vector<PrecisionT>> &,
                           std::vector<std::size_t> &) const override {
        PLB200_ABORT("Hamiltonian observables as a term of an observable do not support shot measurement.");
    }
    [[nodiscard]] auto getObsName() const -> std::string override {
        std::ostringstream s;
        s << "Hamiltonian: { 'coeffs' : [";
        for (std::size_t i = 0; i < coeffs_.size(); i++) s << (i ? ", " : "") << coeffs_[i];
        s << "], 'observables' : [";
        for (std::size_t i = 0; i < obs_.size(); i++) s << (i ? ", " : "") << obs_[i]->getObsName();
        s << "]}";
        return s.str();
    }

  private:
    [[nodiscard]] bool isEqual(const Observable<StateVectorT> &other) const override {
        const auto &o = static_cast<const Hamiltonian &>(other);
        if (coeffs_ != o.coeffs_ || obs_.size() != o.obs_.size()) return false;
        for (std::size_t i = 0; i < obs_.size(); i++)
            if (*obs_[i] != *o.obs_[i]) return false;
        return true;
    }
    std::vector<PrecisionT> coeffs_;
    std::vector<ObsPtr> obs_;
};

} // namespace Pennylane::LightningB200::Observables
