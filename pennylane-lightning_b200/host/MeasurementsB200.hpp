// Measurements<StateVectorB200> — mirrors MeasurementsBase (core/measurements/MeasurementsBase.hpp:
// 58-147: seed handling, expval/var/probs/generate_samples) and the backend overloads bound for LGPU
// (lightning_gpu/bindings/LGPUBindings.hpp:65-150; MeasurementsGPU.hpp:449-530 Pauli words).
#pragma once
#include <optional>
#include <string>
#include <vector>

#include "ObservablesB200.hpp"
#include "StateVectorB200.hpp"

namespace Pennylane::LightningB200::Measures {

template <class StateVectorT> class Measurements {
  public:
    using PrecisionT = typename StateVectorT::PrecisionT;
    using ComplexT = typename StateVectorT::ComplexT;
    using ObservableT = Observables::Observable<StateVectorT>;

    explicit Measurements(StateVectorT &statevector) : sv_{statevector} {}

    void setSeed(const std::optional<std::size_t> &seed = std::nullopt) { seed_ = seed; }

    auto probs() -> std::vector<PrecisionT> {
        std::vector<double> p(sv_.getLength());
        PLB200_ABI(plb200_probs(sv_.handle(), nullptr, -1, p.data()));
        return {p.begin(), p.end()};
    }
    auto probs(const std::vector<std::size_t> &wires) -> std::vector<PrecisionT> {
        PLB200_ABORT_IF(wires.size() > sv_.getNumQubits(), "Invalid number of wires");
        const auto w = detail::to_i64(wires);
        std::vector<double> p(std::size_t{1} << wires.size());
        PLB200_ABI(plb200_probs(sv_.handle(), w.data(), static_cast<int64_t>(w.size()), p.data()));
        return {p.begin(), p.end()};
    }
    auto expval(const ObservableT &obs) -> PrecisionT {
        double r = 0;
        PLB200_ABI(plb200_expval_obs(sv_.handle(), obs.handle(), &r));
        return static_cast<PrecisionT>(r);
    }
    auto var(const ObservableT &obs) -> PrecisionT {
        double r = 0;
        PLB200_ABI(plb200_var_obs(sv_.handle(), obs.handle(), &r));
        return static_cast<PrecisionT>(r);
    }
    auto expval(const std::string &operation, const std::vector<std::size_t> &wires) -> PrecisionT {
        const auto w = detail::to_i64(wires);
        double r = 0;
        PLB200_ABI(plb200_expval_named(sv_.handle(), operation.c_str(), w.data(), static_cast<int64_t>(w.size()), &r));
        return static_cast<PrecisionT>(r);
    }
    auto var(const std::string &operation, const std::vector<std::size_t> &wires) -> PrecisionT {
        const auto w = detail::to_i64(wires);
        double r = 0;
        PLB200_ABI(plb200_var_named(sv_.handle(), operation.c_str(), w.data(), static_cast<int64_t>(w.size()), &r));
        return static_cast<PrecisionT>(r);
    }
    auto expval(const std::vector<ComplexT> &matrix, const std::vector<std::size_t> &wires) -> PrecisionT {
        PLB200_ABORT_IF(matrix.size() != (std::size_t{1} << (2 * wires.size())),
                        "The size of matrix does not match with the given number of wires");
        const auto w = detail::to_i64(wires);
        const auto m = detail::to_c128(matrix.data(), matrix.size());
        double r = 0;
        PLB200_ABI(plb200_expval_matrix(sv_.handle(), m.data(), w.data(), static_cast<int64_t>(w.size()), &r));
        return static_cast<PrecisionT>(r);
    }
    auto var(const std::vector<ComplexT> &matrix, const std::vector<std::size_t> &wires) -> PrecisionT {
        PLB200_ABORT_IF(matrix.size() != (std::size_t{1} << (2 * wires.size())),
                        "The size of matrix does not match with the given number of wires");
        const auto w = detail::to_i64(wires);
        const auto m = detail::to_c128(matrix.data(), matrix.size());
        double r = 0;
        PLB200_ABI(plb200_var_matrix(sv_.handle(), m.data(), w.data(), static_cast<int64_t>(w.size()), &r));
        return static_cast<PrecisionT>(r);
    }
    // sum_k coeffs[k] <pauli_words[k]> in ONE fused launch (LGPU: custatevecComputeExpectationsOnPauliBasis)
    auto expval(const std::vector<std::string> &pauli_words, const std::vector<std::vector<std::size_t>> &tgts,
                const std::vector<PrecisionT> &coeffs) -> PrecisionT {
        PLB200_ABORT_IF(pauli_words.size() != tgts.size() || pauli_words.size() != coeffs.size(),
                        "pauli_words, wires and coeffs must have the same size");
        std::vector<const char *> words;
        std::vector<int64_t> flat, off{0};
        for (std::size_t k = 0; k < pauli_words.size(); k++) {
            words.push_back(pauli_words[k].c_str());
            flat.insert(flat.end(), tgts[k].begin(), tgts[k].end());
            off.push_back(static_cast<int64_t>(flat.size()));
        }
        if (flat.capacity() == 0) flat.reserve(1);
        const auto c = detail::to_f64(coeffs);
        double r = 0;
        PLB200_ABI(plb200_expval_pauli_words(sv_.handle(), words.data(), flat.data(), off.data(), c.data(),
                                             static_cast<int64_t>(words.size()), &r));
        return static_cast<PrecisionT>(r);
    }
    auto generate_samples(std::size_t num_samples) -> std::vector<std::size_t> {
        std::vector<uint64_t> out(num_samples * sv_.getNumQubits());
        PLB200_ABI(plb200_generate_samples(sv_.handle(), nullptr, -1, static_cast<int64_t>(num_samples), seed_arg(),
                                           out.data()));
        return {out.begin(), out.end()};
    }
    auto generate_samples(const std::vector<std::size_t> &wires, std::size_t num_samples) -> std::vector<std::size_t> {
        const auto w = detail::to_i64(wires);
        std::vector<uint64_t> out(num_samples * wires.size());
        PLB200_ABI(plb200_generate_samples(sv_.handle(), w.data(), static_cast<int64_t>(w.size()),
                                           static_cast<int64_t>(num_samples), seed_arg(), out.data()));
        return {out.begin(), out.end()};
    }

  private:
    int64_t seed_arg() const { return seed_.has_value() ? static_cast<int64_t>(seed_.value()) : int64_t{-1}; }
    StateVectorT &sv_;
    std::optional<std::size_t> seed_{std::nullopt};
};

} // namespace Pennylane::LightningB200::Measures
