// Measurements<StateVectorB200> — mirrors MeasurementsBase (core/measurements/MeasurementsBase.hpp:
// 58-147: seed handling, expval/var/probs/generate_samples) and the backend overloads bound for LGPU
// (lightning_gpu/bindings/LGPUBindings.hpp:65-150; MeasurementsGPU.hpp:449-530 Pauli words).
#pragma once
#include <numeric>
#include <optional>
#include <string>
#include <unordered_map>
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
    // CSR overloads (lightning_gpu/bindings/LGPUBindings.hpp:65-150): <psi|A|psi> and its variance for a sparse
    // matrix over the full index space, through the engine's own CSR kernel
    auto expval(const int64_t *indptr, std::size_t indptr_size, const int64_t *indices, const ComplexT *data,
                std::size_t nnz) -> PrecisionT {
        const auto d = detail::to_c128(data, nnz);
        double r = 0;
        PLB200_ABI(plb200_expval_sparse(sv_.handle(), indptr, indices, d.data(), static_cast<int64_t>(indptr_size) - 1, &r));
        return static_cast<PrecisionT>(r);
    }
    auto var(const int64_t *indptr, std::size_t indptr_size, const int64_t *indices, const ComplexT *data,
             std::size_t nnz) -> PrecisionT {
        const auto d = detail::to_c128(data, nnz);
        double r = 0;
        PLB200_ABI(plb200_var_sparse(sv_.handle(), indptr, indices, d.data(), static_cast<int64_t>(indptr_size) - 1, &r));
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

    // ---- shot-based API (MeasurementsBase.hpp:159-521), built on generate_samples + applyInPlaceShots
    auto expval(const ObservableT &obs, const std::size_t &num_shots, const std::vector<std::size_t> &shot_range = {})
        -> PrecisionT {
        PrecisionT result{0.0};
        const auto name = obs.getObsName();
        if (name.find("SparseHamiltonian") != std::string::npos) {
            PLB200_ABORT("SparseHamiltonian observables do not support shot measurement.");
        } else if (name.find("Hamiltonian") != std::string::npos) {
            const auto coeffs = obs.getCoeffs();
            const auto terms = obs.getObs();
            for (std::size_t k = 0; k < coeffs.size(); k++) result += coeffs[k] * expval(*terms[k], num_shots, shot_range);
        } else {
            const auto s = measure_with_samples(obs, num_shots, shot_range);
            result = std::accumulate(s.begin(), s.end(), 0.0);
            result /= s.size();
        }
        return result;
    }
    auto measure_with_samples(const ObservableT &obs, const std::size_t &num_shots,
                              const std::vector<std::size_t> &shot_range) -> std::vector<PrecisionT> {
        const std::size_t num_qubits = sv_.getTotalNumQubits();
        std::vector<std::size_t> obs_wires;
        std::vector<std::vector<PrecisionT>> eigenValues;
        const auto sub_samples = _sample_state(obs, num_shots, shot_range, obs_wires, eigenValues);
        const std::size_t num_samples = shot_range.empty() ? num_shots : shot_range.size();
        std::vector<PrecisionT> obs_samples(num_samples, 0);
        std::vector<PrecisionT> eigenVals = eigenValues[0];
        for (std::size_t i = 1; i < eigenValues.size(); i++) {
            std::vector<PrecisionT> next;
            for (auto a : eigenVals)
                for (auto b : eigenValues[i]) next.push_back(a * b); // kronProd
            eigenVals.swap(next);
        }
        for (std::size_t i = 0; i < num_samples; i++) {
            std::size_t idx = 0, wire_idx = 0;
            for (auto &w : obs_wires) {
                idx += sub_samples[i * num_qubits + w] << (obs_wires.size() - 1 - wire_idx);
                wire_idx++;
            }
            obs_samples[i] = eigenVals[idx];
        }
        return obs_samples;
    }
    auto var(const ObservableT &obs, const std::size_t &num_shots) -> PrecisionT {
        PrecisionT result{0.0};
        const auto name = obs.getObsName();
        if (name.find("SparseHamiltonian") != std::string::npos) {
            PLB200_ABORT("SparseHamiltonian observables do not support shot measurement.");
        } else if (name.find("Hamiltonian") != std::string::npos) {
            const auto coeffs = obs.getCoeffs();
            const auto terms = obs.getObs();
            for (std::size_t k = 0; k < coeffs.size(); k++) result += coeffs[k] * coeffs[k] * var(*terms[k], num_shots);
        } else {
            const auto s = measure_with_samples(obs, num_shots, {});
            const auto square_mean = std::accumulate(s.begin(), s.end(), 0.0) / s.size();
            const auto mean_square =
                std::accumulate(s.begin(), s.end(), 0.0, [](PrecisionT acc, PrecisionT e) { return acc + e * e; }) /
                s.size();
            result = mean_square - square_mean * square_mean;
        }
        return result;
    }
    auto probs(const ObservableT &obs, std::size_t num_shots) -> std::vector<PrecisionT> {
        PLB200_ABORT_IF(obs.getObsName().find("Hamiltonian") != std::string::npos,
                        "Hamiltonian and Sparse Hamiltonian do not support samples().");
        std::vector<std::size_t> obs_wires;
        std::vector<std::vector<PrecisionT>> eigenvalues;
        StateVectorT sv(sv_);
        obs.applyInPlaceShots(sv, eigenvalues, obs_wires);
        Measurements measure(sv);
        measure.setSeed(seed_);
        if (num_shots > 0) return measure.probs(obs_wires, num_shots);
        return measure.probs(obs_wires);
    }
    auto probs(const std::vector<std::size_t> &wires, std::size_t num_shots) -> std::vector<PrecisionT> {
        const auto counts_map = counts(num_shots);
        const std::size_t num_wires = sv_.getTotalNumQubits();
        std::vector<PrecisionT> prob_shots(std::size_t{1} << wires.size(), 0.0);
        for (const auto &it : counts_map) {
            std::size_t bitVal = 0;
            for (std::size_t bit = 0; bit < wires.size(); bit++)
                bitVal += ((it.first >> (num_wires - 1 - wires[bit])) & std::size_t{1}) << (wires.size() - 1 - bit);
            prob_shots[bitVal] += it.second / static_cast<PrecisionT>(num_shots);
        }
        return prob_shots;
    }
    auto probs(std::size_t num_shots) -> std::vector<PrecisionT> {
        const auto counts_map = counts(num_shots);
        std::vector<PrecisionT> prob_shots(std::size_t{1} << sv_.getTotalNumQubits(), 0.0);
        for (const auto &it : counts_map) prob_shots[it.first] = it.second / static_cast<PrecisionT>(num_shots);
        return prob_shots;
    }
    auto sample(const ObservableT &obs, const std::size_t &num_shots) -> std::vector<PrecisionT> {
        PLB200_ABORT_IF(obs.getObsName().find("Hamiltonian") != std::string::npos,
                        "Hamiltonian and Sparse Hamiltonian do not support samples().");
        return measure_with_samples(obs, num_shots, {});
    }
    auto sample(const std::size_t &num_shots) -> std::vector<std::size_t> {
        Measurements measure(sv_);
        measure.setSeed(seed_);
        return measure.generate_samples(num_shots);
    }
    auto counts(const ObservableT &obs, const std::size_t &num_shots) -> std::unordered_map<PrecisionT, std::size_t> {
        std::unordered_map<PrecisionT, std::size_t> outcome_map;
        const auto s = sample(obs, num_shots);
        for (std::size_t i = 0; i < num_shots; i++) outcome_map[s[i]] += 1;
        return outcome_map;
    }
    auto counts(const std::size_t &num_shots) -> std::unordered_map<std::size_t, std::size_t> {
        std::unordered_map<std::size_t, std::size_t> outcome_map;
        const auto s = sample(num_shots);
        const std::size_t num_wires = sv_.getTotalNumQubits();
        for (std::size_t i = 0; i < num_shots; i++) {
            std::size_t key = 0;
            for (std::size_t j = 0; j < num_wires; j++) key += s[i * num_wires + j] << (num_wires - 1 - j);
            outcome_map[key] += 1;
        }
        return outcome_map;
    }

  private:
    auto _sample_state(const ObservableT &obs, const std::size_t &num_shots, const std::vector<std::size_t> &shot_range,
                       std::vector<std::size_t> &obs_wires, std::vector<std::vector<PrecisionT>> &eigenValues)
        -> std::vector<std::size_t> {
        const std::size_t num_qubits = sv_.getTotalNumQubits();
        StateVectorT sv(sv_); // device-to-device copy
        obs.applyInPlaceShots(sv, eigenValues, obs_wires);
        Measurements measure(sv);
        measure.setSeed(seed_);
        auto samples = measure.generate_samples(num_shots);
        if (!shot_range.empty()) {
            std::vector<std::size_t> sub(shot_range.size() * num_qubits);
            std::size_t shot_idx = 0;
            for (const auto &i : shot_range) {
                for (std::size_t j = i * num_qubits; j < (i + 1) * num_qubits; j++)
                    sub[shot_idx * num_qubits + j - i * num_qubits] = samples[j];
                shot_idx++;
            }
            return sub;
        }
        return samples;
    }
    int64_t seed_arg() const { return seed_.has_value() ? static_cast<int64_t>(seed_.value()) : int64_t{-1}; }
    StateVectorT &sv_;
    std::optional<std::size_t> seed_{std::nullopt};
};

} // namespace Pennylane::LightningB200::Measures
