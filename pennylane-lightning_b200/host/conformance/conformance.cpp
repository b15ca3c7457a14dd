// Conformance of StateVectorB200 against the REFERENCE's own C++ template contract.
//
// This TU includes the reference's headers from /root/reference UNMODIFIED —
//   core/observables/Observables.hpp        (Observable, NamedObsBase, HermitianObsBase, TensorProdObsBase, HamiltonianBase)
//   core/measurements/MeasurementsBase.hpp  (CRTP base incl. the whole shot-based API, :159-521)
//   core/algorithms/JacobianData.hpp        (OpsData, JacobianData)
//   core/algorithms/AdjointJacobianBase.hpp (CRTP base with applyOperations / applyObservables helpers)
// — and instantiates every one of them with StateVectorB200<float|double>.  The backend-side finals a
// maintainer would add next to them (the analogue of lightning_gpu/observables/ObservablesGPU.hpp,
// measurements/MeasurementsGPU.hpp and algorithms/AdjointJacobianGPU.hpp) are the ~150 lines below: they
// only forward to the C ABI.  main() then runs sections transcribed from the reference's own typed suites
// (core/algorithms/tests/Test_AdjointJacobian.cpp:75-260, core/measurements/tests/Test_MeasurementsBase.cpp)
// through the REFERENCE base-class code paths on a B200, Catch2-free, printing one line per check.
//
// Built only where /root/reference exists (make -C host conformance); the binary travels to the GPU box.
#include <array>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <optional>
#include <random>
#include <unordered_map>

#include "StateVectorB200.hpp"
// reference headers
#include "AdjointJacobianBase.hpp"
#include "JacobianData.hpp"
#include "MeasurementsBase.hpp"
#include "Observables.hpp"

namespace B200 = Pennylane::LightningB200;
namespace Conf {
using Pennylane::Algorithms::JacobianData;
using Pennylane::Algorithms::OpsData;

// an object that owns the engine-side twin of a reference observable
struct EngineObs {
    plb200_obs *h = nullptr;
    virtual ~EngineObs() {
        if (h) plb200_obs_destroy(h);
    }
};

template <class SV> class NamedObs final : public Pennylane::Observables::NamedObsBase<SV>, public EngineObs {
    using Base = Pennylane::Observables::NamedObsBase<SV>;
    using PrecisionT = typename SV::PrecisionT;

  public:
    NamedObs(std::string name, std::vector<std::size_t> wires, std::vector<PrecisionT> params = {})
        : Base{name, wires, params} {
        const auto w = B200::detail::to_i64(wires);
        const auto p = B200::detail::to_f64(params);
        PLB200_ABI(plb200_obs_named(&h, name.c_str(), w.data(), static_cast<int64_t>(w.size()), p.data(),
                                    static_cast<int64_t>(p.size())));
    }
};
template <class SV> class HermitianObs final : public Pennylane::Observables::HermitianObsBase<SV>, public EngineObs {
    using Base = Pennylane::Observables::HermitianObsBase<SV>;

  public:
    using MatrixT = typename Base::MatrixT;
    HermitianObs(MatrixT matrix, std::vector<std::size_t> wires) : Base{matrix, wires} {
        const auto w = B200::detail::to_i64(wires);
        const auto m = B200::detail::to_c128(matrix.data(), matrix.size());
        PLB200_ABI(plb200_obs_hermitian(&h, m.data(), w.data(), static_cast<int64_t>(w.size())));
    }
};
template <class SV> class TensorProdObs final : public Pennylane::Observables::TensorProdObsBase<SV>, public EngineObs {
    using Base = Pennylane::Observables::TensorProdObsBase<SV>;

  public:
    template <typename... Ts> explicit TensorProdObs(Ts &&...arg) : Base{arg...} {
        std::vector<const plb200_obs *> hs;
        for (const auto &o : this->obs_) hs.push_back(dynamic_cast<const EngineObs &>(*o).h);
        PLB200_ABI(plb200_obs_tensor(&h, hs.data(), static_cast<int64_t>(hs.size())));
    }
};
template <class SV> class Hamiltonian final : public Pennylane::Observables::HamiltonianBase<SV>, public EngineObs {
    using Base = Pennylane::Observables::HamiltonianBase<SV>;
    using PrecisionT = typename SV::PrecisionT;

  public:
    template <typename T1, typename T2> Hamiltonian(T1 &&coeffs, T2 &&obs) : Base{std::forward<T1>(coeffs), std::forward<T2>(obs)} {
        std::vector<const plb200_obs *> hs;
        for (const auto &o : this->obs_) hs.push_back(dynamic_cast<const EngineObs &>(*o).h);
        const auto c = B200::detail::to_f64(this->coeffs_);
        PLB200_ABI(plb200_obs_hamiltonian(&h, c.data(), hs.data(), static_cast<int64_t>(hs.size())));
    }
    // sum_k c_k O_k |psi> in one out-of-place engine pass (reference: ObservablesLQubit.hpp:156-251)
    void applyInPlace(SV &sv) const override { PLB200_ABI(plb200_obs_apply(h, sv.handle())); }
};

// Derived of the reference's MeasurementsBase: the five primitives the base forwards to; everything
// shot-based (expval/var/probs/sample/counts with num_shots) is INHERITED from the reference header.
template <class SV> class Measurements final : public Pennylane::Measures::MeasurementsBase<SV, Measurements<SV>> {
    using Base = Pennylane::Measures::MeasurementsBase<SV, Measurements<SV>>;
    using PrecisionT = typename SV::PrecisionT;

  public:
    explicit Measurements(SV &sv) : Base{sv} {}
    using Base::expval;
    using Base::probs;
    using Base::var;
    auto probs() -> std::vector<PrecisionT> {
        std::vector<double> p(this->_statevector.getLength());
        PLB200_ABI(plb200_probs(this->_statevector.handle(), nullptr, -1, p.data()));
        return {p.begin(), p.end()};
    }
    auto probs(const std::vector<std::size_t> &wires) -> std::vector<PrecisionT> {
        const auto w = B200::detail::to_i64(wires);
        std::vector<double> p(std::size_t{1} << wires.size());
        PLB200_ABI(plb200_probs(this->_statevector.handle(), w.data(), static_cast<int64_t>(w.size()), p.data()));
        return {p.begin(), p.end()};
    }
    auto expval(const Pennylane::Observables::Observable<SV> &obs) -> PrecisionT {
        double r = 0;
        PLB200_ABI(plb200_expval_obs(this->_statevector.handle(), dynamic_cast<const EngineObs &>(obs).h, &r));
        return static_cast<PrecisionT>(r);
    }
    auto var(const Pennylane::Observables::Observable<SV> &obs) -> PrecisionT {
        double r = 0;
        PLB200_ABI(plb200_var_obs(this->_statevector.handle(), dynamic_cast<const EngineObs &>(obs).h, &r));
        return static_cast<PrecisionT>(r);
    }
    auto generate_samples(std::size_t num_samples) -> std::vector<std::size_t> {
        const std::size_t nq = this->_statevector.getNumQubits();
        std::vector<uint64_t> out(num_samples * nq);
        const int64_t seed = this->_deviceseed.has_value() ? static_cast<int64_t>(*this->_deviceseed) : -1;
        PLB200_ABI(plb200_generate_samples(this->_statevector.handle(), nullptr, -1, static_cast<int64_t>(num_samples), seed,
                                           out.data()));
        return {out.begin(), out.end()};
    }
};

// Derived of the reference's AdjointJacobianBase: one engine call for the whole sweep.
template <class SV> class AdjointJacobian final : public Pennylane::Algorithms::AdjointJacobianBase<SV, AdjointJacobian<SV>> {
    using PrecisionT = typename SV::PrecisionT;

  public:
    void adjointJacobian(std::span<PrecisionT> jac, const JacobianData<SV> &jd, const SV &ref_data, bool apply_operations = false) {
        const auto &obs = jd.getObservables();
        const auto &tp = jd.getTrainableParams();
        if (!jd.hasTrainableParams()) return;
        PL_ABORT_IF_NOT(jac.size() == tp.size() * obs.size(),
                        "The size of preallocated jacobian must be same as the number of trainable parameters times the "
                        "number of observables provided.");
        const OpsData<SV> &ops = jd.getOperations();
        B200::detail::OpsBlob blob;
        for (std::size_t i = 0; i < ops.getSize(); i++)
            blob.add<PrecisionT>(ops.getOpsName()[i], ops.getOpsWires()[i], ops.getOpsInverses()[i], ops.getOpsParams()[i],
                                 ops.getOpsControlledWires()[i], ops.getOpsControlledValues()[i], ops.getOpsMatrices()[i]);
        const auto v = blob.view();
        std::vector<const plb200_obs *> hs;
        for (const auto &o : obs) hs.push_back(dynamic_cast<const EngineObs &>(*o).h);
        const auto t = B200::detail::to_i64(tp);
        std::vector<double> out(jac.size());
        PLB200_ABI(plb200_adjoint_jacobian(ref_data.handle(), hs.data(), static_cast<int64_t>(hs.size()), &v, t.data(),
                                           static_cast<int64_t>(t.size()), apply_operations, out.data()));
        for (std::size_t i = 0; i < out.size(); i++) jac[i] = static_cast<PrecisionT>(out[i]);
    }
};
} // namespace Conf

// explicit instantiation of the reference templates over both precisions: the compiled fact
template class Pennylane::Observables::NamedObsBase<B200::StateVectorB200<float>>;
template class Pennylane::Observables::NamedObsBase<B200::StateVectorB200<double>>;
template class Pennylane::Observables::HermitianObsBase<B200::StateVectorB200<float>>;
template class Pennylane::Observables::HermitianObsBase<B200::StateVectorB200<double>>;
template class Pennylane::Observables::TensorProdObsBase<B200::StateVectorB200<float>>;
template class Pennylane::Observables::TensorProdObsBase<B200::StateVectorB200<double>>;
template class Pennylane::Observables::HamiltonianBase<B200::StateVectorB200<float>>;
template class Pennylane::Observables::HamiltonianBase<B200::StateVectorB200<double>>;
template class Pennylane::Algorithms::OpsData<B200::StateVectorB200<float>>;
template class Pennylane::Algorithms::OpsData<B200::StateVectorB200<double>>;
template class Pennylane::Algorithms::JacobianData<B200::StateVectorB200<float>>;
template class Pennylane::Algorithms::JacobianData<B200::StateVectorB200<double>>;
template class Conf::Measurements<B200::StateVectorB200<float>>;
template class Conf::Measurements<B200::StateVectorB200<double>>;
template class Conf::AdjointJacobian<B200::StateVectorB200<float>>;
template class Conf::AdjointJacobian<B200::StateVectorB200<double>>;

// ------------------------------------------------------------------------------------------------
namespace {
int g_fail = 0, g_pass = 0;
void check(bool ok, const std::string &what) {
    std::printf("%s %s\n", ok ? "PASS" : "FAIL", what.c_str());
    (ok ? g_pass : g_fail)++;
}
bool approx(double a, double b, double tol) { return std::abs(a - b) <= tol * std::max(1.0, std::abs(b)); }

// createNonTrivialState of core/utils/TestHelpers.hpp:436-466: RX(p) RY(p) on wire k, p = 0.7 - 0.2 k
template <class SV> SV nontrivial_state(std::size_t n) {
    using P = typename SV::PrecisionT;
    SV sv(n);
    std::vector<std::string> gates;
    std::vector<std::vector<std::size_t>> wires;
    std::vector<std::vector<P>> phase;
    P p = 0.7;
    for (std::size_t q = 0; q < n; q++) {
        gates.emplace_back("RX"), gates.emplace_back("RY");
        wires.push_back({q}), wires.push_back({q});
        phase.push_back({p}), phase.push_back({p});
        p -= static_cast<P>(0.2);
    }
    sv.applyOperations(gates, wires, std::vector<bool>(2 * n, false), phase);
    return sv;
}

template <class SV> void run(const char *tag) {
    using P = typename SV::PrecisionT;
    using C = typename SV::ComplexT;
    using namespace Conf;
    const double tol = sizeof(P) == 8 ? 1e-12 : 1e-5;
    const std::string T = std::string(" [") + tag + "]";
    const std::vector<P> param{static_cast<P>(-M_PI / 7), static_cast<P>(M_PI / 5), static_cast<P>(2 * M_PI / 3)};
    AdjointJacobian<SV> adj;

    { // Test_AdjointJacobian.cpp:85-108 "Throws an exception when size mismatches"
        const std::vector<std::size_t> tp{0, 1};
        const auto obs = std::make_shared<NamedObs<SV>>("PauliZ", std::vector<std::size_t>{0});
        std::vector<P> jac(1 * tp.size() - 1, 0);
        auto ops = OpsData<SV>({"RX"}, {{static_cast<P>(0.742)}}, {{0}}, {false});
        std::vector<C> cdata(2);
        cdata[0] = C{1, 0};
        SV psi(cdata.data(), cdata.size());
        JacobianData<SV> tape{3, psi.getLength(), psi.getData(), {obs}, ops, tp};
        bool thrown = false;
        try {
            adj.adjointJacobian(std::span{jac}, tape, psi, true);
        } catch (const Pennylane::Util::LightningException &e) {
            thrown = std::string(e.what()).find("The size of preallocated jacobian must be same as") != std::string::npos;
        }
        check(thrown, "adjoint: size mismatch throws LightningException with the reference text" + T);
    }
    { // :131-165 "Op=PhaseShift, Obs=Y" with N-1 controls
        bool ok = true;
        for (std::size_t nq : {2, 3, 4})
            for (const auto &p : param) {
                std::vector<std::vector<std::size_t>> controls{std::vector<std::size_t>(nq - 1)};
                std::iota(controls[0].begin(), controls[0].end(), 0);
                std::vector<std::vector<bool>> cvals{std::vector<bool>(nq - 1, true)};
                const auto obs = std::make_shared<NamedObs<SV>>("PauliY", std::vector<std::size_t>{nq - 1});
                auto ops = OpsData<SV>({"PhaseShift"}, {{p}}, {{nq - 1}}, {false}, {{}}, controls, cvals);
                std::vector<C> cdata(std::size_t{1} << nq);
                cdata[cdata.size() - 2] = cdata[cdata.size() - 1] = C{static_cast<P>(M_SQRT1_2), 0};
                SV psi(cdata.data(), cdata.size());
                std::vector<P> jac(1, 0);
                JacobianData<SV> tape{3, psi.getLength(), psi.getData(), {obs}, ops, {0}};
                adj.adjointJacobian(std::span{jac}, tape, psi, true);
                ok = ok && approx(jac[0], std::cos(p), 10 * tol);
            }
        check(ok, "adjoint: controlled PhaseShift / PauliY = cos(p)" + T);
    }
    { // :167-219 RX/Z = -sin p, RY/X = cos p
        bool ok = true;
        for (const auto &p : param) {
            std::vector<C> cdata(2);
            cdata[0] = C{1, 0};
            SV psi(cdata.data(), cdata.size());
            std::vector<P> jac(1, 0);
            const auto z = std::make_shared<NamedObs<SV>>("PauliZ", std::vector<std::size_t>{0});
            JacobianData<SV> t1{3, psi.getLength(), psi.getData(), {z}, OpsData<SV>({"RX"}, {{p}}, {{0}}, {false}), {0}};
            adj.adjointJacobian(std::span{jac}, t1, psi, true);
            ok = ok && approx(jac[0], -std::sin(p), 10 * tol);
            const auto x = std::make_shared<NamedObs<SV>>("PauliX", std::vector<std::size_t>{0});
            JacobianData<SV> t2{3, psi.getLength(), psi.getData(), {x}, OpsData<SV>({"RY"}, {{p}}, {{0}}, {false}), {0}};
            adj.adjointJacobian(std::span{jac}, t2, psi, true);
            ok = ok && approx(jac[0], std::cos(p), 10 * tol);
        }
        check(ok, "adjoint: RX/Z = -sin(p), RY/X = cos(p)" + T);
    }
    { // :221-260 RX, Obs = [Z0, Z1] and a Hamiltonian / TensorProd through the reference classes
        std::vector<C> cdata(4);
        cdata[0] = C{1, 0};
        SV psi(cdata.data(), cdata.size());
        const auto z0 = std::make_shared<NamedObs<SV>>("PauliZ", std::vector<std::size_t>{0});
        const auto z1 = std::make_shared<NamedObs<SV>>("PauliZ", std::vector<std::size_t>{1});
        std::vector<P> jac(2, 0);
        JacobianData<SV> tape{1, psi.getLength(), psi.getData(), {z0, z1}, OpsData<SV>({"RX"}, {{param[0]}}, {{0}}, {false}), {0}};
        adj.adjointJacobian(std::span{jac}, tape, psi, true);
        check(approx(jac[0], -std::sin(param[0]), 10 * tol) && std::abs(jac[1]) < 10 * tol, "adjoint: RX with observables [Z0, Z1]" + T);
        // H = 0.3 Z0 (x) Z1 + 0.7 Z0: d/dp = -(0.3 + 0.7) sin p on |00>
        auto tp = std::make_shared<TensorProdObs<SV>>(z0, z1);
        using ObsPtr = std::shared_ptr<Pennylane::Observables::Observable<SV>>;
        auto ham = std::make_shared<Hamiltonian<SV>>(std::vector<P>{static_cast<P>(0.3), static_cast<P>(0.7)},
                                                      std::vector<ObsPtr>{tp, z0});
        std::vector<P> j1(1, 0);
        JacobianData<SV> t2{1, psi.getLength(), psi.getData(), {ham}, OpsData<SV>({"RX"}, {{param[0]}}, {{0}}, {false}), {0}};
        adj.adjointJacobian(std::span{j1}, t2, psi, true);
        check(approx(j1[0], -std::sin(param[0]), 10 * tol), "adjoint: Hamiltonian(TensorProd, Named) built from the reference classes" + T);
    }
    { // Test_MeasurementsBase.cpp: probabilities / expval / var of the 3-qubit createNonTrivialState
        SV sv = nontrivial_state<SV>(3);
        Measurements<SV> m(sv);
        // reference literals: Test_MeasurementsBase.cpp:94-118 (probs), :466-470 (expval of PauliX/Y/Z on wire 0..2)
        const std::vector<double> probs_ref{0.67078706, 0.03062806, 0.0870997, 0.00397696, 0.17564072, 0.00801973, 0.02280642, 0.00104134};
        const auto p = m.probs();
        bool ok = p.size() == 8;
        for (std::size_t i = 0; ok && i < 8; i++) ok = std::abs(p[i] - probs_ref[i]) < 1e-6;
        check(ok, "measurements: probs() of createNonTrivialState(3) = reference literals" + T);
        NamedObs<SV> x0("PauliX", {0}), y1("PauliY", {1}), z2("PauliZ", {2});
        // reference literals (default.qubit): Test_MeasurementsBase.cpp:466-470 (expval), :890-894 (var)
        const double X0 = 0.49272486, Y1 = -0.47942553, Z2 = 0.9126678, VZ2 = 0.1670374;
        auto bloch = [&](double a) { return std::array<double, 3>{a == 0.7 ? X0 : 0, a == 0.5 ? Y1 : 0, a == 0.3 ? Z2 : 0}; };
        check(std::abs(m.expval(x0) - X0) < 1e-6 && std::abs(m.expval(y1) - Y1) < 1e-6 && std::abs(m.expval(z2) - Z2) < 1e-6,
              "measurements: expval(NamedObs) = reference literals" + T);
        check(std::abs(m.var(z2) - VZ2) < 1e-6, "measurements: var(NamedObs) = reference literal" + T);
        // INHERITED shot-based estimators (MeasurementsBase.hpp:159-260): basis rotation via the reference's
        // NamedObsBase::applyInPlaceShots on a COPY of our state, then our sampler
        m.setSeed(1337);
        const std::size_t shots = 200000;
        const double ex = m.expval(x0, shots, {}), ey = m.expval(y1, shots, {}), ez = m.expval(z2, shots, {});
        check(std::abs(ex - bloch(0.7)[0]) < 0.01 && std::abs(ey - bloch(0.5)[1]) < 0.01 && std::abs(ez - bloch(0.3)[2]) < 0.01,
              "measurements: shot-based expval inherited from the reference base (X, Y, Z rotations)" + T);
        const double vz = m.var(z2, shots);
        check(std::abs(vz - VZ2) < 0.01, "measurements: shot-based var inherited from the reference base" + T);
        auto pz = m.probs({0, 1}, shots);
        bool okp = pz.size() == 4;
        const auto pex = m.probs({0, 1});
        for (std::size_t i = 0; okp && i < 4; i++) okp = std::abs(pz[i] - pex[i]) < 0.01;
        check(okp, "measurements: shot-based probs(wires, shots) inherited from the reference base" + T);
        TensorProdObs<SV> xz(std::make_shared<NamedObs<SV>>("PauliX", std::vector<std::size_t>{0}),
                             std::make_shared<NamedObs<SV>>("PauliZ", std::vector<std::size_t>{2}));
        check(std::abs(m.expval(xz, shots, {}) - bloch(0.7)[0] * bloch(0.3)[2]) < 0.01,
              "measurements: shot-based expval of a TensorProdObs (reference TensorProdObsBase::applyInPlaceShots)" + T);
        auto cnt = m.counts(z2, 1000);
        std::size_t total = 0;
        for (auto &kv : cnt) total += kv.second;
        check(total == 1000 && cnt.size() <= 2, "measurements: counts(obs, shots) inherited from the reference base" + T);
    }
}
} // namespace

int main() {
    int n = 0;
    if (plb200_device_count(&n) != 0 || n == 0) {
        std::printf("SKIP no CUDA device (templates instantiated at build time; run on a B200 for the sections)\n");
        return 0;
    }
    try {
        run<B200::StateVectorB200<double>>("double");
        run<B200::StateVectorB200<float>>("float");
    } catch (const std::exception &e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf("conformance: %d passed, %d failed\n", g_pass, g_fail);
    return g_fail ? 1 : 0;
}
