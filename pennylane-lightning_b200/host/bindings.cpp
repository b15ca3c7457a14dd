// pybind11 module `lightning_b200_ops`: the Python-visible surface of lightning_gpu_ops
// (core/bindings/Bindings.hpp:933-1010 + lightning_gpu/bindings/LGPUBindings.hpp) on top of the
// B200 host classes, so pennylane_lightning/lightning_gpu/*.py and lightning_base/*.py can drive
// the engine unchanged.  The reference binds with nanobind (fetched at configure time,
// CMakeLists.txt:101-105); nanobind is not available offline, pybind11 is — names, argument order
// and return types are kept identical.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <set>

#include "AdjointJacobianB200.hpp"
#include "MeasurementsB200.hpp"
#include "ObservablesB200.hpp"
#include "StateVectorB200.hpp"

namespace py = pybind11;
using namespace Pennylane::LightningB200;
using namespace Pennylane::LightningB200::Observables;
using namespace Pennylane::LightningB200::Measures;
using namespace Pennylane::LightningB200::Algorithms;

namespace {

const char *const kGateNames[] = {
    "Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T", "PhaseShift", "RX", "RY", "RZ", "Rot",
    "CNOT", "CY", "CZ", "SWAP", "IsingXX", "IsingXY", "IsingYY", "IsingZZ", "ControlledPhaseShift", "CRX", "CRY",
    "CRZ", "CRot", "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus", "PSWAP", "Toffoli", "CSWAP",
    "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus", "MultiRZ", "GlobalPhase", "PCPhase"};
const std::set<std::string> kControlledGateNames = {
    "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T", "PhaseShift", "RX", "RY", "RZ", "Rot", "SWAP", "IsingXX",
    "IsingXY", "IsingYY", "IsingZZ", "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus",
    "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus", "PSWAP", "MultiRZ", "GlobalPhase", "PCPhase"};

template <class PrecisionT>
using np_arr_c = py::array_t<std::complex<PrecisionT>, py::array::c_style | py::array::forcecast>;

template <class PrecisionT> void registerPrecision(py::module_ &m, const std::string &bits) {
    using SV = StateVectorB200<PrecisionT>;
    using ComplexT = std::complex<PrecisionT>;
    using ObsT = Observable<SV>;
    using ObsPtr = std::shared_ptr<ObsT>;

    // ------------------------------------------------------------------ StateVectorC{64,128}
    auto sv = py::class_<SV>(m, ("StateVectorC" + bits).c_str());
    sv.def(py::init<std::size_t>())
        .def(py::init<std::size_t, const DevTag<int> &>())
        .def(py::init([](const np_arr_c<PrecisionT> &arr) {
            return new SV(arr.data(), static_cast<std::size_t>(arr.size()));
        }))
        .def("__len__", &SV::getLength)
        .def("size", &SV::getLength)
        .def("dataLength", &SV::getLength)
        .def("numQubits", &SV::getNumQubits)
        .def("getCurrentGPU", [](const SV &s) { return s.getDevTag().getDeviceID(); })
        .def("GetNumGPUs", [](const SV &) { return DevicePool<int>::getTotalDevices(); })
        .def("kernelLaunches", &SV::kernelLaunches)
        .def("pendingOps", &SV::pendingOps, "gates queued by the per-gate calls and not yet applied")
        .def("resetStateVector", [](SV &s, bool async) { s.resetStateVector(async); }, py::arg("async") = false)
        .def("setBasisState",
             [](SV &s, const std::vector<std::size_t> &state, const std::vector<std::size_t> &wires, bool async) {
                 s.setBasisState(state, wires, async);
             },
             py::arg("state"), py::arg("wires"), py::arg("async") = false)
        .def("setStateVector",
             [](SV &s, const np_arr_c<PrecisionT> &state, const std::vector<std::size_t> &wires, bool async) {
                 s.setStateVector(state.data(), static_cast<std::size_t>(state.size()), wires, async);
             },
             py::arg("state"), py::arg("wires"), py::arg("async") = false)
        .def("updateData", [](SV &s, const np_arr_c<PrecisionT> &state) {
            s.updateData(state.data(), static_cast<std::size_t>(state.size()));
        })
        .def("collapse", &SV::collapse)
        .def("normalize", &SV::normalize)
        .def("DeviceToHost",
             [](const SV &s, py::array_t<ComplexT, py::array::c_style> &out, bool async) {
                 PLB200_ABORT_IF_NOT(static_cast<std::size_t>(out.size()) == s.getLength(),
                                     "Sizes do not match for Host and GPU data");
                 s.CopyGpuDataToHost(out.mutable_data(), s.getLength(), async);
             },
             py::arg("host_sv"), py::arg("async") = false)
        .def("getState",
             [](const SV &s, py::array_t<ComplexT, py::array::c_style> &out) {
                 PLB200_ABORT_IF_NOT(static_cast<std::size_t>(out.size()) == s.getLength(),
                                     "Sizes do not match for Host and GPU data");
                 s.CopyGpuDataToHost(out.mutable_data(), s.getLength(), false);
             })
        .def("HostToDevice",
             [](SV &s, const np_arr_c<PrecisionT> &in, bool async) {
                 s.CopyHostDataToGpu(in.data(), static_cast<std::size_t>(in.size()), async);
             },
             py::arg("host_sv"), py::arg("async") = false)
        .def("DeviceToDevice", [](SV &s, const SV &other, bool async) { s.CopyGpuDataToGpuIn(other, async); },
             py::arg("other"), py::arg("async") = false)
        .def("applyMatrix",
             [](SV &s, const np_arr_c<PrecisionT> &matrix, const std::vector<std::size_t> &wires, bool inverse) {
                 PLB200_ABORT_IF(static_cast<std::size_t>(matrix.size()) != (std::size_t{1} << (2 * wires.size())),
                                 "The size of matrix does not match with the given number of wires");
                 s.applyMatrix(matrix.data(), wires, inverse);
             },
             py::arg("matrix"), py::arg("wires"), py::arg("inverse") = false)
        .def("applyControlledMatrix",
             [](SV &s, const np_arr_c<PrecisionT> &matrix, const std::vector<std::size_t> &cw,
                const std::vector<bool> &cv, const std::vector<std::size_t> &wires, bool inverse) {
                 PLB200_ABORT_IF(static_cast<std::size_t>(matrix.size()) != (std::size_t{1} << (2 * wires.size())),
                                 "The size of matrix does not match with the given number of wires");
                 s.applyControlledMatrix(matrix.data(), cw, cv, wires, inverse);
             },
             py::arg("matrix"), py::arg("controlled_wires"), py::arg("controlled_values"), py::arg("wires"),
             py::arg("inverse") = false)
        .def("applyPauliRot", &SV::applyPauliRot)
        // fallback `apply` overloads (LGPUBindings.hpp:354-391, caller lightning_gpu/_state_vector.py:384-390)
        .def("apply",
             [](SV &s, const std::string &name, const std::vector<std::size_t> &wires, bool inv,
                const std::vector<std::vector<PrecisionT>> &params, const np_arr_c<PrecisionT> &matrix) {
                 std::vector<PrecisionT> flat;
                 for (const auto &p : params) flat.insert(flat.end(), p.begin(), p.end());
                 std::vector<ComplexT> mat(matrix.data(), matrix.data() + matrix.size());
                 if (SV::isNativeGate(name)) s.applyOperation(name, wires, inv, flat);
                 else s.applyOperation(name, wires, inv, std::vector<PrecisionT>{}, mat);
             })
        .def("apply", [](SV &s, const std::string &name, const std::vector<std::size_t> &cw, const std::vector<bool> &cv,
                         const std::vector<std::size_t> &wires, bool inv, const std::vector<PrecisionT> &params) {
            s.applyOperation(name, cw, cv, wires, inv, params);
        });
    // one method per gate name, both overloads (Bindings.hpp:223-276)
    for (const char *g : kGateNames) {
        const std::string name(g);
        sv.def(g, [name](SV &s, const std::vector<std::size_t> &wires, bool inverse,
                         const std::vector<PrecisionT> &params) { s.applyOperation(name, wires, inverse, params); },
               py::arg("wires"), py::arg("inverse") = false, py::arg("params") = std::vector<PrecisionT>{});
        if (kControlledGateNames.count(name))
            sv.def(g,
                   [name](SV &s, const std::vector<std::size_t> &cw, const std::vector<bool> &cv,
                          const std::vector<std::size_t> &wires, bool inverse, const std::vector<PrecisionT> &params) {
                       s.applyOperation(name, cw, cv, wires, inverse, params);
                   },
                   py::arg("controlled_wires"), py::arg("controlled_values"), py::arg("wires"),
                   py::arg("inverse") = false, py::arg("params") = std::vector<PrecisionT>{});
    }

    // ------------------------------------------------------------------ observables
    py::module_ obs = m.def_submodule("observables", "Submodule for observables classes.");
    py::class_<ObsT, ObsPtr>(obs, ("ObservableC" + bits).c_str(), py::module_local());
    py::class_<NamedObs<SV>, std::shared_ptr<NamedObs<SV>>, ObsT>(obs, ("NamedObsC" + bits).c_str(), py::module_local())
        .def(py::init([](const std::string &name, const std::vector<std::size_t> &wires) {
            return std::make_shared<NamedObs<SV>>(name, wires);
        }))
        .def("__repr__", &NamedObs<SV>::getObsName)
        .def("get_wires", &NamedObs<SV>::getWires)
        .def("__eq__", [](const NamedObs<SV> &a, py::handle other) {
            return py::isinstance<NamedObs<SV>>(other) && a == other.cast<const NamedObs<SV> &>();
        });
    py::class_<HermitianObs<SV>, std::shared_ptr<HermitianObs<SV>>, ObsT>(obs, ("HermitianObsC" + bits).c_str(),
                                                                         py::module_local())
        .def(py::init([](const np_arr_c<PrecisionT> &matrix, const std::vector<std::size_t> &wires) {
            return std::make_shared<HermitianObs<SV>>(std::vector<ComplexT>(matrix.data(), matrix.data() + matrix.size()),
                                                      wires);
        }))
        .def("__repr__", &HermitianObs<SV>::getObsName)
        .def("get_wires", &HermitianObs<SV>::getWires)
        .def("get_matrix", [](const HermitianObs<SV> &o) { return o.getMatrix(); })
        .def("__eq__", [](const HermitianObs<SV> &a, py::handle other) {
            return py::isinstance<HermitianObs<SV>>(other) && a == other.cast<const HermitianObs<SV> &>();
        });
    py::class_<TensorProdObs<SV>, std::shared_ptr<TensorProdObs<SV>>, ObsT>(obs, ("TensorProdObsC" + bits).c_str(),
                                                                           py::module_local())
        .def(py::init([](const std::vector<ObsPtr> &o) { return std::make_shared<TensorProdObs<SV>>(o); }))
        .def("__repr__", &TensorProdObs<SV>::getObsName)
        .def("get_wires", &TensorProdObs<SV>::getWires)
        .def("get_ops", &TensorProdObs<SV>::getObs)
        .def("__eq__", [](const TensorProdObs<SV> &a, py::handle other) {
            return py::isinstance<TensorProdObs<SV>>(other) && a == other.cast<const TensorProdObs<SV> &>();
        });
    py::class_<Hamiltonian<SV>, std::shared_ptr<Hamiltonian<SV>>, ObsT>(obs, ("HamiltonianC" + bits).c_str(),
                                                                       py::module_local())
        .def(py::init([](const py::array_t<PrecisionT, py::array::c_style | py::array::forcecast> &coeffs,
                         const std::vector<ObsPtr> &o) {
            return std::make_shared<Hamiltonian<SV>>(std::vector<PrecisionT>(coeffs.data(), coeffs.data() + coeffs.size()),
                                                     o);
        }))
        .def("__repr__", &Hamiltonian<SV>::getObsName)
        .def("get_wires", &Hamiltonian<SV>::getWires)
        .def("get_ops", &Hamiltonian<SV>::getObs)
        .def("get_coeffs", &Hamiltonian<SV>::getCoeffs)
        .def("__eq__", [](const Hamiltonian<SV> &a, py::handle other) {
            return py::isinstance<Hamiltonian<SV>>(other) && a == other.cast<const Hamiltonian<SV> &>();
        });

    using SpH = SparseHamiltonian<SV>;
    using idx_arr = py::array_t<int64_t, py::array::c_style | py::array::forcecast>;
    py::class_<SpH, std::shared_ptr<SpH>, ObsT>(obs, ("SparseHamiltonianC" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_c<PrecisionT> &data, const idx_arr &indices, const idx_arr &offsets,
                         const std::vector<std::size_t> &wires) {
            return std::make_shared<SpH>(std::vector<ComplexT>(data.data(), data.data() + data.size()),
                                         std::vector<int64_t>(indices.data(), indices.data() + indices.size()),
                                         std::vector<int64_t>(offsets.data(), offsets.data() + offsets.size()), wires);
        }))
        .def("__repr__", &SpH::getObsName)
        .def("get_wires", &SpH::getWires)
        .def("__eq__", [](const SpH &a, py::handle other) {
            return py::isinstance<SpH>(other) && a == other.cast<const SpH &>();
        });

    // ------------------------------------------------------------------ MeasurementsC{64,128}
    using M = Measurements<SV>;
    py::class_<M>(m, ("MeasurementsC" + bits).c_str())
        .def(py::init<SV &>(), py::keep_alive<1, 2>())
        .def("set_random_seed", [](M &mm, std::size_t seed) { mm.setSeed(seed); })
        .def("probs", [](M &mm) { return py::array_t<PrecisionT>(py::cast(mm.probs())); })
        .def("probs", [](M &mm, const std::vector<std::size_t> &wires) {
            return py::array_t<PrecisionT>(py::cast(mm.probs(wires)));
        })
        .def("expval", [](M &mm, const ObsPtr &o) { return mm.expval(*o); })
        .def("var", [](M &mm, const ObsPtr &o) { return mm.var(*o); })
        .def("expval", [](M &mm, const std::string &op, const std::vector<std::size_t> &wires) {
            return mm.expval(op, wires);
        })
        .def("var", [](M &mm, const std::string &op, const std::vector<std::size_t> &wires) { return mm.var(op, wires); })
        .def("expval", [](M &mm, const np_arr_c<PrecisionT> &matrix, const std::vector<std::size_t> &wires) {
            return mm.expval(std::vector<ComplexT>(matrix.data(), matrix.data() + matrix.size()), wires);
        })
        .def("var", [](M &mm, const np_arr_c<PrecisionT> &matrix, const std::vector<std::size_t> &wires) {
            return mm.var(std::vector<ComplexT>(matrix.data(), matrix.data() + matrix.size()), wires);
        })
        .def("expval", [](M &mm, const std::vector<std::string> &words, const std::vector<std::vector<std::size_t>> &wires,
                          const py::array_t<PrecisionT, py::array::c_style | py::array::forcecast> &coeffs) {
            return mm.expval(words, wires, std::vector<PrecisionT>(coeffs.data(), coeffs.data() + coeffs.size()));
        })
        // shot-based C++ API of MeasurementsBase (not bound by the reference; exposed for the parity tests)
        .def("expval", [](M &mm, const idx_arr &indptr, const idx_arr &indices, const np_arr_c<PrecisionT> &data) {
            return mm.expval(indptr.data(), static_cast<std::size_t>(indptr.size()), indices.data(), data.data(),
                             static_cast<std::size_t>(data.size()));
        }, "Expected value of a sparse Hamiltonian (CSR).")
        .def("var", [](M &mm, const idx_arr &indptr, const idx_arr &indices, const np_arr_c<PrecisionT> &data) {
            return mm.var(indptr.data(), static_cast<std::size_t>(indptr.size()), indices.data(), data.data(),
                          static_cast<std::size_t>(data.size()));
        }, "Variance of a sparse Hamiltonian (CSR).")
        .def("expval_shots", [](M &mm, const ObsPtr &o, std::size_t shots, const std::vector<std::size_t> &range) {
            return mm.expval(*o, shots, range);
        }, py::arg("obs"), py::arg("num_shots"), py::arg("shot_range") = std::vector<std::size_t>{})
        .def("var_shots", [](M &mm, const ObsPtr &o, std::size_t shots) { return mm.var(*o, shots); })
        .def("probs_shots", [](M &mm, const std::vector<std::size_t> &wires, std::size_t shots) {
            return py::array_t<PrecisionT>(py::cast(mm.probs(wires, shots)));
        })
        .def("probs_shots", [](M &mm, const ObsPtr &o, std::size_t shots) {
            return py::array_t<PrecisionT>(py::cast(mm.probs(*o, shots)));
        })
        .def("sample_obs", [](M &mm, const ObsPtr &o, std::size_t shots) {
            return py::array_t<PrecisionT>(py::cast(mm.sample(*o, shots)));
        })
        .def("counts", [](M &mm, std::size_t shots) { return mm.counts(shots); })
        .def("generate_samples", [](M &mm, std::size_t num_wires, std::size_t num_shots) {
            auto s = mm.generate_samples(num_shots);
            py::array_t<std::size_t> out({num_shots, num_wires});
            std::copy(s.begin(), s.end(), out.mutable_data());
            return out;
        })
        .def("generate_samples", [](M &mm, const std::vector<std::size_t> &wires, std::size_t num_shots) {
            auto s = mm.generate_samples(wires, num_shots);
            py::array_t<std::size_t> out({num_shots, wires.size()});
            std::copy(s.begin(), s.end(), out.mutable_data());
            return out;
        });

    // ------------------------------------------------------------------ algorithms
    py::module_ alg = m.def_submodule("algorithms", "Submodule for the algorithms functionality.");
    using Ops = OpsData<SV>;
    auto make_ops = [](const std::vector<std::string> &names, const std::vector<std::vector<PrecisionT>> &params,
                       const std::vector<std::vector<std::size_t>> &wires, const std::vector<bool> &inverses,
                       const std::vector<np_arr_c<PrecisionT>> &mats, const std::vector<std::vector<std::size_t>> &cw,
                       const std::vector<std::vector<bool>> &cv) {
        std::vector<std::vector<ComplexT>> conv(mats.size());
        for (std::size_t i = 0; i < mats.size(); i++)
            conv[i] = std::vector<ComplexT>(mats[i].data(), mats[i].data() + mats[i].size());
        return Ops(names, params, wires, inverses, conv, cw, cv);
    };
    py::class_<Ops>(alg, ("OpsStructC" + bits).c_str(), py::module_local())
        .def(py::init(make_ops))
        .def("__repr__", [](const Ops &ops) {
            std::ostringstream s;
            s << "Operations: [";
            for (std::size_t i = 0; i < ops.getSize(); i++) s << (i ? ", " : "") << ops.getOpsName()[i];
            s << "]";
            return s.str();
        });
    alg.def(("create_ops_listC" + bits).c_str(), make_ops, "Create a list of operations from data.");
    using Adj = AdjointJacobian<SV>;
    auto call_adj = [](Adj &adj, const SV &svec, const std::vector<ObsPtr> &observables, const Ops &operations,
                       const std::vector<std::size_t> &trainableParams) {
        std::vector<PrecisionT> jac(observables.size() * trainableParams.size(), PrecisionT{0});
        const JacobianData<SV> jd{operations.getTotalNumParams(), svec.getLength(), svec.getData(), observables,
                                  operations, trainableParams};
        adj.adjointJacobian(std::span<PrecisionT>{jac}, jd, svec, false);
        return py::array_t<PrecisionT>(py::cast(jac));
    };
    using Vjp = VectorJacobianProduct<SV>;
    py::class_<Vjp>(alg, ("VectorJacobianProductC" + bits).c_str(), py::module_local())
        .def(py::init<>())
        .def("__call__", [](Vjp &v, const SV &svec, const Ops &operations, const np_arr_c<PrecisionT> &dy,
                            const std::vector<std::size_t> &trainableParams) {
            std::vector<ComplexT> out(trainableParams.size(), ComplexT{});
            const JacobianData<SV> jd{operations.getTotalNumParams(), svec.getLength(), svec.getData(), {}, operations,
                                      trainableParams};
            v(std::span<ComplexT>{out}, jd, std::span<const ComplexT>{dy.data(), static_cast<std::size_t>(dy.size())}, svec);
            return py::array_t<ComplexT>(py::cast(out));
        }, "Vector Jacobian Product method.");
    py::class_<Adj>(alg, ("AdjointJacobianC" + bits).c_str(), py::module_local())
        .def(py::init<>())
        .def("__call__", call_adj, "Adjoint Jacobian method.")
        .def("batched", [](Adj &adj, const SV &svec, const std::vector<ObsPtr> &observables, const Ops &operations,
                           const std::vector<std::size_t> &trainableParams) {
            std::vector<PrecisionT> jac(observables.size() * trainableParams.size(), PrecisionT{0});
            const JacobianData<SV> jd{operations.getTotalNumParams(), svec.getLength(), svec.getData(), observables,
                                      operations, trainableParams};
            adj.batchAdjointJacobian(std::span<PrecisionT>{jac}, jd, svec, false);
            return py::array_t<PrecisionT>(py::cast(jac));
        }, "Adjoint Jacobian with the observables split over the GPUs of the box.");
}

} // namespace

PYBIND11_MODULE(lightning_b200_ops, m) {
    m.doc() = "B200-native state-vector engine behind pennylane-lightning's binding surface";
    py::register_exception<Pennylane::LightningB200::Util::LightningException>(m, "LightningException",
                                                                               PyExc_RuntimeError);
    m.def("backend_info", []() {
        py::dict d;
        d["NAME"] = "lightning.b200";
        return d;
    });
    m.def("compile_info", []() {
        py::dict d;
        d["cpu.arch"] = "x86_64";
        d["compiler.name"] = "nvcc+g++";
        d["compiler.version"] = __VERSION__;
        d["cuda.arch"] = "sm_100a";
        d["engine"] = plb200_version();
        d["AVX2"] = false;
        d["AVX512F"] = false;
        return d;
    });
    m.def("runtime_info", []() {
        py::dict d;
        int n = 0;
        plb200_device_count(&n);
        d["cuda_devices"] = n;
        d["AVX"] = false, d["AVX2"] = false, d["AVX512F"] = false;
        return d;
    });
    m.def("is_gpu_supported", [](int device) {
        int arch = 0;
        return plb200_device_arch(device, &arch) == 0 && arch >= 100;
    }, py::arg("device_number") = 0);
    m.def("get_gpu_arch", [](int device) {
        int arch = 0;
        PLB200_ABI(plb200_device_arch(device, &arch));
        return std::make_pair(arch / 10, arch % 10);
    }, py::arg("device_number") = 0);
    py::class_<DevTag<int>>(m, "DevTag")
        .def(py::init<>())
        .def(py::init<int>())
        .def(py::init([](int device, std::uintptr_t stream) { return DevTag<int>(device, reinterpret_cast<void *>(stream)); }))
        .def("getDeviceID", &DevTag<int>::getDeviceID)
        .def("getStreamID", [](const DevTag<int> &t) { return reinterpret_cast<std::uintptr_t>(t.getStreamID()); })
        .def("refresh", &DevTag<int>::refresh);
    py::class_<DevicePool<int>>(m, "DevPool")
        .def(py::init<>())
        .def("getActiveDevices", [](DevicePool<int> &) { return 0; })
        .def("isActive", &DevicePool<int>::isActive)
        .def("isInactive", &DevicePool<int>::isInactive)
        .def("acquireDevice", &DevicePool<int>::acquireDevice)
        .def("releaseDevice", &DevicePool<int>::releaseDevice)
        .def("syncDevice", &DevicePool<int>::syncDevice)
        .def("refresh", &DevicePool<int>::refresh)
        .def_static("getTotalDevices", &DevicePool<int>::getTotalDevices)
        .def_static("getDeviceUIDs", []() { return std::vector<std::string>{}; })
        .def_static("setDeviceID", [](int) {});
    registerPrecision<float>(m, "64");
    registerPrecision<double>(m, "128");
}
