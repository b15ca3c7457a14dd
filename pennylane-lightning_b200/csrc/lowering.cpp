// Host-side lowering: named gate / matrix / generator -> canonical ops (common.hpp).
//
// Gate definitions follow the reference's dense matrices in core/gates/Gates.hpp:38-1385
// (getRX :329, getRY :346, getRZ :363, getRot :390, getCRZ :474, ... getPSWAP) and the
// kernels' semantics in GateImplementationsLM.hpp (file:line cited per gate below).  Nothing
// here touches the device; matrices are built in double and narrowed at launch for c64.
#include "common.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>

namespace plb200 {
namespace {

const cd I1{0.0, 1.0};

struct GateInfo {
    int wires;  // -1 = any (>=1)
    int params; //
};

const std::map<std::string, GateInfo> &gate_table() {
    // names/wires/params: core/gates/Constant.hpp:60-135,230-300,350-430
    static const std::map<std::string, GateInfo> t = {
        {"Identity", {1, 0}},
        {"PauliX", {1, 0}},
        {"PauliY", {1, 0}},
        {"PauliZ", {1, 0}},
        {"Hadamard", {1, 0}},
        {"S", {1, 0}},
        {"SX", {1, 0}},
        {"T", {1, 0}},
        {"PhaseShift", {1, 1}},
        {"RX", {1, 1}},
        {"RY", {1, 1}},
        {"RZ", {1, 1}},
        {"Rot", {1, 3}},
        {"CNOT", {2, 0}},
        {"CY", {2, 0}},
        {"CZ", {2, 0}},
        {"SWAP", {2, 0}},
        {"IsingXX", {2, 1}},
        {"IsingXY", {2, 1}},
        {"IsingYY", {2, 1}},
        {"IsingZZ", {2, 1}},
        {"ControlledPhaseShift", {2, 1}},
        {"CRX", {2, 1}},
        {"CRY", {2, 1}},
        {"CRZ", {2, 1}},
        {"CRot", {2, 3}},
        {"SingleExcitation", {2, 1}},
        {"SingleExcitationMinus", {2, 1}},
        {"SingleExcitationPlus", {2, 1}},
        {"PSWAP", {2, 1}},
        {"Toffoli", {3, 0}},
        {"CSWAP", {3, 0}},
        {"DoubleExcitation", {4, 1}},
        {"DoubleExcitationMinus", {4, 1}},
        {"DoubleExcitationPlus", {4, 1}},
        {"MultiRZ", {-1, 1}},
        {"GlobalPhase", {-1, 1}},
        {"PCPhase", {-1, 2}},
    };
    return t;
}

// gates that exist in ControlledGateOperation (core/gates/GateOperation.hpp:76-110)
bool controlled_gate_known(const std::string &n) {
    static const char *names[] = {"PauliX", "PauliY", "PauliZ", "Hadamard", "S", "SX", "T",
                                  "PhaseShift", "RX", "RY", "RZ", "Rot", "SWAP", "IsingXX",
                                  "IsingXY", "IsingYY", "IsingZZ", "SingleExcitation",
                                  "SingleExcitationMinus", "SingleExcitationPlus",
                                  "DoubleExcitation", "DoubleExcitationMinus",
                                  "DoubleExcitationPlus", "PSWAP", "MultiRZ", "GlobalPhase",
                                  "PCPhase"};
    for (auto *s : names)
        if (n == s) return true;
    return false;
}

std::vector<cd> eye(int dim) {
    std::vector<cd> m(static_cast<size_t>(dim) * dim, 0.0);
    for (int i = 0; i < dim; i++) m[static_cast<size_t>(i) * dim + i] = 1.0;
    return m;
}

std::vector<cd> dagger(const std::vector<cd> &m, int dim) {
    std::vector<cd> r(m.size());
    for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++)
            r[static_cast<size_t>(i) * dim + j] = std::conj(m[static_cast<size_t>(j) * dim + i]);
    return r;
}

// |1><1| (x) U on (ctrl = MSB, targets) -> 2*dim matrix
std::vector<cd> controlled(const std::vector<cd> &u, int dim) {
    int D = 2 * dim;
    auto m = eye(D);
    for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++)
            m[static_cast<size_t>(dim + i) * D + dim + j] = u[static_cast<size_t>(i) * dim + j];
    return m;
}

std::vector<cd> rot_matrix(double phi, double theta, double omega) {
    // Gates.hpp:390-410 getRot: RZ(omega) RY(theta) RZ(phi)
    const double c = std::cos(theta / 2), s = std::sin(theta / 2);
    return {std::exp(-I1 * ((phi + omega) / 2)) * c, -std::exp(I1 * ((phi - omega) / 2)) * s,
            std::exp(-I1 * ((phi - omega) / 2)) * s, std::exp(I1 * ((phi + omega) / 2)) * c};
}

} // namespace

bool gate_known(const std::string &name) { return gate_table().count(name) != 0; }
int gate_num_params(const std::string &name) {
    auto it = gate_table().find(name);
    return it == gate_table().end() ? -1 : it->second.params;
}

// Dense matrix of a fixed-size named gate (not MultiRZ/GlobalPhase/PCPhase with many wires,
// which are lowered directly).  `inverse` is folded in here.
std::vector<cd> named_gate_matrix(const std::string &name, const std::vector<double> &p,
                                  bool inverse, int64_t n_wires) {
    const double isq2 = 1.0 / std::sqrt(2.0);
    auto P = [&](size_t i) { return inverse ? -p.at(i) : p.at(i); };
    std::vector<cd> m;
    int dim = 0;
    bool self_handled_inverse = true; // angles negated above; non-param gates fixed below
    if (name == "Identity") {
        dim = 1 << n_wires;
        m = eye(dim);
    } else if (name == "PauliX") {
        dim = 2, m = {0, 1, 1, 0};
    } else if (name == "PauliY") {
        dim = 2, m = {0, -I1, I1, 0};
    } else if (name == "PauliZ") {
        dim = 2, m = {1, 0, 0, -1};
    } else if (name == "Hadamard") {
        dim = 2, m = {isq2, isq2, isq2, -isq2};
    } else if (name == "S") {
        dim = 2, m = {1, 0, 0, inverse ? -I1 : I1};
    } else if (name == "T") {
        dim = 2, m = {1, 0, 0, std::exp(I1 * (inverse ? -M_PI / 4 : M_PI / 4))};
    } else if (name == "SX") {
        // Gates.hpp getSX: 0.5 [[1+i, 1-i],[1-i, 1+i]]
        cd a{0.5, 0.5}, b{0.5, -0.5};
        if (inverse) a = std::conj(a), b = std::conj(b);
        dim = 2, m = {a, b, b, a};
    } else if (name == "PhaseShift") {
        dim = 2, m = {1, 0, 0, std::exp(I1 * P(0))};
    } else if (name == "RX") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        dim = 2, m = {c, -I1 * s, -I1 * s, c};
    } else if (name == "RY") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        dim = 2, m = {c, -s, s, c};
    } else if (name == "RZ") {
        dim = 2, m = {std::exp(-I1 * (P(0) / 2)), 0, 0, std::exp(I1 * (P(0) / 2))};
    } else if (name == "Rot") {
        // inverse = Rot(-omega,-theta,-phi): GateImplementationsLM.hpp:1257-1280
        dim = 2;
        m = inverse ? rot_matrix(-p.at(2), -p.at(1), -p.at(0)) : rot_matrix(p.at(0), p.at(1), p.at(2));
    } else if (name == "CNOT") {
        dim = 4, m = controlled({0, 1, 1, 0}, 2);
    } else if (name == "CY") {
        dim = 4, m = controlled({0, -I1, I1, 0}, 2);
    } else if (name == "CZ") {
        dim = 4, m = controlled({1, 0, 0, -1}, 2);
    } else if (name == "SWAP") {
        dim = 4, m = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    } else if (name == "IsingXX") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        cd z = -I1 * s;
        dim = 4, m = {c, 0, 0, z, 0, c, z, 0, 0, z, c, 0, z, 0, 0, c};
    } else if (name == "IsingXY") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        cd z = I1 * s;
        dim = 4, m = {1, 0, 0, 0, 0, c, z, 0, 0, z, c, 0, 0, 0, 0, 1};
    } else if (name == "IsingYY") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        cd z = I1 * s;
        dim = 4, m = {c, 0, 0, z, 0, c, -z, 0, 0, -z, c, 0, z, 0, 0, c};
    } else if (name == "IsingZZ") {
        cd e0 = std::exp(-I1 * (P(0) / 2)), e1 = std::exp(I1 * (P(0) / 2));
        dim = 4, m = {e0, 0, 0, 0, 0, e1, 0, 0, 0, 0, e1, 0, 0, 0, 0, e0};
    } else if (name == "ControlledPhaseShift") {
        dim = 4, m = eye(4);
        m[15] = std::exp(I1 * P(0));
    } else if (name == "CRX") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        dim = 4, m = controlled({c, -I1 * s, -I1 * s, c}, 2);
    } else if (name == "CRY") {
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        dim = 4, m = controlled({c, -s, s, c}, 2);
    } else if (name == "CRZ") {
        dim = 4, m = controlled({std::exp(-I1 * (P(0) / 2)), 0, 0, std::exp(I1 * (P(0) / 2))}, 2);
    } else if (name == "CRot") {
        dim = 4;
        m = controlled(inverse ? rot_matrix(-p.at(2), -p.at(1), -p.at(0))
                               : rot_matrix(p.at(0), p.at(1), p.at(2)),
                       2);
    } else if (name == "SingleExcitation" || name == "SingleExcitationMinus" ||
               name == "SingleExcitationPlus") {
        // GateImplementationsLM.hpp:1612-1734
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        cd e = 1.0;
        if (name == "SingleExcitationMinus") e = std::exp(-I1 * (P(0) / 2));
        if (name == "SingleExcitationPlus") e = std::exp(I1 * (P(0) / 2));
        dim = 4, m = {e, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, e};
    } else if (name == "PSWAP") {
        cd e = std::exp(I1 * P(0));
        dim = 4, m = {1, 0, 0, 0, 0, 0, e, 0, 0, e, 0, 0, 0, 0, 0, 1};
    } else if (name == "Toffoli") {
        dim = 8, m = controlled(controlled({0, 1, 1, 0}, 2), 4);
    } else if (name == "CSWAP") {
        dim = 8, m = controlled({1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}, 4);
    } else if (name == "DoubleExcitation" || name == "DoubleExcitationMinus" ||
               name == "DoubleExcitationPlus") {
        // GateImplementationsLM.hpp:1857-1984: Givens rotation on (|0011>, |1100>)
        double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
        cd e = 1.0;
        if (name == "DoubleExcitationMinus") e = std::exp(-I1 * (P(0) / 2));
        if (name == "DoubleExcitationPlus") e = std::exp(I1 * (P(0) / 2));
        dim = 16, m.assign(256, 0.0);
        for (int i = 0; i < 16; i++) m[i * 16 + i] = e;
        m[3 * 16 + 3] = c, m[3 * 16 + 12] = -s, m[12 * 16 + 3] = s, m[12 * 16 + 12] = c;
    } else {
        fail("Gate operation does not exist for " + name);
    }
    (void)self_handled_inverse;
    (void)dim;
    return m;
}

namespace {

uint64_t bit_of_wire(int64_t n, int64_t w) { return uint64_t{1} << (n - 1 - w); }

void check_wires(int64_t n, const std::vector<int64_t> &wires,
                 const std::vector<int64_t> &ctrl_wires,
                 const std::vector<uint8_t> &ctrl_values) {
    PLB_CHECK(ctrl_wires.size() == ctrl_values.size(),
              "`controlled_wires` must have the same size as `controlled_values`.");
    uint64_t seen = 0;
    for (auto w : ctrl_wires) {
        PLB_CHECK(w >= 0 && w < n, "Invalid wire index");
        seen |= bit_of_wire(n, w);
    }
    for (auto w : wires) {
        PLB_CHECK(w >= 0 && w < n, "Invalid wire index");
        PLB_CHECK((seen & bit_of_wire(n, w)) == 0 || std::find(ctrl_wires.begin(), ctrl_wires.end(), w) ==
                                                         ctrl_wires.end(),
                  "`controlled_wires` and `target wires` must be disjoint.");
        PLB_CHECK(std::find(ctrl_wires.begin(), ctrl_wires.end(), w) == ctrl_wires.end(),
                  "`controlled_wires` and `target wires` must be disjoint.");
    }
}

void controls_to_mask(int64_t n, const std::vector<int64_t> &cw, const std::vector<uint8_t> &cv,
                      uint64_t *cmask, uint64_t *cval) {
    *cmask = 0, *cval = 0;
    for (size_t i = 0; i < cw.size(); i++) {
        uint64_t b = bit_of_wire(n, cw[i]);
        *cmask |= b;
        if (cv[i]) *cval |= b;
    }
}

// Core structural analysis of a (possibly non-unitary) 2^k matrix on lsb-first bits `tbits`.
std::vector<COp> analyse(std::vector<cd> m, std::vector<int> tbits, uint64_t cmask, uint64_t cval,
                         bool allow_project_controls) {
    std::vector<COp> out;
    int k = static_cast<int>(tbits.size());
    // 1. implied controls: bit j such that the matrix is identity when bit j == v (both on rows
    //    and columns) and has no coupling between the two halves.
    bool changed = true;
    while (changed && k > 0) {
        changed = false;
        int dim = 1 << k;
        for (int j = 0; j < k && !changed; j++) {
            for (int v = 0; v < 2 && !changed; v++) {
                bool ok = true;
                for (int r = 0; r < dim && ok; r++)
                    for (int c = 0; c < dim && ok; c++) {
                        int rb = (r >> j) & 1, cb = (c >> j) & 1;
                        cd e = m[static_cast<size_t>(r) * dim + c];
                        if (rb != cb) {
                            if (e != cd(0.0)) ok = false;
                        } else if (rb == v) {
                            if (e != (r == c ? cd(1.0) : cd(0.0))) ok = false;
                        }
                    }
                if (!ok) continue;
                // bit j becomes a control with value 1-v; keep the (1-v) block
                int nd = dim >> 1;
                std::vector<cd> r2(static_cast<size_t>(nd) * nd);
                auto expand = [&](int x) {
                    int lo = x & ((1 << j) - 1), hi = x >> j;
                    return (hi << (j + 1)) | ((1 - v) << j) | lo;
                };
                for (int r = 0; r < nd; r++)
                    for (int c = 0; c < nd; c++)
                        r2[static_cast<size_t>(r) * nd + c] = m[static_cast<size_t>(expand(r)) * dim + expand(c)];
                uint64_t b = uint64_t{1} << tbits[j];
                cmask |= b;
                if (v == 0) cval |= b;
                tbits.erase(tbits.begin() + j);
                m.swap(r2);
                k--;
                changed = true;
            }
        }
    }
    (void)allow_project_controls;
    int dim = 1 << k;
    // 2. diagonal?
    bool is_diag = true;
    for (int r = 0; r < dim && is_diag; r++)
        for (int c = 0; c < dim; c++)
            if (r != c && m[static_cast<size_t>(r) * dim + c] != cd(0.0)) {
                is_diag = false;
                break;
            }
    if (is_diag) {
        bool is_id = true;
        for (int r = 0; r < dim; r++)
            if (m[static_cast<size_t>(r) * dim + r] != cd(1.0)) is_id = false;
        if (is_id) return out; // identity
        COp op;
        op.kind = OP_DIAG;
        op.tbits = tbits;
        op.cmask = cmask, op.cval = cval;
        op.diag.resize(dim);
        for (int r = 0; r < dim; r++) op.diag[r] = m[static_cast<size_t>(r) * dim + r];
        out.push_back(std::move(op));
        return out;
    }
    // 3. direct sum of <=2x2 blocks?
    std::vector<int> partner(dim, -1);
    bool blocky = (k <= 4);
    for (int r = 0; r < dim && blocky; r++) {
        for (int c = 0; c < dim; c++) {
            if (r == c) continue;
            if (m[static_cast<size_t>(r) * dim + c] != cd(0.0) || m[static_cast<size_t>(c) * dim + r] != cd(0.0)) {
                if (partner[r] == -1) partner[r] = c;
                else if (partner[r] != c) {
                    blocky = false;
                    break;
                }
            }
        }
    }
    if (blocky)
        for (int r = 0; r < dim; r++)
            if (partner[r] >= 0 && partner[partner[r]] != r) blocky = false;
    if (blocky) {
        COp pairs;
        pairs.kind = OP_PAIRS;
        pairs.tbits = tbits;
        pairs.cmask = cmask, pairs.cval = cval;
        std::vector<cd> rest(dim, 1.0);
        bool rest_nontrivial = false;
        for (int r = 0; r < dim; r++) {
            if (partner[r] < 0) {
                rest[r] = m[static_cast<size_t>(r) * dim + r];
                if (rest[r] != cd(1.0)) rest_nontrivial = true;
            } else if (r < partner[r]) {
                int c = partner[r];
                Block2 b;
                b.a = r, b.b = c;
                b.m[0] = m[static_cast<size_t>(r) * dim + r], b.m[1] = m[static_cast<size_t>(r) * dim + c];
                b.m[2] = m[static_cast<size_t>(c) * dim + r], b.m[3] = m[static_cast<size_t>(c) * dim + c];
                pairs.blocks.push_back(b);
            }
        }
        if (pairs.blocks.size() <= 8) {
            out.push_back(std::move(pairs));
            if (rest_nontrivial) {
                COp d;
                d.kind = OP_DIAG;
                d.tbits = tbits;
                d.cmask = cmask, d.cval = cval;
                d.diag = rest;
                out.push_back(std::move(d));
            }
            return out;
        }
    }
    COp d;
    d.kind = OP_DENSE;
    d.tbits = tbits;
    d.cmask = cmask, d.cval = cval;
    d.mat = std::move(m);
    out.push_back(std::move(d));
    return out;
}

std::vector<int> wires_to_tbits(int64_t n, const std::vector<int64_t> &wires) {
    // matrix bit j (lsb-first) <-> wires[k-1-j] <-> state bit n-1-wire
    int k = static_cast<int>(wires.size());
    std::vector<int> t(k);
    for (int j = 0; j < k; j++) t[j] = static_cast<int>(n - 1 - wires[k - 1 - j]);
    return t;
}

} // namespace

std::vector<COp> lower_matrix(int64_t n, const std::vector<cd> &matrix,
                              const std::vector<int64_t> &wires,
                              const std::vector<int64_t> &ctrl_wires,
                              const std::vector<uint8_t> &ctrl_values, bool inverse,
                              bool allow_nonunitary) {
    PLB_CHECK(!wires.empty(), "Number of wires must be larger than 0");
    check_wires(n, wires, ctrl_wires, ctrl_values);
    {
        std::vector<int64_t> s(wires);
        std::sort(s.begin(), s.end());
        PLB_CHECK(std::adjacent_find(s.begin(), s.end()) == s.end(), "Wires must be unique");
    }
    PLB_CHECK(wires.size() <= 20, "applyMatrix supports at most 20 target wires");
    const size_t dim = size_t{1} << wires.size();
    PLB_CHECK(matrix.size() == dim * dim,
              "The size of matrix does not match with the given number of wires");
    uint64_t cmask, cval;
    controls_to_mask(n, ctrl_wires, ctrl_values, &cmask, &cval);
    (void)allow_nonunitary;
    if (wires.size() > 10) { // too large to analyse: straight dense
        COp d;
        d.kind = OP_DENSE;
        d.tbits = wires_to_tbits(n, wires);
        d.cmask = cmask, d.cval = cval;
        d.mat = inverse ? dagger(matrix, static_cast<int>(dim)) : matrix;
        PLB_CHECK(d.k() <= 11, "applyMatrix: dense matrices on more than 11 wires are not supported");
        return {d};
    }
    auto ops = analyse(inverse ? dagger(matrix, static_cast<int>(dim)) : matrix, wires_to_tbits(n, wires), cmask, cval, false);
    // the dense kernels stage a 2^k column in shared memory: reject what cannot execute HERE (validation time),
    // not after the matrix has been copied to the device
    for (const COp &op : ops)
        PLB_CHECK(op.kind != OP_DENSE || op.k() <= 11, "applyMatrix: dense matrices on more than 11 wires are not supported");
    return ops;
}

std::vector<COp> lower_gate(int64_t n, const GateCall &g) {
    if (!g.matrix.empty() && !gate_known(g.name))
        return lower_matrix(n, g.matrix, g.wires, g.ctrl_wires, g.ctrl_values, g.inverse);
    // "PauliRot[XYZ]": applyPauliRot travelling through a tape (the lazy gate queue of the C++ mirror)
    if (g.name.rfind("PauliRot[", 0) == 0 && g.name.back() == ']') {
        PLB_CHECK(g.params.size() == 1, "PauliRot needs one parameter");
        PLB_CHECK(g.ctrl_wires.empty(), "Controlled gate operation does not exist for PauliRot");
        return lower_pauli_rot(n, g.wires, g.inverse, g.params[0], g.name.substr(9, g.name.size() - 10));
    }
    auto it = gate_table().find(g.name);
    PLB_CHECK(it != gate_table().end(), "Gate operation does not exist for " + g.name);
    const GateInfo gi = it->second;
    if (!g.ctrl_wires.empty())
        PLB_CHECK(controlled_gate_known(g.name),
                  "Controlled gate operation does not exist for " + g.name);
    check_wires(n, g.wires, g.ctrl_wires, g.ctrl_values);
    PLB_CHECK(gi.wires < 0 || static_cast<int64_t>(g.wires.size()) == gi.wires,
              "The number of wires does not match the gate " + g.name);
    PLB_CHECK(static_cast<int>(g.params.size()) == gi.params,
              "The number of parameters does not match the gate " + g.name);
    PLB_CHECK(!g.wires.empty() || g.name == "GlobalPhase", "Number of wires must be larger than 0");
    uint64_t cmask, cval;
    controls_to_mask(n, g.ctrl_wires, g.ctrl_values, &cmask, &cval);

    if (g.name == "Identity") return {};
    if (g.name == "GlobalPhase") {
        // GateImplementationsLM.hpp:2044-2111: multiply by exp(-i phi) (controlled: on the
        // control subspace only); wires are irrelevant.
        COp d;
        d.kind = OP_DIAG;
        d.cmask = cmask, d.cval = cval;
        d.diag = {std::exp(-I1 * (g.inverse ? -g.params[0] : g.params[0]))};
        return {d};
    }
    if (g.name == "MultiRZ") {
        // GateImplementationsLM.hpp:1988-2042: exp(-i theta/2 Z^{(x)k}); parity form
        double th = g.inverse ? -g.params[0] : g.params[0];
        COp d;
        d.kind = OP_DIAG;
        d.cmask = cmask, d.cval = cval;
        d.parity = true;
        for (auto w : g.wires) {
            PLB_CHECK((d.pmask & bit_of_wire(n, w)) == 0, "Wires must be unique");
            d.pmask |= bit_of_wire(n, w);
        }
        d.pd[0] = std::exp(-I1 * (th / 2));
        d.pd[1] = std::exp(I1 * (th / 2));
        return {d};
    }
    if (g.name == "PCPhase") {
        // GateImplementationsLM.hpp:2113-2162
        const int64_t k = static_cast<int64_t>(g.wires.size());
        PLB_CHECK(k <= 24, "PCPhase supports at most 24 wires");
        const double dimf = std::round(g.params[1]);
        PLB_CHECK(dimf >= 0 && dimf <= std::ldexp(1.0, static_cast<int>(n)),
                  "The dimension of the PCPhase gate must be a positive integer and less than or "
                  "equal to statevector size.");
        const size_t dsz = static_cast<size_t>(dimf);
        const double ph = g.inverse ? -g.params[0] : g.params[0];
        const cd up{std::cos(ph), std::sin(ph)};
        COp d;
        d.kind = OP_DIAG;
        d.tbits = wires_to_tbits(n, g.wires);
        d.cmask = cmask, d.cval = cval;
        d.diag.resize(size_t{1} << k);
        for (size_t i = 0; i < d.diag.size(); i++) d.diag[i] = i < dsz ? up : std::conj(up);
        return {d};
    }
    auto m = named_gate_matrix(g.name, g.params, g.inverse, static_cast<int64_t>(g.wires.size()));
    {
        std::vector<int64_t> s(g.wires);
        std::sort(s.begin(), s.end());
        PLB_CHECK(std::adjacent_find(s.begin(), s.end()) == s.end(), "Wires must be unique");
    }
    return analyse(std::move(m), wires_to_tbits(n, g.wires), cmask, cval, false);
}

PauliWordMask pauli_word_mask(int64_t n, const std::string &word,
                              const std::vector<int64_t> &wires) {
    PLB_CHECK(wires.size() == word.size(), "wires and word have incompatible dimensions.");
    PauliWordMask p;
    for (size_t i = 0; i < word.size(); i++) {
        PLB_CHECK(wires[i] >= 0 && wires[i] < n, "Invalid wire index");
        const uint64_t b = bit_of_wire(n, wires[i]);
        PLB_CHECK(((p.x | p.z) & b) == 0 || word[i] == 'I', "Wires must be unique");
        switch (word[i]) {
        case 'I':
            break;
        case 'X':
            p.x |= b;
            break;
        case 'Y':
            p.x |= b, p.z |= b, p.ny++;
            break;
        case 'Z':
            p.z |= b;
            break;
        default:
            fail("Invalid Pauli word character");
        }
    }
    return p;
}

std::vector<COp> lower_pauli_rot(int64_t n, const std::vector<int64_t> &wires, bool inverse,
                                 double theta, const std::string &word) {
    // exp(-i theta/2 P) = cos(theta/2) I - i sin(theta/2) P, GateImplementationsLM.hpp:575-629.
    // P|j> = i^{ny} (-1)^{popcount(j & z)} |j ^ x>.  For x == 0 this is a parity-diagonal; else
    // amplitudes pair up as (j, j^x) and every pair gets a 2x2 whose off-diagonal phases depend
    // on popcount(j & z): expressed as OP_PAIRS over the bits of x|z when small, else as a
    // dedicated kernel op (parity-paired form carried in COp::parity + blocks).
    PLB_CHECK(wires.size() == word.size(), "wires and word have incompatible dimensions.");
    const PauliWordMask p = pauli_word_mask(n, word, wires);
    const double th = inverse ? -theta : theta;
    const double c = std::cos(th / 2), s = std::sin(th / 2);
    if (p.x == 0) {
        if (p.z == 0) { // identity word: global phase exp(-i th/2)
            COp d;
            d.kind = OP_DIAG;
            d.diag = {std::exp(-I1 * (th / 2))};
            return {d};
        }
        COp d;
        d.kind = OP_DIAG;
        d.parity = true;
        d.pmask = p.z;
        d.pd[0] = std::exp(-I1 * (th / 2));
        d.pd[1] = std::exp(I1 * (th / 2));
        return {d};
    }
    // General word: OP_PAIRS in "parity-paired" form.  One block; a = 0, b = x (as state offsets,
    // pivot = lowest set bit of x is the inserted bit), phases depend on parity of (j & z).
    //   new[j]   = c a_j     - i s * conj-phase ... handled in-kernel:
    //   (P psi)[j] = i^{ny} (-1)^{pc((j^x) & z)} psi[j^x]
    COp op;
    op.kind = OP_PAIRS;
    op.parity = true; // marks parity-paired form
    op.pmask = p.z;
    Block2 b;
    b.a = 0, b.b = 0;
    // store: m[0] = c, m[1] = -i s i^{ny};  kernel applies signs
    cd iy = 1.0;
    for (int q = 0; q < (p.ny & 3); q++) iy *= I1;
    b.m[0] = c, b.m[1] = -I1 * s * iy, b.m[2] = 0, b.m[3] = 0;
    op.blocks.push_back(b);
    // tbits carries the bits of x (flip mask), lsb-first
    for (int bit = 0; bit < 64; bit++)
        if (p.x >> bit & 1) op.tbits.push_back(bit);
    return {op};
}

// -------------------------------------------------------------------------------------
// Generators (GateImplementationsLM.hpp:2181-2952, PauliGenerator.hpp:31-59)
// -------------------------------------------------------------------------------------
namespace {
struct GenInfo {
    int wires;
    double scale;
};
const std::map<std::string, GenInfo> &gen_table() {
    static const std::map<std::string, GenInfo> t = {
        {"PhaseShift", {1, 1.0}},
        {"RX", {1, -0.5}},
        {"RY", {1, -0.5}},
        {"RZ", {1, -0.5}},
        {"IsingXX", {2, -0.5}},
        {"IsingXY", {2, 0.5}},
        {"IsingYY", {2, -0.5}},
        {"IsingZZ", {2, -0.5}},
        {"CRX", {2, -0.5}},
        {"CRY", {2, -0.5}},
        {"CRZ", {2, -0.5}},
        {"ControlledPhaseShift", {2, 1.0}},
        {"SingleExcitation", {2, -0.5}},
        {"SingleExcitationMinus", {2, -0.5}},
        {"SingleExcitationPlus", {2, -0.5}},
        {"DoubleExcitation", {4, -0.5}},
        {"DoubleExcitationMinus", {4, -0.5}},
        {"DoubleExcitationPlus", {4, 0.5}},
        {"PSWAP", {2, 1.0}},
        {"MultiRZ", {-1, -0.5}},
        {"GlobalPhase", {-1, -1.0}},
    };
    return t;
}
bool controlled_gen_known(const std::string &n) {
    return n != "CRX" && n != "CRY" && n != "CRZ" && n != "ControlledPhaseShift" &&
           gen_table().count(n);
}

std::vector<cd> generator_matrix(const std::string &name) {
    const cd X[4] = {0, 1, 1, 0}, Y[4] = {0, -I1, I1, 0}, Z[4] = {1, 0, 0, -1};
    auto kron2 = [](const cd *a, const cd *b) {
        std::vector<cd> m(16);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++)
                    for (int l = 0; l < 2; l++) m[(i * 2 + k) * 4 + (j * 2 + l)] = a[i * 2 + j] * b[k * 2 + l];
        return m;
    };
    auto proj1 = [](const std::vector<cd> &u, int dim) { // |1><1| (x) u, zero elsewhere
        int D = 2 * dim;
        std::vector<cd> m(static_cast<size_t>(D) * D, 0.0);
        for (int i = 0; i < dim; i++)
            for (int j = 0; j < dim; j++) m[static_cast<size_t>(dim + i) * D + dim + j] = u[i * dim + j];
        return m;
    };
    if (name == "PhaseShift") return {0, 0, 0, 1};
    if (name == "RX") return {X, X + 4};
    if (name == "RY") return {Y, Y + 4};
    if (name == "RZ") return {Z, Z + 4};
    if (name == "IsingXX") return kron2(X, X);
    if (name == "IsingYY") return kron2(Y, Y);
    if (name == "IsingZZ") return kron2(Z, Z);
    if (name == "IsingXY" || name == "PSWAP") { // swap 01<->10, zero 00,11
        std::vector<cd> m(16, 0.0);
        m[1 * 4 + 2] = 1, m[2 * 4 + 1] = 1;
        return m;
    }
    if (name == "CRX") return proj1({X, X + 4}, 2);
    if (name == "CRY") return proj1({Y, Y + 4}, 2);
    if (name == "CRZ") return proj1({Z, Z + 4}, 2);
    if (name == "ControlledPhaseShift") {
        std::vector<cd> m(16, 0.0);
        m[15] = 1;
        return m;
    }
    if (name == "SingleExcitation" || name == "SingleExcitationMinus" ||
        name == "SingleExcitationPlus") {
        std::vector<cd> m(16, 0.0);
        m[1 * 4 + 2] = -I1, m[2 * 4 + 1] = I1;
        if (name == "SingleExcitationMinus") m[0] = 1, m[15] = 1;
        if (name == "SingleExcitationPlus") m[0] = -1, m[15] = -1;
        return m;
    }
    if (name == "DoubleExcitation" || name == "DoubleExcitationMinus" ||
        name == "DoubleExcitationPlus") {
        std::vector<cd> m(256, 0.0);
        if (name != "DoubleExcitation")
            for (int i = 0; i < 16; i++) m[i * 16 + i] = 1;
        m[3 * 16 + 3] = 0, m[12 * 16 + 12] = 0;
        if (name == "DoubleExcitationPlus")
            m[3 * 16 + 12] = I1, m[12 * 16 + 3] = -I1;
        else
            m[3 * 16 + 12] = -I1, m[12 * 16 + 3] = I1;
        return m;
    }
    fail("Generator operation does not exist for " + name);
}
} // namespace

std::vector<COp> lower_generator(int64_t n, const GateCall &g, double *scale) {
    auto it = gen_table().find(g.name);
    PLB_CHECK(it != gen_table().end(), "Generator operation does not exist for " + g.name);
    if (!g.ctrl_wires.empty())
        PLB_CHECK(controlled_gen_known(g.name),
                  "Controlled generator operation does not exist for " + g.name);
    check_wires(n, g.wires, g.ctrl_wires, g.ctrl_values);
    PLB_CHECK(it->second.wires < 0 || static_cast<int64_t>(g.wires.size()) == it->second.wires,
              "The number of wires does not match the generator " + g.name);
    *scale = it->second.scale;
    uint64_t cmask, cval;
    controls_to_mask(n, g.ctrl_wires, g.ctrl_values, &cmask, &cval);
    std::vector<COp> out;
    if (g.name == "GlobalPhase") {
        // identity on the control subspace (no-op uncontrolled)
    } else if (g.name == "MultiRZ") {
        COp d;
        d.kind = OP_DIAG;
        d.cmask = cmask, d.cval = cval;
        d.parity = true;
        for (auto w : g.wires) d.pmask |= bit_of_wire(n, w);
        d.pd[0] = 1.0, d.pd[1] = -1.0;
        out.push_back(d);
    } else {
        // Generator matrices contain exact zeros/ones; the structural analysis must not turn
        // zero blocks into "implied controls" wrongly — analyse() only extracts identity
        // blocks, which is still correct for non-unitary matrices.
        out = analyse(generator_matrix(g.name), wires_to_tbits(n, g.wires), cmask, cval, false);
    }
    if (cmask) {
        COp p;
        p.kind = OP_PROJECT;
        p.cmask = cmask, p.cval = cval;
        out.push_back(p);
    }
    return out;
}

bool generator_as_pauli(int64_t n, const GateCall &g, PauliWordMask *out, double *scale) {
    auto it = gen_table().find(g.name);
    if (it == gen_table().end()) return false;
    if (it->second.wires >= 0 && static_cast<int64_t>(g.wires.size()) != it->second.wires) return false;
    PauliWordMask p;
    controls_to_mask(n, g.ctrl_wires, g.ctrl_values, &p.cmask, &p.cval);
    auto W = [&](size_t i) { return bit_of_wire(n, g.wires.at(i)); };
    const std::string &nm = g.name;
    *scale = it->second.scale;
    if (nm == "RX") p.x = W(0);
    else if (nm == "RY") p.x = W(0), p.z = W(0), p.ny = 1;
    else if (nm == "RZ") p.z = W(0);
    else if (nm == "PhaseShift") p.cmask |= W(0), p.cval |= W(0);
    else if (nm == "IsingXX") p.x = W(0) | W(1);
    else if (nm == "IsingYY") p.x = W(0) | W(1), p.z = p.x, p.ny = 2;
    else if (nm == "IsingZZ") p.z = W(0) | W(1);
    else if (nm == "CRX") p.cmask |= W(0), p.cval |= W(0), p.x = W(1);
    else if (nm == "CRY") p.cmask |= W(0), p.cval |= W(0), p.x = W(1), p.z = W(1), p.ny = 1;
    else if (nm == "CRZ") p.cmask |= W(0), p.cval |= W(0), p.z = W(1);
    else if (nm == "ControlledPhaseShift") p.cmask |= W(0) | W(1), p.cval |= W(0) | W(1);
    else if (nm == "MultiRZ") {
        for (size_t i = 0; i < g.wires.size(); i++) p.z |= W(i);
    } else if (nm == "GlobalPhase") {
        // identity (x) control projector
    } else
        return false;
    *out = p;
    return true;
}

} // namespace plb200
