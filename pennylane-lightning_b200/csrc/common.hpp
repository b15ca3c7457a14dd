// Shared host-side declarations of the plb200 engine (libplb200.so).
#pragma once
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace plb200 {

using cd = std::complex<double>;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] inline void fail(const std::string &msg) { throw Error(msg); }
#define PLB_CHECK(cond, msg)                                                   \
    do {                                                                       \
        if (!(cond)) ::plb200::fail(msg);                                      \
    } while (0)

// ---------------------------------------------------------------------------------------
// Canonical operations.  Every named gate / matrix / generator of the reference is lowered on
// the host (in double precision) into a short list of these; the CUDA side has one kernel
// family per kind.  Bit positions are state-index bits (bit b <-> wire n-1-b).
// ---------------------------------------------------------------------------------------
enum OpKind : int {
    OP_PAIRS = 0,   // direct sum of 2x2 blocks on pairs (a,b) of the 2^k target subspace
    OP_DIAG = 1,    // diagonal: table over the k target bits, or parity form
    OP_DENSE = 2,   // dense 2^k x 2^k matrix
    OP_PROJECT = 3, // zero every amplitude whose control bits do not match (generators)
};

struct Block2 {
    uint32_t a, b; // local indices inside the 2^k target space (lsb-first over tbits)
    cd m[4];       // row-major 2x2 acting on (amp[a], amp[b])
};

struct COp {
    OpKind kind = OP_DIAG;
    std::vector<int> tbits;   // target bits, tbits[0] = least-significant matrix bit
    uint64_t cmask = 0;       // control bits
    uint64_t cval = 0;        // required values on the control bits
    std::vector<Block2> blocks; // OP_PAIRS
    std::vector<cd> diag;       // OP_DIAG table (size 2^k), k = tbits.size()
    bool parity = false;        // OP_DIAG parity form: pd[popcount(i & pmask) & 1]
    uint64_t pmask = 0;
    cd pd[2];
    std::vector<cd> mat; // OP_DENSE row-major
    // True when the op leaves amplitudes outside its control subspace untouched and acts as a
    // unitary/diagonal inside: always the case except OP_PROJECT.
    int k() const { return static_cast<int>(tbits.size()); }
    // bits whose value changes the *action* non-diagonally (needed inside a fused tile)
    uint64_t nondiag_mask() const {
        uint64_t m = 0;
        if (kind == OP_PAIRS || kind == OP_DENSE)
            for (int b : tbits) m |= (uint64_t{1} << b);
        return m;
    }
};

struct GateCall {
    std::string name;
    std::vector<int64_t> wires;
    std::vector<int64_t> ctrl_wires;
    std::vector<uint8_t> ctrl_values;
    std::vector<double> params;
    bool inverse = false;
    std::vector<cd> matrix; // optional explicit matrix (row-major)
};

// lowering.cpp -------------------------------------------------------------------------
// Gate -> canonical ops.  Throws plb200::Error with the reference's message on bad input.
std::vector<COp> lower_gate(int64_t n, const GateCall &g);
std::vector<COp> lower_matrix(int64_t n, const std::vector<cd> &matrix,
                              const std::vector<int64_t> &wires,
                              const std::vector<int64_t> &ctrl_wires,
                              const std::vector<uint8_t> &ctrl_values, bool inverse,
                              bool allow_nonunitary = false);
std::vector<COp> lower_pauli_rot(int64_t n, const std::vector<int64_t> &wires, bool inverse,
                                 double theta, const std::string &word);
// Generator of a named gate: canonical ops computing G|psi> (controls -> projector (x) G) and
// the scaling factor the reference returns (GateImplementationsLM.hpp:2181-2952).
std::vector<COp> lower_generator(int64_t n, const GateCall &g, double *scale);
// If the generator of `g` (including controls) is c * (controlled) Pauli word, describe it:
// flip mask x, sign mask z, number of Y factors, control mask/value.  Returns false otherwise.
struct PauliWordMask {
    uint64_t x = 0, z = 0;
    int ny = 0;
    uint64_t cmask = 0, cval = 0;
};
bool generator_as_pauli(int64_t n, const GateCall &g, PauliWordMask *out, double *scale);
bool gate_known(const std::string &name);
int gate_num_params(const std::string &name); // -1 unknown
std::vector<cd> named_gate_matrix(const std::string &name, const std::vector<double> &params,
                                  bool inverse, int64_t n_wires);
PauliWordMask pauli_word_mask(int64_t n, const std::string &word,
                              const std::vector<int64_t> &wires);

} // namespace plb200
