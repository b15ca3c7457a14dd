// Error reporting of the B200 backend, mirroring core/utils/Error.hpp:23-139 of the reference:
// every failure is a thrown LightningException whose text is
//   "[file][Line:n][Method:f]: Error in PennyLane Lightning: <msg>"
// so tests that match substrings of <msg> behave the same.
#pragma once
#include <exception>
#include <sstream>
#include <string>

namespace Pennylane::LightningB200::Util {

class LightningException : public std::exception {
  public:
    explicit LightningException(std::string msg) : msg_(std::move(msg)) {}
    [[nodiscard]] const char *what() const noexcept override { return msg_.c_str(); }

  private:
    std::string msg_;
};

[[noreturn]] inline void Abort(const std::string &message, const char *file, int line, const char *func) {
    std::stringstream s;
    s << "[" << file << "][Line:" << line << "][Method:" << func << "]: Error in PennyLane Lightning: " << message;
    throw LightningException(s.str());
}

} // namespace Pennylane::LightningB200::Util

#define PLB200_ABORT(msg) ::Pennylane::LightningB200::Util::Abort(msg, __FILE__, __LINE__, __func__)
#define PLB200_ABORT_IF(cond, msg)                                                                       \
    if (cond) PLB200_ABORT(msg)
#define PLB200_ABORT_IF_NOT(cond, msg)                                                                   \
    if (!(cond)) PLB200_ABORT(msg)
// status code of the C ABI -> exception carrying plb200_last_error()
#define PLB200_ABI(call)                                                                                 \
    do {                                                                                                 \
        if ((call) != 0) PLB200_ABORT(std::string(plb200_last_error()));                                 \
    } while (0)
