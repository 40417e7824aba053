// oracle/shims -- functional stand-in for srrg_core's system_utils.h (branch marchless, un-vendored).
// TEST INFRASTRUCTURE ONLY (see oracle/shims/Eigen/Core).  getTime(): wall-clock seconds, what the reference's
// CHRONOMETER macros (src/types/definitions.h:147-151) accumulate.
#pragma once
#include <chrono>
#include <fstream>
#include <string>
namespace srrg_core {
inline double getTime() {
  return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}
inline bool isAccessible(const std::string& filename) { return std::ifstream(filename.c_str()).good(); }
inline std::string getTimestamp() { return std::to_string(getTime()); }
}  // namespace srrg_core
