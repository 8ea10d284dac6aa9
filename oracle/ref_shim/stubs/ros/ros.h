// TEST INFRASTRUCTURE: stand-in for <ros/ros.h>.  trackdlo.cpp uses only the rosconsole logging
// macros (trackdlo.cpp:426,434,931-981); they are routed to tdlo_ref_shim::log so that the
// harness (oracle/ref_harness.cpp) can recover what the reference only logs: the iteration count
// ("Iteration until convergence: k") and the tracking_step state ("Tail occluded", ...).
//
// <math.h>/<stdlib.h>: in the real build ros.h drags the C headers in (boost), which is what makes
// the reference's unqualified abs(double) (trackdlo.cpp:228,337,340,345,348) resolve to the
// floating-point overload; reproduce that here.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <cstdio>
#include <sstream>
#include <string>

namespace tdlo_ref_shim {
void log(int level, const std::string& msg);      // defined by the harness
inline void logf(int level, const char* fmt) { log(level, std::string(fmt)); }
template <typename... A> inline void logf(int level, const char* fmt, A... a) {
    char buf[512];
    std::snprintf(buf, sizeof buf, fmt, a...);
    log(level, std::string(buf));
}
}  // namespace tdlo_ref_shim

#define ROS_INFO(...) ::tdlo_ref_shim::logf(0, __VA_ARGS__)
#define ROS_WARN(...) ::tdlo_ref_shim::logf(1, __VA_ARGS__)
#define ROS_ERROR(...) ::tdlo_ref_shim::logf(2, __VA_ARGS__)
#define ROS_INFO_STREAM(x) do { std::stringstream tdlo_ss_; tdlo_ss_ << x; ::tdlo_ref_shim::log(0, tdlo_ss_.str()); } while (0)
#define ROS_WARN_STREAM(x) do { std::stringstream tdlo_ss_; tdlo_ss_ << x; ::tdlo_ref_shim::log(1, tdlo_ss_.str()); } while (0)
#define ROS_ERROR_STREAM(x) do { std::stringstream tdlo_ss_; tdlo_ss_ << x; ::tdlo_ref_shim::log(2, tdlo_ss_.str()); } while (0)
