// Internal header of libmvtools_b200.so (not installed; the public ABI is
// include/mvtools_b200.h).
#pragma once

#include "../../include/mvtools_b200.h"

namespace mvb200 {

// sets the thread-local message returned by mvb200_last_error(); returns 1
int set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));

int plane_dim_luma(int src, int level, int ratio, int pad);

} // namespace mvb200
