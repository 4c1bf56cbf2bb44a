// Shared host-side helpers: thread-local error string and status plumbing for the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/foley_b200.h"

namespace foley {

std::string& last_error_ref();

inline foley_status fail(foley_status code, const std::string& msg) {
    last_error_ref() = msg;
    return code;
}

#define FOLEY_CUDA_OK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ::foley::fail(FOLEY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

}  // namespace foley
