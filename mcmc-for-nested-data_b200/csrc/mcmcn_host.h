// mcmcn_host.h -- host-side helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>

#include "mcmcn.h"

namespace mcmcn {

void set_error(const char* fmt, ...);

}  // namespace mcmcn

// Return MCMCN_ERR_CUDA with the runtime's message if a CUDA call failed.
#define CK(call)                                                                       \
    do {                                                                               \
        const cudaError_t mcmcn_e_ = (call);                                           \
        if (mcmcn_e_ != cudaSuccess) {                                                 \
            mcmcn::set_error("%s failed: %s", #call, cudaGetErrorString(mcmcn_e_));    \
            return MCMCN_ERR_CUDA;                                                     \
        }                                                                              \
    } while (0)
