// mcmcn_nvrtc.cu -- user objectives compiled at run time (placeholder until the NVRTC path lands).
#include "mcmcn_host.h"

extern "C" {

int mcmcn_user_objective_compile(const char*, int32_t, int32_t, int32_t, int32_t, void**) {
    mcmcn::set_error("user objectives (NVRTC) are not built into this library yet");
    return MCMCN_ERR_UNSUPPORTED;
}

int mcmcn_user_objective_free(void*) { return MCMCN_OK; }

}  // extern "C"
