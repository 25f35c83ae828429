// mcmcn_nvrtc.cu -- user objectives compiled at run time (north star (1)).
//
// The user supplies CUDA source for one device function, mcmc_obj_loglik (contract in
// include/mcmcn.h).  It is compiled by NVRTC for sm_100a together with the step-path header
// (embedded in this library at build time), the resulting cubin is loaded with
// cudaLibraryLoadData and the same kernel templates the registry objectives use are looked up
// by their lowered names.  libnvrtc is opened with dlopen on first use, so the library itself
// loads without it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mcmcn_host.h"
#include "mcmcn_registry.h"
#include "build/embedded_headers.inc"

namespace mcmcn {

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* so = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*DestroyProgram)(nvrtcProgram*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*AddNameExpression)(nvrtcProgram, const char*);
    int (*GetLoweredName)(nvrtcProgram, const char*, const char**);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    const char* (*GetErrorString)(int);
};

static Nvrtc* nvrtc() {
    static Nvrtc api;
    static bool tried = false;
    if (tried) return api.so ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* n : names) {
        api.so = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.so) break;
    }
    if (!api.so) return nullptr;
#define MCMCN_SYM(field, name)                                        \
    *(void**)(&api.field) = dlsym(api.so, name);                      \
    if (!api.field) { dlclose(api.so); api.so = nullptr; return nullptr; }
    MCMCN_SYM(CreateProgram, "nvrtcCreateProgram")
    MCMCN_SYM(DestroyProgram, "nvrtcDestroyProgram")
    MCMCN_SYM(CompileProgram, "nvrtcCompileProgram")
    MCMCN_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    MCMCN_SYM(GetProgramLog, "nvrtcGetProgramLog")
    MCMCN_SYM(AddNameExpression, "nvrtcAddNameExpression")
    MCMCN_SYM(GetLoweredName, "nvrtcGetLoweredName")
    MCMCN_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    MCMCN_SYM(GetCUBIN, "nvrtcGetCUBIN")
    MCMCN_SYM(GetErrorString, "nvrtcGetErrorString")
#undef MCMCN_SYM
    return &api;
}

struct UserObjective {
    KernelSet set;
    cudaLibrary_t lib;
    std::vector<char> cubin;
};

const KernelSet* user_kernel_set(const void* handle) {
    return handle ? &static_cast<const UserObjective*>(handle)->set : nullptr;
}

}  // namespace mcmcn

using namespace mcmcn;

extern "C" {

int mcmcn_user_objective_compile(const char* source, int32_t n_params, int32_t obs_floats, int32_t hdr_floats,
                                 int32_t precision, void** out_handle) {
    if (!source || !out_handle) { set_error("null source / handle"); return MCMCN_ERR_INVALID; }
    if (n_params < 1 || n_params > MCMCN_MAX_PARAMS || obs_floats < 0 || hdr_floats < 0 || (obs_floats & 3) || (hdr_floats & 3) ||
        (precision != 32 && precision != 64)) {
        set_error("user objective: need 1 <= P <= %d, obs_floats and hdr_floats multiples of 4, precision 32 or 64", MCMCN_MAX_PARAMS);
        return MCMCN_ERR_INVALID;
    }
    Nvrtc* rt = nvrtc();
    if (!rt) { set_error("libnvrtc could not be loaded: %s", dlerror() ? dlerror() : "not found"); return MCMCN_ERR_UNSUPPORTED; }

    const char* real = precision == 32 ? "float" : "double";
    const int cw = precision == 32 ? 4 : 2;
    char prologue[512];
    snprintf(prologue, sizeof(prologue),
             "#define MCMCN_USER_P %d\n#define MCMCN_USER_OBS %d\n#define MCMCN_USER_HDR %d\ntypedef %s mcmc_real;\n"
             "#line 1 \"user_objective.cu\"\n", n_params, obs_floats, hdr_floats, real);
    std::string src = std::string(prologue) + source + "\n#include \"mcmcn_device.cuh\"\n";

    nvrtcProgram prog = nullptr;
    const char* hdr_src[] = {kEmbeddedMcmcnH, kEmbeddedDeviceCuh};
    const char* hdr_names[] = {"mcmcn.h", "mcmcn_device.cuh"};
    int rc = rt->CreateProgram(&prog, src.c_str(), "mcmcn_user.cu", 2, hdr_src, hdr_names);
    if (rc) { set_error("nvrtcCreateProgram: %s", rt->GetErrorString(rc)); return MCMCN_ERR_NVRTC; }

    char expr[12][160];
    int n_expr = 0;
    for (int f = 0; f < 4; ++f) snprintf(expr[n_expr++], 160, "mcmcn::sweep_kernel<mcmcn::UserObj, %d, %s, 3, %d, 128>", cw, real, f);
    snprintf(expr[n_expr++], 160, "mcmcn::sweep_kernel<mcmcn::UserObj, %d, %s, 2, -1>", cw, real);
    snprintf(expr[n_expr++], 160, "mcmcn::sweep_kernel<mcmcn::UserObj, 1, %s, 1, -1>", real);
    snprintf(expr[n_expr++], 160, "mcmcn::eval_kernel<mcmcn::UserObj, %d, %s>", cw, real);
    snprintf(expr[n_expr++], 160, "mcmcn::eval_kernel<mcmcn::UserObj, 1, %s>", real);
    snprintf(expr[n_expr++], 160, "mcmcn::pointwise_kernel<mcmcn::UserObj, %s>", real);
    for (int i = 0; i < n_expr; ++i) rt->AddNameExpression(prog, expr[i]);

    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    rc = rt->CompileProgram(prog, 4, opts);
    if (rc) {
        size_t n = 0;
        rt->GetProgramLogSize(prog, &n);
        std::string log(n ? n : 1, '\0');
        if (n) rt->GetProgramLog(prog, &log[0]);
        set_error("user objective failed to compile (%s):\n%.900s", rt->GetErrorString(rc), log.c_str());
        rt->DestroyProgram(&prog);
        return MCMCN_ERR_NVRTC;
    }
    UserObjective* uo = new UserObjective();
    size_t nb = 0;
    rt->GetCUBINSize(prog, &nb);
    uo->cubin.resize(nb);
    rt->GetCUBIN(prog, uo->cubin.data());
    std::string lowered[12];
    for (int i = 0; i < n_expr; ++i) {
        const char* nm = nullptr;
        rt->GetLoweredName(prog, expr[i], &nm);
        lowered[i] = nm ? nm : "";
    }
    rt->DestroyProgram(&prog);

    cudaError_t e = cudaLibraryLoadData(&uo->lib, uo->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) { set_error("cudaLibraryLoadData: %s", cudaGetErrorString(e)); delete uo; return MCMCN_ERR_CUDA; }
    cudaKernel_t k[12];
    for (int i = 0; i < n_expr; ++i) {
        e = cudaLibraryGetKernel(&k[i], uo->lib, lowered[i].c_str());
        if (e != cudaSuccess) {
            set_error("kernel %s not found in the compiled objective: %s", expr[i], cudaGetErrorString(e));
            cudaLibraryUnload(uo->lib);
            delete uo;
            return MCMCN_ERR_CUDA;
        }
    }
    KernelSet& s = uo->set;
    memset(&s, 0, sizeof(s));
    s.objective = MCMCN_OBJ_USER;
    s.P = n_params;
    s.K = 0;
    s.precision = precision;
    s.c_wide = cw;
    for (int f = 0; f < 4; ++f) s.sweep_fast[f] = (sweep_fn)(void*)k[f];
    s.sweep_wide = (sweep_fn)(void*)k[4];
    s.sweep_one = (sweep_fn)(void*)k[5];
    s.eval_wide = (sweep_fn)(void*)k[6];
    s.eval_one = (sweep_fn)(void*)k[7];
    s.pointwise = (pointwise_fn)(void*)k[8];
    s.elem_bytes = precision == 32 ? 4 : 8;
    s.park_doubles = 7;
    *out_handle = uo;
    return MCMCN_OK;
}

int mcmcn_user_objective_free(void* handle) {
    if (!handle) return MCMCN_OK;
    UserObjective* uo = static_cast<UserObjective*>(handle);
    cudaLibraryUnload(uo->lib);
    delete uo;
    return MCMCN_OK;
}

}  // extern "C"
