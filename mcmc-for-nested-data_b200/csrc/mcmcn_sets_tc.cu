// mcmcn_sets_tc.cu -- instantiations of the tcgen05 step kernel (mcmcn_tc.cuh).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
#include "mcmcn_tc.cuh"
namespace mcmcn {
// f = MCMCN_F_PARTIAL | MCMCN_F_COUNT for the production variants, -1 for the general kernel;
// uniform208: every group pads to 208 observations (see sweep_tc_kernel)
sweep_fn tc_sweep_kernel(int f, bool uniform208, int k_blocks) {
    if (k_blocks == 2) return tc_sweep_kernel_two_blocks(f);
    if (uniform208) {
        switch (f) {
            case 0: return sweep_tc_kernel<0, true>;
            case 1: return sweep_tc_kernel<1, true>;
            case 2: return sweep_tc_kernel<2, true>;
            case 3: return sweep_tc_kernel<3, true>;
            default: return sweep_tc_kernel<-1, true>;
        }
    }
    switch (f) {
        case 0: return sweep_tc_kernel<0, false>;
        case 1: return sweep_tc_kernel<1, false>;
        case 2: return sweep_tc_kernel<2, false>;
        case 3: return sweep_tc_kernel<3, false>;
        default: return sweep_tc_kernel<-1, false>;
    }
}
eval_tc_fn tc_eval_kernel(int k_blocks) { return k_blocks == 2 ? tc_eval_kernel_two_blocks() : eval_tc_kernel<1>; }
}  // namespace mcmcn
