// mcmcn_sets_tc.cu -- instantiations of the tcgen05 step kernel (mcmcn_tc.cuh).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
#include "mcmcn_tc.cuh"
namespace mcmcn {
// f = MCMCN_F_PARTIAL | MCMCN_F_COUNT for the production variants, -1 for the general kernel
sweep_fn tc_sweep_kernel(int f) {
    switch (f) {
        case 0: return sweep_tc_kernel<0>;
        case 1: return sweep_tc_kernel<1>;
        case 2: return sweep_tc_kernel<2>;
        case 3: return sweep_tc_kernel<3>;
        default: return sweep_tc_kernel<-1>;
    }
}
}  // namespace mcmcn
