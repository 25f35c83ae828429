// mcmcn_sets_tc2.cu -- instantiations of the tcgen05 step kernel for K = 9..16 coefficients (two K blocks
// of 8: accumulator chunks of 96 observations, 7 MMAs per chunk; mcmcn_tc.cuh).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
#include "mcmcn_tc.cuh"
namespace mcmcn {
sweep_fn tc_sweep_kernel_two_blocks(int f) {
    switch (f) {
        case 0: return sweep_tc_kernel<0, false, 2>;
        case 1: return sweep_tc_kernel<1, false, 2>;
        case 2: return sweep_tc_kernel<2, false, 2>;
        case 3: return sweep_tc_kernel<3, false, 2>;
        default: return sweep_tc_kernel<-1, false, 2>;
    }
}
eval_tc_fn tc_eval_kernel_two_blocks() { return eval_tc_kernel<2>; }
}  // namespace mcmcn
