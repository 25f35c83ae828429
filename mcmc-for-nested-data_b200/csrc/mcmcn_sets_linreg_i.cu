// mcmcn_sets_linreg_i.cu -- kernel instantiations (see mcmcn_registry.h): K = 15, 16 coefficients.
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<15>, 15, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<15>, 15, 64, double, 1),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<16>, 16, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<16>, 16, 64, double, 1),
};
const KernelSet* sets_linreg_i(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
