// mcmcn_sets_linreg_h.cu -- kernel instantiations (see mcmcn_registry.h): K = 13, 14 coefficients.
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<13>, 13, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<13>, 13, 64, double, 1),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<14>, 14, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<14>, 14, 64, double, 1),
};
const KernelSet* sets_linreg_h(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
