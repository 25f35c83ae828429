// mcmcn_sets_linreg_g.cu -- kernel instantiations (see mcmcn_registry.h): K = 11, 12 coefficients.
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<11>, 11, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<11>, 11, 64, double, 1),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<12>, 12, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<12>, 12, 64, double, 1),
};
const KernelSet* sets_linreg_g(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
