// mcmcn_sets_linreg_d.cu -- kernel instantiations (see mcmcn_registry.h).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<3>, 3, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<3>, 3, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<5>, 5, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<5>, 5, 64, double, 2),
};
const KernelSet* sets_linreg_d(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
