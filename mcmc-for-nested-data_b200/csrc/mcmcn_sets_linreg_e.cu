// mcmcn_sets_linreg_e.cu -- kernel instantiations (see mcmcn_registry.h).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<6>, 6, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<6>, 6, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<7>, 7, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<7>, 7, 64, double, 2),
};
const KernelSet* sets_linreg_e(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
