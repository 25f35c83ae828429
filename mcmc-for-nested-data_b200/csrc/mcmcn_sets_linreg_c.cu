// mcmcn_sets_linreg_c.cu -- kernel instantiations (see mcmcn_registry.h).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<8>, 8, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<8>, 8, 64, double, 2),
};
const KernelSet* sets_linreg_c(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
