// mcmcn_sets_linreg_f.cu -- kernel instantiations (see mcmcn_registry.h): K = 9, 10 coefficients.
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<9>, 9, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<9>, 9, 64, double, 1),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<10>, 10, 32, float, 2),
    MCMCN_SET(MCMCN_OBJ_LINEAR_REGRESSION, LinReg<10>, 10, 64, double, 1),
};
const KernelSet* sets_linreg_f(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
