// mcmcn_sets_gauss.cu -- kernel instantiations (see mcmcn_registry.h).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<1>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<1>, 0, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<2>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<2>, 0, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<3>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<3>, 0, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<4>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<4>, 0, 64, double, 2),
};
const KernelSet* sets_gauss(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
