// mcmcn_sets_gauss_b.cu -- kernel instantiations (see mcmcn_registry.h): Gaussian distribution with 5..8 parameters.
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<5>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<5>, 0, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<6>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<6>, 0, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<7>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<7>, 0, 64, double, 2),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<8>, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_GAUSSIAN_DISTRIBUTION, GaussDist<8>, 0, 64, double, 2),
};
const KernelSet* sets_gauss_b(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
