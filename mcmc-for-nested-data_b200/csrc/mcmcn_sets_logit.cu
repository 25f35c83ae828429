// mcmcn_sets_logit.cu -- kernel instantiations (see mcmcn_registry.h).
#include <cuda_runtime.h>
#include "mcmcn_registry.h"
namespace mcmcn {
static const KernelSet kSets[] = {
    MCMCN_SET(MCMCN_OBJ_BERNOULLI_LOGIT, Logit, 0, 32, float, 4),
    MCMCN_SET(MCMCN_OBJ_BERNOULLI_LOGIT, Logit, 0, 64, double, 2),
};
const KernelSet* sets_logit(int* n) { *n = (int)(sizeof(kSets) / sizeof(kSets[0])); return kSets; }
}  // namespace mcmcn
