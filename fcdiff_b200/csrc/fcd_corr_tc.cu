// K1 tensor-core path (tcgen05 + TMEM, split-TF32).  Placeholder until the
// UMMA kernel lands: reports "shape not supported" so fcd_corr_fisherz uses the
// SIMT Gram kernel.
#include "fcd_corr.cuh"

namespace fcd {

int corr_gram_tc(const float*, int, int, int, double*, int64_t, int, int, cudaStream_t) { return 1; }

}  // namespace fcd
