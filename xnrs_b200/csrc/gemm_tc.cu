// tcgen05 (5th-gen tensor core) GEMM path.  Placeholder dispatcher until the TMA/TMEM kernel lands:
// reports "not taken" so xnrs_gemm falls through to the exact-fp32 SIMT kernel (still CUDA, never CPU).
#include "gemm.cuh"

namespace xnrs {
int gemm_tensorcore(const GemmArgs &, int, cudaStream_t, int *) { return 0; }
}  // namespace xnrs
