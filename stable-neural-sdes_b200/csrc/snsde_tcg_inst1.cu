// Explicit instantiations of the general tcgen05 kernel (split over units so they compile in parallel).
// One accumulator chain per product (CH = 1): see AccRegion in snsde_tc_common.cuh.
#include "snsde_tcg_kernel.cuh"
namespace snsde {
template cudaError_t tcg_launch<32, 1, 1, 0>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<32, 1, 1, 1>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<32, 1, 1, 2>(const TcgParams&, int, size_t, cudaStream_t);
// M-split (CTA-pair) launches for hidden 129..256
template cudaError_t tcg_launch<8, 1, 1, 0, true>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<8, 1, 1, 1, true>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 1, 1, 0, true>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 1, 1, 1, true>(const TcgParams&, int, size_t, cudaStream_t);
}  // namespace snsde
