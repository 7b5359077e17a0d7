// Explicit instantiations of the general tcgen05 kernel (split over units so they compile in parallel).
// One accumulator chain per product (CH = 1): see AccRegion in snsde_tc_common.cuh.
#include "snsde_tcg_kernel.cuh"
namespace snsde {
template cudaError_t tcg_launch<8, 1, 2, 0>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<8, 1, 2, 1>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 1, 2, 0>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 1, 2, 1>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<8, 1, 2, 2>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 1, 2, 2>(const TcgParams&, int, size_t, cudaStream_t);
}  // namespace snsde
