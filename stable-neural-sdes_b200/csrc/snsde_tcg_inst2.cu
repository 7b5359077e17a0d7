// Explicit instantiations of the general tcgen05 kernel (split over units so they compile in parallel).
#include "snsde_tcg_kernel.cuh"
namespace snsde {
template cudaError_t tcg_launch<8, 2, 2, 0>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<8, 2, 2, 1>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 2, 2, 0>(const TcgParams&, int, size_t, cudaStream_t);
template cudaError_t tcg_launch<16, 2, 2, 1>(const TcgParams&, int, size_t, cudaStream_t);
}  // namespace snsde
