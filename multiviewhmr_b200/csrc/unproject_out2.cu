// Output format 2 of the L1-gather fused kernel (0: (B,C,N) as the reference, 1: channels-last-3D,
// 2: fused max_pool3d(2)) — see unproject_kernel.cuh.
#include "unproject_kernel.cuh"

namespace mvhmr {
int launch_unproject_gather_out2(const UnprojParams &p, bool bf, int method, unsigned grid, size_t smem, void *stream)
{
    return launch_unproject_gather<2>(p, bf, method, dim3(grid), smem, (cudaStream_t)stream);
}
}  // namespace mvhmr
