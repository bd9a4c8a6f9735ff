// Output format 3 of the L1-gather fused kernel: the aggregate (or no volume at all) plus the online-softmax
// records of the fused 3-D soft-argmax — see unproject_kernel.cuh.
#include "unproject_kernel.cuh"

namespace mvhmr {
int launch_unproject_gather_out3(const UnprojParams &p, bool bf, int method, unsigned grid, size_t smem, void *stream)
{
    return launch_unproject_gather<3>(p, bf, method, dim3(grid), smem, (cudaStream_t)stream);
}
}  // namespace mvhmr
