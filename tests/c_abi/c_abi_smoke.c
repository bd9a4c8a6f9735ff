/* The header must be valid C99 and the library must link from a plain C program: this is what a
 * cgo / JNI / FFI binding on the reference side would see.  No device work is done here. */
#include <stdio.h>
#include <string.h>
#include "mvhmr_b200.h"

int main(void)
{
    mvhmr_grid_t grid;
    int rc;
    memset(&grid, 0, sizeof grid);
    if (mvhmr_abi_version() != MVHMR_ABI_VERSION) return 1;
    if (mvhmr_last_error() == NULL) return 2;
    /* argument validation happens before any CUDA call */
    {
        const float *fake = (const float *)(size_t)16;   /* never dereferenced: validation fails first */
        rc = mvhmr_unproject_aggregate(fake, MVHMR_F32, MVHMR_LAYOUT_NCHW, fake, fake, (float *)(size_t)16, 1, 4, 32, 64, 64, 32, 32, 32,
                                       7 /* no such method */, 0, 1, 0, 32768, 0, 32768, 0, NULL, 0, NULL);
    }
    if (rc != MVHMR_ERR_INVALID_ARGUMENT) return 3;
    if (strstr(mvhmr_last_error(), "Unknown aggregation_method") == NULL) return 4;
    rc = mvhmr_soft_argmax3d_grid(NULL, &grid, NULL, 1, 1, 4, 4, 4, 64, NULL, 0, NULL);
    if (rc != MVHMR_ERR_INVALID_ARGUMENT) return 5;
    if (mvhmr_packed_bytes(MVHMR_BF16, 32, 32, 96, 96) != (size_t)32 * 4 * 100 * 100 * 16) return 6;
    if (mvhmr_unproject_backward_workspace_bytes(MVHMR_F32, 1, 1, 4, 2, 2, MVHMR_SUM) != (size_t)6 * 6 * 4 * 4) return 7;
    printf("c abi ok, version %d\n", mvhmr_abi_version());
    return 0;
}
