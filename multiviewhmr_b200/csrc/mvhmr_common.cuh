// Shared helpers of the sm_100a kernels: error reporting for the C ABI and the
// IEEE-exact scalar building blocks that reproduce the reference's fp32
// rounding order (SURVEY.md §7 hard parts 1-2).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/mvhmr_b200.h"

namespace mvhmr {

// thread-local last-error buffer (defined in abi.cu)
char *err_buf();
constexpr int kErrBufLen = 512;

inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), kErrBufLen, fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(MVHMR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return MVHMR_OK;
}

// [X Y Z 1] · P[r,:] as the reference's sgemm evaluates it: one rounded product
// followed by three FMAs in k order (utils/multiview.py:105).
__device__ __forceinline__ float proj_row(float X, float Y, float Z, float p0, float p1, float p2, float p3) {
    float acc = __fmul_rn(X, p0);
    acc = __fmaf_rn(Y, p1, acc);
    acc = __fmaf_rn(Z, p2, acc);
    return __fadd_rn(acc, p3);          // fma(1, p3, acc) == acc + p3
}

// rot[i,:] · d for the K=3 sgemm of utils/volumetric.py:110.
__device__ __forceinline__ float rot_row(float r0, float r1, float r2, float d0, float d1, float d2) {
    float acc = __fmul_rn(r0, d0);
    acc = __fmaf_rn(r1, d1, acc);
    return __fmaf_rn(r2, d2, acc);
}

}  // namespace mvhmr
