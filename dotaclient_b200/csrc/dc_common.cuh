// Shared helpers for the dotaclient_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dotaclient_b200.h"

// Records a message retrievable through dc_last_error() (thread-local).
void dc_set_error(const char *fmt, ...);

#define DC_REQUIRE(cond, code, ...)        \
    do {                                   \
        if (!(cond)) {                     \
            dc_set_error(__VA_ARGS__);     \
            return (code);                 \
        }                                  \
    } while (0)

#define DC_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            dc_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                         cudaGetErrorString(e__));                                         \
            return (int)e__;                                                               \
        }                                                                                  \
    } while (0)

#define DC_LAUNCH_OK() DC_CUDA(cudaGetLastError())

static inline cudaStream_t dc_cu_stream(dc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached per device; 148 on B200).
int dc_sm_count();

// csrc/gemm_tf32x3.cu: split-K tcgen05 3xTF32 GEMM writing `ksplit` partial [M, N] products (library-internal).
int dc_gemm_tf32x3_splitk(const float *A, int lda, const float *B, int ldb, float *part, int64_t M, int N, int K, int ksplit,
                          bool first_call, cudaStream_t st);

// csrc/encoder.cu: dW_b / db_b (+)= the fixed-order sum of `nblocks` [128][13] partials (library-internal).
int dc_unit_basic_reduce(const float *partial, int nblocks, float *dw_b, float *db_b, int accumulate, cudaStream_t st);

__device__ __forceinline__ float dc_sigmoid(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh via one exp; abs error ~1e-7, saturates cleanly for |x| large.
__device__ __forceinline__ float dc_tanh(float x) { return 1.0f - 2.0f / (__expf(2.0f * x) + 1.0f); }

__device__ __forceinline__ float dc_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double dc_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
