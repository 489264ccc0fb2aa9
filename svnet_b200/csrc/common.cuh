// Shared helpers for the svnet_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/svnet_b200.h"

// ---- error plumbing (thread-local message, see svnet_last_error) ---------------------------------
void svnet_set_error(const char* fmt, ...);

#define SV_REQUIRE(cond, ...)                    \
    do {                                         \
        if (!(cond)) {                           \
            svnet_set_error(__VA_ARGS__);        \
            return SVNET_ERR_ARG;                \
        }                                        \
    } while (0)

#define SV_CHECK_LAUNCH(name)                                                          \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            svnet_set_error("%s: CUDA error %s", name, cudaGetErrorString(e__));       \
            return SVNET_ERR_CUDA;                                                     \
        }                                                                              \
    } while (0)

#define SV_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            svnet_set_error("%s: CUDA error %s", #call, cudaGetErrorString(e__));      \
            return SVNET_ERR_CUDA;                                                     \
        }                                                                              \
    } while (0)

static inline cudaStream_t sv_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int sv_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- device helpers -------------------------------------------------------------------------------
#define SV_FULL 0xffffffffu

// Warp index / a warp-invariant value in a form the compiler's divergence analysis recognises as warp-uniform (a
// broadcast from lane 0).  Branches and loops on `threadIdx.x >> 5` or on a value every lane loaded from the same
// address are convergent, but ptxas cannot know that and wraps every *_sync collective inside them as
// WARPSYNC + op + ENDCOLLECTIVE on sm_100a; on a broadcast value the collectives compile to one instruction.
__device__ __forceinline__ int sv_warp_id() { return __shfl_sync(SV_FULL, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ int sv_uniform(int v) { return __shfl_sync(SV_FULL, v, 0); }

// k-th channel of the kNN feature vector [s | v(x=0) | v(x=1) | v(x=2)] of row r (sv_util.py:100)
__device__ __forceinline__ float sv_feat(const svnet_view& in, long r, int c)
{
    if (c < in.Cs) return __ldg(in.s + r * in.lds + c);
    c -= in.Cs;
    int x = c / in.Cv;
    int d = c - x * in.Cv;
    return __ldg(in.v + r * in.ldv + x * in.xs + d);
}

__device__ __forceinline__ float sv_act(float t, int act)
{
    if (act == SVNET_ACT_LEAKY) return t > 0.0f ? t : __fmul_rn(0.2f, t);
    if (act == SVNET_ACT_RELU) return t > 0.0f ? t : 0.0f;
    return t;
}

__device__ __forceinline__ float sv_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
