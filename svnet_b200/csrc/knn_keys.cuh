// Ordering keys and warp-level sorted lists shared by the kNN kernels (knn.cu, knn_tc.cu).
#pragma once
#include "common.cuh"

namespace svknn {

// Ordering keys: (score desc, index asc) as one unsigned 64-bit compare.
//   hi = order-preserving map of the fp32 score (-0 is folded into +0 first, so that float equality
//        and key equality agree), lo = ~index (smaller index -> larger key).  Empty slots are key 0,
//        which ranks below every real candidate (even -inf).
typedef unsigned long long kkey_t;

__device__ __forceinline__ kkey_t make_key(float p, int j)
{
    const unsigned f = __float_as_uint(p + 0.0f);
    const unsigned hi = (f & 0x80000000u) ? ~f : (f | 0x80000000u);
    return ((kkey_t)hi << 32) | (unsigned)(~j);
}
__device__ __forceinline__ float key_score(kkey_t k)
{
    const unsigned hi = (unsigned)(k >> 32);
    if (hi == 0u) return -INFINITY;
    return __uint_as_float((hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi);
}
__device__ __forceinline__ int key_index(kkey_t k) { return (int)(~(unsigned)k); }

__device__ __forceinline__ kkey_t shfl_key(kkey_t k, int src)
{
    const unsigned lo = __shfl_sync(SV_FULL, (unsigned)k, src);
    const unsigned hi = __shfl_sync(SV_FULL, (unsigned)(k >> 32), src);
    return ((kkey_t)hi << 32) | lo;
}
__device__ __forceinline__ kkey_t shfl_up_key(kkey_t k)
{
    const unsigned lo = __shfl_up_sync(SV_FULL, (unsigned)k, 1);
    const unsigned hi = __shfl_up_sync(SV_FULL, (unsigned)(k >> 32), 1);
    return ((kkey_t)hi << 32) | lo;
}
__device__ __forceinline__ kkey_t shfl_xor_key(kkey_t k, int m)
{
    const unsigned lo = __shfl_xor_sync(SV_FULL, (unsigned)k, m);
    const unsigned hi = __shfl_xor_sync(SV_FULL, (unsigned)(k >> 32), m);
    return ((kkey_t)hi << 32) | lo;
}

template <int R>
struct TopK {
    kkey_t k[R];   // sorted, best (largest key) at position 0; position = r*32 + lane
};

template <int R>
__device__ __forceinline__ void topk_insert(TopK<R>& L, kkey_t c, int lane)
{
    int P = 0;  // number of entries that rank before the candidate
#pragma unroll
    for (int r = 0; r < R; ++r) P += __popc(__ballot_sync(SV_FULL, L.k[r] > c));
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        kkey_t up = shfl_up_key(L.k[r]);
        if (r > 0) {
            const kkey_t prev = shfl_key(L.k[r - 1], 31);
            if (lane == 0) up = prev;
        }
        const int pos = r * 32 + lane;
        if (pos > P) L.k[r] = up;
        else if (pos == P) L.k[r] = c;
    }
}


}  // namespace svknn
