// Vector branch of the fused SV edge convolution, shared by the specialised edge kernels
// (edge_fast.cu: XNOR/popcount and fp32 scalar branches; edge_tc.cu: tcgen05 scalar branch).
#pragma once
#include "common.cuh"

namespace {

// ---- P5: vector branch from the per-point P|Q table: w_e = (P_j - P_i) + Q_i, VectorBN, gate, mean over
// the edges.  Channels beyond a multiple of 32 are spread over lane groups (group g takes the edges
// e = g, g+G, ...; partial sums meet by shuffle), so CVO = 10 keeps 30 lanes busy instead of 10.
// Tolerance-level arithmetic (norms and mean pools are not bit-pinned): rsqrt /
// fast division instead of the IEEE sequences.
__device__ __forceinline__ float fast_sqrt(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int CVO, int VB = 0>
__device__ __forceinline__ void vector_branch(const svnet_edge_params& p, long r, int b, long cbase, const int* nidx, int k, int lane,
                                              float* partial = nullptr)
{
    // `partial` ([3][CVO] floats, shared memory): this warp holds only `k` of the point's edges -- the raw sums
    // go there and the caller combines them (vector_branch_combine); otherwise k is the point's edge count
    constexpr int LDP = 2 * CVO;
    const float inv_k = 1.0f / (float)k;
    constexpr int FULL = CVO / 32;                 // passes with 32 channels, one edge group
    constexpr int R = CVO % 32;                    // remainder channels
    constexpr int G = R > 0 ? 32 / R : 1;          // edge groups of the remainder pass
#pragma unroll
    for (int pass = 0; pass < FULL + (R > 0 ? 1 : 0); ++pass) {
        const bool rem = pass == FULL;
        const int g = rem ? lane / R : 0;
        const int c = rem ? FULL * 32 + lane % (R > 0 ? R : 1) : pass * 32 + lane;
        const int ng = rem ? G : 1;
        const bool active = !rem || g < G;
        float sum[3] = {0.0f, 0.0f, 0.0f};
        if (active) {
            const float* pi = p.PQ + r * 3 * LDP + c;
            const float* pq0 = p.PQ + cbase * 3 * LDP + c;      // row offsets inside a cloud fit 32 bits (launch guard)
            const float d_i[3] = {__ldg(pi + CVO) - __ldg(pi), __ldg(pi + LDP + CVO) - __ldg(pi + LDP),
                                  __ldg(pi + 2 * LDP + CVO) - __ldg(pi + 2 * LDP)};      // Q_i - P_i
            const float a2 = __ldg(p.bn2_a + c), c2 = __ldg(p.bn2_c + c);
            if constexpr (VB > 0) {
                // VB neighbour rows in flight per round (the gathers are L2 latency: one exposure per round)
                for (int e0 = g; e0 < k; e0 += ng * VB) {
                    float w[VB][3];
#pragma unroll
                    for (int i = 0; i < VB; ++i) {
                        const int e = e0 + i * ng;
                        const float* pj = pq0 + (unsigned)nidx[e < k ? e : 0] * (unsigned)(3 * LDP);
#pragma unroll
                        for (int x = 0; x < 3; ++x) w[i][x] = __ldg(pj + x * LDP);
                    }
#pragma unroll
                    for (int i = 0; i < VB; ++i) {
                        if (e0 + i * ng < k) {
                            const float w0 = w[i][0] + d_i[0], w1 = w[i][1] + d_i[1], w2 = w[i][2] + d_i[2];
                            const float s2 = fmaf(w2, w2, fmaf(w1, w1, w0 * w0));
                            const float t = fmaf(c2, fast_rcp(fast_sqrt(s2) + 1e-6f), a2);
                            sum[0] = fmaf(w0, t, sum[0]);
                            sum[1] = fmaf(w1, t, sum[1]);
                            sum[2] = fmaf(w2, t, sum[2]);
                        }
                    }
                }
            } else {
#pragma unroll 4
            for (int e = g; e < k; e += ng) {
                const float* pj = pq0 + (unsigned)nidx[e] * (unsigned)(3 * LDP);
                const float w0 = __ldg(pj) + d_i[0], w1 = __ldg(pj + LDP) + d_i[1], w2 = __ldg(pj + 2 * LDP) + d_i[2];
                const float s2 = fmaf(w2, w2, fmaf(w1, w1, w0 * w0));
                const float t = fmaf(c2, fast_rcp(fast_sqrt(s2) + 1e-6f), a2);      // (n a2 + c2) / n,  n = |w| + 1e-6
                sum[0] = fmaf(w0, t, sum[0]);
                sum[1] = fmaf(w1, t, sum[1]);
                sum[2] = fmaf(w2, t, sum[2]);
            }
            }
        }
        if (rem && G > 1) {
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                float tot = sum[x];
#pragma unroll
                for (int gg = 1; gg < G; ++gg) tot += __shfl_sync(SV_FULL, sum[x], (lane % (R > 0 ? R : 1)) + gg * R);
                sum[x] = tot;
            }
        }
        if (active && g == 0) {
            if (partial) {
#pragma unroll
                for (int x = 0; x < 3; ++x) partial[x * CVO + c] = sum[x];
            } else {
                const float gt = p.gate[(long)b * CVO + c] * inv_k;
#pragma unroll
                for (int x = 0; x < 3; ++x) p.out.v[r * p.out.ldv + x * p.out.xs + c] = sum[x] * gt;
            }
        }
    }
}

// sums of NW warps' partial vector sums -> gated mean over all ktot edges of the point
template <int CVO, int NW>
__device__ __forceinline__ void vector_branch_combine(const svnet_edge_params& p, long r, int b, const float* partial, int pstride,
                                                      int ktot, int lane)
{
    const float inv_k = 1.0f / (float)ktot;
    for (int i = lane; i < 3 * CVO; i += 32) {
        const int x = i / CVO, c = i - x * CVO;
        float t = partial[i];
#pragma unroll
        for (int w = 1; w < NW; ++w) t += partial[w * pstride + i];
        p.out.v[r * p.out.ldv + x * p.out.xs + c] = t * (p.gate[(long)b * CVO + c] * inv_k);
    }
}

}  // namespace
