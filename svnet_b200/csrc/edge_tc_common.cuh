// Pieces shared by the tensor-core edge kernels (edge_tc.cu: binary linear1 as fp8 UMMAs; edge_fp_tc.cu: full-precision
// linear1 as three-plane bf16 UMMAs): mbarrier / descriptor / tensor-memory wrappers and the vector branch on the
// per-point float4 table.
#pragma once
#include "common.cuh"
#include "edge_vector.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)      // bounded: a protocol mistake traps
{
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
        if (++spins > (1 << 24)) __trap();
    }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// ---- vector branch on the float4 table: w_e = P_j + (Q_i - P_i), VectorBN, gate, mean over the edges (see
// edge_vector.cuh for the arithmetic; here every gather is one 16-byte load and VB rows are in flight per round).
// The first round's gathers are issued by prefetch() before the tile barrier, so that their latency hides
// behind the barrier and the MMA issue.
template <typename S, int CVO>
struct VBranch {
    static constexpr int VB = S::BT;
    static constexpr int FULL = CVO / 32, R = CVO % 32, G = R > 0 ? 32 / R : 1, NPASS = FULL + (R > 0 ? 1 : 0);
    struct Lane {
        bool rem, active;
        int graw, g, c, ng, cc;
    };
    static __device__ __forceinline__ Lane lane_of(int pass, int lane)
    {
        Lane L;
        L.rem = pass == FULL;
        L.graw = L.rem ? lane / (R > 0 ? R : 1) : 0;
        L.active = !L.rem || L.graw < G;
        L.g = L.active ? L.graw : 0;            // idle lanes walk the same rounds (the index shuffles are warp-wide)
        L.c = L.rem ? FULL * 32 + lane % (R > 0 ? R : 1) : pass * 32 + lane;
        L.ng = L.rem ? G : 1;
        L.cc = L.active ? L.c : 0;
        return L;
    }
    float4 w0[VB], pi[NPASS], qi[NPASS];
    float a2[NPASS], c2[NPASS], gt[NPASS];

    __device__ __forceinline__ void init(const svnet_edge_params& p, int lane)
    {
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
            const Lane L = lane_of(pass, lane);
            a2[pass] = __ldg(p.bn2_a + L.cc);
            c2[pass] = __ldg(p.bn2_c + L.cc);
        }
    }

    __device__ __forceinline__ void load_round(float4 (&w)[VB], const float4* pcol, int my_j, int e0, int ng) const
    {
#pragma unroll
        for (int i = 0; i < VB; ++i) {
            const int e = e0 + i * ng;
            const unsigned j = (unsigned)__shfl_sync(SV_FULL, my_j, e < S::EPW ? e : 0);
            w[i] = __ldg(pcol + (size_t)j * S::NC);
        }
    }
    __device__ __forceinline__ void prefetch(const svnet_edge_params& p, int b, const float4* tabc, const float4* trow, int my_j, int lane)
    {
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
            const Lane L = lane_of(pass, lane);
            pi[pass] = __ldg(trow + L.cc);
            qi[pass] = __ldg(trow + S::TQ0 + L.cc);
            gt[pass] = __ldg(p.gate + (long)b * CVO + L.cc);
        }
        const Lane L = lane_of(0, lane);
        load_round(w0, tabc + L.cc, my_j, L.g, L.ng);
    }
    __device__ __forceinline__ void run(const svnet_edge_params& p, long r, int b, const float4* tabc, const float4* trow, int my_j,
                                        int ktot, int lane, float* partial) const
    {
        const float inv_k = 1.0f / (float)ktot;
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
            const Lane L = lane_of(pass, lane);
            float sum[3] = {0.0f, 0.0f, 0.0f};
            const float d0 = qi[pass].x - pi[pass].x, d1 = qi[pass].y - pi[pass].y, d2 = qi[pass].z - pi[pass].z;
            const float a2p = a2[pass], c2p = c2[pass];
            const float4* pcol = tabc + L.cc;
            auto consume = [&](const float4 (&w)[VB], int e0) {
#pragma unroll
                for (int i = 0; i < VB; ++i) {
                    if (e0 + i * L.ng < S::EPW) {
                        const float w0_ = w[i].x + d0, w1 = w[i].y + d1, w2 = w[i].z + d2;
                        const float s2 = fmaf(w2, w2, fmaf(w1, w1, w0_ * w0_));
                        const float t = fmaf(c2p, fast_rcp(fast_sqrt(s2) + 1e-6f), a2p);      // (n a2 + c2) / n,  n = |w| + 1e-6
                        sum[0] = fmaf(w0_, t, sum[0]);
                        sum[1] = fmaf(w1, t, sum[1]);
                        sum[2] = fmaf(w2, t, sum[2]);
                    }
                }
            };
            int e0 = L.g;
            if (pass == 0) {
                consume(w0, e0);
                e0 += L.ng * VB;
            }
#pragma unroll 1
            for (; e0 < S::EPW; e0 += L.ng * VB) {
                float4 w[VB];
                load_round(w, pcol, my_j, e0, L.ng);
                consume(w, e0);
            }
            if (L.rem && G > 1) {
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    float tot = sum[x];
#pragma unroll
                    for (int gg = 1; gg < G; ++gg) tot += __shfl_sync(SV_FULL, sum[x], (lane % (R > 0 ? R : 1)) + gg * R);
                    sum[x] = tot;
                }
            }
            if (L.active && L.graw == 0) {
                if (partial) {
#pragma unroll
                    for (int x = 0; x < 3; ++x) partial[x * CVO + L.c] = sum[x];
                } else {
                    const float g = gt[pass] * inv_k;
#pragma unroll
                    for (int x = 0; x < 3; ++x) p.out.v[r * p.out.ldv + x * p.out.xs + L.c] = sum[x] * g;
                }
            }
        }
    }
};


}  // namespace
