// tcgen05 / TMEM path for the binary-weight vector linear of the per-point SVBlocks (conv5:
// 98 304 x 83 x 170 per batch) with the VectorBN + gate epilogue
// (reference models/sv_layers.py:43-49,94-100,192-194).
//
// Formulation (transposed so that the epilogue is thread-local and stores are coalesced):
//     D[c][x*32 + p] = sum_k  Wsign[c][k] * v[p][x][k]          M = channels, N = 3 x 32 points
//   * "A" operand  = sign(W) as bf16 +-1, [M_TILES*128 rows][Kpad], staged once per CTA; full-precision weights
//                    (fp SV models, `sign_w` == 0) as three bf16 planes hi+mid+lo like the activations, and the six
//                    plane products down to 2^-16 |a||w| (h*l, l*h, m*m, h*m, m*h, h*h), small terms first
//   * "B" operand  = activations, split EXACTLY into three bf16 planes hi+mid+lo (8+8+8 mantissa
//                    bits), [96 rows][Kpad] per plane; every plane*weight product is exact, the
//                    tensor core only reorders the fp32 summation
//   * accumulators = fp32 in tensor memory: one 128 x 96 tile per 128 channels
//   * one elected thread issues  tcgen05.mma.cta_group::1.kind::f16  (M=128, N=96, K=16) and commits
//     to an mbarrier; the epilogue reads TMEM with tcgen05.ld (lane = channel), so a thread holds
//     (x, y, z) of the same (point, channel): VectorBN + gate, then stores that are contiguous across
//     the warp (32 consecutive channels).
// Shared-memory operands use the canonical K-major no-swizzle layout of the UMMA descriptors:
// 8 rows x 16 bytes core matrices, rows 16 B apart, 8-row groups 128 B apart (SBO), the next 8
// K-elements KBS bytes further (LBO).
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int NTH = 256;
constexpr int PTS = 32;              // points per tile (two CTAs per SM overlap staging / epilogue)
constexpr int NCOL = 3 * PTS;        // 96 accumulator columns per 128-channel tile
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;    // descriptor version for sm_100
    return d;                  // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// smem: Ws [KB][MT*128 rows][8] bf16 | Bs [3 planes][KB][96 rows (+1 pad)][8] bf16 | mbarrier | tmem base
template <int WP>
__global__ void __launch_bounds__(NTH, WP == 1 ? 2 : 1) vlinear_tcgen05_kernel(svnet_gemm_params p, int Kpad, int MT, int ntiles,
                                                                              int* __restrict__ err_flag)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int KB = Kpad / 8;                       // 8-element (16 B) k-blocks
    const int WROWS = MT * 128;
    const uint32_t KBS_W = WROWS * 16;             // bytes between k-blocks of the weight operand
    const uint32_t KBS_B = NCOL * 16 + 16;         // bytes between k-blocks of one activation plane (+16: bank spread)
    unsigned char* Ws = smraw;                     // WP weight planes
    const size_t wplane_bytes = (size_t)KB * KBS_W;
    unsigned char* Bs = Ws + (size_t)WP * wplane_bytes;
    const size_t plane_bytes = (size_t)KB * KBS_B;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Bs + 3 * plane_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = sv_warp_id();

    // ---- one-time: weights -> bf16 (+-1, or three exact planes) in the canonical layout; barrier; tensor memory ----
    for (int i = tid; i < WROWS * KB; i += NTH) {
        const int row = i % WROWS, kb = i / WROWS;
        uint32_t w4[WP][4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint32_t pl[WP];
#pragma unroll
            for (int q = 0; q < WP; ++q) pl[q] = 0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = kb * 8 + h * 2 + e;
                float w = 0.0f;
                if (row < p.N && k < p.K) w = __ldg(p.W + (long)row * p.ldw + k);
                if (WP == 1) {
                    const unsigned short v = (w > 0.0f) ? 0x3F80 : ((w < 0.0f) ? 0xBF80 : 0);
                    pl[0] |= (uint32_t)v << (16 * e);
                } else {
                    const uint32_t hb = __float_as_uint(w) & 0xFFFF0000u;
                    const float r1 = w - __uint_as_float(hb);
                    const uint32_t mb = __float_as_uint(r1) & 0xFFFF0000u;
                    const float r2 = r1 - __uint_as_float(mb);
                    const uint32_t lb = __float_as_uint(r2) & 0xFFFF0000u;
                    pl[0] |= (hb >> 16) << (16 * e);
                    pl[WP > 1 ? 1 : 0] |= (mb >> 16) << (16 * e);
                    pl[WP > 2 ? 2 : 0] |= (lb >> 16) << (16 * e);
                }
            }
#pragma unroll
            for (int q = 0; q < WP; ++q) w4[q][h] = pl[q];
        }
#pragma unroll
        for (int q = 0; q < WP; ++q)
            *reinterpret_cast<uint4*>(Ws + q * wplane_bytes + (size_t)kb * KBS_W + (size_t)row * 16) =
                make_uint4(w4[q][0], w4[q][1], w4[q][2], w4[q][3]);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptor: D fp32, A/B bf16, both K-major, N = NCOL, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((128u >> 4) << 24);
    const long npoints = p.M / 3;
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long p0 = (long)tile * PTS;
        // ---- stage the activations: thread -> (row = x*32 + p, k-block); exact 3-way bf16 split ----
        for (int i = tid; i < NCOL * KB; i += NTH) {
            const int row = i / KB, kb = i - row * KB;     // consecutive threads: consecutive 32 B of one row
            const int x = row / PTS, pi = row - x * PTS;
            const long pnt = p0 + pi;
            const float* src = p.A + pnt * p.lda_g + (long)x * p.lda_x + kb * 8;
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = (pnt < npoints && kb * 8 + e < p.K) ? __ldg(src + e) : 0.0f;
            uint32_t hw[4], mw[4], lw[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                uint32_t hh = 0, mm = 0, ll = 0;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float av = a[h * 2 + e];
                    const uint32_t hb = __float_as_uint(av) & 0xFFFF0000u;
                    const float r1 = av - __uint_as_float(hb);
                    const uint32_t mb = __float_as_uint(r1) & 0xFFFF0000u;
                    const float r2 = r1 - __uint_as_float(mb);
                    const uint32_t lb = __float_as_uint(r2) & 0xFFFF0000u;
                    hh |= (hb >> 16) << (16 * e);
                    mm |= (mb >> 16) << (16 * e);
                    ll |= (lb >> 16) << (16 * e);
                }
                hw[h] = hh; mw[h] = mm; lw[h] = ll;
            }
            unsigned char* dst = Bs + (size_t)kb * KBS_B + (size_t)row * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(dst + plane_bytes) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            *reinterpret_cast<uint4*>(dst + 2 * plane_bytes) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> tensor-core reads
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

        // ---- one thread issues all MMAs of the tile, then commits to the mbarrier ----
        if (tid == 0) {
            // (weight plane, activation plane) products, small terms first; sign weights: one weight plane
            constexpr int NPROD = WP == 1 ? 3 : 6;
            const int apl1[3] = {2, 1, 0};
            const int wpl3[6] = {0, 2, 1, 0, 1, 0}, apl3[6] = {2, 0, 1, 1, 0, 0};      // h*l, l*h, m*m, h*m, m*h, h*h
            for (int mt = 0; mt < MT; ++mt) {
                uint32_t first = 1;
                for (int pr = 0; pr < NPROD; ++pr) {
                    const int wq = WP == 1 ? 0 : wpl3[pr], aq = WP == 1 ? apl1[pr] : apl3[pr];
                    for (int ks = 0; ks < Kpad / 16; ++ks) {
                        const uint64_t adesc = make_desc(smem_u32(Ws) + (uint32_t)(wq * wplane_bytes) + (uint32_t)(2 * ks) * KBS_W +
                                                             (uint32_t)mt * 128 * 16, KBS_W, 128);
                        const uint64_t bdesc = make_desc(smem_u32(Bs) + (uint32_t)(aq * plane_bytes) + (uint32_t)(2 * ks) * KBS_B,
                                                         KBS_B, 128);
                        umma_bf16(tmem_base + (uint32_t)(mt * NCOL), adesc, bdesc, idesc, first ? 0u : 1u);
                        first = 0;
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
        }
        // ---- wait for the accumulators (bounded spin: a mistake must not hang the GPU) ----
        {
            uint32_t done = 0;
            long spins = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                    : "=r"(done)
                    : "r"(smem_u32(mbar)), "r"(phase)
                    : "memory");
                if (++spins > (1L << 26)) { if (err_flag) atomicExch(err_flag, 1); __trap(); }   // fail loudly, never hang
            }
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

        // ---- epilogue: warp w reads TMEM lanes 32*(w%4).. of channel tile w/4; lane = channel ----
        const int mt = warp >> 2;
        if (mt < MT) {
            const int c = mt * 128 + (warp & 3) * 32 + lane;
            const bool cok = c < p.N;
            const float cs = (cok && p.colscale) ? p.colscale[c] : 1.0f;
            const float a2 = (cok && p.vbn) ? p.bn_a[c] : 0.0f, c2 = (cok && p.vbn) ? p.bn_c[c] : 0.0f;
            const uint32_t tbase = tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(mt * NCOL);
            const unsigned gpc = (unsigned)p.groups_per_cloud;
            // the gate depends on (cloud, channel) only: one load per tile when the tile lies inside one cloud
            // (always, when the rows per cloud are a multiple of the 32-point tile)
            const long last_pt = (p0 + PTS - 1 < npoints ? p0 + PTS - 1 : npoints - 1);
            const bool one_cloud = ((unsigned)p0 / gpc) == ((unsigned)last_pt / gpc);
            const float gate_tile = (p.gate && cok && one_cloud) ? __ldg(p.gate + (long)((unsigned)p0 / gpc) * p.N + c) : 1.0f;
            for (int q0 = 0; q0 < PTS; q0 += 16) {
                float vx[16], vy[16], vz[16];
                tmem_ld16(tbase + q0, vx);
                tmem_ld16(tbase + PTS + q0, vy);
                tmem_ld16(tbase + 2 * PTS + q0, vz);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const long pnt = p0 + q0 + e;
                    if (!cok || pnt >= npoints) continue;
                    const float w0 = __fmul_rn(vx[e], cs), w1 = __fmul_rn(vy[e], cs), w2 = __fmul_rn(vz[e], cs);
                    if (!p.vbn) {   // plain table (P|Q of the fused edge kernel)
                        if (p.c4) {   // one (x, y, z, 0) column per (point, channel): 512 contiguous bytes per warp
                            *reinterpret_cast<float4*>(p.C + pnt * p.ldc_g + 4 * c) = make_float4(w0, w1, w2, 0.0f);
                            continue;
                        }
                        float* cq = p.C + pnt * p.ldc_g + c;
                        cq[0] = w0; cq[p.ldc_x] = w1; cq[2L * p.ldc_x] = w2;
                        continue;
                    }
                    // VectorBN: v * (bn(n) / n), n = |v| + 1e-6 (sv_layers.py:94-100).  Tolerance-level arithmetic
                    // (norms are not bit-pinned): rsqrt / fast division instead of the IEEE sequences.
                    const float s2 = fmaf(w2, w2, fmaf(w1, w1, w0 * w0));
                    const float nrm = (s2 > 0.0f ? s2 * rsqrtf(s2) : 0.0f) + 1e-6f;
                    float sfac = a2 + __fdividef(c2, nrm);
                    if (p.gate) sfac *= one_cloud ? gate_tile : __ldg(p.gate + (long)((unsigned)pnt / gpc) * p.N + c);
                    float* cp = p.C + pnt * p.ldc_g + c;
                    cp[0] = w0 * sfac;
                    cp[p.ldc_x] = w1 * sfac;
                    cp[2L * p.ldc_x] = w2 * sfac;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();   // accumulators and activation planes may be overwritten by the next tile
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
}

}  // namespace

// Returns 1 if handled, 0 if the caller should use another kernel, < 0 on error.
int svnet_vlinear_tcgen05_dispatch(const svnet_gemm_params* p, cudaStream_t st)
{
    const char* on = getenv("SVNET_TCGEN05");
    if (on && on[0] == '0') return 0;                         // SVNET_TCGEN05=0 falls back to mma.sync / CUDA cores
    if (p->G != 3 || p->M % 3 != 0) return 0;
    if (p->K > 96 || p->N > 256 || p->bias || p->act != SVNET_ACT_NONE || (!p->vbn && p->bn_a)) return 0;
    if (p->c4 && (p->vbn || (p->ldc_g & 3) || (reinterpret_cast<uintptr_t>(p->C) & 15))) return 0;
    if (!p->sign_w) {
        // full-precision weights: three weight planes.  Only where a caller needs this kernel's epilogues (the float4
        // table layout, VectorBN + gate); plain tables stay on the fp32 chain kernel (their bits are part of the
        // fp models' parity record), SVNET_VLINEAR_FP_TC=0 switches the path off.
        const char* fp = getenv("SVNET_VLINEAR_FP_TC");
        if (fp && fp[0] == '0') return 0;
        if (!p->c4 && !p->vbn) return 0;
    }
    const int WP = p->sign_w ? 1 : 3;
    const int Kpad = (p->K + 15) / 16 * 16;
    const int MT = (p->N + 127) / 128;
    const long npoints = p->M / 3;
    const int ntiles = (int)((npoints + PTS - 1) / PTS);
    const size_t smem = (size_t)WP * (Kpad / 8) * (MT * 128) * 16 + (size_t)3 * (Kpad / 8) * (NCOL * 16 + 16) + 64;
    if (smem > 220 * 1024) return 0;
    const int per_sm = WP == 1 ? 2 : 1;
    const int grid = ntiles < 148 * per_sm ? ntiles : 148 * per_sm;
    if (WP == 1) {
        SV_CUDA(cudaFuncSetAttribute(vlinear_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vlinear_tcgen05_kernel<1><<<grid, NTH, smem, st>>>(*p, Kpad, MT, ntiles, nullptr);
    } else {
        SV_CUDA(cudaFuncSetAttribute(vlinear_tcgen05_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        vlinear_tcgen05_kernel<3><<<grid, NTH, smem, st>>>(*p, Kpad, MT, ntiles, nullptr);
    }
    SV_CHECK_LAUNCH("svnet_linear_rows(tcgen05)");
    return 1;
}
