// K1-TC -- kNN graph construction with the pairwise scores on tcgen05 tensor cores.
// Replaces sv_util.knn (reference models/utils/sv_util.py:19-25) for k <= 48, 64 <= N <= 4096, C <= 160;
// other shapes stay on the CUDA-core kernel in knn.cu.  Results are bit-identical to
// oracle/svnet_oracle.c:orc_knn: the tensor cores only FILTER, every decision that the filter
// cannot certify is re-taken with the oracle's exact fp32 fmaf chain.
//
//   pack kernel : features -> three bf16 planes hi/mid/lo (an exact split of the fp32 value) in the
//                 canonical K-major UMMA layout, 128 rows x 16 channels per 12 KB chunk, plus the
//                 exact squared norms xx (chain fmaf, channel ascending).
//   knn kernel  : one CTA = 128 query rows (UMMA M) of one cloud.  Query chunks stay resident in
//                 shared memory; candidate chunks stream through a ring of cp.async.bulk copies;
//                 one thread issues tcgen05.mma (6 plane products per chunk, error ~2^-20 |a||b|)
//                 into a double-buffered 128x128 fp32 accumulator in tensor memory; 8 scanner warps
//                 read their rows with tcgen05.ld (lane = query row) -- no shuffles, no atomics:
//       pass A  : 64 running group maxima per row of a LOWER bound of the score; the k-th largest
//                 group maximum is a lower bound L of the row's k-th best score (k distinct
//                 candidates reach it) and leaves ~k+5..10 survivors.  A lower bound may be loose, so this
//                 pass runs only the three largest plane products (h*h, h*m, m*h: error ~2^-15 |a||b|, its
//                 own eps) and streams only the hi and mid planes -- half the tensor work of pass B;
//       pass B  : candidates whose UPPER bound reaches L are queued (approximate score, index);
//       finish  : a warp sorts a row's survivors by approximate score; neighbours in that order
//                 whose scores differ by more than the error bound are certainly ordered; runs of
//                 closer scores (and exact ties) are re-scored with the exact chain and re-sorted
//                 by (score desc, index asc).  Rows that overflow the queue take a brute-force
//                 exact path.
// Error model: |p_ij - q_ij| <= EPS * (xx_i + xx_j), p = oracle chain score, q = tensor-core score.
// Dropped plane products contribute < 2^-19, fp32 accumulation (tensor core and chain) < 2^-17 of
// |f_i||f_j|; see knn_tc_eps() for the bound that is used.  Scores are handled in half scale:
//   q/2 = dot - (xx_i + xx_j)/2.
#include "common.cuh"
#include "knn_keys.cuh"
#include <stdlib.h>

namespace {
using namespace svknn;

constexpr int TM = 128;                        // query rows per CTA == UMMA M
constexpr int TNB = 256;                       // candidates per block == UMMA N
constexpr int KCH = 16;                        // channels per chunk == UMMA K for bf16
constexpr int A_KB = TM * 16;                  // resident query operand: one 8-channel k-block (descriptor LBO)
constexpr int A_PLANE = 2 * A_KB;
constexpr int A_CHUNK = 3 * A_PLANE;           // hi | mid | lo, 12 KB
constexpr int B_KB = TNB * 16;                 // streamed candidate operand, 256 rows
constexpr int B_PLANE = 2 * B_KB;
constexpr int B_CHUNK = 3 * B_PLANE;           // 24 KB
constexpr int NSCAN = 256;                     // scanner threads: warps 0..7
constexpr int NTHREADS = 320;                  // + warp 8 (MMA issue) + warp 9 (bulk-copy producer)
constexpr int CAPH = 64;                       // survivor queue capacity per (row, column half)
constexpr int GROUPS = 64;
constexpr int GM_BYTES = GROUPS * TM * 4;      // group maxima between the passes
#ifndef FIN_MIN_BLOCKS
#define FIN_MIN_BLOCKS 5
#endif
#ifndef FIN_WIDE_MIN_BLOCKS
#define FIN_WIDE_MIN_BLOCKS 4                  // two entries per lane: 64 registers per thread
#endif
constexpr int FIN_WARPS = 8;                   // finish kernel: warps per CTA, one warp = one row
constexpr int KMAX = 160;                      // padded channels
constexpr int MAX_STAGES = 5;
constexpr int KNN_PASSA3_MAX_N = 1024;         // threshold pass with three plane products up to this cloud size
constexpr int KNN_TC_MAX_K = 48;               // k <= 32: approximate-order finish; 33..48: every survivor re-scored
                                               // (the 64-group threshold gets loose as k approaches 64)

__device__ unsigned long long g_knn_tc_stats[12];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / bulk copy / tcgen05 wrappers -------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded wait: a protocol mistake must trap, never hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;    // descriptor version for sm_100
    return d;                  // no swizzle, K-major canonical layout
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void scan_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(NSCAN) : "memory"); }

// ---- pack: exact 3-way bf16 split into the canonical operand layout + exact norms --------------
// pack[b][cb][kc][plane][kb][row 0..255][8 bf16]; xx[b][cb*256 + row] (+inf for rows >= N)
// One CTA = 32 rows: coalesced row loads (lane = channel) into a shared tile, then thread (row, k-block)
// splits 8 channels and writes one 16-byte piece per plane (32 consecutive rows = 512 contiguous bytes);
// warp 0 also runs the sequential norm chains from the tile.
constexpr int PACK_ROWS = 32;
__global__ void __launch_bounds__(256) knn_pack_kernel(svnet_view in, int N, int NCB, int NKC, unsigned char* __restrict__ pack,
                                                       float* __restrict__ xx, float* __restrict__ xf, int C4)
{
    extern __shared__ float tile[];                 // [32 rows][Kpad + 1]
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * PACK_ROWS;          // first (padded) row of this CTA within the cloud
    const int C = in.Cs + 3 * in.Cv;
    const int Kpad = NKC * KCH, ld = Kpad + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = sv_warp_id();
    // channel -> address pieces of this lane's channels
    for (int c = lane; c < Kpad; c += 32) {
        const float* src = nullptr;
        long stride = 0;
        if (c < in.Cs) { src = in.s + c; stride = in.lds; }
        else if (c < C) {
            const int cc = c - in.Cs;
            const int x = (cc >= in.Cv ? 1 : 0) + (cc >= 2 * in.Cv ? 1 : 0);
            src = in.v + (long)x * in.xs + (cc - x * in.Cv);
            stride = in.ldv;
        }
#pragma unroll
        for (int q = 0; q < PACK_ROWS / 8; ++q) {
            const int rr = warp * (PACK_ROWS / 8) + q;
            const int r = r0 + rr;
            tile[rr * ld + c] = (src && r < N) ? __ldg(src + ((long)b * N + r) * stride) : 0.0f;
        }
    }
    __syncthreads();
    if (warp == 0) {
        const int r = r0 + lane;
        float nrm = 0.0f;          // channel ascending chain, as oracle/svnet_oracle.c:orc_knn (zero padding is exact)
        const float* tr = tile + lane * ld;
#pragma unroll 8
        for (int c = 0; c < C; ++c) nrm = __fmaf_rn(tr[c], tr[c], nrm);
        xx[(size_t)b * NCB * TNB + r] = r < N ? nrm : INFINITY;
    }
    {   // fp32 copy of the rows in chain order, C4 = C rounded up to 4 floats per row (zero padded): the finish
        // kernel's exact chains gather candidate rows from it with 16-byte loads
        const int q4n = C4 >> 2;
        float4* xfd = reinterpret_cast<float4*>(xf + ((size_t)b * NCB * TNB + r0) * C4);
        for (int item = tid; item < PACK_ROWS * q4n; item += 256) {
            const int rr = item / q4n, q = item - rr * q4n;
            const float* a = tile + rr * ld + 4 * q;
            xfd[item] = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
    const int cb = r0 / TNB, rin = r0 % TNB;
    unsigned char* dst0 = pack + ((size_t)(b * NCB + cb) * NKC) * B_CHUNK + (size_t)rin * 16;
    for (int item = tid; item < PACK_ROWS * 2 * NKC; item += 256) {
        const int rr = item & 31, kbg = item >> 5;              // consecutive threads: consecutive rows
        const int kc = kbg >> 1, kb = kbg & 1;
        const float* a = tile + rr * ld + kbg * 8;
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
        unsigned char* d = dst0 + (size_t)kc * B_CHUNK + kb * B_KB + rr * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(d + B_PLANE) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
        *reinterpret_cast<uint4*>(d + 2 * B_PLANE) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

__device__ __forceinline__ float exact_score(float dot, float xxi, float xxj)
{
    const float inner = -2.0f * dot;
    return __fsub_rn(__fsub_rn(-xxj, inner), xxi);   // ((-xx_j) - (-2 dot)) - xx_i, sv_util.py:20-22
}

struct knn_tc_args {
    svnet_view in;
    int N, k, NCB, NKC, stages;
    float eps;                   // |p_ij - q_ij| <= eps * (xx_i + xx_j), six plane products (pass B, finish)
    float epsA;                  // the same bound for the three-product scores of pass A
    int passA3;                  // pass A with three plane products (0: six, tuning aid SVNET_KNN_PASSA=6)
    const unsigned char* pack;
    const float* xx;             // [B][NCB*256]
    const float* xf;             // [B][NCB*256][C4] fp32 rows in chain order (zero padded)
    int C4;
    float2* gq;                  // survivor queues [B*N rows][2 halves][CAPH] of (half-scale score, index bits)
    int* gqcnt;                  // [B*N][2]
    int stats;                   // != 0: accumulate g_knn_tc_stats (same-address atomics: profiling runs only)
    int32_t* idx32;
    int64_t* idx64;
};

__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_kernel(knn_tc_args p)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int NKC = p.NKC, S = p.stages, NB = p.NCB;
    unsigned char* As = smraw;                                        // NKC query chunks (128 rows), resident
    unsigned char* Ring = As + (size_t)NKC * A_CHUNK;                 // S candidate chunks (256 rows)
    float* xxs = reinterpret_cast<float*>(Ring + (size_t)S * B_CHUNK);   // [NCB*256]
    float* gm = xxs + NB * TNB;                                       // [64 groups][128 rows]
    float* thr = gm + GROUPS * TM;                                    // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr + TM);           // barA | full[S] | empty[S] | tfull[2] | tempty[2]
    uint64_t* barA = bars;
    uint64_t* full = bars + 1;
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = sv_warp_id();
    const int rb = blockIdx.x, b = blockIdx.y;
    const int i0 = rb * TM;
    const long base = (long)b * p.N;
    const int k = p.k;

    if (tid == 0) {
        mbar_init(barA, 1);
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(tfull + t, 1); mbar_init(tempty + t, NSCAN / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(2 * TNB));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const unsigned char* cloud_pack = p.pack + (size_t)b * NB * NKC * B_CHUNK;
    const int nblk = 2 * NB;     // pass A then pass B over the candidate blocks

    if (warp == 9) {
        // ================= producer: bulk copies of operand chunks =================
        if (lane == 0) {
            // query rows = one half of a 256-row block of the same pack: 6 pieces of 2 KB per chunk
            mbar_expect_tx(barA, (uint32_t)(NKC * A_CHUNK));
            const unsigned char* src = cloud_pack + (size_t)(rb >> 1) * NKC * B_CHUNK + (size_t)(rb & 1) * A_KB;
            for (int kc = 0; kc < NKC; ++kc)
                for (int pl = 0; pl < 3; ++pl)
                    for (int kb = 0; kb < 2; ++kb)
                        bulk_g2s(As + (size_t)kc * A_CHUNK + pl * A_PLANE + kb * A_KB,
                                 src + (size_t)kc * B_CHUNK + pl * B_PLANE + kb * B_KB, A_KB, barA);
            int t = 0, s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nblk; ++it) {
                const int cb = it >= NB ? it - NB : it;
                const unsigned char* bsrc = cloud_pack + (size_t)cb * NKC * B_CHUNK;
                for (int kc = 0; kc < NKC; ++kc, ++t) {
                    if (t >= S) mbar_wait(empty + s, ph ^ 1u);
                    const uint32_t bytes = (it >= NB || !p.passA3) ? B_CHUNK : 2 * B_PLANE;     // pass A reads the hi and mid planes only
                    mbar_expect_tx(full + s, bytes);
                    bulk_g2s(Ring + (size_t)s * B_CHUNK, bsrc + (size_t)kc * B_CHUNK, bytes, full + s);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // D fp32, A/B bf16, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TNB >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint64_t adesc0 = make_desc(smem_u32(As), A_KB, 128);
            const uint64_t bdesc0 = make_desc(smem_u32(Ring), B_KB, 128);
            mbar_wait(barA, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nblk; ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait(tempty + buf, (uint32_t)(((it >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t dcol = tmem_base + (uint32_t)(buf * TNB);
                for (int kc = 0; kc < NKC; ++kc) {
                    mbar_wait(full + s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint64_t ad = adesc0 + (uint64_t)((kc * A_CHUNK) >> 4);
                    const uint64_t bd = bdesc0 + (uint64_t)((s * B_CHUNK) >> 4);
                    constexpr uint64_t A1 = A_PLANE >> 4, A2 = (2 * A_PLANE) >> 4, B1 = B_PLANE >> 4, B2 = (2 * B_PLANE) >> 4;
                    // plane products, small terms first: h*l, l*h, m*m, h*m, m*h, h*h (pass A: the last three)
                    if (it >= NB || !p.passA3) {
                        umma_bf16(dcol, ad, bd + B2, idesc, kc == 0 ? 0u : 1u);
                        umma_bf16(dcol, ad + A2, bd, idesc, 1u);
                        umma_bf16(dcol, ad + A1, bd + B1, idesc, 1u);
                        umma_bf16(dcol, ad, bd + B1, idesc, 1u);
                    } else {
                        umma_bf16(dcol, ad, bd + B1, idesc, kc == 0 ? 0u : 1u);
                    }
                    umma_bf16(dcol, ad + A1, bd, idesc, 1u);
                    umma_bf16(dcol, ad, bd, idesc, 1u);
                    umma_commit(empty + s);     // the stage may be refilled once these MMAs have read it
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                umma_commit(tfull + buf);       // accumulators of this block are complete
            }
        }
    } else {
        // ================= scanners: warps 0..7, lane quarter = warp & 3, column half = warp >> 2 =================
        const int q4 = warp & 3, half = warp >> 2;
        const int row = q4 * 32 + lane;
        long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0;
        for (int j = tid; j < NB * TNB; j += NSCAN) xxs[j] = __ldg(p.xx + (size_t)b * NB * TNB + j);
        scan_bar();
        const bool row_ok = (i0 + row) < p.N;
        const float xi = xxs[i0 + row];
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * (TNB / 2));
        const float CA = -0.5f * (1.0f + p.epsA), CB = -0.5f * (1.0f - p.eps);

        float m[32];
#pragma unroll
        for (int g = 0; g < 32; ++g) m[g] = -INFINITY;
        float R = INFINITY;
        const float hxi = row_ok ? 0.5f * xi : 0.0f;
        int qn = 0;                                   // this thread's queue fill (row, half)
        const uint32_t stg0 = smem_u32(gm) + (uint32_t)tid * 4u;      // pass B staging: [32 slots][NSCAN threads] floats
        float2* myq = p.gq + ((base + i0 + (row_ok ? row : 0)) * 2 + half) * CAPH;

        for (int it = 0; it < nblk; ++it) {
            const int buf = it & 1;
            const bool passB = it >= NB;
            const int cb = passB ? it - NB : it;
            if (it == NB) {
                tc1 = clock64();
                // ---- between the passes: threshold from the k-th largest of the 64 group maxima ----
#pragma unroll
                for (int g = 0; g < 32; ++g) gm[(half * 32 + g) * TM + row] = m[g];
                scan_bar();
                if (warp < 4) {
                    float v[GROUPS];
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g) v[g] = gm[g * TM + row];
#pragma unroll
                    for (int size = 2; size <= GROUPS; size <<= 1)
#pragma unroll
                        for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
                            for (int i = 0; i < GROUPS; ++i) {
                                const int l = i ^ stride;
                                if (l > i) {
                                    const float a = v[i], c = v[l];
                                    const bool desc = (i & size) == 0;
                                    v[i] = desc ? fmaxf(a, c) : fminf(a, c);
                                    v[l] = desc ? fminf(a, c) : fmaxf(a, c);
                                }
                            }
                    float Lp = v[0];
#pragma unroll
                    for (int i = 1; i < GROUPS; ++i)
                        if (i == k - 1) Lp = v[i];
                    // upper bound (pass B, eps) >= lower bound of the k-th score (pass A, epsA)
                    //   <=>  V_ij >= Lp - (eps + epsA)/2 * xx_i (minus fp32 slack)
                    const float r = Lp - 0.5f * (p.eps + p.epsA) * xi - 9.5367431640625e-7f * (fabsf(Lp) + xi);
                    thr[row] = row_ok ? r : INFINITY;
                }
                scan_bar();
                R = thr[row];
                tc2 = clock64();
            }
            mbar_wait(tfull + buf, (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int j0 = cb * TNB + half * (TNB / 2);
#pragma unroll
            for (int part = 0; part < 4; ++part) {
                float d[32];
                tmem_ld32(trow + (uint32_t)(buf * TNB + part * 32), d);
                if (part == 3) {
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty + buf);      // this warp has drained the accumulator buffer
                }
                const float* xp = xxs + j0 + part * 32;
                if (!passB) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 x4 = *reinterpret_cast<const float4*>(xp + c);
                        m[c + 0] = fmaxf(m[c + 0], fmaf(x4.x, CA, d[c + 0]));
                        m[c + 1] = fmaxf(m[c + 1], fmaf(x4.y, CA, d[c + 1]));
                        m[c + 2] = fmaxf(m[c + 2], fmaf(x4.z, CA, d[c + 2]));
                        m[c + 3] = fmaxf(m[c + 3], fmaf(x4.w, CA, d[c + 3]));
                    }
                } else {
                    // survivors are rare (~2 % of the scores): the unrolled test costs six instructions per score -- the
                    // passing accumulator goes to this thread's column of a shared staging area (the group-maxima region,
                    // free after the threshold) and sets its bit in a mask; the few set bits are then turned into
                    // (score, index) entries of the thread-private survivor queue (row, half) in global memory: no
                    // atomics, ascending candidate order.  (Round 1 tested and stored every score with a predicated
                    // 8-byte global store: 14 instructions per score.)
                    float xr[32];
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 x4 = *reinterpret_cast<const float4*>(xp + c);
                        xr[c] = x4.x; xr[c + 1] = x4.y; xr[c + 2] = x4.z; xr[c + 3] = x4.w;
                    }
                    uint32_t sa = stg0;
                    unsigned mask = 0u;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float t = fmaf(xr[c], CB, d[c]);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %2, %3;\n\t@p st.shared.f32 [%0], %4;\n\t@p add.u32 %0, %0, %5;\n\t"
                            "@p or.b32 %1, %1, %6;\n\t}\n"
                            : "+r"(sa), "+r"(mask)
                            : "f"(t), "f"(R), "f"(d[c]), "n"(NSCAN * 4), "r"(1u << c));
                    }
                    uint32_t ra = stg0;
                    while (mask) {
                        const int c = __ffs(mask) - 1;
                        mask &= mask - 1u;
                        float dv;
                        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(dv) : "r"(ra));
                        ra += NSCAN * 4;
                        const int j = j0 + part * 32 + c;
                        const float val = fmaf(xp[c], -0.5f, dv) - hxi;      // q/2: small for near neighbours
                        if (qn < CAPH) myq[qn] = make_float2(val, __int_as_float(j));
                        ++qn;
                    }
                }
            }
        }
        if (row_ok) p.gqcnt[(base + i0 + row) * 2 + half] = qn;
        tc3 = clock64();
        if (tid == 0 && p.stats) {
            atomicAdd(&g_knn_tc_stats[4], 1ull);
            atomicAdd(&g_knn_tc_stats[5], (unsigned long long)(tc1 - tc0));
            atomicAdd(&g_knn_tc_stats[6], (unsigned long long)(tc2 - tc1));
            atomicAdd(&g_knn_tc_stats[7], (unsigned long long)(tc3 - tc2));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * TNB));
}

// staged query rows of the exact re-scoring: stride = multiple of 4 floats with an odd number of float4s
// (16-byte loads, lanes reading different rows fall into different bank groups)
__host__ __device__ __forceinline__ int fin_stride(int C)
{
    const int q = (C + 3) >> 2;
    return 4 * (q | 1);
}

// bitonic sorts, descending, of 32 keys (one per lane)
__device__ __forceinline__ void warp_sort32_desc(kkey_t& key, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
        const bool desc = (lane & size) == 0;
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const kkey_t o = shfl_xor_key(key, stride);
            const bool keep_larger = (((lane & stride) == 0) == desc);
            if ((key > o) != keep_larger) key = o;
        }
    }
}
__device__ __forceinline__ void warp_sort32_desc_u32(unsigned& key, int lane)
{
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
        const bool desc = (lane & size) == 0;
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned o = __shfl_xor_sync(SV_FULL, key, stride);
            const bool keep_larger = (((lane & stride) == 0) == desc);
            key = keep_larger ? max(key, o) : min(key, o);
        }
    }
}

// 64 keys in two registers per lane (position = slot * 32 + lane), descending
__device__ __forceinline__ void warp_sort64_desc_u32(unsigned& k0, unsigned& k1, int lane)
{
    warp_sort32_desc_u32(k0, lane);
    warp_sort32_desc_u32(k1, lane);
    const unsigned r = __shfl_sync(SV_FULL, k1, 31 - lane);     // k0 ++ reverse(k1) is bitonic: one half-cleaner, then
    k1 = min(k0, r);                                            // each half is bitonic again and the upper half dominates
    k0 = max(k0, r);
    warp_sort32_desc_u32(k0, lane);
    warp_sort32_desc_u32(k1, lane);
}
__device__ __forceinline__ void warp_sort64_desc(kkey_t& k0, kkey_t& k1, int lane)
{
    warp_sort32_desc(k0, lane);
    warp_sort32_desc(k1, lane);
    const kkey_t r = shfl_key(k1, 31 - lane);
    const kkey_t hi = k0 > r ? k0 : r, lo = k0 > r ? r : k0;
    k0 = hi; k1 = lo;
    warp_sort32_desc(k0, lane);
    warp_sort32_desc(k1, lane);
}
// approximate-order key: order-preserving map of the score with the low 5 bits replaced by the lane
// that holds the entry (the lost bits are far below the error bound of the score); 0 = empty
__device__ __forceinline__ unsigned approx_key(float s, int lane)
{
    const unsigned f = __float_as_uint(s + 0.0f);
    const unsigned hi = (f & 0x80000000u) ? ~f : (f | 0x80000000u);
    return ((hi & ~31u) | (unsigned)lane) | 32u * (hi < 64u);     // never below 32: empty slots are 0..31
}

// Exact oracle-chain score of (query row staged in shared memory, candidate row of the fp32 copy in global memory) as the
// order-preserving 32-bit image make_key() uses.  Channel ascending; the zero padding adds fmaf(0,0,x) == x.
__device__ __forceinline__ unsigned chain_key(const float* __restrict__ arow, const float* __restrict__ brow, int n4, float xxi,
                                              float xxj)
{
    const float4* ap = reinterpret_cast<const float4*>(arow);
    const float4* bp = reinterpret_cast<const float4*>(brow);
    float dot = 0.0f;
#pragma unroll 4
    for (int c4 = 0; c4 < n4; ++c4) {
        const float4 bv = __ldg(bp + c4);
        const float4 av = ap[c4];
        dot = __fmaf_rn(av.x, bv.x, dot);
        dot = __fmaf_rn(av.y, bv.y, dot);
        dot = __fmaf_rn(av.z, bv.z, dot);
        dot = __fmaf_rn(av.w, bv.w, dot);
    }
    return (unsigned)(make_key(exact_score(dot, xxi, xxj), 0) >> 32);
}
__device__ __forceinline__ float key32_score(unsigned hi) { return key_score((kkey_t)hi << 32); }

// brute-force exact selection of one row by one warp (queue overflow / degenerate inputs); k <= 32 R
template <int R>
__device__ void brute_force_row(const knn_tc_args& p, const float* __restrict__ xfc, int n4, int i, const float* arow,
                                const float* __restrict__ xxs, int lane, kkey_t (&out)[R])
{
    TopK<R> L;
#pragma unroll
    for (int r = 0; r < R; ++r) L.k[r] = 0ull;
    const float xi = xxs[i];
    const int k = p.k;
    for (int j0 = 0; j0 < p.N; j0 += 32) {
        const int j = j0 + lane;
        kkey_t c = 0ull;
        if (j < p.N) c = ((kkey_t)chain_key(arow, xfc + (size_t)j * p.C4, n4, xi, xxs[j]) << 32) | (unsigned)(~j);
        kkey_t worst = shfl_key(L.k[0], (k - 1) & 31);
        if (R > 1 && k > 32) worst = shfl_key(L.k[R - 1], (k - 1) & 31);
        unsigned m = __ballot_sync(SV_FULL, c > worst);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            topk_insert<R>(L, shfl_key(c, src), lane);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = L.k[r];
}


// 16-byte async copy global -> shared (no register staging)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// per-warp state of the exact re-scoring
struct fin_env {
    const float* xfc;      // fp32 rows of this cloud
    const float* xxs;
    float* arow;           // the warp's query row in shared memory (Cp floats; doubles as the scratch of the prune)
    int C4, n4, Cp, lane, i;
    float xxi;
    bool arow_loaded;
};

// the query row for the paths that read candidate rows straight from global memory (brute force, heavy ties)
__device__ __forceinline__ void ensure_query_row(fin_env& f)
{
    if (!f.arow_loaded) {
        for (int c4 = f.lane; c4 < f.n4; c4 += 32) cp_async16(f.arow + 4 * c4, f.xfc + (size_t)f.i * f.C4 + 4 * c4);
        cp_async_wait_all();
        f.arow_loaded = true;
    }
    __syncwarp();
}
// exact oracle-chain score keys for the lanes in `flagged` (bit = lane, lane's candidate = jme, its norm = xj); 0 for the
// others.  Every flagged lane reads its candidate row straight from the fp32 copy in global memory (L2) with 16-byte loads
// (round 2, measured against staging the rows in shared memory by cp.async: 138 against 148 us per 32 clouds over the four
// layers, and against pooling the pairs of a CTA's eight rows into a lane-dense pass behind a barrier: 169 us).
__device__ __forceinline__ unsigned exact_keys(fin_env& f, unsigned flagged, int jme, float xj)
{
    ensure_query_row(f);
    unsigned sc = 0u;
    if ((flagged >> f.lane) & 1u) {
        const float4* ap = reinterpret_cast<const float4*>(f.arow);
        const float4* bp = reinterpret_cast<const float4*>(f.xfc + (size_t)jme * f.C4);
        float dot = 0.0f;
#pragma unroll 8
        for (int c4 = 0; c4 < f.n4; ++c4) {
            const float4 bv = __ldg(bp + c4);
            const float4 av = ap[c4];
            dot = __fmaf_rn(av.x, bv.x, dot);
            dot = __fmaf_rn(av.y, bv.y, dot);
            dot = __fmaf_rn(av.z, bv.z, dot);
            dot = __fmaf_rn(av.w, bv.w, dot);
        }
        sc = (unsigned)(make_key(exact_score(dot, f.xxi, xj), 0) >> 32);
    }
    return sc;
}

// Shared memory of the finish kernels: per warp the query row, which doubles as the scratch of the pruned survivor
// lists (32 * S float2; S = sorted positions per lane: 1 for k <= 32, 2 for 32 < k <= 48).
__host__ __device__ inline int fin_warp_floats(int Cp, int S)
{
    const int a = Cp, b = 64 * S;
    return a > b ? a : b;
}
// ---- finish kernel: one warp = one row; warps are independent (no CTA barriers).  High occupancy hides the shuffle
//      and gather latencies that a tensor-core CTA with 8 scanner warps cannot.
//   <= 32 survivors (the normal case, after the prune below if need be): sort by approximate score; neighbours in that
//      order which the error bound does not separate form runs; runs that touch the first k positions are re-scored
//      with the exact chain (candidate rows staged by one 16-byte cp.async per lane from the fp32 copy of the
//      features; measured on the model: 33 - 87 % of the rows re-score 2 - 4 candidates); re-sort by
//      (run, exact score desc, index asc).
//   33 .. 2*CAPH survivors that the prune cannot reduce (heavy ties): every survivor is re-scored exactly by its lane,
//      32 at a time, and merged into the best 32 (always correct: the survivors contain the true top k).
//   queue overflow: brute force over all candidates.
template <int MINB>
__global__ void __launch_bounds__(FIN_WARPS * 32, MINB) knn_finish_kernel(knn_tc_args p)
{
    extern __shared__ __align__(16) float fin_smem[];
    const int lane = threadIdx.x & 31, warp = sv_warp_id();
    const int b = blockIdx.y;
    const int i = blockIdx.x * FIN_WARPS + warp;
    if (i >= p.N) return;
    const long base = (long)b * p.N;
    const long rg = base + i;
    const int k = p.k;
    const size_t prow0 = (size_t)b * p.NCB * TNB;                 // first padded row of this cloud in xx / xf
    const float* xxs = p.xx + prow0;
    fin_env f;
    f.C4 = p.C4; f.n4 = p.C4 >> 2; f.Cp = fin_stride(p.C4); f.lane = lane; f.i = i;
    f.xfc = p.xf + prow0 * p.C4; f.xxs = xxs;
    f.arow = fin_smem + (size_t)warp * fin_warp_floats(f.Cp, 1);
    f.arow_loaded = false;
    float emax = 0.0f;
    const float2* q0 = p.gq + rg * 2 * CAPH;
    const float2 h0 = __ldg(q0 + lane), h1 = __ldg(q0 + CAPH + lane);     // may hold stale data beyond the counts
    const float xxi = __ldg(xxs + i);
    f.xxi = xxi;
    int2 cc = __ldg(reinterpret_cast<const int2*>(p.gqcnt) + rg);
    cc.x = sv_uniform(cc.x); cc.y = sv_uniform(cc.y);             // the same for every lane: keep the branches on them convergent
    const int c0 = cc.x, cnt = cc.x + cc.y;
    const bool brute = !(cc.x <= CAPH && cc.y <= CAPH && cnt >= k);
    bool st_exact = false;
    int jout = 0, nflag = 0;
    // ---- the row's survivors as one entry per lane (the approximate-order path needs at most 32) ----
    float2 ent = make_float2(0.0f, 0.0f);
    int n = cnt;
    bool fast = !brute && cnt <= 32;
    if (fast) {
        // entry e of the row is half0[e] for e < c0, else half1[e - c0] (fetched by shuffle)
        const int sl = (lane - c0) & 31;
        const float a1 = __shfl_sync(SV_FULL, h1.x, sl), b1 = __shfl_sync(SV_FULL, h1.y, sl);
        ent = lane < c0 ? h0 : make_float2(a1, b1);
    } else if (!brute) {
        // ---- 33 .. 128 survivors: the group-maxima threshold of the scan kernel was loose for this row.  With the
        // survivors' own scores the bound is much tighter: T = k-th largest LOWER bound (q - delta) over the
        // survivors is a lower bound of the k-th best exact score; entries whose UPPER bound (q + delta) is below T
        // are out.  Usually that leaves <= 32, which take the approximate-order path. ----
        float2 en[4];
        float up[4];
        unsigned lk[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int e = s4 * 32 + lane;
            en[s4] = make_float2(0.0f, 0.0f);
            up[s4] = -INFINITY;
            lk[s4] = 0u;
            if (e < cnt) {
                en[s4] = __ldg(q0 + (e < c0 ? e : CAPH + e - c0));
                const float dl = 0.5f * p.eps * (xxi + __ldg(xxs + __float_as_int(en[s4].y)));
                const float dlt = dl + 4.76837158203125e-7f * (fabsf(en[s4].x) + dl);      // fp32 slack of these two operations
                up[s4] = en[s4].x + dlt;
                const unsigned fb = __float_as_uint((en[s4].x - dlt) + 0.0f);
                lk[s4] = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);                    // order-preserving, > 0
            }
        }
        // the 32 largest lower bounds: sort each register, then bitonic half-cleaners + merges
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4)
            if (s4 * 32 < cnt) warp_sort32_desc_u32(lk[s4], lane);
        unsigned top = lk[0];
#pragma unroll
        for (int s4 = 1; s4 < 4; ++s4) {
            if (s4 * 32 < cnt) {
                top = max(top, __shfl_sync(SV_FULL, lk[s4], 31 - lane));
                warp_sort32_desc_u32(top, lane);
            }
        }
        const unsigned tk = __shfl_sync(SV_FULL, top, k - 1);       // k <= 32 in this kernel
        const unsigned tb = (tk & 0x80000000u) ? (tk & 0x7FFFFFFFu) : ~tk;
        const float T = __uint_as_float(tb);
        unsigned keep[4];
        int kept = 0;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            keep[s4] = __ballot_sync(SV_FULL, up[s4] >= T);
            kept += __popc(keep[s4]);
        }
        if (kept <= 32 && kept >= k) {
            float2* scratch = reinterpret_cast<float2*>(f.arow);        // the staging area is still unused here
            int before = 0;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                if ((keep[s4] >> lane) & 1u) scratch[before + __popc(keep[s4] & ((1u << lane) - 1u))] = en[s4];
                before += __popc(keep[s4]);
            }
            __syncwarp();
            if (lane < kept) ent = scratch[lane];
            __syncwarp();
            n = kept;
            fast = true;
        }
    }
    if (brute) {
        ensure_query_row(f);
        kkey_t key[1];
        brute_force_row<1>(p, f.xfc, f.n4, i, f.arow, xxs, lane, key);
        jout = key_index(key[0]);
    } else if (fast) {
        const float xe = lane < n ? __ldg(xxs + __float_as_int(ent.y)) : 0.0f;      // norm of the entry's point
        unsigned ak = lane < n ? approx_key(ent.x, lane) : (unsigned)lane;
        warp_sort32_desc_u32(ak, lane);
        // fetch the entry this position now holds
        const int src = (int)(ak & 31u);
        const float sme = __shfl_sync(SV_FULL, ent.x, src);
        const int jme = __float_as_int(__shfl_sync(SV_FULL, ent.y, src));
        float xj = __shfl_sync(SV_FULL, xe, src);
        if (lane >= n) xj = 0.0f;
        jout = jme;
        // ---- neighbours in this order that the error bound does not separate ----
        const float snx = __shfl_down_sync(SV_FULL, sme, 1);
        const float xnx = __shfl_down_sync(SV_FULL, xj, 1);
        bool am = false;
        if (lane + 1 < n) am = (sme - snx) <= 0.5f * p.eps * (2.0f * xxi + xj + xnx);
        const unsigned amb = __ballot_sync(SV_FULL, am);
        unsigned rel = amb & (k >= 32 ? 0xFFFFFFFFu : (1u << k) - 1u);       // pairs e <= k-1, then the runs continuing from them
        for (;;) {
            const unsigned nx = (rel << 1) & amb & ~rel;
            if (!nx) break;
            rel |= nx;
        }
        if (rel) {
            st_exact = true;
            const unsigned flagged = rel | (rel << 1);
            nflag = __popc(flagged);
            const unsigned sc = exact_keys(f, flagged, jme, xj);
            if (p.stats && ((flagged >> lane) & 1u))
                emax = fmaxf(emax, fabsf(key32_score(sc) - 2.0f * sme) / (xxi + xj));     // 2*sme = tensor-core score
            // composite key: (run start asc, exact score desc, index asc); unflagged entries are their own run
            const unsigned starts = ~(rel << 1);
            kkey_t key = 0ull;
            if (lane < n) {
                const int seg = 31 - __clz(starts & ((2u << lane) - 1u));
                key = ((kkey_t)(127 - seg) << 44) | ((kkey_t)sc << 12) | (kkey_t)(4095 - jme);
            }
            warp_sort32_desc(key, lane);
            jout = 4095 - (int)(key & 4095ull);
        }
    } else {
        // ---- more than 32 survivors: exact scores for all of them, 32 at a time, keep the best 32 ----
        st_exact = true;
        ensure_query_row(f);
        kkey_t best = 0ull;
        for (int e0 = 0; e0 < cnt; e0 += 32) {
            const int e = e0 + lane;
            kkey_t key = 0ull;
            if (e < cnt) {
                const float2 en = __ldg(q0 + (e < c0 ? e : CAPH + e - c0));
                const int je = __float_as_int(en.y);
                const unsigned sc = chain_key(f.arow, f.xfc + (size_t)je * f.C4, f.n4, xxi, __ldg(xxs + je));
                key = ((kkey_t)sc << 32) | (unsigned)(~je);     // == make_key(exact score, j)
            }
            warp_sort32_desc(key, lane);
            const kkey_t rev = shfl_key(key, 31 - lane);
            best = best > rev ? best : rev;          // upper half of the bitonic merge: the 32 best of the 64
            warp_sort32_desc(best, lane);
        }
        jout = key_index(best);
    }
    if (lane < k) {
        const long o = rg * k + lane;
        if (p.idx32) p.idx32[o] = jout;
        if (p.idx64) p.idx64[o] = (int64_t)jout;
    }
    if (p.stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(SV_FULL, emax, o));
        if (lane == 0) {
            atomicAdd(&g_knn_tc_stats[0], 1ull);
            if (st_exact) atomicAdd(&g_knn_tc_stats[1], 1ull);
            if (brute) atomicAdd(&g_knn_tc_stats[2], 1ull);
            atomicAdd(&g_knn_tc_stats[3], (unsigned long long)cnt);
            if (cnt > 32 && !brute) atomicAdd(&g_knn_tc_stats[10], 1ull);
            atomicAdd(&g_knn_tc_stats[11], (unsigned long long)nflag);
            if (emax > 0.0f) atomicMax(&g_knn_tc_stats[9], (unsigned long long)__float_as_uint(emax));
        }
    }
}

// ---- finish kernel for 32 < k <= 48 (part segmentation, k = 40): the same path with two sorted positions per lane
//      (64 per row); rows with more than 64 survivors after the prune re-score every survivor ----
__global__ void __launch_bounds__(FIN_WARPS * 32, FIN_WIDE_MIN_BLOCKS) knn_finish_wide_kernel(knn_tc_args p)
{
    extern __shared__ __align__(16) float fin_smem[];
    const int lane = threadIdx.x & 31, warp = sv_warp_id();
    const int b = blockIdx.y;
    const int i = blockIdx.x * FIN_WARPS + warp;
    if (i >= p.N) return;
    const long base = (long)b * p.N;
    const long rg = base + i;
    const int k = p.k;
    const size_t prow0 = (size_t)b * p.NCB * TNB;
    const float* xxs = p.xx + prow0;
    fin_env f;
    f.C4 = p.C4; f.n4 = p.C4 >> 2; f.Cp = fin_stride(p.C4); f.lane = lane; f.i = i;
    f.xfc = p.xf + prow0 * p.C4; f.xxs = xxs;
    f.arow = fin_smem + (size_t)warp * fin_warp_floats(f.Cp, 2);
    f.arow_loaded = false;
    const float xxi = __ldg(xxs + i);
    f.xxi = xxi;
    const float2* q0 = p.gq + rg * 2 * CAPH;
    int2 cc = __ldg(reinterpret_cast<const int2*>(p.gqcnt) + rg);
    cc.x = sv_uniform(cc.x); cc.y = sv_uniform(cc.y);             // the same for every lane: keep the branches on them convergent
    const int c0 = cc.x, cnt = cc.x + cc.y;
    const bool brute = !(cc.x <= CAPH && cc.y <= CAPH && cnt >= k);
    kkey_t best[2] = {0ull, 0ull};
    int nflag = 0;
    if (brute) {
        ensure_query_row(f);
        brute_force_row<2>(p, f.xfc, f.n4, i, f.arow, xxs, lane, best);
    }
    // ---- 65 .. 128 survivors: prune with the survivors' own bounds first (see knn_finish_kernel): T = k-th largest
    // lower bound; entries whose upper bound is below T are out; what is left usually fits the two-entry path ----
    int n = cnt;
    bool pruned = false;
    float2* scratch = reinterpret_cast<float2*>(f.arow);         // the staging area is still unused here (>= 128 floats)
    if (!brute && cnt > 64) {
        float2 en4[4];
        float up[4];
        unsigned lk[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int e = s4 * 32 + lane;
            en4[s4] = make_float2(0.0f, 0.0f);
            up[s4] = -INFINITY;
            lk[s4] = 0u;
            if (e < cnt) {
                en4[s4] = __ldg(q0 + (e < c0 ? e : CAPH + e - c0));
                const float dl = 0.5f * p.eps * (xxi + __ldg(xxs + __float_as_int(en4[s4].y)));
                const float dlt = dl + 4.76837158203125e-7f * (fabsf(en4[s4].x) + dl);
                up[s4] = en4[s4].x + dlt;
                const unsigned fb = __float_as_uint((en4[s4].x - dlt) + 0.0f);
                lk[s4] = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
            }
        }
        warp_sort64_desc_u32(lk[0], lk[1], lane);
        warp_sort64_desc_u32(lk[2], lk[3], lane);
        // the 64 largest of the 128: half-cleaner against the reversed second list, then sort
        const unsigned r0 = __shfl_sync(SV_FULL, lk[3], 31 - lane), r1 = __shfl_sync(SV_FULL, lk[2], 31 - lane);
        unsigned t0 = max(lk[0], r0), t1 = max(lk[1], r1);
        warp_sort64_desc_u32(t0, t1, lane);
        const unsigned tk = __shfl_sync(SV_FULL, (k - 1) < 32 ? t0 : t1, (k - 1) & 31);
        const float T = __uint_as_float((tk & 0x80000000u) ? (tk & 0x7FFFFFFFu) : ~tk);
        unsigned keep[4];
        int kept = 0;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            keep[s4] = __ballot_sync(SV_FULL, up[s4] >= T);
            kept += __popc(keep[s4]);
        }
        if (kept <= 64 && kept >= k) {
            int before = 0;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                if ((keep[s4] >> lane) & 1u) scratch[before + __popc(keep[s4] & ((1u << lane) - 1u))] = en4[s4];
                before += __popc(keep[s4]);
            }
            __syncwarp();
            n = kept;
            pruned = true;
        }
    }
    if (brute) {
    } else if (n <= 64) {
        // ---- approximate-order path with two entries per lane (as knn_finish_kernel, positions 0..63): sort by the
        // tensor-core score; neighbours in that order which the error bound separates are certainly ordered; runs of
        // closer scores that touch the first k positions are re-scored with the exact chain ----
        float2 en[2];
        float xe[2];
        unsigned ak[2];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            const int e = s2 * 32 + lane;
            en[s2] = make_float2(0.0f, 0.0f);
            xe[s2] = 0.0f;
            ak[s2] = (unsigned)(s2 * 32 + lane);                   // empty slots: keys 0..63, below every real key
            if (e < n) {
                en[s2] = pruned ? scratch[e] : __ldg(q0 + (e < c0 ? e : CAPH + e - c0));
                xe[s2] = __ldg(xxs + __float_as_int(en[s2].y));
                const unsigned fb = __float_as_uint(en[s2].x + 0.0f);
                const unsigned hi = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
                ak[s2] = ((hi & ~63u) | (unsigned)(s2 * 32 + lane)) | 64u * (hi < 128u);
            }
        }
        __syncwarp();                                   // the scratch entries are in registers now (exact_keys reuses the area)
        warp_sort64_desc_u32(ak[0], ak[1], lane);
        float sme[2], xj[2];
        int jme[2];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            // fetch the entry that position s2 * 32 + lane now holds: source (slot, lane) in the low 6 bits
            const int sl = (int)(ak[s2] & 31u), ss = (int)((ak[s2] >> 5) & 1u);
            const float a0 = __shfl_sync(SV_FULL, en[0].x, sl), a1 = __shfl_sync(SV_FULL, en[1].x, sl);
            const float b0 = __shfl_sync(SV_FULL, en[0].y, sl), b1 = __shfl_sync(SV_FULL, en[1].y, sl);
            const float x0 = __shfl_sync(SV_FULL, xe[0], sl), x1 = __shfl_sync(SV_FULL, xe[1], sl);
            sme[s2] = ss ? a1 : a0;
            jme[s2] = __float_as_int(ss ? b1 : b0);
            xj[s2] = (s2 * 32 + lane < n) ? (ss ? x1 : x0) : 0.0f;
        }
        // neighbours in this order that the error bound does not separate (position p against p + 1)
        unsigned long long amb = 0ull;
        {
            float snx[2], xnx[2];
            snx[0] = __shfl_down_sync(SV_FULL, sme[0], 1);
            xnx[0] = __shfl_down_sync(SV_FULL, xj[0], 1);
            snx[1] = __shfl_down_sync(SV_FULL, sme[1], 1);
            xnx[1] = __shfl_down_sync(SV_FULL, xj[1], 1);
            const float s10 = __shfl_sync(SV_FULL, sme[1], 0), x10 = __shfl_sync(SV_FULL, xj[1], 0);
            if (lane == 31) { snx[0] = s10; xnx[0] = x10; }
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                bool am = false;
                if (s2 * 32 + lane + 1 < n) am = (sme[s2] - snx[s2]) <= 0.5f * p.eps * (2.0f * xxi + xj[s2] + xnx[s2]);
                amb |= (unsigned long long)__ballot_sync(SV_FULL, am) << (32 * s2);
            }
        }
        unsigned long long rel = amb & ((1ull << k) - 1ull);          // pairs p <= k-1, then the runs continuing from them
        for (;;) {
            const unsigned long long nx = (rel << 1) & amb & ~rel;
            if (!nx) break;
            rel |= nx;
        }
        int jout[2] = {jme[0], jme[1]};
        if (rel) {
            const unsigned long long flagged = rel | (rel << 1);
            nflag = __popcll(flagged);
            unsigned sc[2];
            sc[0] = exact_keys(f, (unsigned)flagged, jme[0], xj[0]);
            sc[1] = exact_keys(f, (unsigned)(flagged >> 32), jme[1], xj[1]);
            // composite key: (run start asc, exact score desc, index asc); unflagged entries are their own run
            const unsigned long long starts = ~(rel << 1);
            kkey_t key[2] = {0ull, 0ull};
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const int pos = s2 * 32 + lane;
                if (pos < n) {
                    const unsigned long long below = pos == 63 ? ~0ull : ((2ull << pos) - 1ull);
                    const int seg = 63 - __clzll((long long)(starts & below));
                    key[s2] = ((kkey_t)(127 - seg) << 44) | ((kkey_t)sc[s2] << 12) | (kkey_t)(4095 - jme[s2]);
                }
            }
            warp_sort64_desc(key[0], key[1], lane);
            jout[0] = 4095 - (int)(key[0] & 4095ull);
            jout[1] = 4095 - (int)(key[1] & 4095ull);
        }
        // hand the ordered indices to the common store below as keys (only the index part is read)
        best[0] = (kkey_t)(unsigned)(~jout[0]);
        best[1] = (kkey_t)(unsigned)(~jout[1]);
    } else {
        ensure_query_row(f);
        for (int e0 = 0; e0 < cnt; e0 += 32) {
            const int e = e0 + lane;
            kkey_t key = 0ull;
            if (e < cnt) {
                const float2 en = __ldg(q0 + (e < c0 ? e : CAPH + e - c0));
                const int je = __float_as_int(en.y);
                const unsigned sc = chain_key(f.arow, f.xfc + (size_t)je * f.C4, f.n4, xxi, __ldg(xxs + je));
                key = ((kkey_t)sc << 32) | (unsigned)(~je);     // == make_key(exact score, j)
            }
            warp_sort32_desc(key, lane);
            // (best[0], best[1]) = the 64 best so far, sorted; merge the new 32 into the lower half, then the halves
            kkey_t rev = shfl_key(key, 31 - lane);
            kkey_t lo = best[1] > rev ? best[1] : rev;
            warp_sort32_desc(lo, lane);
            rev = shfl_key(lo, 31 - lane);
            kkey_t hi = best[0] > rev ? best[0] : rev;
            lo = best[0] > rev ? rev : best[0];
            warp_sort32_desc(hi, lane);
            warp_sort32_desc(lo, lane);
            best[0] = hi; best[1] = lo;
        }
    }
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int pos = w * 32 + lane;
        if (pos < k) {
            const long o = rg * k + pos;
            const int j = key_index(best[w]);
            if (p.idx32) p.idx32[o] = j;
            if (p.idx64) p.idx64[o] = (int64_t)j;
        }
    }
    if (p.stats && lane == 0) {
        atomicAdd(&g_knn_tc_stats[0], 1ull);
        if (nflag || brute || n > 64) atomicAdd(&g_knn_tc_stats[1], 1ull);
        if (brute) atomicAdd(&g_knn_tc_stats[2], 1ull);
        atomicAdd(&g_knn_tc_stats[3], (unsigned long long)cnt);
        if (cnt > 32 && !brute) atomicAdd(&g_knn_tc_stats[10], 1ull);
        atomicAdd(&g_knn_tc_stats[11], (unsigned long long)nflag);
    }
}

size_t knn_tc_smem(int NCB, int NKC, int stages)
{
    return (size_t)NKC * A_CHUNK + (size_t)stages * B_CHUNK + (size_t)NCB * TNB * 4 + GM_BYTES + TM * 4 +
           (1 + 2 * MAX_STAGES + 4) * 8 + 16;
}

// error model of the tensor-core score against the oracle chain, relative to (xx_i + xx_j):
// truncating fp32 accumulation of 6 MMAs per 16 channels, chain rounding, dropped plane products,
// final subtractions; x1.5 safety.  Observed maxima stay below 0.6x of the unscaled model.
float knn_tc_eps(int C)
{
    const int NKC = (C + KCH - 1) / KCH;
    return 1.5f * (NKC * 7.2e-7f + C * 3.0e-8f + 8.0e-7f);
}

// three plane products (h*h, h*m, m*h).  The split truncates: |m| < 2^-7 |a|, |l| < 2^-15 |a| per element, so the
// dropped h*l + l*h + m*m (+ smaller) terms are below (2*2^-15 + 2^-14 + 2^-21) |a||b| <= 6.2e-5 (xx_i + xx_j);
// accumulation as above; x1.5 safety
float knn_tc_eps3(int C)
{
    const int NKC = (C + KCH - 1) / KCH;
    return 1.5f * (6.2e-5f + NKC * 3.6e-7f + C * 3.0e-8f + 8.0e-7f);
}

}  // namespace

struct knn_tc_plan_t {
    int NCB, NKC, stages, C4;
    size_t pack_bytes, xx_bytes, xf_bytes, gq_bytes, cnt_bytes, fin_smem;
    size_t total() const { return pack_bytes + xx_bytes + xf_bytes + gq_bytes + cnt_bytes; }
};

static bool knn_tc_plan(const svnet_view* in, int B, int N, int k, knn_tc_plan_t* pl)
{
    const char* on = getenv("SVNET_KNN_TC");
    if (on && on[0] == '0') return false;
    const int C = in->Cs + 3 * in->Cv;
    if (B < 1 || k > KNN_TC_MAX_K || N > 4096 || C > KMAX || C < 1 || N < 64) return false;
    pl->NCB = sv_cdiv(N, TNB);
    pl->NKC = sv_cdiv(C, KCH);
    int s = MAX_STAGES;
    if (const char* st = getenv("SVNET_KNN_STAGES")) {    // tuning aid: fewer ring stages leave shared memory to co-resident kernels
        const int v = atoi(st);
        if (v >= 2 && v <= MAX_STAGES) s = v;
    }
    const size_t limit = 227 * 1024;
    while (s > 2 && knn_tc_smem(pl->NCB, pl->NKC, s) > limit) --s;
    if (knn_tc_smem(pl->NCB, pl->NKC, s) > limit) return false;
    pl->stages = s;
    const size_t a256 = 255;
    pl->pack_bytes = ((size_t)B * pl->NCB * pl->NKC * B_CHUNK + a256) & ~a256;
    pl->xx_bytes = ((size_t)B * pl->NCB * TNB * sizeof(float) + a256) & ~a256;
    pl->gq_bytes = ((size_t)B * N * 2 * CAPH * sizeof(float2) + a256) & ~a256;
    pl->cnt_bytes = ((size_t)B * N * 2 * sizeof(int) + a256) & ~a256;
    pl->C4 = (C + 3) & ~3;
    pl->xf_bytes = ((size_t)B * pl->NCB * TNB * pl->C4 * sizeof(float) + a256) & ~a256;
    const int Cp = fin_stride(pl->C4);
    pl->fin_smem = (size_t)FIN_WARPS * fin_warp_floats(Cp, k > 32 ? 2 : 1) * sizeof(float);
    return true;
}

// Scratch bytes the tensor-core path needs for this shape; 0 when the shape is not covered.
size_t svnet_knn_tc_workspace(const svnet_view* in, int B, int N, int k)
{
    knn_tc_plan_t pl;
    if (!knn_tc_plan(in, B, N, k, &pl)) return 0;
    return pl.total();
}

// Returns 1 if handled, 0 if the caller should use the CUDA-core kernel, < 0 on error.
int svnet_knn_tc_dispatch(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, void* workspace,
                          size_t workspace_bytes, cudaStream_t st)
{
    knn_tc_plan_t pl;
    if (!workspace || !knn_tc_plan(in, B, N, k, &pl)) return 0;
    if (workspace_bytes < pl.total() || (reinterpret_cast<uintptr_t>(workspace) & 15))
        return 0;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    knn_tc_args a;
    a.in = *in; a.N = N; a.k = k; a.NCB = pl.NCB; a.NKC = pl.NKC; a.stages = pl.stages; a.C4 = pl.C4;
    a.eps = knn_tc_eps(in->Cs + 3 * in->Cv);
    a.epsA = knn_tc_eps3(in->Cs + 3 * in->Cv);
    // the three-product bound is ~6x looser: fine while neighbours are sparse relative to it, but at large N the
    // survivor queues overflow into the brute-force path (N = 4096: kNN 13 ms instead of 4 ms per 32 clouds)
    a.passA3 = N <= KNN_PASSA3_MAX_N ? 1 : 0;
    if (const char* pa = getenv("SVNET_KNN_PASSA")) {
        if (pa[0] == '6') a.passA3 = 0;
        if (pa[0] == '3') a.passA3 = 1;
    }
    if (!a.passA3) a.epsA = a.eps;
    {
        const char* sv = getenv("SVNET_KNN_TC_STATS");
        a.stats = (sv && sv[0] == '1') ? 1 : 0;
    }
    a.pack = ws;
    a.xx = reinterpret_cast<float*>(ws + pl.pack_bytes);
    a.xf = reinterpret_cast<float*>(ws + pl.pack_bytes + pl.xx_bytes);
    a.gq = reinterpret_cast<float2*>(ws + pl.pack_bytes + pl.xx_bytes + pl.xf_bytes);
    a.gqcnt = reinterpret_cast<int*>(ws + pl.pack_bytes + pl.xx_bytes + pl.xf_bytes + pl.gq_bytes);
    a.idx32 = idx32; a.idx64 = idx64;
    knn_pack_kernel<<<dim3(pl.NCB * (TNB / PACK_ROWS), B), 256, (size_t)PACK_ROWS * (pl.NKC * KCH + 1) * sizeof(float), st>>>(*in, N, pl.NCB, pl.NKC, ws, const_cast<float*>(a.xx), const_cast<float*>(a.xf), pl.C4);
    SV_CHECK_LAUNCH("svnet_knn(pack)");
    const size_t smem = knn_tc_smem(pl.NCB, pl.NKC, pl.stages);
    SV_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_tc_kernel<<<dim3(sv_cdiv(N, TM), B), NTHREADS, smem, st>>>(a);
    SV_CHECK_LAUNCH("svnet_knn(tcgen05)");
    if (k <= 32) {
        int mb = FIN_MIN_BLOCKS;
        if (const char* e = getenv("SVNET_KNN_FIN_MB")) mb = atoi(e);      // tuning aid: resident CTAs per SM (register budget)
        const dim3 grid(sv_cdiv(N, FIN_WARPS), B);
        if (mb == 4) knn_finish_kernel<4><<<grid, FIN_WARPS * 32, pl.fin_smem, st>>>(a);
        else if (mb == 5) knn_finish_kernel<5><<<grid, FIN_WARPS * 32, pl.fin_smem, st>>>(a);
        else knn_finish_kernel<6><<<grid, FIN_WARPS * 32, pl.fin_smem, st>>>(a);
    } else {
        SV_CUDA(cudaFuncSetAttribute(knn_finish_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.fin_smem));
        knn_finish_wide_kernel<<<dim3(sv_cdiv(N, FIN_WARPS), B), FIN_WARPS * 32, pl.fin_smem, st>>>(a);
    }
    SV_CHECK_LAUNCH("svnet_knn(finish)");
    return 1;
}

// profiling counters (cumulative since the last reset), see include/svnet_b200.h
extern "C" int svnet_knn_tc_stats(unsigned long long* out12, int reset)
{
    SV_CUDA(cudaMemcpyFromSymbol(out12, g_knn_tc_stats, sizeof(unsigned long long) * 12));
    if (reset) {
        unsigned long long z[12] = {0};
        SV_CUDA(cudaMemcpyToSymbol(g_knn_tc_stats, z, sizeof(z)));
    }
    return SVNET_OK;
}
