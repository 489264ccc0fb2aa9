// K1-TC -- kNN graph construction with the pairwise scores on tcgen05 tensor cores.
// Replaces sv_util.knn (reference models/utils/sv_util.py:19-25) for k <= 24, N <= 4096, C <= 160;
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
//                 candidates reach it) and leaves ~k+5..10 survivors;
//       pass B  : candidates whose UPPER bound reaches L are queued (approximate score, index);
//       finish  : a warp sorts a row's survivors by approximate score; neighbours in that order
//                 whose scores differ by more than the error bound are certainly ordered; runs of
//                 closer scores (and exact ties) are re-scored with the exact chain and re-sorted
//                 by (score desc, index asc).  Rows that overflow the queue take a brute-force
//                 exact path.
// Error model: |p_ij - q_ij| <= EPS * (xx_i + xx_j), p = oracle chain score, q = tensor-core score.
// Dropped plane products contribute < 2^-19, fp32 accumulation (tensor core and chain) < 2^-17 of
// |f_i||f_j|; EPS = 2^-15 leaves a 4x margin.  Scores are handled in half scale:
//   q/2 = dot - (xx_i + xx_j)/2.
#include "common.cuh"
#include "knn_keys.cuh"
#include <stdlib.h>

namespace {
using namespace svknn;

constexpr int TM = 128;                        // query rows per CTA == UMMA M
constexpr int TN = 128;                        // candidates per block == UMMA N
constexpr int KCH = 16;                        // channels per chunk == UMMA K for bf16
constexpr int KB_BYTES = TM * 16;              // one 8-channel k-block of 128 rows (descriptor LBO)
constexpr int PLANE_BYTES = 2 * KB_BYTES;      // one plane of a chunk
constexpr int CHUNK_BYTES = 3 * PLANE_BYTES;   // hi | mid | lo
constexpr int NSCAN = 256;                     // scanner threads: warps 0..7
constexpr int NTHREADS = 320;                  // + warp 8 (MMA issue) + warp 9 (bulk-copy producer)
constexpr int CAPH = 32;                       // survivor queue capacity per (row, column half)
constexpr int QV_LD = CAPH + 1;                // padded strides (bank spread)
constexpr int QJ_LD = CAPH + 2;
constexpr int QUEUE_BYTES = 2 * TM * QV_LD * 4 + 2 * TM * QJ_LD * 2;
constexpr int GROUPS = 64;
constexpr int KMAX = 160;                      // padded channels
constexpr int NU = KMAX / 32;                  // channel slots per lane in the exact re-scoring
constexpr int MAX_STAGES = 8;
constexpr float EPS = 1.0f / 32768.0f;         // 2^-15
constexpr int KNN_TC_MAX_K = 24;
constexpr int UNION_BYTES = QUEUE_BYTES > GROUPS * TM * 4 ? QUEUE_BYTES : GROUPS * TM * 4;

__device__ unsigned long long g_knn_tc_stats[12];   // rows, rows with exact re-scoring, brute-force rows, survivors

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
// pack[b][rb][kc][plane][kb][row 0..127][8 bf16]; xx[b][rb*128 + row] (+inf for rows >= N)
__global__ void __launch_bounds__(128) knn_pack_kernel(svnet_view in, int N, int NRB, int NKC, unsigned char* __restrict__ pack,
                                                       float* __restrict__ xx)
{
    const int rb = blockIdx.x, b = blockIdx.y, rr = threadIdx.x;
    const int r = rb * TM + rr;
    const bool valid = r < N;
    const long row = (long)b * N + r;
    const int C = in.Cs + 3 * in.Cv;
    unsigned char* dst = pack + ((size_t)(b * NRB + rb) * NKC) * CHUNK_BYTES + rr * 16;
    float nrm = 0.0f;
    for (int kc = 0; kc < NKC; ++kc) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            float a[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = kc * KCH + kb * 8 + e;
                a[e] = (valid && c < C) ? sv_feat(in, row, c) : 0.0f;
                nrm = __fmaf_rn(a[e], a[e], nrm);   // zero padding: fmaf(0,0,x) == x
            }
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
            unsigned char* d = dst + (size_t)kc * CHUNK_BYTES + kb * KB_BYTES;
            *reinterpret_cast<uint4*>(d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(d + PLANE_BYTES) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            *reinterpret_cast<uint4*>(d + 2 * PLANE_BYTES) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
    }
    xx[(size_t)b * NRB * TM + r] = valid ? nrm : INFINITY;
}

// ---- exact oracle chain: dot over channels ascending [s | v x0 | v x1 | v x2] ------------------
__device__ __forceinline__ float exact_dot(const svnet_view& in, long rowj, const float* __restrict__ arow)
{
    float dot = 0.0f;
    int c = 0;
    if (in.Cs > 0) {
        const float* ps = in.s + rowj * in.lds;
#pragma unroll 8
        for (int d = 0; d < in.Cs; ++d) dot = __fmaf_rn(arow[c++], __ldg(ps + d), dot);
    }
    if (in.Cv > 0) {
#pragma unroll 1
        for (int x = 0; x < 3; ++x) {
            const float* pv = in.v + rowj * in.ldv + (long)x * in.xs;
#pragma unroll 8
            for (int d = 0; d < in.Cv; ++d) dot = __fmaf_rn(arow[c++], __ldg(pv + d), dot);
        }
    }
    return dot;
}
__device__ __forceinline__ float exact_score(float dot, float xxi, float xxj)
{
    const float inner = -2.0f * dot;
    return __fsub_rn(__fsub_rn(-xxj, inner), xxi);   // ((-xx_j) - (-2 dot)) - xx_i, sv_util.py:20-22
}

struct knn_tc_args {
    svnet_view in;
    int N, k, NRB, NKC, stages;
    const unsigned char* pack;
    const float* xx;
    int32_t* idx32;
    int64_t* idx64;
};

// 4-byte async copy global -> shared (no register staging: many gathers in flight per lane)
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// brute-force exact selection of one row by one warp (queue overflow / degenerate inputs)
__device__ void brute_force_row(const knn_tc_args& p, long base, int i, const float* arow, const float* xxs, int lane,
                                kkey_t& out)
{
    TopK<1> L;      // k <= KNN_TC_MAX_K <= 32
    L.k[0] = 0ull;
    const float xi = xxs[i];
    const int k = p.k;
    for (int j0 = 0; j0 < p.N; j0 += 32) {
        const int j = j0 + lane;
        kkey_t c = 0ull;
        if (j < p.N) c = make_key(exact_score(exact_dot(p.in, base + j, arow), xi, xxs[j]), j);
        const kkey_t worst = shfl_key(L.k[0], k - 1);
        unsigned m = __ballot_sync(SV_FULL, c > worst);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            topk_insert<1>(L, shfl_key(c, src), lane);
        }
    }
    out = L.k[0];
}

// bitonic sort, descending, of four rows in lock-step (ILP across the rows hides the shuffle latency);
// row a holds W*32 keys, key[a][w] at position w*32 + lane
template <int W>
__device__ __forceinline__ void sort4_desc(kkey_t (&key)[4][W], int lane)
{
#pragma unroll
    for (int size = 2; size <= 32 * W; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const bool desc = (((w * 32 + lane) & size) == 0);
                if (stride >= 32) {
                    const int pw = w ^ (stride >> 5);
                    if (pw > w) {
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const kkey_t x = key[a][w], y = key[a][pw];
                            const bool swap = desc ? (x < y) : (x > y);
                            if (swap) { key[a][w] = y; key[a][pw] = x; }
                        }
                    }
                } else {
                    const bool keep_larger = (((lane & stride) == 0) == desc);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const kkey_t o = shfl_xor_key(key[a][w], stride);
                        if ((key[a][w] > o) != keep_larger) key[a][w] = o;
                    }
                }
            }
        }
    }
}

struct fin_ctx {
    const knn_tc_args* p;
    long base;
    int i0, C, Cp, xcap, lane;
    const float* xxs;
    const float* qv;
    const unsigned short* qj;
    float* stage;          // this warp's staging area: 4 query rows, then xcap candidate rows, stride Cp
    int coff[NU];          // channel lane+32u -> offset inside the s row / the v row
    unsigned smask;        // bit u: channel lane+32u lives in the scalar part
    float emax;            // largest observed |p - q| / (xx_i + xx_j)
};

__device__ __forceinline__ const float* fin_src(const fin_ctx& f, long row, int u)
{
    const svnet_view& in = f.p->in;
    return ((f.smask >> u) & 1u) ? in.s + row * in.lds + f.coff[u] : in.v + row * in.ldv + f.coff[u];
}

// Finish four rows (r0 .. r0+3 of the CTA) in lock-step: sort the survivors by approximate score,
// find the neighbours the error bound does not separate, re-score those with the exact chain
// (candidate rows gathered with cp.async into shared memory), re-sort, write the first k indices.
template <int W>
__device__ __forceinline__ void finish_group(fin_ctx& f, int r0, const int (&c0)[4], const int (&cnt)[4], const bool (&ok)[4],
                                             unsigned long long& st_exact)
{
    const int lane = f.lane, k = f.p->k;
    kkey_t key[4][W];
    float xxi[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = r0 + a;
        xxi[a] = ok[a] ? f.xxs[f.i0 + r] : 0.0f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int e = w * 32 + lane;
            kkey_t kk = 0ull;
            if (ok[a] && e < cnt[a]) {
                const int h = e < c0[a] ? 0 : 1, sl = h ? e - c0[a] : e;
                kk = make_key(f.qv[(r * 2 + h) * QV_LD + sl], (int)f.qj[(r * 2 + h) * QJ_LD + sl]);
            }
            key[a][w] = kk;
        }
    }
    sort4_desc<W>(key, lane);

    // ---- neighbours in this order that the error bound does not separate ----
    unsigned long long rel[4];
    bool any = false;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        unsigned long long amb = 0ull;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            kkey_t nx = shfl_key(key[a][w], (lane + 1) & 31);
            if (w + 1 < W) {
                const kkey_t nx2 = shfl_key(key[a][(w + 1 < W) ? w + 1 : w], 0);
                if (lane == 31) nx = nx2;
            }
            const int e = w * 32 + lane;
            bool am = false;
            if (ok[a] && e + 1 < cnt[a]) {
                const float tol = 0.5f * EPS * (2.0f * xxi[a] + f.xxs[key_index(key[a][w])] + f.xxs[key_index(nx)]);
                am = (key_score(key[a][w]) - key_score(nx)) <= tol;
            }
            amb |= (unsigned long long)__ballot_sync(SV_FULL, am) << (32 * w);
        }
        // pairs that can change the first k positions: pairs e <= k-1 and the runs continuing from them
        unsigned long long rl = amb & ((1ull << k) - 1ull);
        for (;;) {
            const unsigned long long nx = (rl << 1) & amb & ~rl;
            if (!nx) break;
            rl |= nx;
        }
        rel[a] = rl;
        any |= rl != 0ull;
    }

    if (any) {
        unsigned long long flagged[4];
        int pre[4][W];       // ordinals of the flagged entries in (a, w, lane) order
        int T = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            flagged[a] = rel[a] | (rel[a] << 1);
            if (rel[a]) st_exact++;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                pre[a][w] = T;
                T += __popc((unsigned)(flagged[a] >> (32 * w)));
            }
        }
        unsigned sc[4][W];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int w = 0; w < W; ++w) sc[a][w] = 0u;
        float* arow = f.stage;
        float* exb = f.stage + 4 * f.Cp;
        const unsigned lt = (1u << lane) - 1u;
        for (int ch0 = 0; ch0 < T; ch0 += f.xcap) {
            __syncwarp();
            if (ch0 == 0) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (rel[a]) {
#pragma unroll
                        for (int u = 0; u < NU; ++u)
                            if (lane + 32 * u < f.C) cp_async4(arow + a * f.Cp + lane + 32 * u, fin_src(f, f.base + f.i0 + r0 + a, u));
                    }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const unsigned mfull = (unsigned)(flagged[a] >> (32 * w));
                    unsigned m = mfull;
                    const int jmine = key_index(key[a][w]);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int o = pre[a][w] + __popc(mfull & ((1u << src) - 1u)) - ch0;
                        if (o < 0 || o >= f.xcap) continue;
                        const int j = __shfl_sync(SV_FULL, jmine, src);
#pragma unroll
                        for (int u = 0; u < NU; ++u)
                            if (lane + 32 * u < f.C) cp_async4(exb + o * f.Cp + lane + 32 * u, fin_src(f, f.base + j, u));
                    }
                }
            cp_async_wait_all();
            __syncwarp();
            // exact chains (channel ascending), the four rows interleaved
#pragma unroll
            for (int w = 0; w < W; ++w) {
                bool mine[4];
                const float* bp[4];
                float dot[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const unsigned mfull = (unsigned)(flagged[a] >> (32 * w));
                    const int o = pre[a][w] + __popc(mfull & lt) - ch0;
                    mine[a] = ((mfull >> lane) & 1u) && o >= 0 && o < f.xcap;
                    bp[a] = exb + (mine[a] ? o : 0) * f.Cp;
                    dot[a] = 0.0f;
                }
                if (mine[0] || mine[1] || mine[2] || mine[3]) {
                    for (int c = 0; c < f.C; ++c) {
#pragma unroll
                        for (int a = 0; a < 4; ++a)
                            if (mine[a]) dot[a] = __fmaf_rn(arow[a * f.Cp + c], bp[a][c], dot[a]);
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        if (mine[a]) {
                            const int j = key_index(key[a][w]);
                            const float xj = f.xxs[j];
                            const float pe = exact_score(dot[a], xxi[a], xj);
                            sc[a][w] = (unsigned)(make_key(pe, 0) >> 32);
                            const float qa = 2.0f * key_score(key[a][w]) - xxi[a];      // tensor-core score
                            f.emax = fmaxf(f.emax, fabsf(pe - qa) / (xxi[a] + xj));
                        }
                }
            }
        }
        // composite keys: (run start asc, exact score desc, index asc); unflagged entries are their own run.
        // Rows without ambiguity keep (approximate score, index) keys: sorting them again changes nothing.
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (!rel[a]) continue;
            const unsigned long long starts = ~(rel[a] << 1);      // e starts a run iff pair (e-1, e) is not ambiguous
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int e = w * 32 + lane;
                if (e < cnt[a]) {
                    const int j = key_index(key[a][w]);
                    const int seg = 63 - __clzll((long long)(starts & ((2ull << e) - 1ull)));
                    key[a][w] = ((kkey_t)(127 - seg) << 44) | ((kkey_t)sc[a][w] << 12) | (kkey_t)(4095 - j);
                } else {
                    key[a][w] = 0ull;
                }
            }
        }
        sort4_desc<W>(key, lane);
    }

#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (!ok[a]) continue;
        const bool comp = any && rel[a] != 0ull;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int pos = w * 32 + lane;
            if (pos < k) {
                const long o = (f.base + f.i0 + r0 + a) * k + pos;
                const int j = comp ? 4095 - (int)(key[a][w] & 4095ull) : key_index(key[a][w]);
                if (f.p->idx32) f.p->idx32[o] = j;
                if (f.p->idx64) f.p->idx64[o] = (int64_t)j;
            }
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) knn_tc_kernel(knn_tc_args p)
{
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int NKC = p.NKC, S = p.stages, NB = p.NRB;
    unsigned char* As = smraw;                                        // NKC chunks, resident
    unsigned char* Ring = As + (size_t)NKC * CHUNK_BYTES;             // S chunks
    float* xxs = reinterpret_cast<float*>(Ring + (size_t)S * CHUNK_BYTES);   // [NRB*128]
    float* un = xxs + NB * TM;                                        // union: gm [64][128]  |  survivor queues
    float* gm = un;
    float* qv = un;                                                   // [128 rows][2 halves][QV_LD]
    unsigned short* qj = reinterpret_cast<unsigned short*>(un + 2 * TM * QV_LD);   // [128][2][QJ_LD]
    float* thr = un + UNION_BYTES / 4;                                // [128]
    int* qcnt = reinterpret_cast<int*>(thr + TM);                     // [128][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(qcnt + 2 * TM);      // barA | full[S] | empty[S] | tfull[2] | tempty[2]
    uint64_t* barA = bars;
    uint64_t* full = bars + 1;
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(2 * TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const unsigned char* cloud_pack = p.pack + (size_t)b * NB * NKC * CHUNK_BYTES;
    const int nblk = 2 * NB;     // pass A then pass B over the candidate blocks

    if (warp == 9) {
        // ================= producer: bulk copies of operand chunks =================
        if (lane == 0) {
            mbar_expect_tx(barA, (uint32_t)(NKC * CHUNK_BYTES));
            const unsigned char* src = cloud_pack + (size_t)rb * NKC * CHUNK_BYTES;
            for (int kc = 0; kc < NKC; ++kc) bulk_g2s(As + (size_t)kc * CHUNK_BYTES, src + (size_t)kc * CHUNK_BYTES, CHUNK_BYTES, barA);
            int t = 0, s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nblk; ++it) {
                const int cb = it >= NB ? it - NB : it;
                const unsigned char* bsrc = cloud_pack + (size_t)cb * NKC * CHUNK_BYTES;
                for (int kc = 0; kc < NKC; ++kc, ++t) {
                    if (t >= S) mbar_wait(empty + s, ph ^ 1u);
                    mbar_expect_tx(full + s, CHUNK_BYTES);
                    bulk_g2s(Ring + (size_t)s * CHUNK_BYTES, bsrc + (size_t)kc * CHUNK_BYTES, CHUNK_BYTES, full + s);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // D fp32, A/B bf16, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint64_t adesc0 = make_desc(smem_u32(As), KB_BYTES, 128);
            const uint64_t bdesc0 = make_desc(smem_u32(Ring), KB_BYTES, 128);
            mbar_wait(barA, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nblk; ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait(tempty + buf, (uint32_t)(((it >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t dcol = tmem_base + (uint32_t)(buf * TN);
                for (int kc = 0; kc < NKC; ++kc) {
                    mbar_wait(full + s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint64_t ad = adesc0 + (uint64_t)((kc * CHUNK_BYTES) >> 4);
                    const uint64_t bd = bdesc0 + (uint64_t)((s * CHUNK_BYTES) >> 4);
                    constexpr uint64_t P1 = PLANE_BYTES >> 4, P2 = (2 * PLANE_BYTES) >> 4;
                    // plane products, small terms first: h*l, l*h, m*m, h*m, m*h, h*h
                    umma_bf16(dcol, ad, bd + P2, idesc, kc == 0 ? 0u : 1u);
                    umma_bf16(dcol, ad + P2, bd, idesc, 1u);
                    umma_bf16(dcol, ad + P1, bd + P1, idesc, 1u);
                    umma_bf16(dcol, ad, bd + P1, idesc, 1u);
                    umma_bf16(dcol, ad + P1, bd, idesc, 1u);
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
        long long tc0 = clock64(), tc1 = 0, tc2 = 0, tc3 = 0, tc4 = 0;
        for (int j = tid; j < NB * TM; j += NSCAN) xxs[j] = __ldg(p.xx + (size_t)b * NB * TM + j);
        scan_bar();
        const bool row_ok = (i0 + row) < p.N;
        const float xi = xxs[i0 + row];
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * 64);
        constexpr float CA = -0.5f * (1.0f + EPS), CB = -0.5f * (1.0f - EPS);

        float m[32];
#pragma unroll
        for (int g = 0; g < 32; ++g) m[g] = -INFINITY;
        float R = INFINITY;
        int qn = 0;                                   // this thread's queue fill (row, half)
        float* myqv = qv + (row * 2 + half) * QV_LD;
        unsigned short* myqj = qj + (row * 2 + half) * QJ_LD;

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
                    // hi_ij >= lower bound of the k-th score  <=>  V_ij >= Lp - EPS*xx_i (minus fp32 slack)
                    const float r = Lp - EPS * xi - 9.5367431640625e-7f * (fabsf(Lp) + xi);
                    thr[row] = row_ok ? r : INFINITY;
                }
                scan_bar();      // thresholds visible; gm (aliased by the queues) no longer read
                R = thr[row];
                tc2 = clock64();
            }
            mbar_wait(tfull + buf, (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            float d0[32], d1[32];
            tmem_ld32(trow + (uint32_t)(buf * TN), d0);
            tmem_ld32(trow + (uint32_t)(buf * TN + 32), d1);
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + buf);      // this warp has drained the accumulator buffer
            const int j0 = cb * TN + half * 64;
            if (!passB) {
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 x0 = *reinterpret_cast<const float4*>(xxs + j0 + c);
                    const float4 x1 = *reinterpret_cast<const float4*>(xxs + j0 + 32 + c);
                    m[c + 0] = fmaxf(m[c + 0], fmaxf(fmaf(x0.x, CA, d0[c + 0]), fmaf(x1.x, CA, d1[c + 0])));
                    m[c + 1] = fmaxf(m[c + 1], fmaxf(fmaf(x0.y, CA, d0[c + 1]), fmaf(x1.y, CA, d1[c + 1])));
                    m[c + 2] = fmaxf(m[c + 2], fmaxf(fmaf(x0.z, CA, d0[c + 2]), fmaf(x1.z, CA, d1[c + 2])));
                    m[c + 3] = fmaxf(m[c + 3], fmaxf(fmaf(x0.w, CA, d0[c + 3]), fmaf(x1.w, CA, d1[c + 3])));
                }
            } else {
                // thread-private queue (row, half): predicated stores, no atomics
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xxs + j0 + c);
                    const float xs4[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float dv = (c + e) < 32 ? d0[(c + e) & 31] : d1[(c + e) & 31];
                        if (fmaf(xs4[e], CB, dv) >= R) {
                            if (qn < CAPH) {
                                myqv[qn] = fmaf(xs4[e], -0.5f, dv);
                                myqj[qn] = (unsigned short)(j0 + c + e);
                            }
                            ++qn;
                        }
                    }
                }
            }
        }
        qcnt[row * 2 + half] = qn;
        scan_bar();     // all queues complete; operand buffers are free (every MMA has completed)
        tc3 = clock64();

        // ================= finish: one warp = 16 rows, four at a time =================
        fin_ctx f;
        f.p = &p; f.base = base; f.i0 = i0; f.lane = lane;
        f.C = p.in.Cs + 3 * p.in.Cv;
        f.Cp = f.C | 1;
        f.xxs = xxs; f.qv = qv; f.qj = qj;
        const int stage_floats = (int)(((size_t)(NKC + S) * CHUNK_BYTES) / (8 * sizeof(float)));
        f.stage = reinterpret_cast<float*>(As) + warp * stage_floats;
        f.xcap = (stage_floats - 4 * f.Cp) / f.Cp;
        f.emax = 0.0f;
        f.smask = 0u;
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int c = lane + 32 * u;
            f.coff[u] = 0;
            if (c < p.in.Cs) { f.smask |= 1u << u; f.coff[u] = c; }
            else if (c < f.C) { const int cc = c - p.in.Cs, x = cc / p.in.Cv; f.coff[u] = x * p.in.xs + (cc - x * p.in.Cv); }
        }
        unsigned long long st_exact = 0, st_brute = 0, st_surv = 0, st_rows = 0;
        for (int g = 0; g < 4; ++g) {
            const int r0 = warp * 16 + g * 4;
            if (i0 + r0 >= p.N) break;
            int c0[4], cnt[4];
            bool ok[4], brute[4];
            bool wide = false;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const bool exists = (i0 + r0 + a) < p.N;
                c0[a] = qcnt[(r0 + a) * 2];
                const int c1 = qcnt[(r0 + a) * 2 + 1];
                cnt[a] = c0[a] + c1;
                const bool usable = c0[a] <= CAPH && c1 <= CAPH && cnt[a] >= k;
                ok[a] = exists && usable;
                brute[a] = exists && !usable;
                wide |= ok[a] && cnt[a] > 32;
                if (exists) { st_rows++; st_surv += (unsigned long long)cnt[a]; }
            }
            if (wide) finish_group<2>(f, r0, c0, cnt, ok, st_exact);
            else finish_group<1>(f, r0, c0, cnt, ok, st_exact);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (!brute[a]) continue;
                const int i = i0 + r0 + a;
                __syncwarp();
                for (int c = lane; c < f.C; c += 32) f.stage[c] = sv_feat(p.in, base + i, c);
                __syncwarp();
                kkey_t out;
                brute_force_row(p, base, i, f.stage, xxs, lane, out);
                if (lane < k) {
                    const long o = (base + i) * k + lane;
                    if (p.idx32) p.idx32[o] = key_index(out);
                    if (p.idx64) p.idx64[o] = (int64_t)key_index(out);
                }
                st_brute++;
            }
        }
        tc4 = clock64();
        if (tid == 0) {
            atomicAdd(&g_knn_tc_stats[4], 1ull);
            atomicAdd(&g_knn_tc_stats[5], (unsigned long long)(tc1 - tc0));
            atomicAdd(&g_knn_tc_stats[6], (unsigned long long)(tc2 - tc1));
            atomicAdd(&g_knn_tc_stats[7], (unsigned long long)(tc3 - tc2));
            atomicAdd(&g_knn_tc_stats[8], (unsigned long long)(tc4 - tc3));
        }
        float em = f.emax;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) em = fmaxf(em, __shfl_xor_sync(SV_FULL, em, o));
        if (lane == 0) {
            atomicAdd(&g_knn_tc_stats[0], st_rows);
            atomicAdd(&g_knn_tc_stats[1], st_exact);
            atomicAdd(&g_knn_tc_stats[2], st_brute);
            atomicAdd(&g_knn_tc_stats[3], st_surv);
            atomicMax(&g_knn_tc_stats[9], (unsigned long long)__float_as_uint(em));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * TN));
}

size_t knn_tc_smem(int NRB, int NKC, int stages)
{
    return (size_t)(NKC + stages) * CHUNK_BYTES + (size_t)NRB * TM * 4 + UNION_BYTES + TM * 4 + 2 * TM * 4 +
           (1 + 2 * MAX_STAGES + 4) * 8 + 16;
}

}  // namespace

static bool knn_tc_plan(const svnet_view* in, int B, int N, int k, int* NRB, int* NKC, int* stages, size_t* pack_bytes,
                        size_t* xx_bytes)
{
    const char* on = getenv("SVNET_KNN_TC");
    if (on && on[0] == '0') return false;
    const int C = in->Cs + 3 * in->Cv;
    if (B < 1 || k > KNN_TC_MAX_K || N > 4096 || C > KMAX || C < 1 || N < 64) return false;
    *NRB = sv_cdiv(N, TM);
    *NKC = sv_cdiv(C, KCH);
    int s = MAX_STAGES;
    const size_t limit = 227 * 1024;
    while (s > 2 && knn_tc_smem(*NRB, *NKC, s) > limit) --s;
    if (knn_tc_smem(*NRB, *NKC, s) > limit) return false;
    *stages = s;
    *pack_bytes = (size_t)B * *NRB * *NKC * CHUNK_BYTES;
    *xx_bytes = (size_t)B * *NRB * TM * sizeof(float);
    return true;
}

// Scratch bytes the tensor-core path needs for this shape; 0 when the shape is not covered.
size_t svnet_knn_tc_workspace(const svnet_view* in, int B, int N, int k)
{
    int NRB, NKC, stages;
    size_t pb, xb;
    if (!knn_tc_plan(in, B, N, k, &NRB, &NKC, &stages, &pb, &xb)) return 0;
    return pb + xb;
}

// Returns 1 if handled, 0 if the caller should use the CUDA-core kernel, < 0 on error.
int svnet_knn_tc_dispatch(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, void* workspace,
                          size_t workspace_bytes, cudaStream_t st)
{
    int NRB, NKC, stages;
    size_t pack_bytes, xx_bytes;
    if (!workspace || !knn_tc_plan(in, B, N, k, &NRB, &NKC, &stages, &pack_bytes, &xx_bytes)) return 0;
    if (workspace_bytes < pack_bytes + xx_bytes || (reinterpret_cast<uintptr_t>(workspace) & 15)) return 0;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* xx = reinterpret_cast<float*>(ws + pack_bytes);
    knn_pack_kernel<<<dim3(NRB, B), TM, 0, st>>>(*in, N, NRB, NKC, ws, xx);
    SV_CHECK_LAUNCH("svnet_knn(pack)");
    knn_tc_args a;
    a.in = *in; a.N = N; a.k = k; a.NRB = NRB; a.NKC = NKC; a.stages = stages;
    a.pack = ws; a.xx = xx; a.idx32 = idx32; a.idx64 = idx64;
    const size_t smem = knn_tc_smem(NRB, NKC, stages);
    SV_CUDA(cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_tc_kernel<<<dim3(NRB, B), NTHREADS, smem, st>>>(a);
    SV_CHECK_LAUNCH("svnet_knn(tcgen05)");
    return 1;
}

// debug: rows, rows re-scored exactly, brute-force rows, survivors (cumulative since the last reset)
extern "C" int svnet_knn_tc_stats(unsigned long long* out12, int reset)
{
    SV_CUDA(cudaMemcpyFromSymbol(out12, g_knn_tc_stats, sizeof(unsigned long long) * 12));
    if (reset) {
        unsigned long long z[12] = {0};
        SV_CUDA(cudaMemcpyToSymbol(g_knn_tc_stats, z, sizeof(z)));
    }
    return SVNET_OK;
}
