// K1 -- kNN graph construction in one kernel.  Replaces sv_util.knn (reference
// models/utils/sv_util.py:19-25), which materialises a BxNxN matrix (bmm + 3 elementwise passes) and
// calls torch.topk.  Here the score tile lives in registers, candidates that beat the row's current
// k-th score are queued in shared memory, and the owning warp merges them into a sorted list held in
// registers (one entry per lane).  Nothing but the (B,N,k) indices is written to HBM.
//
// Arithmetic contract (must match oracle/svnet_oracle.c:orc_knn bit for bit):
//   dot_ij = chain fmaf over channels c ascending, starting from 0
//   xx_i   = the same chain with both operands f_i
//   p_ij   = ((-xx_j) - (-2*dot_ij)) - xx_i
//   order  = larger p first; equal p -> smaller index first
// (zero-padded channels add fmaf(0,0,acc) == acc, so padding a chunk does not change any bit.)
#include "common.cuh"
#include "knn_keys.cuh"

namespace {
using namespace svknn;

constexpr int TI = 64;    // query rows per CTA
constexpr int TJ = 128;   // candidates per tile (== queue capacity per row: a tile can never overflow it)
constexpr int NT = 256;   // threads
constexpr int NW = NT / 32;
constexpr int ROWS_PER_WARP = TI / NW;  // 8
constexpr int MERGE_MIN = 7;            // >= this many survivors in a 32-chunk: sort-merge instead of inserting

__device__ __forceinline__ int swz(int c, int r) { return r ^ ((c & 7) << 2); }

// 4-byte async copy global -> shared; nbytes == 0 writes zeros (padding rows / channels)
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gsrc, int nbytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(nbytes));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// compare-exchange with lane ^ j
__device__ __forceinline__ void cex(kkey_t& k, int j, bool keep_better)
{
    const kkey_t o = shfl_xor_key(k, j);
    if ((k > o) != keep_better) k = o;
}

template <int R, int KC>
__global__ void __launch_bounds__(NT, 2) knn_kernel(svnet_view in, int N, int k, int32_t* __restrict__ idx32,
                                                    int64_t* __restrict__ idx64)
{
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                   // [2][KC][TI]   swizzled, double buffered
    float* Bs = As + 2 * KC * TI;                       // [2][KC][TJ]   swizzled, double buffered
    float* xxi = Bs + 2 * KC * TJ;                      // [TI]
    float* xxj = xxi + TI;                              // [TJ]
    float* thr = xxj + TJ;                              // [TI] current k-th score per row
    int* qcnt = reinterpret_cast<int*>(thr + TI);       // [TI]
    float* qv = reinterpret_cast<float*>(qcnt + TI);    // [TI][TJ] queued scores
    unsigned char* qj = reinterpret_cast<unsigned char*>(qv + TI * TJ);   // [TI][TJ] queued column within the tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int i0 = blockIdx.x * TI;
    const long base = (long)b * N;
    const int C = in.Cs + 3 * in.Cv;
    const int ty = tid >> 4, tx = tid & 15;  // micro tile: rows ty*4..+3, cols tx*4..+3 and 64+tx*4..+3

    TopK<R> L[ROWS_PER_WARP];
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr)
#pragma unroll
        for (int r = 0; r < R; ++r) L[rr].k[r] = 0ull;
    if (tid < TI) { thr[tid] = -INFINITY; qcnt[tid] = 0; }

    const int ntiles = (N + TJ - 1) / TJ;
    const int nch = (C + KC - 1) / KC;
    const int nstages = ntiles * nch;

    // stage s = (candidate tile jt, channel chunk ci): async copy of one chunk of the query rows and of
    // the candidate rows into buffer `buf` (lane = channel, one row per warp iteration)
    auto issue = [&](int s, int buf) {
        if (lane < KC) {
            const int jt = s / nch, c = (s - jt * nch) * KC + lane, j0 = jt * TJ;
            const float* cptr = in.s ? in.s : in.v;   // any valid address for the zero-fill form
            long cstride = 0;
            bool cok = false;
            if (c < in.Cs) { cptr = in.s + c; cstride = in.lds; cok = true; }
            else if (c < C) {
                const int cc = c - in.Cs, x = cc / in.Cv, d = cc - x * in.Cv;
                cptr = in.v + x * in.xs + d; cstride = in.ldv; cok = true;
            }
            float* Ad = As + buf * KC * TI + lane * TI;
            float* Bd = Bs + buf * KC * TJ + lane * TJ;
            const long step = (long)NW * cstride;
            const float* pa = cptr + (base + i0 + warp) * cstride;
#pragma unroll 4
            for (int r = warp; r < TI; r += NW, pa += step) {
                const bool ok = cok && (i0 + r < N);
                cp_async4_zfill(Ad + swz(lane, r), ok ? pa : cptr, ok ? 4 : 0);
            }
            const float* pb = cptr + (base + j0 + warp) * cstride;
#pragma unroll 4
            for (int r = warp; r < TJ; r += NW, pb += step) {
                const bool ok = cok && (j0 + r < N);
                cp_async4_zfill(Bd + swz(lane, r), ok ? pb : cptr, ok ? 4 : 0);
            }
        }
    };

    float acc[4][8];
    float nrm = 0.0f;  // xx_j for tid < TJ; xx_i for TJ <= tid < TJ+TI (first tile only)
    issue(0, 0);
    for (int s = 0; s < nstages; ++s) {
        const int buf = s & 1;
        const int jt = s / nch, ci = s - jt * nch;
        const int j0 = jt * TJ;
        cp_async_wait_all();
        __syncthreads();                 // stage s has landed for everyone; buffer buf^1 is free again
        if (s + 1 < nstages) issue(s + 1, buf ^ 1);
        if (ci == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[a][q] = 0.0f;
            nrm = 0.0f;
        }
        const float* Ab = As + buf * KC * TI;
        const float* Bb = Bs + buf * KC * TJ;
        // squared norms, sequential chain over channels (zero padding is exact)
        if (tid < TJ) {
#pragma unroll
            for (int cc = 0; cc < KC; ++cc) { const float t = Bb[cc * TJ + swz(cc, tid)]; nrm = __fmaf_rn(t, t, nrm); }
        } else if (jt == 0 && tid < TJ + TI) {
#pragma unroll
            for (int cc = 0; cc < KC; ++cc) { const float t = Ab[cc * TI + swz(cc, tid - TJ)]; nrm = __fmaf_rn(t, t, nrm); }
        }
#pragma unroll
        for (int cc = 0; cc < KC; ++cc) {
            const float4 a4 = *reinterpret_cast<const float4*>(&Ab[cc * TI + swz(cc, ty * 4)]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bb[cc * TJ + swz(cc, tx * 4)]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bb[cc * TJ + swz(cc, 64 + tx * 4)]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[a][q] = __fmaf_rn(av[a], bv[q], acc[a][q]);
        }
        if (ci != nch - 1) continue;     // more channel chunks of this tile to come

        if (tid < TJ) xxj[tid] = nrm;
        else if (jt == 0 && tid < TJ + TI) xxi[tid - TJ] = nrm;
        __syncthreads();
        // ---- scores; candidates that reach the row's current k-th score go to the row's queue ----
        {
            float xj[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) xj[q] = -xxj[(q < 4) ? (tx * 4 + q) : (64 + tx * 4 + (q - 4))];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = ty * 4 + a;
                const float xi = xxi[r];
                const float tv = thr[r];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int col = (q < 4) ? (tx * 4 + q) : (64 + tx * 4 + (q - 4));
                    const float inner = -2.0f * acc[a][q];
                    const float p = __fsub_rn(__fsub_rn(xj[q], inner), xi);
                    if (p >= tv && j0 + col < N) {
                        const int slot = atomicAdd(&qcnt[r], 1);
                        qv[r * TJ + slot] = p;
                        qj[r * TJ + slot] = (unsigned char)col;
                    }
                }
            }
        }
        __syncthreads();
        // ---- drain: warp owns rows warp*8 .. warp*8+7.  The eight rows are processed in lock-step,
        //      branch-free, so that their (latency-bound, shuffle-heavy) updates overlap ----
        {
            const int r0 = warp * ROWS_PER_WARP;
            int cnts[ROWS_PER_WARP];
            int maxc = 0;
#pragma unroll
            for (int rr = 0; rr < ROWS_PER_WARP; ++rr) { cnts[rr] = qcnt[r0 + rr]; maxc = max(maxc, cnts[rr]); }
            for (int q0 = 0; q0 < maxc; q0 += 32) {
                kkey_t c[ROWS_PER_WARP];
                unsigned m[ROWS_PER_WARP];
                bool any_merge = false;
#pragma unroll
                for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                    const bool have = q0 + lane < cnts[rr];
                    const int qi = (r0 + rr) * TJ + q0 + lane;
                    c[rr] = have ? make_key(qv[qi], j0 + (int)qj[qi]) : 0ull;
                    const kkey_t worst = shfl_key(L[rr].k[R - 1], (k - 1) & 31);
                    m[rr] = __ballot_sync(SV_FULL, c[rr] > worst);
                    any_merge |= __popc(m[rr]) >= MERGE_MIN;
                }
                if (R == 1 && any_merge) {
                    // batched bitonic sort of the 8 candidate vectors, then merge into the 8 lists
#pragma unroll
                    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
                        for (int j = size >> 1; j > 0; j >>= 1) {
                            const bool keep = ((lane & size) == 0) == ((lane & j) == 0);
#pragma unroll
                            for (int rr = 0; rr < ROWS_PER_WARP; ++rr) cex(c[rr], j, keep);
                        }
#pragma unroll
                    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                        const kkey_t rv = shfl_key(c[rr], 31 - lane);
                        if (rv > L[rr].k[0]) L[rr].k[0] = rv;
                    }
#pragma unroll
                    for (int j = 16; j > 0; j >>= 1)
#pragma unroll
                        for (int rr = 0; rr < ROWS_PER_WARP; ++rr) cex(L[rr].k[0], j, (lane & j) == 0);
                    continue;
                }
                // insertion rounds: one candidate per row per round.  A candidate that no longer beats
                // the (meanwhile improved) k-th entry lands at a position >= k, which is never read.
                for (;;) {
                    unsigned anym = 0u;
#pragma unroll
                    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) anym |= m[rr];
                    if (!anym) break;
#pragma unroll
                    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                        const bool had = m[rr] != 0u;
                        const int src = had ? (__ffs(m[rr]) - 1) : 0;
                        m[rr] &= m[rr] - 1;
                        kkey_t cc = shfl_key(c[rr], src);
                        if (!had) cc = 0ull;
                        topk_insert<R>(L[rr], cc, lane);
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
                const kkey_t worst = shfl_key(L[rr].k[R - 1], (k - 1) & 31);
                if (lane == 0 && cnts[rr] > 0) { qcnt[r0 + rr] = 0; thr[r0 + rr] = key_score(worst); }
            }
        }
        // (the next stage's __syncthreads orders these writes before the next push phase)
    }
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        const int i = i0 + warp * ROWS_PER_WARP + rr;
        if (i >= N) continue;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int pos = r * 32 + lane;
            if (pos < k) {
                const long o = (base + i) * k + pos;
                if (idx32) idx32[o] = key_index(L[rr].k[r]);
                if (idx64) idx64[o] = (int64_t)key_index(L[rr].k[r]);
            }
        }
    }
}

template <int R, int KC>
int launch_knn(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, cudaStream_t st)
{
    const size_t smem = sizeof(float) * (2 * KC * TI + 2 * KC * TJ + TI + TJ + TI + TI + TI * TJ) + TI * TJ;
    dim3 grid(sv_cdiv(N, TI), B);
    SV_CUDA(cudaFuncSetAttribute(knn_kernel<R, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<R, KC><<<grid, NT, smem, st>>>(*in, N, k, idx32, idx64);
    SV_CHECK_LAUNCH("svnet_knn");
    return SVNET_OK;
}

}  // namespace

size_t svnet_knn_tc_workspace(const svnet_view* in, int B, int N, int k);
int svnet_knn_tc_dispatch(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, void* workspace,
                          size_t workspace_bytes, cudaStream_t st);

extern "C" size_t svnet_knn_workspace_bytes(const svnet_view* in, int B, int N, int k)
{
    if (!in || B < 1 || N < 1 || k < 1 || k > N) return 0;
    return svnet_knn_tc_workspace(in, B, N, k);
}

extern "C" int svnet_knn(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, void* stream)
{
    return svnet_knn_ws(in, B, N, k, idx32, idx64, nullptr, 0, stream);
}

extern "C" int svnet_knn_ws(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, void* workspace,
                            size_t workspace_bytes, void* stream)
{
    SV_REQUIRE(in != nullptr, "svnet_knn: null view");
    SV_REQUIRE(B >= 0 && N >= 1, "svnet_knn: bad B=%d N=%d", B, N);
    SV_REQUIRE(k >= 1 && k <= N, "svnet_knn: k=%d out of range for N=%d (selected index k out of range)", k, N);
    SV_REQUIRE(k <= 128, "svnet_knn: k=%d > 128 unsupported", k);
    SV_REQUIRE(in->Cs + 3 * in->Cv >= 1, "svnet_knn: empty feature");
    SV_REQUIRE(in->Cs == 0 || in->s, "svnet_knn: null s");
    SV_REQUIRE(in->Cv == 0 || in->v, "svnet_knn: null v");
    SV_REQUIRE(idx32 || idx64, "svnet_knn: no output buffer");
    if (B == 0) return SVNET_OK;
    const int R = (k + 31) / 32;
    const bool small = (in->Cs + 3 * in->Cv) <= 8;
    cudaStream_t st = sv_stream(stream);
    {   // tensor-core filter + exact re-scoring (knn_tc.cu) for the shapes it covers
        const int handled = svnet_knn_tc_dispatch(in, B, N, k, idx32, idx64, workspace, workspace_bytes, st);
        if (handled != 0) return handled < 0 ? handled : SVNET_OK;
    }
    if (small) {
        if (R == 1) return launch_knn<1, 8>(in, B, N, k, idx32, idx64, st);
        if (R == 2) return launch_knn<2, 8>(in, B, N, k, idx32, idx64, st);
        if (R == 3) return launch_knn<3, 8>(in, B, N, k, idx32, idx64, st);
        return launch_knn<4, 8>(in, B, N, k, idx32, idx64, st);
    }
    if (R == 1) return launch_knn<1, 32>(in, B, N, k, idx32, idx64, st);
    if (R == 2) return launch_knn<2, 32>(in, B, N, k, idx32, idx64, st);
    if (R == 3) return launch_knn<3, 32>(in, B, N, k, idx32, idx64, st);
    return launch_knn<4, 32>(in, B, N, k, idx32, idx64, st);
}
