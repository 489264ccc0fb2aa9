// K1 -- kNN graph construction in one kernel: tiled pairwise scores in shared memory + per-row
// top-k kept in registers of the owning warp.  Replaces sv_util.knn (reference
// models/utils/sv_util.py:19-25), which materialises a BxNxN matrix and calls torch.topk.
//
// Arithmetic contract (must match oracle/svnet_oracle.c:orc_knn bit for bit):
//   dot_ij = chain fmaf over channels c ascending, starting from 0
//   xx_i   = the same chain with both operands f_i
//   p_ij   = ((-xx_j) - (-2*dot_ij)) - xx_i
//   order  = larger p first; equal p -> smaller index first
#include "common.cuh"

namespace {

constexpr int TI = 64;    // query rows per CTA
constexpr int TJ = 128;   // candidates per tile
constexpr int KC = 32;    // channels per smem chunk
constexpr int NT = 256;   // threads
constexpr int ROWS_PER_WARP = TI / (NT / 32);  // 8

__device__ __forceinline__ int swz(int c, int r) { return r ^ ((c & 7) << 2); }

template <int R>
struct TopK {
    float v[R];
    int i[R];
};

template <int R>
__device__ __forceinline__ void topk_insert(TopK<R>& L, float cv, int cj, int lane)
{
    // position = number of entries that rank before the candidate
    int P = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        bool better = (L.v[r] > cv) || (L.v[r] == cv && L.i[r] < cj);
        P += __popc(__ballot_sync(SV_FULL, better));
    }
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        float uv = __shfl_up_sync(SV_FULL, L.v[r], 1);
        int ui = __shfl_up_sync(SV_FULL, L.i[r], 1);
        if (r > 0) {
            float pv = __shfl_sync(SV_FULL, L.v[r - 1], 31);
            int pi = __shfl_sync(SV_FULL, L.i[r - 1], 31);
            if (lane == 0) { uv = pv; ui = pi; }
        }
        int pos = r * 32 + lane;
        if (pos > P) { L.v[r] = uv; L.i[r] = ui; }
        else if (pos == P) { L.v[r] = cv; L.i[r] = cj; }
    }
}

template <int R>
__global__ void __launch_bounds__(NT, 2) knn_kernel(svnet_view in, int N, int k, int32_t* __restrict__ idx32,
                                                 int64_t* __restrict__ idx64)
{
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                 // [KC][TI]   swizzled
    float* Bs = As + KC * TI;         // [KC][TJ]   swizzled
    float* D = Bs + KC * TJ;          // [TI][TJ]
    float* xxi = D + TI * TJ;         // [TI]
    float* xxj = xxi + TI;            // [TJ]

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
        for (int r = 0; r < R; ++r) { L[rr].v[r] = -INFINITY; L[rr].i[r] = 0x7fffffff; }

    const int ntiles = (N + TJ - 1) / TJ;
    for (int jt = 0; jt < ntiles; ++jt) {
        const int j0 = jt * TJ;
        float acc[4][8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[a][q] = 0.0f;
        float nrm = 0.0f;  // xx_j for tid < TJ; xx_i for TJ <= tid < TJ+TI (first tile only)

        for (int c0 = 0; c0 < C; c0 += KC) {
            __syncthreads();  // previous chunk fully consumed
            // load chunk: warp loads one row's KC channels (lane = channel)
            const int c = c0 + lane;
            for (int r = warp; r < TI + TJ; r += NT / 32) {
                float val = 0.0f;
                if (r < TI) {
                    int i = i0 + r;
                    if (c < C && i < N) val = sv_feat(in, base + i, c);
                    As[lane * TI + swz(lane, r)] = val;
                } else {
                    int j = j0 + (r - TI);
                    if (c < C && j < N) val = sv_feat(in, base + j, c);
                    Bs[lane * TJ + swz(lane, r - TI)] = val;
                }
            }
            __syncthreads();
            const int kc = min(KC, C - c0);
            // squared norms, sequential chain over channels
            if (tid < TJ) {
                for (int cc = 0; cc < kc; ++cc) { float t = Bs[cc * TJ + swz(cc, tid)]; nrm = __fmaf_rn(t, t, nrm); }
            } else if (jt == 0 && tid < TJ + TI) {
                for (int cc = 0; cc < kc; ++cc) { float t = As[cc * TI + swz(cc, tid - TJ)]; nrm = __fmaf_rn(t, t, nrm); }
            }
#pragma unroll 8
            for (int cc = 0; cc < kc; ++cc) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[cc * TI + swz(cc, ty * 4)]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cc * TJ + swz(cc, tx * 4)]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cc * TJ + swz(cc, 64 + tx * 4)]);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[a][q] = __fmaf_rn(av[a], bv[q], acc[a][q]);
            }
        }
        if (tid < TJ) xxj[tid] = nrm;
        else if (jt == 0 && tid < TJ + TI) xxi[tid - TJ] = nrm;
        __syncthreads();
        // scores -> D
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = ty * 4 + a;
            const float xi = xxi[r];
            float p[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int cidx = (q < 4) ? (tx * 4 + q) : (64 + tx * 4 + (q - 4));
                const float inner = -2.0f * acc[a][q];
                const float t = __fsub_rn(-xxj[cidx], inner);
                p[q] = __fsub_rn(t, xi);
            }
            *reinterpret_cast<float4*>(&D[r * TJ + tx * 4]) = make_float4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<float4*>(&D[r * TJ + 64 + tx * 4]) = make_float4(p[4], p[5], p[6], p[7]);
        }
        __syncthreads();
        // selection: warp owns rows warp*8 .. warp*8+7
#pragma unroll
        for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
            const int r = warp * ROWS_PER_WARP + rr;
            if (i0 + r >= N) continue;  // warp-uniform
#pragma unroll
            for (int q = 0; q < TJ / 32; ++q) {
                const int j = j0 + q * 32 + lane;
                const float p = D[r * TJ + q * 32 + lane];
                float wv = __shfl_sync(SV_FULL, L[rr].v[R - 1], (k - 1) & 31);
                int wi = __shfl_sync(SV_FULL, L[rr].i[R - 1], (k - 1) & 31);
                bool pass = (j < N) && ((p > wv) || (p == wv && j < wi));
                unsigned m = __ballot_sync(SV_FULL, pass);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float cv = __shfl_sync(SV_FULL, p, src);
                    const int cj = __shfl_sync(SV_FULL, j, src);
                    if ((cv > wv) || (cv == wv && cj < wi)) {  // warp-uniform
                        topk_insert<R>(L[rr], cv, cj, lane);
                        wv = __shfl_sync(SV_FULL, L[rr].v[R - 1], (k - 1) & 31);
                        wi = __shfl_sync(SV_FULL, L[rr].i[R - 1], (k - 1) & 31);
                    }
                }
            }
        }
    }
    // write indices
#pragma unroll
    for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
        const int i = i0 + warp * ROWS_PER_WARP + rr;
        if (i >= N) continue;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int pos = r * 32 + lane;
            if (pos < k) {
                const long o = (base + i) * k + pos;
                if (idx32) idx32[o] = L[rr].i[r];
                if (idx64) idx64[o] = (int64_t)L[rr].i[r];
            }
        }
    }
}

}  // namespace

extern "C" int svnet_knn(const svnet_view* in, int B, int N, int k, int32_t* idx32, int64_t* idx64, void* stream)
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
    const size_t smem = sizeof(float) * (KC * TI + KC * TJ + TI * TJ + TI + TJ);
    dim3 grid(sv_cdiv(N, TI), B);
    const int R = (k + 31) / 32;
    cudaStream_t st = sv_stream(stream);
#define LAUNCH(RR)                                                                                          \
    do {                                                                                                    \
        SV_CUDA(cudaFuncSetAttribute(knn_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        knn_kernel<RR><<<grid, NT, smem, st>>>(*in, N, k, idx32, idx64);                                    \
    } while (0)
    if (R == 1) LAUNCH(1);
    else if (R == 2) LAUNCH(2);
    else if (R == 3) LAUNCH(3);
    else LAUNCH(4);
#undef LAUNCH
    SV_CHECK_LAUNCH("svnet_knn");
    return SVNET_OK;
}
