// First edge layer, fully fused: get_graph_feature[_cross] + init_scalar + SVBlock(FP) + svpool.
// Reference: models/sv_dgcnn_cls.py:49-53, models/sv_pointnet_cls.py:35-39,
// models/utils/sv_util.py:28-88,118-132, models/sv_layers.py:111-129,172-196.
//
// One warp per centre point, one lane per edge (k > 32 takes several rounds).  Every edge quantity
// lives in registers; the scalar branch reproduces the oracle's sequential fmaf chains so the
// pooled scalars are bit-identical; the vector mean over k is a shuffle tree (tolerance-level).
#include "common.cuh"

namespace {

constexpr int WARPS = 8;

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SV_FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SV_FULL, v, o);
    return v;
}

// packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): two IEEE-rounded operations per issue slot, the same bits
// per element as the scalar instructions
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
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

template <int NV>
__global__ void __launch_bounds__(WARPS * 32) edge_xyz_kernel(svnet_edge_xyz_params p)
{
    constexpr int KU = 6 * NV;
    extern __shared__ float sm[];
    float* W1 = sm;                      // [Cout][KU]
    float* a1 = W1 + p.Cout * KU;        // [Cout]
    float* c1 = a1 + p.Cout;
    float* W2 = c1 + p.Cout;             // [Cvo][NV]
    float* a2 = W2 + p.Cvo * NV;
    float* c2 = a2 + p.Cvo;
    float* Wi = c2 + p.Cvo;              // [3][NV]
    float* Wz = Wi + 3 * NV;             // [3][NV]
    for (int i = threadIdx.x; i < p.Cout * KU; i += blockDim.x) W1[i] = p.W1[i];
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) { a1[i] = p.bn1_a[i]; c1[i] = p.bn1_c[i]; }
    for (int i = threadIdx.x; i < p.Cvo * NV; i += blockDim.x) W2[i] = p.W2[i];
    for (int i = threadIdx.x; i < p.Cvo; i += blockDim.x) { a2[i] = p.bn2_a[i]; c2[i] = p.bn2_c[i]; }
    for (int i = threadIdx.x; i < 3 * NV; i += blockDim.x) { Wi[i] = p.Winit[i]; Wz[i] = p.Wz[i]; }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long r = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= (long)p.B * p.N) return;
    const int b = (int)(r / p.N);
    const float xi[3] = {p.xyz[r * 3], p.xyz[r * 3 + 1], p.xyz[r * 3 + 2]};

    float smax[2] = {-INFINITY, -INFINITY};  // out o = lane, lane+32
    float vsum[3] = {0.0f, 0.0f, 0.0f};      // out c = lane

    for (int e0 = 0; e0 < p.k; e0 += 32) {
        const int e = e0 + lane;
        const bool valid = e < p.k;
        const long j = (long)b * p.N + (valid ? p.idx[r * p.k + e] : 0);
        const float xj[3] = {p.xyz[j * 3], p.xyz[j * 3 + 1], p.xyz[j * 3 + 2]};
        float ve[3][NV];
#pragma unroll
        for (int a = 0; a < 3; ++a) { ve[a][0] = __fsub_rn(xj[a], xi[a]); ve[a][1] = xi[a]; }
        if (NV == 3) {
            ve[0][NV - 1] = __fsub_rn(__fmul_rn(xj[1], xi[2]), __fmul_rn(xj[2], xi[1]));
            ve[1][NV - 1] = __fsub_rn(__fmul_rn(xj[2], xi[0]), __fmul_rn(xj[0], xi[2]));
            ve[2][NV - 1] = __fsub_rn(__fmul_rn(xj[0], xi[1]), __fmul_rn(xj[1], xi[0]));
        }
        float u[KU];
        // init_scalar and the block's own v2s: same form, different weights
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const float* W = pass == 0 ? Wi : Wz;
            float z[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    float zz = 0.0f;
#pragma unroll
                    for (int d = 0; d < NV; ++d) zz = __fmaf_rn(ve[a][d], W[m * NV + d], zz);
                    z[a][m] = zz;
                }
#pragma unroll
            for (int d = 0; d < NV; ++d)
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    float q = __fmul_rn(ve[0][d], z[0][m]);
                    q = __fmaf_rn(ve[1][d], z[1][m], q);
                    q = __fmaf_rn(ve[2][d], z[2][m], q);
                    u[pass * 3 * NV + d * 3 + m] = q;
                }
        }
        // scalar branch: linear1 (fp) -> bn1 -> leaky -> max over edges
        for (int o = 0; o < p.Cout; ++o) {
            float y = 0.0f;
#pragma unroll
            for (int t = 0; t < KU; ++t) y = __fmaf_rn(u[t], W1[o * KU + t], y);
            y = __fadd_rn(__fmul_rn(y, a1[o]), c1[o]);
            y = y > 0.0f ? y : __fmul_rn(0.2f, y);
            y = warp_max(valid ? y : -INFINITY);
            if ((o & 31) == lane) smax[o >> 5] = fmaxf(smax[o >> 5], y);
        }
        // vector branch: linear2 (fp) -> VectorBN -> sum over edges
        for (int c = 0; c < p.Cvo; ++c) {
            float w[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float t = 0.0f;
#pragma unroll
                for (int d = 0; d < NV; ++d) t = __fmaf_rn(ve[a][d], W2[c * NV + d], t);
                w[a] = t;
            }
            const float n = __fadd_rn(
                __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2]))),
                1e-6f);
            const float nb = __fadd_rn(__fmul_rn(n, a2[c]), c2[c]);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float t = valid ? __fmul_rn(__fdiv_rn(w[a], n), nb) : 0.0f;
                t = warp_sum(t);
                if (c == lane) vsum[a] += t;
            }
        }
    }
    for (int o = lane; o < p.Cout; o += 32) p.out.s[r * p.out.lds + o] = smax[o >> 5];
    if (lane < p.Cvo) {
        const float g = p.gate[(long)b * p.Cvo + lane];
#pragma unroll
        for (int a = 0; a < 3; ++a) p.out.v[r * p.out.ldv + a * p.out.xs + lane] = (vsum[a] / (float)p.k) * g;
    }
}

// Shape-specialised variant: 8 centre points per warp, 4 lanes per point, every lane walks its share of
// the point's edges sequentially and keeps the running max / sum of all outputs in registers; only
// two shuffle steps per output at the very end.  All 32 lanes stay busy (the generic kernel above
// uses one lane per edge: 20 of 32 at k=20) and there is no per-edge cross-lane reduction.
// Scalar branch: the sequential chains of outputs (o, o + 1) advance together as packed fp32 pairs (FFMA2: the same
// IEEE fma per element, half the issue slots); weights sit in shared memory as [o / 2][t][o & 1] so that one 16-byte
// load feeds two chain steps of a pair; LeakyReLU(y) = max(y, 0.2 y) (slope < 1: identical bits, signed zeros included).
// Vector branch: tolerance-level arithmetic as in edge_vector.cuh (approximate sqrt / reciprocal).
template <int NV, int COUT, int CVO>
__global__ void __launch_bounds__(WARPS * 32, 2) edge_xyz_fast_kernel(svnet_edge_xyz_params p)
{
    constexpr int KU = 6 * NV;
    static_assert(COUT % 2 == 0 && KU % 2 == 0, "packed pairs");
    __shared__ __align__(16) float W1[COUT * KU];                 // [o / 2][t][o & 1]
    __shared__ __align__(8) float a1[COUT], c1[COUT];
    __shared__ float W2[CVO * NV], a2[CVO], c2[CVO], Wi[3 * NV], Wz[3 * NV];
    for (int i = threadIdx.x; i < COUT * KU; i += blockDim.x) {
        const int o = i / KU, t = i - o * KU;
        W1[((o >> 1) * KU + t) * 2 + (o & 1)] = p.W1[i];
    }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) { a1[i] = p.bn1_a[i]; c1[i] = p.bn1_c[i]; }
    for (int i = threadIdx.x; i < CVO * NV; i += blockDim.x) W2[i] = p.W2[i];
    for (int i = threadIdx.x; i < CVO; i += blockDim.x) { a2[i] = p.bn2_a[i]; c2[i] = p.bn2_c[i]; }
    for (int i = threadIdx.x; i < 3 * NV; i += blockDim.x) { Wi[i] = p.Winit[i]; Wz[i] = p.Wz[i]; }
    __syncthreads();

    const int lane = threadIdx.x & 31, sub = lane & 3;
    const long r = ((long)blockIdx.x * WARPS + (threadIdx.x >> 5)) * 8 + (lane >> 2);
    const bool rok = r < (long)p.B * p.N;
    const long rr = rok ? r : 0;
    const int b = (int)(rr / p.N);
    const float xi[3] = {p.xyz[rr * 3], p.xyz[rr * 3 + 1], p.xyz[rr * 3 + 2]};

    float smax[COUT], vsum[3][CVO];
#pragma unroll
    for (int o = 0; o < COUT; ++o) smax[o] = -INFINITY;
#pragma unroll
    for (int c = 0; c < CVO; ++c) { vsum[0][c] = 0.0f; vsum[1][c] = 0.0f; vsum[2][c] = 0.0f; }

    // neighbour coordinates arrive one edge ahead, indices two edges ahead (the gathers are L2 latency)
    const int* irow = p.idx + rr * p.k;
    const float* xc = p.xyz + (long)b * p.N * 3;
    int j1 = sub + 4 < p.k ? __ldg(irow + sub + 4) : 0;
    float xn[3];
    {
        const int j0 = sub < p.k ? __ldg(irow + sub) : 0;
        xn[0] = __ldg(xc + j0 * 3); xn[1] = __ldg(xc + j0 * 3 + 1); xn[2] = __ldg(xc + j0 * 3 + 2);
    }
    for (int e = sub; e < p.k; e += 4) {
        asm volatile("" ::: "memory");   // keep the (loop-invariant) weight loads inside the loop: no register blow-up
        const float xj[3] = {xn[0], xn[1], xn[2]};
        xn[0] = __ldg(xc + j1 * 3); xn[1] = __ldg(xc + j1 * 3 + 1); xn[2] = __ldg(xc + j1 * 3 + 2);
        j1 = e + 8 < p.k ? __ldg(irow + e + 8) : 0;
        float ve[3][NV];
#pragma unroll
        for (int a = 0; a < 3; ++a) { ve[a][0] = __fsub_rn(xj[a], xi[a]); ve[a][1] = xi[a]; }
        if (NV == 3) {
            ve[0][NV - 1] = __fsub_rn(__fmul_rn(xj[1], xi[2]), __fmul_rn(xj[2], xi[1]));
            ve[1][NV - 1] = __fsub_rn(__fmul_rn(xj[2], xi[0]), __fmul_rn(xj[0], xi[2]));
            ve[2][NV - 1] = __fsub_rn(__fmul_rn(xj[0], xi[1]), __fmul_rn(xj[1], xi[0]));
        }
        float u[KU];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const float* W = pass == 0 ? Wi : Wz;
            float z[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    float zz = 0.0f;
#pragma unroll
                    for (int d = 0; d < NV; ++d) zz = __fmaf_rn(ve[a][d], W[m * NV + d], zz);
                    z[a][m] = zz;
                }
#pragma unroll
            for (int d = 0; d < NV; ++d)
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    float q = __fmul_rn(ve[0][d], z[0][m]);
                    q = __fmaf_rn(ve[1][d], z[1][m], q);
                    q = __fmaf_rn(ve[2][d], z[2][m], q);
                    u[pass * 3 * NV + d * 3 + m] = q;
                }
        }
        f32x2 up[KU];
#pragma unroll
        for (int t = 0; t < KU; ++t) up[t] = pk2(u[t], u[t]);
        const f32x2 slope = pk2(0.2f, 0.2f);
#pragma unroll
        for (int o2 = 0; o2 < COUT / 2; ++o2) {
            f32x2 y = pk2(0.0f, 0.0f);
#pragma unroll
            for (int t = 0; t < KU; t += 2) {
                const float4 w = *reinterpret_cast<const float4*>(W1 + (o2 * KU + t) * 2);
                y = fma2(up[t], pk2(w.x, w.y), y);
                y = fma2(up[t + 1], pk2(w.z, w.w), y);
            }
            const float2 a = *reinterpret_cast<const float2*>(a1 + 2 * o2), c = *reinterpret_cast<const float2*>(c1 + 2 * o2);
            // BN as separately rounded multiply and add (scalar intrinsics: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2)
            float y0, y1, l0, l1;
            upk2(y, y0, y1);
            y0 = __fadd_rn(__fmul_rn(y0, a.x), c.x);
            y1 = __fadd_rn(__fmul_rn(y1, a.y), c.y);
            upk2(mul2(slope, pk2(y0, y1)), l0, l1);
            smax[2 * o2] = fmaxf(smax[2 * o2], fmaxf(y0, l0));
            smax[2 * o2 + 1] = fmaxf(smax[2 * o2 + 1], fmaxf(y1, l1));
        }
#pragma unroll
        for (int c = 0; c < CVO; ++c) {
            float w[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float t = 0.0f;
#pragma unroll
                for (int d = 0; d < NV; ++d) t = __fmaf_rn(ve[a][d], W2[c * NV + d], t);
                w[a] = t;
            }
            // (n a2 + c2) / n with n = |w| + 1e-6
            const float s2 = fmaf(w[2], w[2], fmaf(w[1], w[1], w[0] * w[0]));
            const float t = fmaf(c2[c], fast_rcp(fast_sqrt(s2) + 1e-6f), a2[c]);
#pragma unroll
            for (int a = 0; a < 3; ++a) vsum[a][c] = fmaf(w[a], t, vsum[a][c]);
        }
    }
    // combine the four lanes of a point; lane `sub` stores outputs o = sub, sub+4, ...
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        float m = smax[o];
        m = fmaxf(m, __shfl_xor_sync(SV_FULL, m, 1));
        m = fmaxf(m, __shfl_xor_sync(SV_FULL, m, 2));
        if (rok && (o & 3) == sub) p.out.s[r * p.out.lds + o] = m;
    }
    const float inv_k = 1.0f / (float)p.k;
#pragma unroll
    for (int c = 0; c < CVO; ++c) {
        const float g = p.gate[(long)b * CVO + c];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float t = vsum[a][c];
            t += __shfl_xor_sync(SV_FULL, t, 1);
            t += __shfl_xor_sync(SV_FULL, t, 2);
            if (rok && ((a * CVO + c) & 3) == sub) p.out.v[r * p.out.ldv + a * p.out.xs + c] = (t * inv_k) * g;
        }
    }
}

}  // namespace

extern "C" int svnet_edge_xyz_fwd(const svnet_edge_xyz_params* p, void* stream)
{
    SV_REQUIRE(p, "svnet_edge_xyz_fwd: null params");
    SV_REQUIRE(p->xyz && p->idx && p->Winit && p->Wz && p->W1 && p->bn1_a && p->bn1_c && p->W2 && p->bn2_a &&
                   p->bn2_c && p->gate && p->out.s && p->out.v,
               "svnet_edge_xyz_fwd: null pointer");
    SV_REQUIRE(p->nv == 2 || p->nv == 3, "svnet_edge_xyz_fwd: nv=%d", p->nv);
    SV_REQUIRE(p->B >= 0 && p->N >= 1 && p->k >= 1, "svnet_edge_xyz_fwd: bad shape");
    SV_REQUIRE(p->Cout >= 1 && p->Cout <= 64 && p->Cvo >= 1 && p->Cvo <= 32,
               "svnet_edge_xyz_fwd: Cout=%d (<=64) Cvo=%d (<=32) unsupported", p->Cout, p->Cvo);
    const long P = (long)p->B * p->N;
    if (P == 0) return SVNET_OK;
    if (P > 0) {
        const int gridf = sv_cdiv(P, WARPS * 8);
        bool done = true;
        if (p->nv == 2 && p->Cout == 32 && p->Cvo == 10) edge_xyz_fast_kernel<2, 32, 10><<<gridf, WARPS * 32, 0, sv_stream(stream)>>>(*p);
        else if (p->nv == 2 && p->Cout == 32 && p->Cvo == 16) edge_xyz_fast_kernel<2, 32, 16><<<gridf, WARPS * 32, 0, sv_stream(stream)>>>(*p);
        else if (p->nv == 3 && p->Cout == 32 && p->Cvo == 10) edge_xyz_fast_kernel<3, 32, 10><<<gridf, WARPS * 32, 0, sv_stream(stream)>>>(*p);
        else done = false;
        if (done) {
            SV_CHECK_LAUNCH("svnet_edge_xyz_fwd(fast)");
            return SVNET_OK;
        }
    }
    const size_t smem = sizeof(float) * ((size_t)p->Cout * 6 * p->nv + 2 * p->Cout + (size_t)p->Cvo * p->nv + 2 * p->Cvo + 6 * p->nv);
    const int grid = sv_cdiv(P, WARPS);
    if (p->nv == 2) edge_xyz_kernel<2><<<grid, WARPS * 32, smem, sv_stream(stream)>>>(*p);
    else edge_xyz_kernel<3><<<grid, WARPS * 32, smem, sv_stream(stream)>>>(*p);
    SV_CHECK_LAUNCH("svnet_edge_xyz_fwd");
    return SVNET_OK;
}
