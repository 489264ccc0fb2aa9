// Tensor-core path for the binary-weight / fp32-activation linears of the vector branch
// (reference models/sv_layers.py:43-49 with bw set and ba unset: linear2 of every binary SVBlock,
// and the per-point P|Q tables of the fused edge kernel).
//
// The activations are split exactly into three bf16 planes  a = hi + mid + lo  (8 + 8 + 8 mantissa
// bits); the weights are +-1, so every product of a plane with a weight is exact and the tensor cores
// (mma.sync m16n8k16, fp32 accumulate) only reorder the fp32 summation.  Rows come in xyz groups of
// three; the three components of a point go to three different MMA row tiles so that one thread ends
// up holding (x, y, z) of the same (point, channel) -- the VectorBN + gate epilogue
// (sv_layers.py:94-100,194) is then thread-local.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int NTH = 128;          // 4 warps
constexpr int WCOLS = 48;         // columns per warp work item (6 n8 tiles)
constexpr int NT8 = WCOLS / 8;

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// smem: W bf16 [Npad][KS] | A planes 3 x [PT*48][KS] bf16,  KS = Kpad + 8 halfs
__global__ void __launch_bounds__(NTH) signlinear_tc_kernel(svnet_gemm_params p, int Kpad, int Npad, int PT, int CB,
                                                            int ntiles)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int KS = Kpad + 8;
    unsigned short* Ws = reinterpret_cast<unsigned short*>(smraw);
    unsigned short* As = Ws + (size_t)Npad * KS;   // plane pl, row r: As[(pl*PT*48 + r)*KS + k]
    const int tid = threadIdx.x, lane = tid & 31, warp = sv_warp_id();
    const int g = lane >> 2, q = lane & 3;

    // ---- weights: sign -> bf16 +-1 (0 for exact zeros), zero padded ----
    for (int n = warp; n < Npad; n += NTH / 32)
        for (int k = lane; k < KS; k += 32) {
            unsigned short v = 0;
            if (n < p.N && k < p.K) {
                const float w = __ldg(p.W + (long)n * p.ldw + k);
                v = (w > 0.0f) ? 0x3F80 : ((w < 0.0f) ? 0xBF80 : 0);
            }
            Ws[(size_t)n * KS + k] = v;
        }
    const long npoints = p.M / 3;
    const int wp = warp / CB, wc = warp - wp * CB;   // this warp's point tile and column block within the CTA pass

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long p0 = (long)tile * PT * 16;
        __syncthreads();   // previous pass fully consumed (and W visible on the first pass)
        // ---- stage A: rows ordered [point tile][x][16 points]; exact 3-way bf16 split ----
        const int rows = PT * 48;
        // four rows per warp iteration, lanes over k (coalesced); all loads issued before any use
        for (int rb = warp * 4; rb < rows; rb += (NTH / 32) * 4) {
            float a[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rb + u;
                const int pt = r / 48, rr = r - pt * 48, x = rr >> 4, pi = rr & 15;
                const long pnt = p0 + pt * 16 + pi;
                const float* arow = p.A + pnt * p.lda_g + (long)x * p.lda_x;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const int k = lane + 32 * kk;
                    a[u][kk] = (r < rows && pnt < npoints && k < p.K) ? __ldg(arow + k) : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = rb + u;
                if (r >= rows) break;
                unsigned short* d0 = As + (size_t)r * KS;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const int k = lane + 32 * kk;
                    if (k >= Kpad) break;
                    const float av = a[u][kk];
                    const unsigned hb = __float_as_uint(av) & 0xFFFF0000u;
                    const float r1 = av - __uint_as_float(hb);
                    const unsigned mb = __float_as_uint(r1) & 0xFFFF0000u;
                    const float r2 = r1 - __uint_as_float(mb);
                    d0[k] = (unsigned short)(hb >> 16);
                    d0[(size_t)rows * KS + k] = (unsigned short)(mb >> 16);
                    d0[(size_t)2 * rows * KS + k] = (unsigned short)(__float_as_uint(r2) >> 16);
                }
            }
        }
        __syncthreads();
        for (int cb0 = 0; cb0 < Npad; cb0 += CB * WCOLS) {
            const int c0 = cb0 + wc * WCOLS;
            if (c0 >= Npad) continue;   // warp-uniform
            float acc[3][NT8][4];
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int j = 0; j < NT8; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[x][j][e] = 0.0f;
            for (int k0 = 0; k0 < Kpad; k0 += 16) {
                uint32_t bf[NT8][2];
#pragma unroll
                for (int j = 0; j < NT8; ++j) {
                    const int n = c0 + j * 8 + g;
                    const unsigned short* wr = Ws + (size_t)min(n, Npad - 1) * KS + k0 + 2 * q;
                    bf[j][0] = *reinterpret_cast<const uint32_t*>(wr);
                    bf[j][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
                    if (n >= Npad) { bf[j][0] = 0u; bf[j][1] = 0u; }
                }
#pragma unroll
                for (int pl = 2; pl >= 0; --pl) {   // small planes first
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        const unsigned short* ar = As + ((size_t)pl * rows + wp * 48 + x * 16 + g) * KS + k0 + 2 * q;
                        uint32_t af[4];
                        af[0] = *reinterpret_cast<const uint32_t*>(ar);
                        af[1] = *reinterpret_cast<const uint32_t*>(ar + 8 * KS);
                        af[2] = *reinterpret_cast<const uint32_t*>(ar + 8);
                        af[3] = *reinterpret_cast<const uint32_t*>(ar + 8 * KS + 8);
#pragma unroll
                        for (int j = 0; j < NT8; ++j) mma_bf16(acc[x][j], af, bf[j]);
                    }
                }
            }
            // ---- epilogue: thread holds (x,y,z) of points g and g+8, columns c0 + 8j + 2q + {0,1} ----
#pragma unroll
            for (int j = 0; j < NT8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = c0 + j * 8 + 2 * q + (e & 1);
                    const long pnt = p0 + wp * 16 + g + ((e >> 1) ? 8 : 0);
                    if (n >= p.N || pnt >= npoints) continue;
                    float w[3];
#pragma unroll
                    for (int x = 0; x < 3; ++x) w[x] = p.colscale ? __fmul_rn(acc[x][j][e], p.colscale[n]) : acc[x][j][e];
                    float* cp = p.C + pnt * p.ldc_g + n;
                    if (p.vbn) {
                        const float nrm = __fadd_rn(
                            __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2]))),
                            1e-6f);
                        const float nb = __fadd_rn(__fmul_rn(nrm, p.bn_a[n]), p.bn_c[n]);
                        const float gt = p.gate ? p.gate[(pnt / p.groups_per_cloud) * p.N + n] : 1.0f;
#pragma unroll
                        for (int x = 0; x < 3; ++x) {
                            float t = __fmul_rn(__fdiv_rn(w[x], nrm), nb);
                            if (p.gate) t = __fmul_rn(t, gt);
                            cp[(long)x * p.ldc_x] = t;
                        }
                    } else {
#pragma unroll
                        for (int x = 0; x < 3; ++x) cp[(long)x * p.ldc_x] = w[x];
                    }
                }
        }
    }
}

}  // namespace

// Returns 1 if handled, 0 if the caller should use the CUDA-core kernel, < 0 on error.
int svnet_signlinear_tc_dispatch(const svnet_gemm_params* p, cudaStream_t st)
{
    if (!p->sign_w || p->G != 3 || p->M % 3 != 0) return 0;
    if (p->bias || p->act != SVNET_ACT_NONE || (!p->vbn && p->bn_a)) return 0;
    if (p->K < 32 || p->K > 96 || p->N > 1024) return 0;   // tiny K: the CUDA-core kernel is as fast
    const char* off = getenv("SVNET_NO_TC");
    if (off && off[0] == '1') return 0;
    const int Kpad = (p->K + 15) / 16 * 16;
    const int Npad = (p->N + 7) / 8 * 8;
    const int ncb = (Npad + WCOLS - 1) / WCOLS;
    const int CB = ncb >= 4 ? 4 : (ncb >= 2 ? 2 : 1);
    const int PT = 4 / CB;
    const long npoints = p->M / 3;
    const int ntiles = (int)((npoints + PT * 16 - 1) / (PT * 16));
    const size_t smem = sizeof(unsigned short) * ((size_t)Npad * (Kpad + 8) + (size_t)3 * PT * 48 * (Kpad + 8));
    if (smem > 200 * 1024) return 0;
    SV_CUDA(cudaFuncSetAttribute(signlinear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int grid = ntiles < 148 * per_sm ? ntiles : 148 * per_sm;
    signlinear_tc_kernel<<<grid, NTH, smem, st>>>(*p, Kpad, Npad, PT, CB, ntiles);
    SV_CHECK_LAUNCH("svnet_linear_rows(tensor core)");
    return 1;
}
