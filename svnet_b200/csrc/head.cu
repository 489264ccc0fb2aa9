// K4 -- classification head, one CTA per cloud: up to three chained layers
//   [binarised Linear (sign-pack + XNOR/popcount) | fp Linear] (+bias) -> BatchNorm affine -> activation
// with the activations kept in shared memory (reference models/sv_dgcnn_cls.py:76-80,
// models/sv_pointnet_cls.py:76-81; Linear: models/sv_layers.py:29-53).
// Binary layers are integer-exact; small fp layers use one sequential fmaf chain per output (the
// oracle's order), large fp layers a warp-per-output shuffle reduction.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int NT = 512;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SV_FULL, v, o);
    return v;
}

__global__ void __launch_bounds__(NT) head_kernel(svnet_head_params p)
{
    extern __shared__ __align__(16) float sm[];
    // two ping-pong activation buffers + sign words
    int maxc = p.K0;
    for (int l = 0; l < p.nlayers; ++l) maxc = max(maxc, p.layer[l].Cout);
    float* act0 = sm;
    float* act1 = act0 + maxc;
    uint32_t* bits = reinterpret_cast<uint32_t*>(act1 + maxc);
    const int maxw = (maxc + 31) / 32;
    uint32_t* mask = bits + maxw;
    float* wst = reinterpret_cast<float*>(mask + maxw);      // staged weights of a small fp layer, row stride K + 1
    __shared__ int nvalid_s;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < p.K0; c += NT) act0[c] = p.x[(long)b * p.ldx + c];
    __syncthreads();
    float* cur = act0;
    float* nxt = act1;
    int K = p.K0;
    for (int l = 0; l < p.nlayers; ++l) {
        const svnet_head_layer& L = p.layer[l];
        const int Cout = L.Cout;
        if (L.W1b) {
            // sign(x + beta) -> words (sv_layers.py:36-39)
            const int Kw = (K + 31) / 32;
            if (tid == 0) nvalid_s = 0;
            __syncthreads();
            for (int w = warp; w < Kw; w += NT / 32) {
                const int c = w * 32 + lane;
                const float t = (c < K) ? __fadd_rn(cur[c], L.beta[c]) : 0.0f;
                const unsigned pos = __ballot_sync(SV_FULL, t > 0.0f);
                const unsigned nz = __ballot_sync(SV_FULL, t != 0.0f);
                if (lane == 0) { bits[w] = pos; mask[w] = nz; atomicAdd(&nvalid_s, __popc(nz)); }
            }
            __syncthreads();
            const int nvalid = nvalid_s;
            for (int o = tid; o < Cout; o += NT) {
                int mism = 0;
#pragma unroll 32
                for (int w = 0; w < Kw; ++w) mism += __popc((bits[w] ^ __ldg(L.W1b + (long)w * Cout + o)) & mask[w]);
                float y = __fmul_rn((float)(nvalid - 2 * mism), L.scale ? L.scale[o] : 1.0f);
                if (L.bias) y = __fadd_rn(y, L.bias[o]);
                if (L.bn_a) y = __fadd_rn(__fmul_rn(y, L.bn_a[o]), L.bn_c[o]);
                nxt[o] = sv_act(y, L.act);
            }
        } else if ((long)K * Cout <= 32768) {
            // small fp layer: one sequential fmaf chain per output (the oracle's order).  The weights are
            // first staged in shared memory by all threads (coalesced, every load in flight); the chains
            // then run from shared memory instead of Cout serial walks over global rows.
            const int ldw = K + 1;                      // odd stride: the Cout chains hit different banks
            for (int i = tid; i < K * Cout; i += NT) {
                const int o = i / K, c = i - o * K;
                float wv = __ldg(L.W + i);
                if (L.sign_w) wv = (wv > 0.0f) ? 1.0f : ((wv < 0.0f) ? -1.0f : 0.0f);
                wst[o * ldw + c] = wv;
            }
            __syncthreads();
            for (int o = tid; o < Cout; o += NT) {
                const float* w = wst + o * ldw;
                float acc = 0.0f;
#pragma unroll 8
                for (int c = 0; c < K; ++c) acc = __fmaf_rn(cur[c], w[c], acc);
                if (L.scale) acc = __fmul_rn(acc, L.scale[o]);
                if (L.bias) acc = __fadd_rn(acc, L.bias[o]);
                if (L.bn_a) acc = __fadd_rn(__fmul_rn(acc, L.bn_a[o]), L.bn_c[o]);
                nxt[o] = sv_act(acc, L.act);
            }
        } else {
            for (int o = warp; o < Cout; o += NT / 32) {
                const float* w = L.W + (long)o * K;
                float acc = 0.0f;
                for (int c = lane; c < K; c += 32) {
                    float wv = __ldg(w + c);
                    if (L.sign_w) wv = (wv > 0.0f) ? 1.0f : ((wv < 0.0f) ? -1.0f : 0.0f);
                    acc = fmaf(cur[c], wv, acc);
                }
                acc = warp_sum(acc);
                if (lane == 0) {
                    if (L.scale) acc *= L.scale[o];
                    if (L.bias) acc += L.bias[o];
                    if (L.bn_a) acc = acc * L.bn_a[o] + L.bn_c[o];
                    nxt[o] = sv_act(acc, L.act);
                }
            }
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
        K = Cout;
    }
    for (int c = tid; c < K; c += NT) p.out[(long)b * p.ldo + c] = cur[c];
}

}  // namespace

extern "C" int svnet_head_fwd(const svnet_head_params* p, void* stream)
{
    SV_REQUIRE(p && p->x && p->out, "svnet_head_fwd: null pointer");
    SV_REQUIRE(p->nlayers >= 1 && p->nlayers <= 3 && p->K0 >= 1 && p->B >= 0, "svnet_head_fwd: bad shape");
    int maxc = p->K0, K = p->K0;
    size_t wstage = 0;
    for (int l = 0; l < p->nlayers; ++l) {
        const svnet_head_layer& L = p->layer[l];
        if (!L.W1b && (long)K * L.Cout <= 32768) wstage = std::max(wstage, (size_t)L.Cout * (K + 1));
        SV_REQUIRE(L.Cout >= 1, "svnet_head_fwd: layer %d Cout", l);
        SV_REQUIRE(L.W1b ? (L.beta != nullptr) : (L.W != nullptr), "svnet_head_fwd: layer %d needs W1b+beta or W", l);
        SV_REQUIRE((L.bn_a == nullptr) == (L.bn_c == nullptr), "svnet_head_fwd: layer %d bn_a/bn_c", l);
        maxc = L.Cout > maxc ? L.Cout : maxc;
        K = L.Cout;
    }
    SV_REQUIRE(p->ldx >= p->K0 && p->ldo >= K, "svnet_head_fwd: ldx/ldo too small");
    if (p->B == 0) return SVNET_OK;
    const size_t smem = sizeof(float) * ((size_t)2 * maxc + 2 * ((maxc + 31) / 32) + wstage);
    SV_REQUIRE(smem <= 200 * 1024, "svnet_head_fwd: layer too wide");
    if (smem > 48 * 1024) SV_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_kernel<<<p->B, NT, smem, sv_stream(stream)>>>(*p);
    SV_CHECK_LAUNCH("svnet_head_fwd");
    return SVNET_OK;
}
