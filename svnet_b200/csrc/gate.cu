// SVBlock gate: g = sigmoid(G2 relu(G1 mean(s)))  (reference models/sv_layers.py:156-161,179-183).
// One CTA per cloud; the mean over the block's input rows is reduced in a fixed order (no float
// atomics), then the two tiny linears run warp-per-output.
//
//   rows variant : mean over materialised rows
//   edge variant : the rows are the N*k edges [s_j - s_i | s_i] of get_graph_feature_sv
//                  (sv_util.py:114); sum_e s_j = sum_j indeg(j) s_j, so only the per-point table and
//                  the kNN indices are read
//   xyz variant  : the rows are init_scalar([x_j - x_i | x_i (| x_j x x_i)]) of the first layer
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SV_FULL, v, o);
    return v;
}

// One tiny linear, warp per output: a warp takes OB consecutive outputs at a time so that the weight loads of OB rows
// are in flight together (one L2 latency per group instead of one per output: the gate of conv5, 85 + 170 outputs of
// 256 / 85 inputs, took 17 of the kernel's 37 us that way).  Per output the summation order is unchanged: lane-strided
// partial sums, then the shuffle tree.
template <int OB, typename F>
__device__ __forceinline__ void gate_linear(const float* x, int Cin, const float* __restrict__ W, int Cout, F&& store)
{
    const int lane = threadIdx.x & 31, warp = sv_warp_id();
    for (int j0 = warp * OB; j0 < Cout; j0 += (NT / 32) * OB) {
        float acc[OB];
#pragma unroll
        for (int u = 0; u < OB; ++u) acc[u] = 0.0f;
#pragma unroll 2
        for (int c = lane; c < Cin; c += 32) {
            const float xv = x[c];
#pragma unroll
            for (int u = 0; u < OB; ++u)
                if (j0 + u < Cout) acc[u] = fmaf(xv, __ldg(W + (long)(j0 + u) * Cin + c), acc[u]);
        }
#pragma unroll
        for (int u = 0; u < OB; ++u) acc[u] = warp_sum(acc[u]);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < OB; ++u)
                if (j0 + u < Cout) store(j0 + u, acc[u]);
        }
    }
}

// mean[Cin] in smem -> gate[Co] in global.  h is smem scratch [H].
__device__ void gate_mlp(const float* mean, int Cin, const float* __restrict__ G1, const float* __restrict__ G2, int H,
                         int Co, float* h, float* __restrict__ gate)
{
    gate_linear<8>(mean, Cin, G1, H, [&](int j, float acc) { h[j] = acc > 0.0f ? acc : 0.0f; });
    __syncthreads();
    gate_linear<8>(h, H, G2, Co, [&](int o, float acc) { gate[o] = sv_sigmoid(acc); });
}

// smem: mean[Cs] | h[H] | part[8][32]
__global__ void __launch_bounds__(NT) gate_rows_kernel(const float* __restrict__ s, int lds, int Cs, long rows,
                                                       const float* __restrict__ G1, const float* __restrict__ G2,
                                                       int H, int Co, float* __restrict__ gate)
{
    extern __shared__ float sm[];
    float* mean = sm;
    float* h = mean + Cs;
    float* part = h + H;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, rg = sv_warp_id();
    const float* p = s + (long)b * rows * lds;
    for (int c0 = 0; c0 < Cs; c0 += 32) {
        const int c = c0 + lane;
        float acc = 0.0f;
        if (c < Cs)
            for (long r = rg; r < rows; r += 8) acc += p[r * lds + c];
        part[rg * 32 + lane] = acc;
        __syncthreads();
        if (rg == 0 && c < Cs) {
            float t = part[lane];
            for (int g = 1; g < 8; ++g) t += part[g * 32 + lane];
            mean[c] = t / (float)rows;
        }
        __syncthreads();
    }
    gate_mlp(mean, Cs, G1, G2, H, Co, h, gate + (long)b * Co);
}

// Cluster variant for many rows per cloud (conv5: N rows x 256 channels): a thread-block cluster of
// CL CTAs per cloud splits the rows, partial column sums meet in distributed shared memory and rank
// 0 finishes (fixed summation order -> deterministic).  smem: psum[Cs] | mean[Cs] | h[H] | part[8][32]
constexpr int CL = 8;
__global__ void __launch_bounds__(NT) gate_rows_cluster_kernel(const float* __restrict__ s, int lds, int Cs, long rows,
                                                               const float* __restrict__ G1,
                                                               const float* __restrict__ G2, int H, int Co,
                                                               float* __restrict__ gate)
{
    extern __shared__ float sm[];
    float* psum = sm;
    float* mean = psum + Cs;
    float* h = mean + Cs;
    float* part = h + H;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x / CL;
    const int lane = threadIdx.x & 31, rg = sv_warp_id();
    const long r_lo = rows * rank / CL, r_hi = rows * (rank + 1) / CL;
    const float* p = s + (long)b * rows * lds;
    // one pass over this CTA's rows per 256-column super block: 8 independent accumulators per thread keep
    // 8 x unroll loads in flight (the per-column summation order is unchanged: warp rg adds its rows in
    // order, the warps are added in order, then the ranks)
    for (int cb = 0; cb < Cs; cb += 256) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
#pragma unroll 4
        for (long r = r_lo + rg; r < r_hi; r += 8) {
            const float* pr = p + r * lds + cb + lane;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (cb + 32 * u + lane < Cs) acc[u] += pr[32 * u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = cb + 32 * u + lane;
            __syncthreads();
            part[rg * 32 + lane] = acc[u];
            __syncthreads();
            if (rg == 0 && c < Cs) {
                float t = part[lane];
                for (int g = 1; g < 8; ++g) t += part[g * 32 + lane];
                psum[c] = t;
            }
        }
    }
    cluster.sync();
    if (rank == 0) {
        for (int c = threadIdx.x; c < Cs; c += NT) {
            float t = 0.0f;
            for (unsigned q = 0; q < CL; ++q) t += cluster.map_shared_rank(psum, q)[c];
            mean[c] = t / (float)rows;
        }
        __syncthreads();
    }
    cluster.sync();   // remote psum must stay alive until rank 0 has read it
    if (rank == 0) gate_mlp(mean, Cs, G1, G2, H, Co, h, gate + (long)b * Co);
}

// smem: mean[2Cs] | h[H] | part[2][8][32] | indeg[N]
__global__ void __launch_bounds__(NT) gate_edge_kernel(svnet_view in, const int32_t* __restrict__ idx, int N, int k,
                                                       const float* __restrict__ G1, const float* __restrict__ G2,
                                                       int H, int Co, float* __restrict__ gate)
{
    extern __shared__ float sm[];
    const int Cs = in.Cs;
    float* mean = sm;
    float* h = mean + 2 * Cs;
    float* part = h + H;
    int* indeg = reinterpret_cast<int*>(part + 2 * 8 * 32);
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, rg = sv_warp_id();
    for (int i = threadIdx.x; i < N; i += NT) indeg[i] = 0;
    __syncthreads();
    const int32_t* ib = idx + (long)b * N * k;
    for (long t = threadIdx.x; t < (long)N * k; t += NT) atomicAdd(&indeg[ib[t]], 1);
    __syncthreads();
    const float* p = in.s + (long)b * N * in.lds;
    for (int c0 = 0; c0 < Cs; c0 += 32) {
        const int c = c0 + lane;
        float accd = 0.0f, accp = 0.0f;
        if (c < Cs)
            for (int r = rg; r < N; r += 8) {
                float t = p[(long)r * in.lds + c];
                accd = fmaf((float)indeg[r], t, accd);
                accp += t;
            }
        part[rg * 32 + lane] = accd;
        part[256 + rg * 32 + lane] = accp;
        __syncthreads();
        if (rg == 0 && c < Cs) {
            float td = part[lane], tp = part[256 + lane];
            for (int g = 1; g < 8; ++g) { td += part[g * 32 + lane]; tp += part[256 + g * 32 + lane]; }
            const float mp = tp / (float)N;
            mean[c] = td / ((float)N * (float)k) - mp;   // mean_e (s_j - s_i)
            mean[Cs + c] = mp;                           // mean_e s_i
        }
        __syncthreads();
    }
    gate_mlp(mean, 2 * Cs, G1, G2, H, Co, h, gate + (long)b * Co);
}

// Cluster variant of gate_edge: CL CTAs per cloud.  Each CTA histograms its slice of the kNN indices,
// the partial histograms are summed through distributed shared memory for the CTA's slice of rows,
// partial column sums meet in rank 0 (fixed order -> deterministic).
// smem: psum[2Cs] | mean[2Cs] | h[H] | part[2][8][32] | hist[N] | indeg[N/CL + 1]
__global__ void __launch_bounds__(NT) gate_edge_cluster_kernel(svnet_view in, const int32_t* __restrict__ idx, int N, int k,
                                                               const float* __restrict__ G1,
                                                               const float* __restrict__ G2, int H, int Co,
                                                               float* __restrict__ gate)
{
    extern __shared__ float sm[];
    const int Cs = in.Cs;
    float* psum = sm;
    float* mean = psum + 2 * Cs;
    float* h = mean + 2 * Cs;
    float* part = h + H;
    int* hist = reinterpret_cast<int*>(part + 2 * 8 * 32);
    int* indeg = hist + N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x / CL;
    const int lane = threadIdx.x & 31, rg = sv_warp_id();
    for (int i = threadIdx.x; i < N; i += NT) hist[i] = 0;
    __syncthreads();
    const long E = (long)N * k;
    const long e_lo = E * rank / CL, e_hi = E * (rank + 1) / CL;
    const int32_t* ib = idx + (long)b * E;
    {   // the index loads of a thread go out together (they were serialised behind the shared-memory atomics)
        long t = e_lo + threadIdx.x;
        for (; t + 3 * NT < e_hi; t += 4 * NT) {
            const int j0 = __ldg(ib + t), j1 = __ldg(ib + t + NT), j2 = __ldg(ib + t + 2 * NT), j3 = __ldg(ib + t + 3 * NT);
            atomicAdd(&hist[j0], 1); atomicAdd(&hist[j1], 1); atomicAdd(&hist[j2], 1); atomicAdd(&hist[j3], 1);
        }
        for (; t < e_hi; t += NT) atomicAdd(&hist[__ldg(ib + t)], 1);
    }
    cluster.sync();
    // total in-degree of this CTA's rows, gathered from the CL partial histograms
    const int r_lo = (int)((long)N * rank / CL), r_hi = (int)((long)N * (rank + 1) / CL);
    for (int r = r_lo + threadIdx.x; r < r_hi; r += NT) {
        int d = 0;
        for (unsigned q = 0; q < CL; ++q) d += cluster.map_shared_rank(hist, q)[r];
        indeg[r - r_lo] = d;
    }
    cluster.sync();   // every CTA is done reading remote histograms
    const float* p = in.s + (long)b * N * in.lds;
    // one pass over the rows per 128-column super block (same summation order as a column-by-column walk)
    for (int cb = 0; cb < Cs; cb += 128) {
        float accd[4], accp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { accd[u] = 0.0f; accp[u] = 0.0f; }
#pragma unroll 8
        for (int r = r_lo + rg; r < r_hi; r += 8) {
            const float* pr = p + (long)r * in.lds + cb + lane;
            const float dg = (float)indeg[r - r_lo];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cb + 32 * u + lane < Cs) {
                    const float t = pr[32 * u];
                    accd[u] = fmaf(dg, t, accd[u]);
                    accp[u] += t;
                }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = cb + 32 * u + lane;
            __syncthreads();
            part[rg * 32 + lane] = accd[u];
            part[256 + rg * 32 + lane] = accp[u];
            __syncthreads();
            if (rg == 0 && c < Cs) {
                float td = part[lane], tp = part[256 + lane];
                for (int g = 1; g < 8; ++g) { td += part[g * 32 + lane]; tp += part[256 + g * 32 + lane]; }
                psum[c] = td;
                psum[Cs + c] = tp;
            }
        }
    }
    cluster.sync();
    if (rank == 0) {
        for (int c = threadIdx.x; c < Cs; c += NT) {
            float td = 0.0f, tp = 0.0f;
            for (unsigned q = 0; q < CL; ++q) {
                const float* rp = cluster.map_shared_rank(psum, q);
                td += rp[c];
                tp += rp[Cs + c];
            }
            const float mp = tp / (float)N;
            mean[c] = td / ((float)N * (float)k) - mp;
            mean[Cs + c] = mp;
        }
        __syncthreads();
    }
    cluster.sync();
    if (rank == 0) gate_mlp(mean, 2 * Cs, G1, G2, H, Co, h, gate + (long)b * Co);
}

// smem: mean[3nv] | h[H] | part[NT][9] | psum[9]   (launched as a thread-block cluster per cloud)
template <int NV>
__global__ void __launch_bounds__(NT) gate_xyz_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx,
                                                      int N, int k, const float* __restrict__ Winit,
                                                      const float* __restrict__ G1, const float* __restrict__ G2,
                                                      int H, int Co, float* __restrict__ gate)
{
    extern __shared__ float sm[];
    float* mean = sm;
    float* h = mean + 3 * NV;
    float* part = h + H;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank(), csize = cluster.num_blocks();
    const int b = blockIdx.x / csize;
    float W[3][NV];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int d = 0; d < NV; ++d) W[m][d] = Winit[m * NV + d];
    float acc[3 * NV];
#pragma unroll
    for (int q = 0; q < 3 * NV; ++q) acc[q] = 0.0f;
    const float* xb = xyz + (long)b * N * 3;
    const int32_t* ib = idx + (long)b * N * k;
    const long E = (long)N * k;
    const long e_lo = E * crank / csize, e_hi = E * (crank + 1) / csize;
    for (long t = e_lo + threadIdx.x; t < e_hi; t += NT) {
        const int i = (int)(t / k);
        const int j = ib[t];
        float xi[3] = {xb[i * 3], xb[i * 3 + 1], xb[i * 3 + 2]};
        float xj[3] = {xb[j * 3], xb[j * 3 + 1], xb[j * 3 + 2]};
        float ve[3][NV];
#pragma unroll
        for (int a = 0; a < 3; ++a) { ve[a][0] = __fsub_rn(xj[a], xi[a]); ve[a][1] = xi[a]; }
        if (NV == 3) {
            ve[0][NV - 1] = __fsub_rn(__fmul_rn(xj[1], xi[2]), __fmul_rn(xj[2], xi[1]));
            ve[1][NV - 1] = __fsub_rn(__fmul_rn(xj[2], xi[0]), __fmul_rn(xj[0], xi[2]));
            ve[2][NV - 1] = __fsub_rn(__fmul_rn(xj[0], xi[1]), __fmul_rn(xj[1], xi[0]));
        }
        float z[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                float zz = 0.0f;
#pragma unroll
                for (int d = 0; d < NV; ++d) zz = __fmaf_rn(ve[a][d], W[m][d], zz);
                z[a][m] = zz;
            }
#pragma unroll
        for (int d = 0; d < NV; ++d)
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                float q = __fmul_rn(ve[0][d], z[0][m]);
                q = __fmaf_rn(ve[1][d], z[1][m], q);
                q = __fmaf_rn(ve[2][d], z[2][m], q);
                acc[d * 3 + m] += q;
            }
    }
#pragma unroll
    for (int q = 0; q < 3 * NV; ++q) part[threadIdx.x * 9 + q] = acc[q];
    __syncthreads();
    float* psum = part + NT * 9;   // [9] this CTA's partial sums (read by rank 0 through DSMEM)
    if (threadIdx.x < 3 * NV) {
        float t = 0.0f;
        for (int i = 0; i < NT; ++i) t += part[i * 9 + threadIdx.x];
        psum[threadIdx.x] = t;
    }
    cluster.sync();
    if (crank == 0) {
        if (threadIdx.x < 3 * NV) {
            float t = 0.0f;
            for (unsigned q = 0; q < csize; ++q) t += cluster.map_shared_rank(psum, q)[threadIdx.x];
            mean[threadIdx.x] = t / ((float)N * (float)k);
        }
        __syncthreads();
    }
    cluster.sync();
    if (crank == 0) gate_mlp(mean, 3 * NV, G1, G2, H, Co, h, gate + (long)b * Co);
}

}  // namespace

extern "C" int svnet_gate_rows(const float* s, int lds, int Cs, int B, int rows, const float* G1, const float* G2,
                               int H, int Co, float* gate, void* stream)
{
    SV_REQUIRE(s && G1 && G2 && gate, "svnet_gate_rows: null pointer");
    SV_REQUIRE(Cs >= 1 && lds >= Cs && rows >= 1 && H >= 1 && Co >= 1 && B >= 0, "svnet_gate_rows: bad shape");
    if (B == 0) return SVNET_OK;
    if (rows >= 512) {
        const size_t smem_c = sizeof(float) * ((size_t)2 * Cs + H + 256);
        SV_REQUIRE(smem_c <= 200 * 1024, "svnet_gate_rows: Cs=%d too large", Cs);
        if (smem_c > 48 * 1024)
            SV_CUDA(cudaFuncSetAttribute(gate_rows_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(B * CL);
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = smem_c;
        cfg.stream = sv_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        SV_CUDA(cudaLaunchKernelEx(&cfg, gate_rows_cluster_kernel, s, lds, Cs, (long)rows, G1, G2, H, Co, gate));
        SV_CHECK_LAUNCH("svnet_gate_rows(cluster)");
        return SVNET_OK;
    }
    const size_t smem = sizeof(float) * ((size_t)Cs + H + 256);
    SV_REQUIRE(smem <= 200 * 1024, "svnet_gate_rows: Cs=%d too large", Cs);
    if (smem > 48 * 1024)
        SV_CUDA(cudaFuncSetAttribute(gate_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gate_rows_kernel<<<B, NT, smem, sv_stream(stream)>>>(s, lds, Cs, rows, G1, G2, H, Co, gate);
    SV_CHECK_LAUNCH("svnet_gate_rows");
    return SVNET_OK;
}

extern "C" int svnet_gate_edge(const svnet_view* in, const int32_t* idx, int B, int N, int k, const float* G1,
                               const float* G2, int H, int Co, float* gate, void* stream)
{
    SV_REQUIRE(in && in->s && idx && G1 && G2 && gate, "svnet_gate_edge: null pointer");
    SV_REQUIRE(in->Cs >= 1 && N >= 1 && k >= 1 && H >= 1 && Co >= 1 && B >= 0, "svnet_gate_edge: bad shape");
    if (B == 0) return SVNET_OK;
    if (N >= 512) {
        {
            const size_t smem_c = sizeof(float) * ((size_t)4 * in->Cs + H + 512 + N + N / CL + 2);
            SV_REQUIRE(smem_c <= 200 * 1024, "svnet_gate_edge: N=%d / Cs=%d too large", N, in->Cs);
            if (smem_c > 48 * 1024)
                SV_CUDA(cudaFuncSetAttribute(gate_edge_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(B * CL);
            cfg.blockDim = dim3(NT);
            cfg.dynamicSmemBytes = smem_c;
            cfg.stream = sv_stream(stream);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            SV_CUDA(cudaLaunchKernelEx(&cfg, gate_edge_cluster_kernel, *in, idx, N, k, G1, G2, H, Co, gate));
            SV_CHECK_LAUNCH("svnet_gate_edge(cluster)");
            return SVNET_OK;
        }
    }
    const size_t smem = sizeof(float) * ((size_t)2 * in->Cs + H + 512 + N);
    SV_REQUIRE(smem <= 200 * 1024, "svnet_gate_edge: N=%d / Cs=%d too large", N, in->Cs);
    if (smem > 48 * 1024)
        SV_CUDA(cudaFuncSetAttribute(gate_edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gate_edge_kernel<<<B, NT, smem, sv_stream(stream)>>>(*in, idx, N, k, G1, G2, H, Co, gate);
    SV_CHECK_LAUNCH("svnet_gate_edge");
    return SVNET_OK;
}

extern "C" int svnet_gate_xyz(const float* xyz, const int32_t* idx, int B, int N, int k, int nv, const float* Winit,
                              const float* G1, const float* G2, int H, int Co, float* gate, void* stream)
{
    SV_REQUIRE(xyz && idx && Winit && G1 && G2 && gate, "svnet_gate_xyz: null pointer");
    SV_REQUIRE((nv == 2 || nv == 3) && N >= 1 && k >= 1 && H >= 1 && Co >= 1 && B >= 0, "svnet_gate_xyz: bad shape");
    if (B == 0) return SVNET_OK;
    const size_t smem = sizeof(float) * ((size_t)3 * nv + H + NT * 9 + 12);
    const int cl = ((long)N * k >= 4096) ? CL : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * cl);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = sv_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (nv == 2)
        SV_CUDA(cudaLaunchKernelEx(&cfg, gate_xyz_kernel<2>, xyz, idx, N, k, Winit, G1, G2, H, Co, gate));
    else
        SV_CUDA(cudaLaunchKernelEx(&cfg, gate_xyz_kernel<3>, xyz, idx, N, k, Winit, G1, G2, H, Co, gate));
    SV_CHECK_LAUNCH("svnet_gate_xyz");
    return SVNET_OK;
}
