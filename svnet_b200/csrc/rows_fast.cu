// Per-row u = [s | v2s(v)] construction, instruction-lean variants of rows.cu:rows_prep_kernel for
// the per-point SVBlocks of the models (conv5: Cs = 256, Cv = 83; svfuse: Cv = 170):
//   * rows_prep_bits_kernel   : sign / zero-mask words of u + beta (SVBlock scalar branch input,
//                               reference models/sv_layers.py:185-187 with :36-39)
//   * svfuse_pool_kernel      : v2s(v) of SVFuse (sv_layers.py:206-220) reduced on the fly to the
//                               per-cloud column max and mean (sv_dgcnn_cls.py:70-74) -- the fused
//                               (B*N, 3*Cv) table is never written to HBM.
// Same arithmetic and order as rows.cu (frames: sequential fmaf chain over channels; q = v0 z0,
// fma(v1, z1, .), fma(v2, z2, .)), so sign words are bit-identical.  What changes is the bookkeeping:
// rows are staged with cp.async, the chains read 16-byte vectors (zero padded: fmaf(0,0,x) == x),
// the lane -> (channel, frame column) map advances by whole periods (96 values = 32 channels) so no
// division or per-word frame lookups remain in the inner loop, and words leave as 16-byte stores.
#include "common.cuh"

namespace {

constexpr int RW = 4;     // warps per CTA

__device__ __forceinline__ void rf_cp_async4(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void rf_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// per-warp shared layout: v [3 rows][3][CvP] | pad 16 | z [3][12]
__host__ __device__ inline int rf_cvp(int Cv) { return (Cv + 3) & ~3; }
__host__ __device__ inline int rf_warp_floats(int Cv) { return 9 * rf_cvp(Cv) + 16 + 36; }

__device__ __forceinline__ void rf_cp_async8(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}

// stage the vectors of up to three rows (all copies in flight); 8-byte copies when the view allows it
__device__ __forceinline__ void stage_rows(const svnet_view& in, long r0, int ng, float* vsm, int Cv, int CvP, int lane)
{
    const bool wide = (((Cv | in.ldv | in.xs) & 1) == 0) && ((reinterpret_cast<uintptr_t>(in.v) & 7) == 0);
#pragma unroll 1
    for (int x = 0; x < 3; ++x) {
        const float* s0 = in.v + r0 * in.ldv + (long)x * in.xs;
        const float* s1 = s0 + in.ldv;
        const float* s2 = s1 + in.ldv;
        float* d0 = vsm + x * CvP;
        float* d1 = d0 + 3 * CvP;
        float* d2 = d1 + 3 * CvP;
        if (wide) {
            for (int c = 2 * lane; c < Cv; c += 64) {
                rf_cp_async8(d0 + c, s0 + c);
                if (ng > 1) rf_cp_async8(d1 + c, s1 + c);
                if (ng > 2) rf_cp_async8(d2 + c, s2 + c);
            }
        } else {
            for (int c = lane; c < Cv; c += 32) {
                rf_cp_async4(d0 + c, s0 + c);
                if (ng > 1) rf_cp_async4(d1 + c, s1 + c);
                if (ng > 2) rf_cp_async4(d2 + c, s2 + c);
            }
        }
    }
}

// frames z[g][x][m] for the staged rows: 27 lanes, sequential chain (16-byte loads, zero padded)
__device__ __forceinline__ void frames(const float* vsm, const float* wzs, float* zb, int ng, int CvP, const float* zscale,
                                       float* z_out, long r0, int lane)
{
    if (lane < ng * 9) {
        const int g = lane / 9, xm = lane - g * 9, x = xm / 3, m = xm - x * 3;
        const float4* vp = reinterpret_cast<const float4*>(vsm + (g * 3 + x) * CvP);
        const float4* wz = reinterpret_cast<const float4*>(wzs + m * CvP);
        float acc = 0.0f;
        for (int c4 = 0; c4 < CvP / 4; ++c4) {
            const float4 a = vp[c4], w = wz[c4];
            acc = __fmaf_rn(a.x, w.x, acc);
            acc = __fmaf_rn(a.y, w.y, acc);
            acc = __fmaf_rn(a.z, w.z, acc);
            acc = __fmaf_rn(a.w, w.w, acc);
        }
        if (zscale) acc = __fmul_rn(acc, __ldg(zscale + m));
        zb[g * 12 + xm] = acc;
        if (z_out) z_out[(r0 + g) * 9 + xm] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// sign / mask words of [s | q] + beta.  Requires Cs % 32 == 0, Cs <= 512, Kw % 4 == 0.
// ---------------------------------------------------------------------------------------------
// CS_T / CV_T > 0: the widths as compile-time constants (conv5 of the classifier: 256 / 83) -- the word loops unroll, the
// running word index and its slot become constants (no selects), the predicates on ws / nq disappear; 0: runtime widths.
template <int CS_T, int CV_T>
__global__ void __launch_bounds__(RW * 32) rows_prep_bits_kernel(svnet_view in, long rows, const float* __restrict__ Wz,
                                                                 const float* __restrict__ zscale, const float* __restrict__ beta,
                                                                 uint32_t* __restrict__ bits, uint32_t* __restrict__ mask,
                                                                 int32_t* __restrict__ nvalid)
{
    extern __shared__ __align__(16) float rf_smem[];
    constexpr bool STATIC = CS_T > 0 && CV_T > 0;
    const int lane = threadIdx.x & 31, warp = sv_warp_id();
    const int Cs = STATIC ? CS_T : in.Cs, Cv = STATIC ? CV_T : in.Cv, CvP = rf_cvp(Cv);
    const int K = Cs + 3 * Cv, Kw = (K + 31) / 32, KQ = 3 * Cv;
    const int ws = Cs >> 5, nq = Kw - ws;
    float* wzs = rf_smem;                                   // [3][CvP], zero padded
    float* betas = wzs + 3 * CvP;                           // [Kw * 32], zero padded
    float* vsm = betas + Kw * 32 + (size_t)warp * rf_warp_floats(Cv);
    float* zb = vsm + 9 * CvP + 16;
    for (int i = threadIdx.x; i < 3 * CvP; i += RW * 32) {
        const int m = i / CvP, c = i - m * CvP;
        wzs[i] = c < Cv ? __ldg(Wz + m * Cv + c) : 0.0f;
    }
    for (int i = threadIdx.x; i < Kw * 32; i += RW * 32) betas[i] = i < K ? __ldg(beta + i) : 0.0f;
    for (int i = lane; i < 9 * CvP + 16; i += 32) vsm[i] = 0.0f;      // channel padding stays zero
    __syncthreads();

    // lane -> (channel, frame column) of the q values: t = 32 w + lane = 3 dd + m; three words later
    // the pattern repeats with dd + 32
    int ddj[3], mj[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { const int t = 32 * j + lane; ddj[j] = t / 3; mj[j] = t - ddj[j] * 3; }

    const long ngroups = (rows + 2) / 3;
    for (long grp = (long)blockIdx.x * RW + warp; grp < ngroups; grp += (long)gridDim.x * RW) {
        const long r0 = grp * 3;
        const int ng = (int)min(3L, rows - r0);
        stage_rows(in, r0, ng, vsm, Cv, CvP, lane);
        // scalar words of the three rows: loads in flight together with the vector copies
        float spre[3][16];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int w = 0; w < 16; ++w)
                spre[g][w] = (g < ng && w < ws) ? __ldg(in.s + (r0 + g) * in.lds + w * 32 + lane) : 0.0f;
        rf_wait();
        __syncwarp();
        frames(vsm, wzs, zb, ng, CvP, zscale, nullptr, r0, lane);
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            // rows beyond the end (last group only) are computed on stale staging data and not stored: no divergent exit
            // in front of the ballots
            const bool gon = g < ng;
            const long r = r0 + (gon ? g : 0);
            uint4* bout = reinterpret_cast<uint4*>(bits + r * Kw);
            uint4* mout = reinterpret_cast<uint4*>(mask + r * Kw);
            int nval = 0;
            unsigned pw[4], nw[4];
            // ---- scalar words ----
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                if (w < ws) {
                    const float t = __fadd_rn(spre[g][w], betas[w * 32 + lane]);
                    pw[w & 3] = __ballot_sync(SV_FULL, t > 0.0f);
                    nw[w & 3] = __ballot_sync(SV_FULL, t != 0.0f);
                    nval += __popc(nw[w & 3]);
                    if ((w & 3) == 3 && lane == 0 && gon) {
                        bout[w >> 2] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                        mout[w >> 2] = make_uint4(nw[0], nw[1], nw[2], nw[3]);
                    }
                }
            }
            // ---- q words: periods of three words ----
            const float* vr = vsm + g * 3 * CvP;
            const float* z = zb + g * 12;
            float zs[3][3];
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int x = 0; x < 3; ++x) zs[j][x] = z[3 * x + mj[j]];
            // the running word index continues after the ws scalar words; ws % 4 may be non-zero
            int wd = ws;
#pragma unroll(STATIC ? 8 : 1)
            for (int wq = 0; wq < nq; wq += 3) {
                const int ddo = (wq / 3) * 32;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (wq + j < nq) {
                        const int dd = ddj[j] + ddo;
                        const int t = 32 * (wq + j) + lane;
                        float q = __fmul_rn(vr[dd], zs[j][0]);
                        q = __fmaf_rn(vr[CvP + dd], zs[j][1], q);
                        q = __fmaf_rn(vr[2 * CvP + dd], zs[j][2], q);
                        const float tt = t < KQ ? __fadd_rn(q, betas[wd * 32 + lane]) : 0.0f;
                        const unsigned pos = __ballot_sync(SV_FULL, tt > 0.0f);
                        const unsigned nz = __ballot_sync(SV_FULL, tt != 0.0f);
                        nval += __popc(nz);
                        const int slot = wd & 3;
                        // dynamic slot: selects instead of indexing (wd is not a compile-time value here)
                        pw[0] = slot == 0 ? pos : pw[0]; nw[0] = slot == 0 ? nz : nw[0];
                        pw[1] = slot == 1 ? pos : pw[1]; nw[1] = slot == 1 ? nz : nw[1];
                        pw[2] = slot == 2 ? pos : pw[2]; nw[2] = slot == 2 ? nz : nw[2];
                        pw[3] = slot == 3 ? pos : pw[3]; nw[3] = slot == 3 ? nz : nw[3];
                        if (slot == 3 && lane == 0 && gon) {
                            bout[wd >> 2] = make_uint4(pw[0], pw[1], pw[2], pw[3]);
                            mout[wd >> 2] = make_uint4(nw[0], nw[1], nw[2], nw[3]);
                        }
                        ++wd;
                    }
                }
            }
            if (lane == 0 && gon) nvalid[r] = nval;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// v2s(v) of SVFuse reduced to per-cloud column max / sum partials.  grid (parts, B); CTA (part, b)
// takes rows [part * rpp, (part + 1) * rpp) of cloud b.  KQ = 3 * Cv <= 512 (16 words per lane).
// Partials [B][parts][2][KQ] are combined in a fixed order by svfuse_pool_reduce_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int PQW = 16;   // q words per lane held in registers

// CV_T > 0: the width as a compile-time constant (svfuse of the classifier: 170); 0: runtime width
template <int CV_T>
__global__ void __launch_bounds__(RW * 32) svfuse_pool_kernel(svnet_view in, int rows_per_cloud, int rpp, const float* __restrict__ Wz,
                                                              const float* __restrict__ zscale, float* __restrict__ partial)
{
    extern __shared__ __align__(16) float rf_smem[];
    const int lane = threadIdx.x & 31, warp = sv_warp_id();
    const int Cv = CV_T > 0 ? CV_T : in.Cv, CvP = rf_cvp(Cv), KQ = 3 * Cv, nq = (KQ + 31) / 32;
    float* wzs = rf_smem;
    float* red = wzs + 3 * CvP;                              // [RW][2][PQW * 32]
    float* vsm = red + RW * 2 * PQW * 32 + (size_t)warp * rf_warp_floats(Cv);
    float* zb = vsm + 9 * CvP + 16;
    for (int i = threadIdx.x; i < 3 * CvP; i += RW * 32) {
        const int m = i / CvP, c = i - m * CvP;
        wzs[i] = c < Cv ? __ldg(Wz + m * Cv + c) : 0.0f;
    }
    for (int i = lane; i < 9 * CvP + 16; i += 32) vsm[i] = 0.0f;
    __syncthreads();
    int ddj[3], mj[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { const int t = 32 * j + lane; ddj[j] = t / 3; mj[j] = t - ddj[j] * 3; }

    float vmax[PQW], vsum[PQW];
#pragma unroll
    for (int w = 0; w < PQW; ++w) { vmax[w] = -INFINITY; vsum[w] = 0.0f; }

    const int b = blockIdx.y, part = blockIdx.x;
    const long cloud0 = (long)b * rows_per_cloud;
    const int lo = part * rpp, hi = min(rows_per_cloud, lo + rpp);
    // warp w takes the groups w, w + RW, ... of this part (3 rows each)
    for (int g0 = lo + warp * 3; g0 < hi; g0 += RW * 3) {
        const long r0 = cloud0 + g0;
        const int ng = min(3, hi - g0);
        stage_rows(in, r0, ng, vsm, Cv, CvP, lane);
        rf_wait();
        __syncwarp();
        frames(vsm, wzs, zb, ng, CvP, zscale, nullptr, r0, lane);
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            if (g >= ng) break;
            const float* vr = vsm + g * 3 * CvP;
            const float* z = zb + g * 12;
            float zs[3][3];
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int x = 0; x < 3; ++x) zs[j][x] = z[3 * x + mj[j]];
#pragma unroll
            for (int w = 0; w < PQW; ++w) {
                if (w < nq) {
                    const int j = w % 3, dd = ddj[j] + (w / 3) * 32;
                    float q = __fmul_rn(vr[dd], zs[j][0]);
                    q = __fmaf_rn(vr[CvP + dd], zs[j][1], q);
                    q = __fmaf_rn(vr[2 * CvP + dd], zs[j][2], q);
                    vmax[w] = fmaxf(vmax[w], q);
                    vsum[w] += q;
                }
            }
        }
        __syncwarp();
    }
    // ---- combine the warps in a fixed order, write this part's partials ----
#pragma unroll
    for (int w = 0; w < PQW; ++w) {
        red[(warp * 2 + 0) * PQW * 32 + w * 32 + lane] = vmax[w];
        red[(warp * 2 + 1) * PQW * 32 + w * 32 + lane] = vsum[w];
    }
    __syncthreads();
    float* pout = partial + ((size_t)b * gridDim.x + part) * 2 * KQ;
    for (int c = threadIdx.x; c < KQ; c += RW * 32) {
        float mx = red[c], sm = red[PQW * 32 + c];
#pragma unroll
        for (int w = 1; w < RW; ++w) {
            mx = fmaxf(mx, red[(w * 2 + 0) * PQW * 32 + c]);
            sm += red[(w * 2 + 1) * PQW * 32 + c];
        }
        pout[c] = mx;
        pout[KQ + c] = sm;
    }
}

__global__ void svfuse_pool_reduce_kernel(const float* __restrict__ partial, int parts, int KQ, int rows_per_cloud,
                                          float* __restrict__ max_out, float* __restrict__ mean_out, int ldo)
{
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= KQ) return;
    const float* p = partial + (size_t)b * parts * 2 * KQ;
    float mx = p[c], sm = p[KQ + c];
    for (int q = 1; q < parts; ++q) {
        mx = fmaxf(mx, p[(size_t)q * 2 * KQ + c]);
        sm += p[(size_t)q * 2 * KQ + KQ + c];
    }
    if (max_out) max_out[(size_t)b * ldo + c] = mx;
    if (mean_out) mean_out[(size_t)b * ldo + c] = sm / (float)rows_per_cloud;
}

// ---------------------------------------------------------------------------------------------
// Plain sign-pack of a float matrix (Linear `ba`, Conv1d; sv_layers.py:36-39,69): one warp = one (row, 8-word chunk);
// the eight loads of a lane are in flight together, ballots give the words, lanes 0..7 store them (32 contiguous
// bytes per plane).  nvalid (zeroed by the caller) collects the chunks' popcounts with integer atomics.
// Round 2: the generic three-rows-per-warp kernel took 76 us for the seg head's 16 x 1600 per-cloud matrix (one load
// per ballot round) and 33 us for the 32768 x 256 activations of conv9 / conv10.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) signpack_kernel(const float* __restrict__ s, long lds, int K, long rows, const float* __restrict__ beta,
                                                       uint32_t* __restrict__ bits, uint32_t* __restrict__ mask, int32_t* __restrict__ nvalid)
{
    const int lane = threadIdx.x & 31;
    const int Kw = (K + 31) >> 5, nch = (Kw + 7) >> 3;
    const long items = rows * nch;
    for (long it = (long)blockIdx.x * 8 + sv_warp_id(); it < items; it += (long)gridDim.x * 8) {
        const long r = it / nch;
        const int w0 = (int)(it - r * nch) * 8;
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = (w0 + u) * 32 + lane;
            t[u] = c < K ? __fadd_rn(__ldg(s + r * lds + c), __ldg(beta + c)) : 0.0f;
        }
        unsigned mypos = 0u, mynz = 0u;
        int nval = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned pos = __ballot_sync(SV_FULL, t[u] > 0.0f);
            const unsigned nz = __ballot_sync(SV_FULL, t[u] != 0.0f);
            nval += __popc(nz);
            if (lane == u) { mypos = pos; mynz = nz; }
        }
        if (lane < 8 && w0 + lane < Kw) {
            bits[r * Kw + w0 + lane] = mypos;
            mask[r * Kw + w0 + lane] = mynz;
        }
        if (lane == 0) atomicAdd(nvalid + r, nval);
    }
}

}  // namespace

// Returns 1 if handled, 0 if rows.cu must take the call, < 0 on error.
int svnet_rows_prep_fast_dispatch(const svnet_view* in, long rows, const float* Wz, const float* zscale, const float* beta,
                                  uint32_t* bits, uint32_t* mask, int32_t* nvalid, cudaStream_t st)
{
    const int Cs = in->Cs, Cv = in->Cv;
    const int K = Cs + 3 * Cv, Kw = (K + 31) / 32;
    if (Cv == 0) {
        SV_CUDA(cudaMemsetAsync(nvalid, 0, sizeof(int32_t) * (size_t)rows, st));
        const long items = rows * ((Kw + 7) / 8);
        signpack_kernel<<<(int)min((long)sv_cdiv(items, 8), 148L * 32), 256, 0, st>>>(in->s, in->lds, K, rows, beta, bits, mask, nvalid);
        SV_CHECK_LAUNCH("svnet_rows_prep(sign-pack)");
        return 1;
    }
    if (Cv < 1 || (Cs & 31) || Cs > 512 || (Kw & 3) || rows < 1024) return 0;
    if ((reinterpret_cast<uintptr_t>(bits) | reinterpret_cast<uintptr_t>(mask)) & 15) return 0;
    const size_t smem = sizeof(float) * ((size_t)3 * rf_cvp(Cv) + (size_t)Kw * 32 + (size_t)RW * rf_warp_floats(Cv));
    if (smem > 64 * 1024) return 0;
    const long ngroups = (rows + 2) / 3;
    const int grid = (int)min((long)sv_cdiv(ngroups, RW), 148L * 32);
    if (Cs == 256 && Cv == 83) {
        SV_CUDA(cudaFuncSetAttribute(rows_prep_bits_kernel<256, 83>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rows_prep_bits_kernel<256, 83><<<grid, RW * 32, smem, st>>>(*in, rows, Wz, zscale, beta, bits, mask, nvalid);
    } else {
        SV_CUDA(cudaFuncSetAttribute(rows_prep_bits_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rows_prep_bits_kernel<0, 0><<<grid, RW * 32, smem, st>>>(*in, rows, Wz, zscale, beta, bits, mask, nvalid);
    }
    SV_CHECK_LAUNCH("svnet_rows_prep(fast)");
    return 1;
}

extern "C" size_t svnet_svfuse_pool_workspace(int B, int Cv, long rows_per_cloud)
{
    if (B < 1 || Cv < 1 || rows_per_cloud < 1) return 0;
    const int parts = (int)min(32L, (rows_per_cloud + 47) / 48);
    return (size_t)B * parts * 2 * 3 * Cv * sizeof(float);
}

extern "C" int svnet_svfuse_pool(const svnet_view* in, int B, long rows_per_cloud, const float* Wz, const float* zscale,
                                 float* max_out, float* mean_out, int ldo, void* workspace, size_t workspace_bytes, void* stream)
{
    SV_REQUIRE(in && in->v && in->Cv >= 1 && Wz, "svnet_svfuse_pool: null view / Wz");
    SV_REQUIRE(B >= 0 && rows_per_cloud >= 1, "svnet_svfuse_pool: bad shape");
    SV_REQUIRE(max_out || mean_out, "svnet_svfuse_pool: no output requested");
    SV_REQUIRE(3 * in->Cv <= PQW * 32, "svnet_svfuse_pool: 3*Cv = %d > %d unsupported", 3 * in->Cv, PQW * 32);
    SV_REQUIRE(ldo >= 3 * in->Cv, "svnet_svfuse_pool: ldo too small");
    SV_REQUIRE(workspace && workspace_bytes >= svnet_svfuse_pool_workspace(B, in->Cv, rows_per_cloud),
               "svnet_svfuse_pool: workspace too small");
    if (B == 0) return SVNET_OK;
    const int Cv = in->Cv, KQ = 3 * Cv;
    const int parts = (int)min(32L, (rows_per_cloud + 47) / 48);
    const int rpp = (int)((rows_per_cloud + parts - 1) / parts);
    const size_t smem = sizeof(float) * ((size_t)3 * rf_cvp(Cv) + (size_t)RW * 2 * PQW * 32 + (size_t)RW * rf_warp_floats(Cv));
    SV_REQUIRE(smem <= 96 * 1024, "svnet_svfuse_pool: Cv = %d too large", Cv);
    cudaStream_t st = sv_stream(stream);
    if (Cv == 170) {
        SV_CUDA(cudaFuncSetAttribute(svfuse_pool_kernel<170>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        svfuse_pool_kernel<170><<<dim3(parts, B), RW * 32, smem, st>>>(*in, (int)rows_per_cloud, rpp, Wz, zscale,
                                                                       static_cast<float*>(workspace));
    } else {
        SV_CUDA(cudaFuncSetAttribute(svfuse_pool_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        svfuse_pool_kernel<0><<<dim3(parts, B), RW * 32, smem, st>>>(*in, (int)rows_per_cloud, rpp, Wz, zscale,
                                                                     static_cast<float*>(workspace));
    }
    SV_CHECK_LAUNCH("svnet_svfuse_pool");
    svfuse_pool_reduce_kernel<<<dim3(sv_cdiv(KQ, 128), B), 128, 0, st>>>(static_cast<const float*>(workspace), parts, KQ,
                                                                          (int)rows_per_cloud, max_out, mean_out, ldo);
    SV_CHECK_LAUNCH("svnet_svfuse_pool(reduce)");
    return SVNET_OK;
}
