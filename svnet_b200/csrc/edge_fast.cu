// K2 (specialised) -- fused binary SV edge convolution with compile-time layer shapes.
// Same mathematics and bit-level contract as the generic kernel in edge.cu (see the header there and
// include/svnet_b200.h: svnet_svblock_edge_fwd); the shapes of the SV-DGCNN classifier and
// part-segmentation edge layers are template parameters so that every per-lane offset, the number of
// sign words and all inner loops are resolved at compile time (the generic kernel spends 3-5x more
// instructions on address arithmetic than on the arithmetic itself, profiles/r1_v1_*).
//
// One warp per centre point:
//   P1  per edge: s_j - s_i + beta -> sign/mask words by __ballot_sync straight from registers;
//       v_j - v_i -> shared memory (needed by the frame chains)
//   P2  3x3 frames z for all k edges: 9k sequential fmaf chains spread over the lanes
//   P3  q = v_e . z lane-per-channel -> sign/mask words
//   P4  XNOR/popcount linear1 (lanes = output channels), BN, LeakyReLU, max over edges
//   P5  vector branch from the per-point P|Q table, VectorBN, gate, mean over edges
#include "common.cuh"
#include "edge_vector.cuh"
#include <limits.h>

#ifndef EDGE_MB_SMALL
#define EDGE_MB_SMALL 4
#endif
#ifndef EDGE_MB_LARGE
#define EDGE_MB_LARGE 3
#endif

namespace {

template <int CS, int CV, int COUT, int CVO>
struct Shape {
    static constexpr int CVE = 2 * CV;
    static constexpr int K = 2 * CS + 6 * CV;
    static constexpr int KW = (K + 31) / 32;
    static constexpr int TS = CS / 32;                 // s words per half
    static constexpr int NVE = 3 * CV;                 // vector floats per point
    static constexpr int TV = (NVE + 31) / 32;
    static constexpr int QW = KW - 2 * TS;             // words of the q section
    static constexpr int OPT = COUT / 32;
    static constexpr int EB = 20;                       // edges accumulated per popcount pass (k padded to a multiple)
    static constexpr int OPP = (COUT / 32 > 2) ? 2 : COUT / 32;   // output channels per lane per pass
    static constexpr int XS = (CV % 2 == 0) ? CV + 1 : CV;   // odd xyz stride in smem: conflict-free chains
    static constexpr int ES = 3 * XS;                  // floats per staged edge
    static_assert(CS % 32 == 0 && COUT % 32 == 0, "scalar widths must be multiples of 32");
};


template <int CS, int CV, int COUT, int CVO, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS == 8) ? ((COUT <= 64) ? EDGE_MB_SMALL : EDGE_MB_LARGE) : 6) edge_bin_fast_kernel(svnet_edge_params p, int kp)
{
    using S = Shape<CS, CV, COUT, CVO>;
    constexpr int EB = S::EB;
    extern __shared__ __align__(16) float smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- CTA-shared: Wz [3][CVE], W1b [KW][COUT] ----
    float* Wz = smem_raw;
    uint32_t* W1b = reinterpret_cast<uint32_t*>(Wz + 3 * S::CVE + ((3 * S::CVE) & 1));
    constexpr int SHARED = 3 * S::CVE + ((3 * S::CVE) & 1) + S::KW * COUT;
    for (int i = threadIdx.x; i < 3 * S::CVE; i += blockDim.x) Wz[i] = p.Wz[i];
    for (int i = threadIdx.x; i < S::KW * COUT; i += blockDim.x) W1b[i] = p.W1b[i];
    // ---- per warp ----
    const int per_warp = ((2 * S::KW * kp + 2 * kp + 3) & ~3) + kp * 12 + 3 * S::XS + kp * S::ES;
    float* wbase = smem_raw + ((SHARED + 3) & ~3) + (size_t)warp * ((per_warp + 3) & ~3);
    uint32_t* A = reinterpret_cast<uint32_t*>(wbase);            // [KW][kp]   (16B aligned rows: kp % 4 == 0)
    uint32_t* M = A + S::KW * kp;                                 // [KW][kp]
    int* nvalid = reinterpret_cast<int*>(M + S::KW * kp);         // [kp]
    int* nidx = nvalid + kp;                                      // [kp]
    float* zb = wbase + ((2 * S::KW * kp + 2 * kp + 3) & ~3);     // [kp][3 m][4]: frame columns, 16-byte aligned
    float* vc = zb + kp * 12;                                     // [3][XS] centre vectors
    float* ves = vc + 3 * S::XS;                                  // [kp][3][XS] neighbour - centre
    // the CTA-wide barrier for the staged weights comes after P0/P1 (which do not read them), so the
    // staging latency hides behind the gathers; warps past the end wait there and leave
    const long r = (long)blockIdx.x * WARPS + warp;
    if (r >= (long)p.B * p.N) {
        __syncthreads();
        return;
    }
    const int b = (int)(r / p.N);
    const long cbase = (long)b * p.N;
    const int k = p.k;

    // ---- per-lane constants ----
    float beta[S::KW];
#pragma unroll
    for (int w = 0; w < S::KW; ++w) beta[w] = (32 * w + lane < S::K) ? __ldg(p.beta + 32 * w + lane) : 0.0f;
    int voff[S::TV], vso[S::TV];
#pragma unroll
    for (int t = 0; t < S::TV; ++t) {
        const int f = 32 * t + lane, x = f / CV, d = f - x * CV;
        voff[t] = (f < S::NVE) ? x * p.in.xs + d : -1;
        vso[t] = x * S::XS + d;
    }

    // ---- P0: centre row ----
    for (int e = lane; e < kp; e += 32) nidx[e] = e < k ? p.idx[r * k + e] : 0;
    float si[S::TS], vi[S::TV];
    const float* srow = p.in.s + r * p.in.lds;
    const float* vrow = p.in.v + r * p.in.ldv;
#pragma unroll
    for (int t = 0; t < S::TS; ++t) si[t] = __ldg(srow + 32 * t + lane);
#pragma unroll
    for (int t = 0; t < S::TV; ++t) {
        vi[t] = voff[t] >= 0 ? __ldg(vrow + voff[t]) : 0.0f;
        if (voff[t] >= 0) vc[vso[t]] = vi[t];
    }
    unsigned cpos[S::TS], cnz[S::TS];
    int ncen = 0;
#pragma unroll
    for (int t = 0; t < S::TS; ++t) {
        const float u = __fadd_rn(si[t], beta[S::TS + t]);
        cpos[t] = __ballot_sync(SV_FULL, u > 0.0f);
        cnz[t] = __ballot_sync(SV_FULL, u != 0.0f);
        ncen += __popc(cnz[t]);
    }
    // zero the padded edge slots (their mask stays 0 so they contribute nothing)
    for (int i = lane; i < S::KW * kp; i += 32) { A[i] = 0u; M[i] = 0u; }
    __syncwarp();

    // ---- P1: gather the neighbour rows into registers, PB rows in flight per round (scalars and vectors
    //      of a round share one latency exposure); S1 sign words by ballot, vector differences to
    //      shared memory.  (cp.async costs four issue slots apiece here -- ptxas pads every LDGSTS
    //      with three dummy LDS -- plus a second pass for the subtraction.) ----
    constexpr int PB = (S::TV + S::TS <= 2) ? EB : EB / 2;
    const float* sbase = p.in.s + cbase * p.in.lds + lane;
    const unsigned lds = (unsigned)p.in.lds, ldv = (unsigned)p.in.ldv;   // row offsets inside a cloud fit 32 bits (launch guard)
    const float* vsrc[S::TV];
#pragma unroll
    for (int t = 0; t < S::TV; ++t) vsrc[t] = p.in.v + cbase * p.in.ldv + (voff[t] >= 0 ? voff[t] : 0);
    for (int eb = 0; eb < k; eb += PB) {
        float sv[PB][S::TS], vv[PB][S::TV];
#pragma unroll
        for (int e = 0; e < PB; ++e) {
            const unsigned j = (unsigned)nidx[eb + e];            // padded slots read row 0
            const float* sj = sbase + j * lds;
#pragma unroll
            for (int t = 0; t < S::TS; ++t) sv[e][t] = __ldg(sj + 32 * t);
#pragma unroll
            for (int t = 0; t < S::TV; ++t) vv[e][t] = __ldg(vsrc[t] + j * ldv);
        }
        // the ballot words are warp-uniform: lane e keeps those of edge eb + e and stores them once
        // (a lane-0 store per word costs an address and a predicated store each)
        unsigned mp[S::TS], mn[S::TS];
        int mnv = 0;
#pragma unroll
        for (int t = 0; t < S::TS; ++t) mp[t] = mn[t] = 0u;
#pragma unroll
        for (int e = 0; e < PB; ++e) {
            if (eb + e < k) {
                int nv = ncen;
#pragma unroll
                for (int t = 0; t < S::TS; ++t) {
                    const float u = __fadd_rn(__fsub_rn(sv[e][t], si[t]), beta[t]);
                    const unsigned pos = __ballot_sync(SV_FULL, u > 0.0f);
                    const unsigned nz = __ballot_sync(SV_FULL, u != 0.0f);
                    nv += __popc(nz);
                    if (lane == e) { mp[t] = pos; mn[t] = nz; }
                }
                if (lane == e) mnv = nv;
#pragma unroll
                for (int t = 0; t < S::TV; ++t)
                    if (voff[t] >= 0) ves[(eb + e) * S::ES + vso[t]] = __fsub_rn(vv[e][t], vi[t]);
            }
        }
        if (lane < PB && eb + lane < k) {
#pragma unroll
            for (int t = 0; t < S::TS; ++t) { A[t * kp + eb + lane] = mp[t]; M[t * kp + eb + lane] = mn[t]; }
            nvalid[eb + lane] = mnv;
        }
    }
    __syncwarp();

    __syncthreads();          // Wz / W1b staged (see the top of the kernel)
    // ---- P2: frames z[e][x][m]: one sequential chain per (edge, x, m), 32 chains per round ----
    const bool use_zscale = p.zscale != nullptr;
    for (int task = lane; task < k * 9; task += 32) {
        const int e = task / 9, xm = task - e * 9, x = xm / 3, m = xm - x * 3;
        const float* dv = ves + e * S::ES + x * S::XS;
        const float* cv = vc + x * S::XS;
        const float* wz = Wz + m * S::CVE;
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < CV; ++d) acc = __fmaf_rn(dv[d], wz[d], acc);
#pragma unroll
        for (int d = 0; d < CV; ++d) acc = __fmaf_rn(cv[d], wz[CV + d], acc);
        if (use_zscale) acc = __fmul_rn(acc, __ldg(p.zscale + m));
        zb[e * 12 + m * 4 + x] = acc;
    }
    __syncwarp();

    // ---- P3: q section sign words.  lane -> channel tq = 32t + lane = 3d + m ----
    {
        int qoff[S::QW], qstr[S::QW], zoff[S::QW];
        bool qok[S::QW];
#pragma unroll
        for (int t = 0; t < S::QW; ++t) {
            const int tq = 32 * t + lane, d = tq / 3, m = tq - d * 3;
            qok[t] = tq < 6 * CV;
            zoff[t] = 4 * m;
            if (d < CV) { qoff[t] = (int)(ves - wbase) + d; qstr[t] = S::ES; }
            else { qoff[t] = (int)(vc - wbase) + (qok[t] ? d - CV : 0); qstr[t] = 0; }
        }
        const float* qsrc[S::QW];
#pragma unroll
        for (int t = 0; t < S::QW; ++t) qsrc[t] = wbase + qoff[t];
        const float* z = zb;
        for (int e0 = 0; e0 < k; e0 += 32) {        // lane e - e0 keeps the words of edge e (see P1)
            const int en = min(32, k - e0);
            unsigned mp[S::QW], mn[S::QW];
            int mnv = 0;
#pragma unroll
            for (int t = 0; t < S::QW; ++t) mp[t] = mn[t] = 0u;
#pragma unroll 2
            for (int el = 0; el < en; ++el, z += 12) {
                const bool me = lane == el;
                int nv = 0;
#pragma unroll
                for (int t = 0; t < S::QW; ++t) {
                    const float* src = qsrc[t];
                    qsrc[t] += qstr[t];
                    const float4 zz = *reinterpret_cast<const float4*>(z + zoff[t]);     // column m of the frame
                    float q = __fmul_rn(src[0], zz.x);
                    q = __fmaf_rn(src[S::XS], zz.y, q);
                    q = __fmaf_rn(src[2 * S::XS], zz.z, q);
                    const float u = qok[t] ? __fadd_rn(q, beta[2 * S::TS + t]) : 0.0f;
                    const unsigned pos = __ballot_sync(SV_FULL, u > 0.0f);
                    const unsigned nz = __ballot_sync(SV_FULL, u != 0.0f);
                    nv += __popc(nz);
                    if (me) { mp[t] = pos; mn[t] = nz; }
                }
                if (me) mnv = nv;
            }
            if (lane < en) {
#pragma unroll
                for (int t = 0; t < S::QW; ++t) { A[(2 * S::TS + t) * kp + e0 + lane] = mp[t]; M[(2 * S::TS + t) * kp + e0 + lane] = mn[t]; }
                nvalid[e0 + lane] += mnv;
            }
        }
    }
    __syncwarp();
    if (p.dbg_bits) {
        for (int i = lane; i < k * S::KW; i += 32) {
            const int e = i / S::KW, w = i - e * S::KW;
            unsigned aw = A[w * kp + e], mw = M[w * kp + e];
#pragma unroll
            for (int t = 0; t < S::TS; ++t)
                if (w == S::TS + t) { aw = cpos[t]; mw = cnz[t]; }
            p.dbg_bits[(r * k + e) * S::KW + w] = aw;
            if (p.dbg_mask) p.dbg_mask[(r * k + e) * S::KW + w] = mw;
        }
    }

    // ---- P4: XNOR/popcount linear1, lanes = output channels (PO per lane per pass; all four groups of a
    //      128-channel layer in one pass, so every sign / mask word is read once) ----
    constexpr int PO = (S::OPT == 4) ? 4 : S::OPP;
    constexpr int NWD = S::KW - S::TS;                       // words that differ per edge
    // PEB edges per accumulation round: at most 20 accumulators, so they stay in registers (40 were
    // spilled inside the loop)
    constexpr int PEB = (PO == 4) ? 4 : (PO == 2) ? EB / 2 : EB;
    constexpr int VW = (PEB % 4 == 0) ? 4 : 2;               // 16-byte loads when the round allows (kp % 4 == 0)
#pragma unroll 1
    for (int ob = 0; ob < S::OPT; ob += PO) {
        // y = leaky(bn(scale * dot)) is a monotone function of the integer dot (every fp32 rounding is
        // monotone), so max over the edges of y is y(max dot) or y(min dot): keep integer extremes and
        // run the float epilogue once per output channel -- bit-identical to taking the max of all y.
        int dmax[PO], dmin[PO];
        int cmis[PO];   // mismatches of the centre words (identical for all edges of this point)
        uint32_t wv[NWD][PO];
#pragma unroll
        for (int oo = 0; oo < PO; ++oo) {
            dmax[oo] = INT_MIN;
            dmin[oo] = INT_MAX;
            cmis[oo] = 0;
#pragma unroll
            for (int t = 0; t < S::TS; ++t)
                cmis[oo] += __popc((cpos[t] ^ W1b[(S::TS + t) * COUT + lane + 32 * (ob + oo)]) & cnz[t]);
#pragma unroll
            for (int w = 0; w < NWD; ++w) wv[w][oo] = W1b[(w < S::TS ? w : w + S::TS) * COUT + lane + 32 * (ob + oo)];
        }
        for (int eb = 0; eb < k; eb += PEB) {
            int acc[PEB][PO];
#pragma unroll
            for (int e = 0; e < PEB; ++e)
#pragma unroll
                for (int oo = 0; oo < PO; ++oo) acc[e][oo] = 0;
#pragma unroll
            for (int w = 0; w < NWD; ++w) {
                const int wd = w < S::TS ? w : w + S::TS;     // the centre words are in cmis
                const uint32_t* Ap = A + wd * kp + eb;
                const uint32_t* Mp = M + wd * kp + eb;
#pragma unroll
                for (int ev = 0; ev < PEB / VW; ++ev) {
                    uint32_t av[VW], mv[VW];
                    if constexpr (VW == 4) {
                        const uint4 a4 = reinterpret_cast<const uint4*>(Ap)[ev], m4 = reinterpret_cast<const uint4*>(Mp)[ev];
                        av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
                        mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
                    } else {
                        const uint2 a2 = reinterpret_cast<const uint2*>(Ap)[ev], m2 = reinterpret_cast<const uint2*>(Mp)[ev];
                        av[0] = a2.x; av[1] = a2.y;
                        mv[0] = m2.x; mv[1] = m2.y;
                    }
#pragma unroll
                    for (int u = 0; u < VW; ++u)
#pragma unroll
                        for (int oo = 0; oo < PO; ++oo) acc[ev * VW + u][oo] += __popc((av[u] ^ wv[w][oo]) & mv[u]);
                }
            }
#pragma unroll
            for (int e = 0; e < PEB; ++e) {
                if (eb + e < k) {
                    const int nv = nvalid[eb + e];
#pragma unroll
                    for (int oo = 0; oo < PO; ++oo) {
                        const int dot = nv - 2 * acc[e][oo];          // the centre words' share comes off after the loop
                        dmax[oo] = max(dmax[oo], dot);
                        dmin[oo] = min(dmin[oo], dot);
                    }
                }
            }
        }
#pragma unroll
        for (int oo = 0; oo < PO; ++oo) {
            const int o = lane + 32 * (ob + oo);
            const float sc = __ldg(p.scale1 + o), a1 = __ldg(p.bn1_a + o), c1 = __ldg(p.bn1_c + o);
            float y0 = __fadd_rn(__fmul_rn(__fmul_rn((float)(dmax[oo] - 2 * cmis[oo]), sc), a1), c1);
            float y1 = __fadd_rn(__fmul_rn(__fmul_rn((float)(dmin[oo] - 2 * cmis[oo]), sc), a1), c1);
            y0 = y0 > 0.0f ? y0 : __fmul_rn(0.2f, y0);
            y1 = y1 > 0.0f ? y1 : __fmul_rn(0.2f, y1);
            p.out.s[r * p.out.lds + o] = fmaxf(y0, y1);
        }
    }

    // ---- P5: vector branch ----
    vector_branch<CVO>(p, r, b, cbase, nidx, k, lane);
}

// ---- full-precision variant (SV-DGCNN fp models, cfg3): same gather / frame phases, the scalar
// branch is a dense fp32 linear  y = (Ya_j - Ya_i) + Yb_i + W1q q  with q staged in shared memory
// (edge index innermost so that 4 edges come with one 16-byte load) ----
template <int CS, int CV, int COUT, int CVO, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS == 8) ? 2 : 4) edge_fp_fast_kernel(svnet_edge_params p, int kp)
{
    using S = Shape<CS, CV, COUT, CVO>;
    constexpr int EB = S::EB;
    constexpr int KQ = 6 * CV;
    extern __shared__ __align__(16) float smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* Wz = smem_raw;                                   // [3][CVE], CTA-shared
    constexpr int SHARED = (3 * S::CVE + 3) & ~3;
    for (int i = threadIdx.x; i < 3 * S::CVE; i += blockDim.x) Wz[i] = p.Wz[i];
    const int per_warp = (KQ * kp + kp * 12 + 3 * S::XS + kp * S::ES + kp + 3) & ~3;
    float* wbase = smem_raw + SHARED + (size_t)warp * per_warp;
    float* qs = wbase;                                      // [KQ][kp]  (16B aligned rows: kp % 4 == 0)
    float* zb = qs + KQ * kp;                               // [kp][3 m][4]: frame columns, 16-byte aligned
    float* vc = zb + kp * 12;                               // [3][XS]
    float* ves = vc + 3 * S::XS;                            // [kp][3][XS]
    int* nidx = reinterpret_cast<int*>(ves + kp * S::ES);   // [kp]
    __syncthreads();

    const long r = (long)blockIdx.x * WARPS + warp;
    if (r >= (long)p.B * p.N) return;
    const int b = (int)(r / p.N);
    const long cbase = (long)b * p.N;
    const int k = p.k;

    int voff[S::TV], vso[S::TV];
#pragma unroll
    for (int t = 0; t < S::TV; ++t) {
        const int f = 32 * t + lane, x = f / CV, d = f - x * CV;
        voff[t] = (f < S::NVE) ? x * p.in.xs + d : -1;
        vso[t] = x * S::XS + d;
    }
    for (int e = lane; e < kp; e += 32) nidx[e] = e < k ? p.idx[r * k + e] : 0;
    float vi[S::TV];
    const float* vrow = p.in.v + r * p.in.ldv;
#pragma unroll
    for (int t = 0; t < S::TV; ++t) {
        vi[t] = voff[t] >= 0 ? __ldg(vrow + voff[t]) : 0.0f;
        if (voff[t] >= 0) vc[vso[t]] = vi[t];
    }
    for (int i = lane; i < KQ * kp; i += 32) qs[i] = 0.0f;   // padded edge slots stay zero
    __syncwarp();

    // ---- P1: neighbour vectors through registers (EB rows in flight), differences to shared memory ----
    {
        const unsigned ldv = (unsigned)p.in.ldv;                 // row offsets inside a cloud fit 32 bits (launch guard)
        const float* vsrc[S::TV];
#pragma unroll
        for (int t = 0; t < S::TV; ++t) vsrc[t] = p.in.v + cbase * p.in.ldv + (voff[t] >= 0 ? voff[t] : 0);
        for (int eb = 0; eb < k; eb += EB) {
            float vv[EB][S::TV];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
                const unsigned j = (unsigned)nidx[eb + e];        // padded slots read row 0
#pragma unroll
                for (int t = 0; t < S::TV; ++t) vv[e][t] = __ldg(vsrc[t] + j * ldv);
            }
#pragma unroll
            for (int e = 0; e < EB; ++e) {
                if (eb + e < k) {
#pragma unroll
                    for (int t = 0; t < S::TV; ++t)
                        if (voff[t] >= 0) ves[(eb + e) * S::ES + vso[t]] = __fsub_rn(vv[e][t], vi[t]);
                }
            }
        }
    }
    __syncwarp();

    // ---- P2: frames (sequential chains, oracle order) ----
    for (int task = lane; task < k * 9; task += 32) {
        const int e = task / 9, xm = task - e * 9, x = xm / 3, m = xm - x * 3;
        const float* dv = ves + e * S::ES + x * S::XS;
        const float* cv = vc + x * S::XS;
        const float* wz = Wz + m * S::CVE;
        float acc = 0.0f;
#pragma unroll
        for (int d = 0; d < CV; ++d) acc = __fmaf_rn(dv[d], wz[d], acc);
#pragma unroll
        for (int d = 0; d < CV; ++d) acc = __fmaf_rn(cv[d], wz[CV + d], acc);
        if (p.zscale) acc = __fmul_rn(acc, __ldg(p.zscale + m));
        zb[e * 12 + m * 4 + x] = acc;
    }
    __syncwarp();

    // ---- P3: q[e][3d+m] -> qs[t][e] ----
    {
        constexpr int QT = (KQ + 31) / 32;
        const float* qsrc[QT];
        int qstr[QT], zoff[QT];
        bool qok[QT];
#pragma unroll
        for (int t = 0; t < QT; ++t) {
            const int tq = 32 * t + lane, d = tq / 3, m = tq - d * 3;
            qok[t] = tq < KQ;
            zoff[t] = 4 * m;
            if (d < CV) { qsrc[t] = ves + d; qstr[t] = S::ES; }
            else { qsrc[t] = vc + (qok[t] ? d - CV : 0); qstr[t] = 0; }
        }
        const float* z = zb;
#pragma unroll 2
        for (int e = 0; e < k; ++e, z += 12) {
#pragma unroll
            for (int t = 0; t < QT; ++t) {
                const float* src = qsrc[t];
                qsrc[t] += qstr[t];
                const float4 zz = *reinterpret_cast<const float4*>(z + zoff[t]);         // column m of the frame
                float q = __fmul_rn(src[0], zz.x);
                q = __fmaf_rn(src[S::XS], zz.y, q);
                q = __fmaf_rn(src[2 * S::XS], zz.z, q);
                if (qok[t]) qs[(32 * t + lane) * kp + e] = q;
            }
        }
    }
    __syncwarp();

    // ---- P4: dense fp32 linear1 + BN + LeakyReLU + max over edges; lanes = output channels ----
#pragma unroll 1
    for (int ob = 0; ob < S::OPT; ob += S::OPP) {
        float smax[S::OPP];
        float yi[S::OPP];
#pragma unroll
        for (int oo = 0; oo < S::OPP; ++oo) {
            const int o = lane + 32 * (ob + oo);
            smax[oo] = -INFINITY;
            yi[oo] = __ldg(p.Yab + r * 2 * COUT + COUT + o) - __ldg(p.Yab + r * 2 * COUT + o);   // Yb_i - Ya_i
        }
        for (int eb = 0; eb < k; eb += EB) {
            float acc[EB][S::OPP];
#pragma unroll
            for (int e = 0; e < EB; ++e) {
                const float* yj = p.Yab + cbase * 2 * COUT + (unsigned)nidx[eb + e] * (unsigned)(2 * COUT) + lane + 32 * ob;
#pragma unroll
                for (int oo = 0; oo < S::OPP; ++oo) acc[e][oo] = __ldg(yj + 32 * oo) + yi[oo];
            }
#pragma unroll 2
            for (int t = 0; t < KQ; ++t) {
                float wv[S::OPP];
#pragma unroll
                for (int oo = 0; oo < S::OPP; ++oo) wv[oo] = __ldg(p.W1q_t + (size_t)t * COUT + lane + 32 * (ob + oo));
                const float4* qp = reinterpret_cast<const float4*>(qs + t * kp + eb);
#pragma unroll
                for (int e4 = 0; e4 < EB / 4; ++e4) {
                    const float4 q4 = qp[e4];
                    const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int oo = 0; oo < S::OPP; ++oo) acc[e4 * 4 + u][oo] = fmaf(qv[u], wv[oo], acc[e4 * 4 + u][oo]);
                }
            }
#pragma unroll
            for (int oo = 0; oo < S::OPP; ++oo) {
                const int o = lane + 32 * (ob + oo);
                const float a1 = __ldg(p.bn1_a + o), c1 = __ldg(p.bn1_c + o);
#pragma unroll
                for (int e = 0; e < EB; ++e) {
                    if (eb + e < k) {
                        float y = __fadd_rn(__fmul_rn(acc[e][oo], a1), c1);
                        y = y > 0.0f ? y : __fmul_rn(0.2f, y);
                        smax[oo] = fmaxf(smax[oo], y);
                    }
                }
            }
        }
#pragma unroll
        for (int oo = 0; oo < S::OPP; ++oo) p.out.s[r * p.out.lds + lane + 32 * (ob + oo)] = smax[oo];
    }

    // ---- P5: vector branch ----
    vector_branch<CVO>(p, r, b, cbase, nidx, k, lane);
}

template <int CS, int CV, int COUT, int CVO>
int launch_fp_fast(const svnet_edge_params* p, cudaStream_t st)
{
    using S = Shape<CS, CV, COUT, CVO>;
    constexpr int EB = S::EB;
    constexpr int KQ = 6 * CV;
    const int kp = ((p->k + EB - 1) / EB) * EB;
    constexpr int SHARED = (3 * S::CVE + 3) & ~3;
    const int per_warp = (KQ * kp + kp * 12 + 3 * S::XS + kp * S::ES + kp + 3) & ~3;
    const long P = (long)p->B * p->N;
    auto smem_for = [&](int warps) { return sizeof(float) * (size_t)(SHARED + warps * per_warp); };
    if (smem_for(8) <= 100 * 1024) {
        const size_t smem = smem_for(8);
        SV_CUDA(cudaFuncSetAttribute(edge_fp_fast_kernel<CS, CV, COUT, CVO, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_fp_fast_kernel<CS, CV, COUT, CVO, 8><<<sv_cdiv(P, 8), 256, smem, st>>>(*p, kp);
    } else {
        const size_t smem = smem_for(4);
        SV_REQUIRE(smem <= 200 * 1024, "svnet_svblock_edge_fwd: k=%d too large for the fused fp kernel", p->k);
        SV_CUDA(cudaFuncSetAttribute(edge_fp_fast_kernel<CS, CV, COUT, CVO, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_fp_fast_kernel<CS, CV, COUT, CVO, 4><<<sv_cdiv(P, 4), 128, smem, st>>>(*p, kp);
    }
    SV_CHECK_LAUNCH("svnet_svblock_edge_fwd(fp fast)");
    return SVNET_OK;
}

template <int CS, int CV, int COUT, int CVO>
int launch_fast(const svnet_edge_params* p, cudaStream_t st)
{
    using S = Shape<CS, CV, COUT, CVO>;
    constexpr int EB = S::EB;
    const int kp = ((p->k + EB - 1) / EB) * EB;
    constexpr int SHARED = 3 * S::CVE + ((3 * S::CVE) & 1) + S::KW * COUT;
    const int per_warp = ((((2 * S::KW * kp + 2 * kp + 3) & ~3) + kp * 12 + 3 * S::XS + kp * S::ES) + 3) & ~3;
    const long P = (long)p->B * p->N;
    auto smem_for = [&](int warps) { return sizeof(float) * (size_t)(((SHARED + 3) & ~3) + warps * per_warp); };
    if (smem_for(8) <= 72 * 1024) {
        const size_t smem = smem_for(8);
        SV_CUDA(cudaFuncSetAttribute(edge_bin_fast_kernel<CS, CV, COUT, CVO, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_bin_fast_kernel<CS, CV, COUT, CVO, 8><<<sv_cdiv(P, 8), 256, smem, st>>>(*p, kp);
    } else {
        const size_t smem = smem_for(4);
        SV_REQUIRE(smem <= 200 * 1024, "svnet_svblock_edge_fwd: k=%d too large for the fused kernel", p->k);
        SV_CUDA(cudaFuncSetAttribute(edge_bin_fast_kernel<CS, CV, COUT, CVO, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        edge_bin_fast_kernel<CS, CV, COUT, CVO, 4><<<sv_cdiv(P, 4), 128, smem, st>>>(*p, kp);
    }
    SV_CHECK_LAUNCH("svnet_svblock_edge_fwd(fast)");
    return SVNET_OK;
}

}  // namespace

// Returns 1 if a specialised kernel handled the layer, 0 if the caller must use the generic kernel,
// < 0 on error.
int svnet_edge_fast_dispatch(const svnet_edge_params* p, cudaStream_t st)
{
    const int cs = p->in.Cs, cv = p->in.Cv, co = p->Cout, cvo = p->Cvo;
    int rc = 1;
    // the specialised kernels address rows inside a cloud with 32-bit element offsets
    const long widest = p->in.ldv > p->in.lds ? p->in.ldv : p->in.lds;
    if ((long)p->N * (widest > 6l * cvo ? widest : 6l * cvo) >= (1l << 31) || (long)p->N * 2 * co >= (1l << 31)) return 0;
    if (!p->binary) {
#define FCASE(A, Bv, C, D) if (cs == A && cv == Bv && co == C && cvo == D) { rc = launch_fp_fast<A, Bv, C, D>(p, st); return rc == SVNET_OK ? 1 : rc; }
        FCASE(32, 10, 32, 10)
        FCASE(32, 10, 64, 21)
        FCASE(64, 21, 128, 42)
        FCASE(32, 16, 32, 16)
        FCASE(32, 16, 64, 24)
        FCASE(64, 24, 128, 40)
#undef FCASE
        return 0;
    }
#define CASE(A, Bv, C, D) if (cs == A && cv == Bv && co == C && cvo == D) { rc = launch_fast<A, Bv, C, D>(p, st); return rc == SVNET_OK ? 1 : rc; }
    CASE(32, 10, 32, 10)    // SV_DGCNN_CLS conv2
    CASE(32, 10, 64, 21)    // SV_DGCNN_CLS conv3
    CASE(64, 21, 128, 42)   // SV_DGCNN_CLS conv4
    CASE(32, 16, 32, 16)    // SV_DGCNN_PSEG conv2
    CASE(32, 16, 64, 24)    // SV_DGCNN_PSEG conv3
    CASE(64, 24, 128, 40)   // SV_DGCNN_PSEG conv4
#undef CASE
    return 0;
}
