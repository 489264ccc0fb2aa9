/*
 * svnet_oracle.c -- CPU restatement of the SVNet inference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for svnet_b200's CUDA kernels.  It is never linked, imported or
 * executed by the product path (svnet_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs call it (through oracle/svnet_oracle.py).
 *
 * It restates, function by function, what the reference computes in eval mode
 * (paths relative to /root/reference):
 *     models/utils/sv_util.py:19-25    knn                      -> orc_knn
 *     models/utils/sv_util.py:28-88    get_graph_feature[_cross]-> orc_graph_feature_xyz
 *     models/utils/sv_util.py:90-116   get_graph_feature_sv     -> orc_graph_feature_sv
 *     models/utils/sv_util.py:118-132  svpool                   -> orc_pool_max / orc_pool_mean
 *     models/sv_layers.py:29-53,64-78  Linear / Conv1d (bw, ba) -> orc_linear, orc_sign_plane
 *     models/sv_layers.py:86-102       VectorBN                 -> orc_vector_bn
 *     models/sv_layers.py:111-129      Vector2Scalar            -> orc_v2s
 *     torch.nn.BatchNorm1d (eval) + LeakyReLU/ReLU              -> orc_bn_act
 *     models/sv_layers.py:179-183      SVBlock gate             -> orc_gate
 *
 * Pinning: the reference has no tests or golden vectors.  This oracle is
 * pinned against outputs of the reference itself, produced in the build container by
 * tests/golden/make_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py).
 *
 * Arithmetic contract (DESIGN.md "numerics"): fp32 throughout.  The reference's summation orders
 * are whatever ATen's sgemm/sum pick; where a result feeds a discontinuity (top-k ranking,
 * sign()) this oracle fixes the natural sequential order below and the CUDA kernels reproduce it
 * bit for bit:
 *     dot products      acc = 0; for c ascending: acc = fmaf(a[c], b[c], acc)
 *     squared norms     the same chain with a == b
 *     kNN score         p_ij = ((-xx_j) - (-2*dot_ij)) - xx_i            (sv_util.py:20-22)
 *     kNN order         larger p first, equal p -> smaller index first
 *     V2S               z[x][m] chain over channels; q[d][m] = v0*z0, then fmaf(v1,z1,.), fmaf(v2,z2,.)
 *     sign input        one fp32 add  u + beta
 * Compile with -ffp-contract=off so that nothing but the explicit fmaf() calls is fused.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_FLAG_BW 1
#define ORC_FLAG_BA 2

int orc_version(void) { return 1; }

static inline float sgnf(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

/* ---------------------------------------------------------------------------------------------
 * kNN (sv_util.py:19-25).  feat is point-major [B][N][C] (the reference passes the transposed
 * view (B,C,N); same numbers).  idx[B][N][k], nearest first, self included.
 * Optional pd_out[B][N][N] receives the score matrix (tests only; may be NULL).
 * ------------------------------------------------------------------------------------------- */
void orc_knn(const float* feat, int B, int N, int C, int k, int64_t* idx, float* pd_out)
{
    #pragma omp parallel
    {
        float* xx = (float*)malloc(sizeof(float) * (size_t)N);
        float* row = (float*)malloc(sizeof(float) * (size_t)N);
        float* bv = (float*)malloc(sizeof(float) * (size_t)k);
        int* bi = (int*)malloc(sizeof(int) * (size_t)k);
        for (int b = 0; b < B; ++b) {
            const float* f = feat + (size_t)b * N * C;
            for (int j = 0; j < N; ++j) {
                float a = 0.0f;
                for (int c = 0; c < C; ++c) a = fmaf(f[(size_t)j * C + c], f[(size_t)j * C + c], a);
                xx[j] = a;
            }
            #pragma omp for schedule(static)
            for (int i = 0; i < N; ++i) {
                const float* fi = f + (size_t)i * C;
                for (int j = 0; j < N; ++j) {
                    const float* fj = f + (size_t)j * C;
                    float dot = 0.0f;
                    for (int c = 0; c < C; ++c) dot = fmaf(fi[c], fj[c], dot);
                    float inner = -2.0f * dot;
                    float t = (-xx[j]) - inner;
                    row[j] = t - xx[i];
                }
                if (pd_out) memcpy(pd_out + ((size_t)b * N + i) * N, row, sizeof(float) * (size_t)N);
                /* insertion into a sorted list: (value desc, index asc) */
                int cnt = 0;
                for (int j = 0; j < N; ++j) {
                    float p = row[j];
                    if (cnt == k && !(p > bv[k - 1])) continue; /* equal value, larger index: loses */
                    int pos = (cnt < k) ? cnt : k - 1;
                    while (pos > 0 && p > bv[pos - 1]) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                    bv[pos] = p; bi[pos] = j;
                    if (cnt < k) ++cnt;
                }
                for (int e = 0; e < k; ++e) idx[((size_t)b * N + i) * k + e] = bi[e];
            }
        }
        free(xx); free(row); free(bv); free(bi);
    }
}

/* ---------------------------------------------------------------------------------------------
 * get_graph_feature (sv_util.py:28-62, first=False) nv == 2:  [x_j - x_i | x_i]
 * get_graph_feature_cross (sv_util.py:64-88)        nv == 3:  [x_j - x_i | x_i | x_j x x_i]
 * xyz [B][N][3]; out [B][N][k][3][nv]  (xyz outer, channel inner, sv_util.py:60,86)
 * ------------------------------------------------------------------------------------------- */
void orc_graph_feature_xyz(const float* xyz, const int64_t* idx, int B, int N, int k, int nv, float* out)
{
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * N; ++r) {
        int b = (int)(r / N);
        const float* xi = xyz + (size_t)r * 3;
        for (int e = 0; e < k; ++e) {
            const float* xj = xyz + ((size_t)b * N + idx[(size_t)r * k + e]) * 3;
            float* o = out + ((size_t)r * k + e) * 3 * nv;
            for (int a = 0; a < 3; ++a) {
                o[a * nv + 0] = xj[a] - xi[a];
                o[a * nv + 1] = xi[a];
            }
            if (nv == 3) { /* torch.cross(feature, x): a x b */
                o[0 * nv + 2] = xj[1] * xi[2] - xj[2] * xi[1];
                o[1 * nv + 2] = xj[2] * xi[0] - xj[0] * xi[2];
                o[2 * nv + 2] = xj[0] * xi[1] - xj[1] * xi[0];
            }
        }
    }
}

/* get_graph_feature_sv (sv_util.py:106-114): s_f = [s_j - s_i | s_i], v_f = [v_j - v_i | v_i] */
void orc_graph_feature_sv(const float* s, const float* v, const int64_t* idx, int B, int N, int k,
                          int Cs, int Cv, float* sf, float* vf)
{
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * N; ++r) {
        int b = (int)(r / N);
        const float* si = s + (size_t)r * Cs;
        const float* vi = v + (size_t)r * 3 * Cv;
        for (int e = 0; e < k; ++e) {
            size_t j = (size_t)b * N + idx[(size_t)r * k + e];
            const float* sj = s + j * Cs;
            const float* vj = v + j * 3 * Cv;
            float* so = sf + ((size_t)r * k + e) * 2 * Cs;
            float* vo = vf + ((size_t)r * k + e) * 3 * 2 * Cv;
            for (int c = 0; c < Cs; ++c) { so[c] = sj[c] - si[c]; so[Cs + c] = si[c]; }
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < Cv; ++c) {
                    vo[a * 2 * Cv + c] = vj[a * Cv + c] - vi[a * Cv + c];
                    vo[a * 2 * Cv + Cv + c] = vi[a * Cv + c];
                }
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Linear / Conv1d(kernel 1), eval mode (sv_layers.py:29-53, 64-78).
 * x [R][K], W [Cout][K], beta [K] or NULL, scale [Cout] or NULL, bias [Cout] or NULL.
 *   ba: x <- sign(x + beta)   bw: W <- sign(W)   y = (x W^T) * scale (+ bias)
 * ------------------------------------------------------------------------------------------- */
void orc_linear(const float* x, const float* W, const float* beta, const float* scale, const float* bias,
                long R, int K, int Cout, int flags, float* y)
{
    const int bw = flags & ORC_FLAG_BW, ba = flags & ORC_FLAG_BA;
    float* Wq = (float*)malloc(sizeof(float) * (size_t)Cout * K);
    for (size_t i = 0; i < (size_t)Cout * K; ++i) Wq[i] = bw ? sgnf(W[i]) : W[i];
    #pragma omp parallel
    {
        float* xr = (float*)malloc(sizeof(float) * (size_t)K);
        #pragma omp for schedule(static)
        for (long r = 0; r < R; ++r) {
            for (int c = 0; c < K; ++c) {
                float t = x[(size_t)r * K + c];
                xr[c] = ba ? sgnf(t + beta[c]) : t;
            }
            for (int o = 0; o < Cout; ++o) {
                const float* w = Wq + (size_t)o * K;
                float acc = 0.0f;
                for (int c = 0; c < K; ++c) acc = fmaf(xr[c], w[c], acc);
                if (bw || ba) acc = acc * (scale ? scale[o] : 1.0f);
                if (bias) acc = acc + bias[o];
                y[(size_t)r * Cout + o] = acc;
            }
        }
        free(xr);
    }
    free(Wq);
}

/* sign(x + beta) as int8 in {-1,0,1} (sv_layers.py:37-39) */
void orc_sign_plane(const float* x, const float* beta, long R, int K, int8_t* out)
{
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < R; ++r)
        for (int c = 0; c < K; ++c) out[(size_t)r * K + c] = (int8_t)sgnf(x[(size_t)r * K + c] + beta[c]);
}

/* ---------------------------------------------------------------------------------------------
 * Vector2Scalar (sv_layers.py:111-129), multi == 3.
 * v [R][3][C]; W [3][C]; scale [3] (binary) or NULL.
 *   z[x][m] = sum_d v[x][d] * W[m][d]   (* scale[m], W <- sign W when binary)   sv_layers.py:116
 *   s[d*3+m] = sum_x v[x][d] * z[x][m]                                         sv_layers.py:117-125
 * z_out [R][3][3] optional (trans_back, sv_layers.py:126-127).
 * ------------------------------------------------------------------------------------------- */
void orc_v2s(const float* v, const float* W, const float* scale, long R, int C, int bw, float* s, float* z_out)
{
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < R; ++r) {
        const float* vr = v + (size_t)r * 3 * C;
        float z[3][3];
        for (int x = 0; x < 3; ++x)
            for (int m = 0; m < 3; ++m) {
                float acc = 0.0f;
                for (int d = 0; d < C; ++d) {
                    float w = W[m * C + d];
                    if (bw) w = sgnf(w);
                    acc = fmaf(vr[x * C + d], w, acc);
                }
                if (bw) acc = acc * scale[m];
                z[x][m] = acc;
            }
        for (int d = 0; d < C; ++d)
            for (int m = 0; m < 3; ++m) {
                float acc = vr[0 * C + d] * z[0][m];
                acc = fmaf(vr[1 * C + d], z[1][m], acc);
                acc = fmaf(vr[2 * C + d], z[2][m], acc);
                s[(size_t)r * 3 * C + d * 3 + m] = acc;
            }
        if (z_out) memcpy(z_out + (size_t)r * 9, z, sizeof(z));
    }
}

/* BatchNorm1d eval (running stats, eps) as ATen's inference transform y = x*a + c with
 * a = w/sqrt(var+eps), c = b - mean*a; then activation: 0 none, 1 LeakyReLU(0.2), 2 ReLU. */
void orc_bn_act(const float* x, const float* w, const float* b, const float* mean, const float* var,
                float eps, long R, int C, int act, float* y)
{
    float* a = (float*)malloc(sizeof(float) * (size_t)C);
    float* c0 = (float*)malloc(sizeof(float) * (size_t)C);
    for (int c = 0; c < C; ++c) {
        float inv = 1.0f / sqrtf(var[c] + eps);
        a[c] = w[c] * inv;
        c0[c] = b[c] - mean[c] * a[c];
    }
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) {
            float t = x[(size_t)r * C + c] * a[c] + c0[c];
            if (act == 1) t = (t > 0.0f) ? t : 0.2f * t;
            else if (act == 2) t = (t > 0.0f) ? t : 0.0f;
            y[(size_t)r * C + c] = t;
        }
    free(a); free(c0);
}

/* VectorBN (sv_layers.py:86-102): n = ||v||_2 over the xyz axis + 1e-6; out = v / n * BN(n). */
void orc_vector_bn(const float* v, const float* w, const float* b, const float* mean, const float* var,
                   float eps, long R, int C, float* out)
{
    float* a = (float*)malloc(sizeof(float) * (size_t)C);
    float* c0 = (float*)malloc(sizeof(float) * (size_t)C);
    for (int c = 0; c < C; ++c) {
        float inv = 1.0f / sqrtf(var[c] + eps);
        a[c] = w[c] * inv;
        c0[c] = b[c] - mean[c] * a[c];
    }
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < R; ++r) {
        const float* vr = v + (size_t)r * 3 * C;
        float* o = out + (size_t)r * 3 * C;
        for (int c = 0; c < C; ++c) {
            float v0 = vr[c], v1 = vr[C + c], v2 = vr[2 * C + c];
            float n = sqrtf(v0 * v0 + v1 * v1 + v2 * v2) + 1e-6f;
            float nb = n * a[c] + c0[c];
            o[c] = v0 / n * nb; o[C + c] = v1 / n * nb; o[2 * C + c] = v2 / n * nb;
        }
    }
    free(a); free(c0);
}

/* SVBlock gate (sv_layers.py:156-161,179-183): per cloud b, g = sigmoid(G2 relu(G1 mean_rows(s))).
 * s [B][rows][Cs]; G1 [H][Cs]; G2 [Co][H]; gate [B][Co]. */
void orc_gate(const float* s, const float* G1, const float* G2, int B, long rows, int Cs, int H, int Co,
              float* gate)
{
    #pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        float* m = (float*)calloc((size_t)Cs, sizeof(float));
        float* h = (float*)malloc(sizeof(float) * (size_t)(H > 0 ? H : 1));
        const float* sb = s + (size_t)b * rows * Cs;
        for (long r = 0; r < rows; ++r)
            for (int c = 0; c < Cs; ++c) m[c] += sb[(size_t)r * Cs + c];
        for (int c = 0; c < Cs; ++c) m[c] = m[c] / (float)rows;
        for (int j = 0; j < H; ++j) {
            float acc = 0.0f;
            for (int c = 0; c < Cs; ++c) acc = fmaf(m[c], G1[(size_t)j * Cs + c], acc);
            h[j] = acc > 0.0f ? acc : 0.0f;
        }
        for (int o = 0; o < Co; ++o) {
            float acc = 0.0f;
            for (int j = 0; j < H; ++j) acc = fmaf(h[j], G2[(size_t)o * H + j], acc);
            gate[(size_t)b * Co + o] = 1.0f / (1.0f + expf(-acc));
        }
        free(m); free(h);
    }
}

/* max / mean over the middle axis of x [A][M][C] -> [A][C]  (svpool, sv_util.py:125-131;
 * adaptive_{max,avg}_pool1d, sv_dgcnn_cls.py:72-73) */
void orc_pool_max(const float* x, long A, long M, long C, float* out)
{
    #pragma omp parallel for schedule(static)
    for (long a = 0; a < A; ++a)
        for (long c = 0; c < C; ++c) {
            float m = x[((size_t)a * M) * C + c];
            for (long j = 1; j < M; ++j) { float t = x[((size_t)a * M + j) * C + c]; if (t > m) m = t; }
            out[(size_t)a * C + c] = m;
        }
}

void orc_pool_mean(const float* x, long A, long M, long C, float* out)
{
    #pragma omp parallel for schedule(static)
    for (long a = 0; a < A; ++a)
        for (long c = 0; c < C; ++c) {
            float m = 0.0f;
            for (long j = 0; j < M; ++j) m += x[((size_t)a * M + j) * C + c];
            out[(size_t)a * C + c] = m / (float)M;
        }
}

/* out[r][x][c] = v[r][x][c] * gate[cloud(r)][c]   (sv_layers.py:194) */
void orc_scale_v(const float* v, const float* gate, int B, long rows, int C, float* out)
{
    #pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)B * rows; ++r) {
        const float* g = gate + (size_t)(r / rows) * C;
        for (int x = 0; x < 3; ++x)
            for (int c = 0; c < C; ++c)
                out[((size_t)r * 3 + x) * C + c] = v[((size_t)r * 3 + x) * C + c] * g[c];
    }
}
