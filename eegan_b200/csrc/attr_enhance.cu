// ATTR_Enhance (models.py:146-180): self-attention of the sentence code over its attribute
// codes — Tk = 1 + attr_num tokens of ntf = D channels per sample:
//     combine = [sent ; attrs]                                   [B*Tk][D]                      (:161-162)
//     q, k, v = combine Wq^T + bq, combine Wk^T + bk, combine Wv^T + bv                         (:163-165)
//     a       = softmax_j(q k^T) * (1 / sqrt(D))                 scale AFTER the softmax        (:166)
//     out     = a v ;  attn_sent = out[0]                                                       (:167-168)
// Everything is a few hundred kFLOP per sample, so the design goal is few, wide launches:
//   fwd  pack (cat) -> one small-GEMM launch for the three projections (grid.z = q/k/v) -> per-sample attention
//   bwd  per-sample attention backward (d_q, d_k, d_v) -> one small-GEMM launch for the token gradients
//        (K runs over the three projections) -> one for dW_q/k/v (grid.z) -> bias gradients
// The small GEMM is a plain 32x32x32 shared-memory tile kernel on the CUDA cores (M = B*Tk = 192 rows: far too small
// for the tensor-core engine's ~10 us launch ramp).  Weight gradients are plain sums in a fixed order: no atomics.
#include "common.cuh"

namespace eegan {

constexpr int AE_MAXTK = 8;

// C_i[m][n] = sum over segments s, k of A_s(m,k) * B_s(n,k) (+ bias_i[n]);  element (m,k) of A at A[m*sAm + k*sAk].
// grid.z = independent problems i (nseg == 1: pointers indexed by z), or nseg K-segments accumulated into one C.
struct SgArgs {
    const float* A[3];
    const float* B[3];
    float* C[3];
    const float* bias[3];
    int M, N, K, nseg;
    long long sAm, sAk, sBn, sBk, ldc;
};

__global__ void __launch_bounds__(256) ae_sgemm_kernel(const SgArgs a) {
    __shared__ float As[32][33], Bs[32][33];  // [k][m], [k][n]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32, z = blockIdx.z;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int s = 0; s < a.nseg; ++s) {
        const int i = a.nseg > 1 ? s : z;
        const float* A = a.A[i];
        const float* B = a.B[i];
        for (int k0 = 0; k0 < a.K; k0 += 32) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = tid + 256 * e;
                // lanes run along the operand's contiguous index
                const int am = a.sAk == 1 ? idx >> 5 : idx & 31, ak = a.sAk == 1 ? idx & 31 : idx >> 5;
                const int bn = a.sBk == 1 ? idx >> 5 : idx & 31, bk = a.sBk == 1 ? idx & 31 : idx >> 5;
                As[ak][am] = (m0 + am < a.M && k0 + ak < a.K) ? __ldg(A + (long long)(m0 + am) * a.sAm + (long long)(k0 + ak) * a.sAk) : 0.f;
                Bs[bk][bn] = (n0 + bn < a.N && k0 + bk < a.K) ? __ldg(B + (long long)(n0 + bn) * a.sBn + (long long)(k0 + bk) * a.sBk) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float a0 = As[k][2 * ty], a1 = As[k][2 * ty + 1], b0 = Bs[k][2 * tx], b1 = Bs[k][2 * tx + 1];
                acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
                acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
            }
            __syncthreads();
        }
    }
    const int i = a.nseg > 1 ? 0 : z;
    float* C = a.C[i];
    const float* bias = a.bias[i];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int m = m0 + 2 * ty + r, n = n0 + 2 * tx + c;
            if (m < a.M && n < a.N) C[(long long)m * a.ldc + n] = acc[r][c] + (bias ? bias[n] : 0.f);
        }
}

static int ae_sgemm(const SgArgs& a, int nz, cudaStream_t st) {
    ae_sgemm_kernel<<<dim3((a.N + 31) / 32, (a.M + 31) / 32, nz), 256, 0, st>>>(a);
    return check_launch("attr_enhance gemm");
}

// combine[b*Tk + t][:] = t == 0 ? sent[b] : attrs[b][t-1]
__global__ void __launch_bounds__(256) ae_pack_kernel(const float* __restrict__ sent, const float* __restrict__ attrs, int B, int D,
                                                      int Tk, float* __restrict__ combine) {
    const long long n = (long long)B * Tk * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long long row = i / D;
        const int t = (int)(row % Tk), b = (int)(row / Tk);
        combine[i] = t == 0 ? sent[(size_t)b * D + d] : attrs[((size_t)b * (Tk - 1) + (t - 1)) * D + d];
    }
}

// per sample: scores, softmax, out = (norm p) v.   qkv is [3][B*Tk][D];  p stash [B][Tk][Tk] (before the norm factor)
__global__ void __launch_bounds__(256) ae_attn_fwd_kernel(const float* __restrict__ qkv, int B, int D, int Tk, float norm,
                                                          float* __restrict__ out, float* __restrict__ pst) {
    extern __shared__ float sm[];
    float* q = sm;                   // [3][Tk][D]
    float* s = q + 3 * Tk * D;       // [Tk][Tk]
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t plane = (size_t)B * Tk * D;
    for (int i = threadIdx.x; i < 3 * Tk * D; i += blockDim.x) {
        const int m = i / (Tk * D), r = i - m * Tk * D;
        q[i] = qkv[m * plane + (size_t)b * Tk * D + r];
    }
    __syncthreads();
    const float* kk = q + Tk * D;
    const float* vv = q + 2 * Tk * D;
    for (int ij = warp; ij < Tk * Tk; ij += nw) {
        const int i = ij / Tk, j = ij - i * Tk;
        float a = 0.f;
        for (int c = lane; c < D; c += 32) a = fmaf(q[i * D + c], kk[j * D + c], a);
        a = warp_sum(a);
        if (lane == 0) s[ij] = a;
    }
    __syncthreads();
    if (threadIdx.x < Tk) {  // softmax over j (:166)
        const int i = threadIdx.x;
        float mx = -INFINITY;
        for (int j = 0; j < Tk; ++j) mx = fmaxf(mx, s[i * Tk + j]);
        float sum = 0.f;
        for (int j = 0; j < Tk; ++j) {
            const float e = expf(s[i * Tk + j] - mx);
            s[i * Tk + j] = e;
            sum += e;
        }
        for (int j = 0; j < Tk; ++j) s[i * Tk + j] /= sum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Tk * D; i += blockDim.x) {
        const int t = i / D, d = i - t * D;
        float a = 0.f;
        for (int j = 0; j < Tk; ++j) a = fmaf(s[t * Tk + j] * norm, vv[j * D + d], a);
        out[(size_t)b * Tk * D + i] = a;
    }
    if (pst && threadIdx.x < Tk * Tk) pst[(size_t)b * Tk * Tk + threadIdx.x] = s[threadIdx.x];
}

// per sample: d_out [Tk][D] -> g = (d_q, d_k, d_v), stored [3][B*Tk][D]
__global__ void __launch_bounds__(256) ae_attn_bwd_kernel(const float* __restrict__ d_attn_sent, const float* __restrict__ d_attn_attrs,
                                                          const float* __restrict__ qkv, const float* __restrict__ pst, int B, int D,
                                                          int Tk, float norm, float* __restrict__ g) {
    extern __shared__ float sm[];
    float* q = sm;                   // [3][Tk][D]
    float* dout = q + 3 * Tk * D;    // [Tk][D]
    float* p = dout + Tk * D;        // [Tk][Tk]
    float* ds = p + Tk * Tk;         // [Tk][Tk]
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t plane = (size_t)B * Tk * D;
    for (int i = threadIdx.x; i < 3 * Tk * D; i += blockDim.x) {
        const int m = i / (Tk * D), r = i - m * Tk * D;
        q[i] = qkv[m * plane + (size_t)b * Tk * D + r];
    }
    for (int i = threadIdx.x; i < Tk * D; i += blockDim.x) {
        float v = d_attn_attrs ? d_attn_attrs[(size_t)b * Tk * D + i] : 0.f;
        if (i < D && d_attn_sent) v += d_attn_sent[(size_t)b * D + i];
        dout[i] = v;
    }
    if (threadIdx.x < Tk * Tk) p[threadIdx.x] = pst[(size_t)b * Tk * Tk + threadIdx.x];
    __syncthreads();
    const float* kk = q + Tk * D;
    const float* vv = q + 2 * Tk * D;
    for (int ij = warp; ij < Tk * Tk; ij += nw) {  // dp[i][j] = norm <d_out_i, v_j>
        const int i = ij / Tk, j = ij - i * Tk;
        float a = 0.f;
        for (int c = lane; c < D; c += 32) a = fmaf(dout[i * D + c], vv[j * D + c], a);
        a = warp_sum(a);
        if (lane == 0) ds[ij] = a * norm;
    }
    __syncthreads();
    if (threadIdx.x < Tk) {  // softmax backward per row
        const int i = threadIdx.x;
        float dot = 0.f;
        for (int j = 0; j < Tk; ++j) dot = fmaf(p[i * Tk + j], ds[i * Tk + j], dot);
        for (int j = 0; j < Tk; ++j) ds[i * Tk + j] = p[i * Tk + j] * (ds[i * Tk + j] - dot);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Tk * D; i += blockDim.x) {
        const int t = i / D, d = i - t * D;
        float aq = 0.f, ak = 0.f, av = 0.f;
        for (int j = 0; j < Tk; ++j) {
            aq = fmaf(ds[t * Tk + j], kk[j * D + d], aq);            // d_q[t] = sum_j ds[t][j] k_j
            ak = fmaf(ds[j * Tk + t], q[j * D + d], ak);             // d_k[t] = sum_i ds[i][t] q_i
            av = fmaf(p[j * Tk + t] * norm, dout[j * D + d], av);    // d_v[t] = sum_i a[i][t] d_out_i
        }
        const size_t o = (size_t)b * Tk * D + i;
        g[o] = aq;
        g[plane + o] = ak;
        g[2 * plane + o] = av;
    }
}

// token gradients [B*Tk][D] -> d_sent [B][D], d_attrs [B][Tk-1][D]; bias gradients db_m[o] = sum over rows of g_m[row][o]
__global__ void __launch_bounds__(256) ae_unpack_kernel(const float* __restrict__ dtok, const float* __restrict__ g, int B, int D, int Tk,
                                                        float* __restrict__ d_sent, float* __restrict__ d_attrs, float* __restrict__ dbq,
                                                        float* __restrict__ dbk, float* __restrict__ dbv) {
    const long long n = (long long)B * Tk * D, i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int d = (int)(i % D);
        const long long row = i / D;
        const int t = (int)(row % Tk), b = (int)(row / Tk);
        if (t == 0) {
            if (d_sent) d_sent[(size_t)b * D + d] = dtok[i];
        } else if (d_attrs) {
            d_attrs[((size_t)b * (Tk - 1) + (t - 1)) * D + d] = dtok[i];
        }
    }
    if (i < 3LL * D) {  // the first 3 D threads also own one bias-gradient entry each
        const int m = (int)(i / D), o = (int)(i % D);
        float* db = m == 0 ? dbq : m == 1 ? dbk : dbv;
        if (db) {
            float s = 0.f;
            const float* gm = g + (size_t)m * B * Tk * D + o;
            for (int r = 0; r < B * Tk; ++r) s += gm[(size_t)r * D];
            db[o] = s;
        }
    }
}

static int ae_check(int B, int D, int Tk) {
    EEGAN_REQUIRE(B > 0 && D > 0 && Tk >= 1 && Tk <= AE_MAXTK, "attr_enhance: B=%d D=%d tokens=%d (1..%d tokens)", B, D, Tk, AE_MAXTK);
    EEGAN_REQUIRE((size_t)(4 * Tk * D + 2 * Tk * Tk) * sizeof(float) <= 200 * 1024, "attr_enhance: D=%d too large for the shared-memory form", D);
    return EEGAN_OK;
}
static int ae_smem_attr() {
    static SmemGrant gf, gb;
    if (int rc = grant_dyn_smem(ae_attn_fwd_kernel, (size_t)200 * 1024, gf, "attr_enhance fwd")) return rc;
    return grant_dyn_smem(ae_attn_bwd_kernel, (size_t)200 * 1024, gb, "attr_enhance bwd");
}

}  // namespace eegan

using namespace eegan;

// scratch for both directions: combine [B*Tk][D] (fwd, kept for the backward) / dtok [B*Tk][D] (bwd)
extern "C" size_t eegan_attr_enhance_workspace_bytes(int B, int D, int attr_num) {
    if (B <= 0 || D <= 0 || attr_num < 0) return 0;
    return align_up((size_t)B * (attr_num + 1) * D * sizeof(float), 256);
}

extern "C" int eegan_attr_enhance_fwd(const float* sent, const float* attrs, const float* Wq, const float* bq, const float* Wk,
                                      const float* bk, const float* Wv, const float* bv, int B, int D, int attr_num, float norm_fact,
                                      float* out, float* qkv, float* p, float* combine, void* stream) {
    const int Tk = attr_num + 1;
    int rc = ae_check(B, D, Tk);
    if (rc) return rc;
    EEGAN_REQUIRE(sent && (attrs || attr_num == 0) && Wq && bq && Wk && bk && Wv && bv && out && qkv && combine,
                  "attr_enhance fwd: null pointer (qkv [3,B*Tk,D] and combine [B*Tk,D] are required scratch / stash)");
    cudaStream_t st = (cudaStream_t)stream;
    rc = ae_smem_attr();
    if (rc) return rc;
    const int rows = B * Tk;
    ae_pack_kernel<<<(unsigned)(((long long)rows * D + 255) / 256), 256, 0, st>>>(sent, attrs, B, D, Tk, combine);
    EEGAN_LAUNCH_CHECK("attr_enhance pack");
    SgArgs a{};
    const size_t plane = (size_t)rows * D;
    const float* W[3] = {Wq, Wk, Wv};
    const float* bb[3] = {bq, bk, bv};
    for (int m = 0; m < 3; ++m) { a.A[m] = combine; a.B[m] = W[m]; a.C[m] = qkv + m * plane; a.bias[m] = bb[m]; }
    a.M = rows; a.N = D; a.K = D; a.nseg = 1; a.sAm = D; a.sAk = 1; a.sBn = D; a.sBk = 1; a.ldc = D;
    rc = ae_sgemm(a, 3, st);
    if (rc) return rc;
    ae_attn_fwd_kernel<<<B, 256, (size_t)(3 * Tk * D + Tk * Tk) * sizeof(float), st>>>(qkv, B, D, Tk, norm_fact, out, p);
    return check_launch("attr_enhance attention");
}

extern "C" int eegan_attr_enhance_bwd(const float* d_attn_sent, const float* d_attn_attrs, const float* combine, const float* qkv,
                                      const float* p, const float* Wq, const float* Wk, const float* Wv, int B, int D, int attr_num,
                                      float norm_fact, float* g, float* dtok, float* d_sent, float* d_attrs, float* dWq, float* dbq,
                                      float* dWk, float* dbk, float* dWv, float* dbv, void* stream) {
    const int Tk = attr_num + 1;
    int rc = ae_check(B, D, Tk);
    if (rc) return rc;
    EEGAN_REQUIRE(combine && qkv && p && Wq && Wk && Wv && g && dtok, "attr_enhance bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    rc = ae_smem_attr();
    if (rc) return rc;
    const int rows = B * Tk;
    const size_t plane = (size_t)rows * D;
    ae_attn_bwd_kernel<<<B, 256, (size_t)(4 * Tk * D + 2 * Tk * Tk) * sizeof(float), st>>>(d_attn_sent, d_attn_attrs, qkv, p, B, D, Tk,
                                                                                          norm_fact, g);
    EEGAN_LAUNCH_CHECK("attr_enhance attention bwd");
    const float* W[3] = {Wq, Wk, Wv};
    if (d_sent || d_attrs) {  // dtok[row][c] = sum_m sum_o g_m[row][o] W_m[o][c]
        SgArgs a{};
        for (int m = 0; m < 3; ++m) { a.A[m] = g + m * plane; a.B[m] = W[m]; }
        a.C[0] = dtok;
        a.M = rows; a.N = D; a.K = D; a.nseg = 3; a.sAm = D; a.sAk = 1; a.sBn = 1; a.sBk = D; a.ldc = D;
        rc = ae_sgemm(a, 1, st);
        if (rc) return rc;
    }
    if (dWq || dWk || dWv) {  // dW_m[o][c] = sum_rows g_m[row][o] combine[row][c]
        SgArgs a{};
        float* dW[3] = {dWq, dWk, dWv};
        EEGAN_REQUIRE(dWq && dWk && dWv, "attr_enhance bwd: the three weight gradients are computed together");
        for (int m = 0; m < 3; ++m) { a.A[m] = g + m * plane; a.B[m] = combine; a.C[m] = dW[m]; }
        a.M = D; a.N = D; a.K = rows; a.nseg = 1; a.sAm = 1; a.sAk = D; a.sBn = 1; a.sBk = D; a.ldc = D;
        rc = ae_sgemm(a, 3, st);
        if (rc) return rc;
    }
    const long long nun = (long long)rows * D > 3LL * D ? (long long)rows * D : 3LL * D;  // the bias gradients ride on the first 3 D threads
    ae_unpack_kernel<<<(unsigned)((nun + 255) / 256), 256, 0, st>>>(dtok, g, B, D, Tk, d_sent, d_attrs, dbq, dbk, dbv);
    return check_launch("attr_enhance unpack");
}
