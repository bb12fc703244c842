// ATTR_Enhance (models.py:146-180): self-attention of the sentence code over its attribute
// codes — Tk = 1 + attr_num tokens of ntf = D channels per sample:
//     combine = [sent ; attrs]                                   [Tk][D]                        (:161-162)
//     q, k, v = combine Wq^T + bq, combine Wk^T + bk, combine Wv^T + bv                         (:163-165)
//     a       = softmax_j(q k^T) * (1 / sqrt(D))                 scale AFTER the softmax        (:166)
//     out     = a v ;  attn_sent = out[0]                                                       (:167-168)
// The reference runs 3 Linear + cat + permute + 2 bmm + softmax + mul (about ten launches of a few microseconds of
// work each); here the forward is ONE launch (CTA = sample, everything in shared memory) and the backward two:
// per-sample gradients (d_q/d_k/d_v, d_sent, d_attrs), then the weight / bias gradients as a deterministic
// reduction over all B * Tk token rows (no atomics).
#include "common.cuh"

namespace eegan {

constexpr int AE_MAXTK = 8;
constexpr int AE_THREADS = 256;

// stash per sample: q, k, v [3][Tk][D] and the softmax p [Tk][Tk] (before the 1/sqrt(D) factor)
__global__ void __launch_bounds__(AE_THREADS) attr_enhance_fwd_kernel(const float* __restrict__ sent, const float* __restrict__ attrs,
                                                                      const float* __restrict__ Wq, const float* __restrict__ bq,
                                                                      const float* __restrict__ Wk, const float* __restrict__ bk,
                                                                      const float* __restrict__ Wv, const float* __restrict__ bv,
                                                                      int D, int Tk, float norm, float* __restrict__ out,
                                                                      float* __restrict__ qkv, float* __restrict__ pst) {
    extern __shared__ float sm[];
    float* tok = sm;                 // [Tk][D]
    float* q = tok + Tk * D;         // [3][Tk][D]: q, k, v
    float* s = q + 3 * Tk * D;       // [Tk][Tk]
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < Tk * D; i += blockDim.x) {
        const int t = i / D, d = i - t * D;
        tok[i] = t == 0 ? sent[(size_t)b * D + d] : attrs[((size_t)b * (Tk - 1) + (t - 1)) * D + d];
    }
    __syncthreads();
    // projections: one warp per (matrix, output channel); lanes split the reduction over the input channels
    for (int mo = warp; mo < 3 * D; mo += nw) {
        const int m = mo / D, o = mo - m * D;
        const float* W = (m == 0 ? Wq : m == 1 ? Wk : Wv) + (size_t)o * D;
        float acc[AE_MAXTK];
#pragma unroll
        for (int t = 0; t < AE_MAXTK; ++t) acc[t] = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float w = __ldg(W + c);
#pragma unroll
            for (int t = 0; t < AE_MAXTK; ++t)
                if (t < Tk) acc[t] = fmaf(w, tok[t * D + c], acc[t]);
        }
        const float bias = (m == 0 ? bq : m == 1 ? bk : bv)[o];
#pragma unroll
        for (int t = 0; t < AE_MAXTK; ++t)
            if (t < Tk) {
                const float v = warp_sum(acc[t]);
                if (lane == 0) q[(m * Tk + t) * D + o] = v + bias;
            }
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
    if (qkv)
        for (int i = threadIdx.x; i < 3 * Tk * D; i += blockDim.x) qkv[(size_t)b * 3 * Tk * D + i] = q[i];
    if (pst && threadIdx.x < Tk * Tk) pst[(size_t)b * Tk * Tk + threadIdx.x] = s[threadIdx.x];
}

// per sample: d_out [Tk][D] (= d_attn_attrs, plus d_attn_sent on row 0) -> g = (d_q, d_k, d_v) [3][Tk][D] and the token
// gradients d_sent / d_attrs = d_q Wq + d_k Wk + d_v Wv
__global__ void __launch_bounds__(AE_THREADS) attr_enhance_bwd_kernel(const float* __restrict__ d_attn_sent,
                                                                      const float* __restrict__ d_attn_attrs,
                                                                      const float* __restrict__ qkv, const float* __restrict__ pst,
                                                                      const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                                      const float* __restrict__ Wv, int D, int Tk, float norm,
                                                                      float* __restrict__ g, float* __restrict__ d_sent,
                                                                      float* __restrict__ d_attrs) {
    extern __shared__ float sm[];
    float* q = sm;                   // [3][Tk][D]
    float* dout = q + 3 * Tk * D;    // [Tk][D]
    float* dg = dout + Tk * D;       // [3][Tk][D]: d_q, d_k, d_v
    float* p = dg + 3 * Tk * D;      // [Tk][Tk]
    float* ds = p + Tk * Tk;         // [Tk][Tk]
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 3 * Tk * D; i += blockDim.x) q[i] = qkv[(size_t)b * 3 * Tk * D + i];
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
        dg[i] = aq;
        dg[Tk * D + i] = ak;
        dg[2 * Tk * D + i] = av;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * Tk * D; i += blockDim.x) g[(size_t)b * 3 * Tk * D + i] = dg[i];
    // token gradients: thread = input channel c (coalesced weight rows), loop over the output channels
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float acc[AE_MAXTK];
#pragma unroll
        for (int t = 0; t < AE_MAXTK; ++t) acc[t] = 0.f;
        for (int o = 0; o < D; ++o) {
            const float wq = __ldg(Wq + (size_t)o * D + c), wk = __ldg(Wk + (size_t)o * D + c), wv = __ldg(Wv + (size_t)o * D + c);
#pragma unroll
            for (int t = 0; t < AE_MAXTK; ++t)
                if (t < Tk) acc[t] = fmaf(dg[t * D + o], wq, fmaf(dg[(Tk + t) * D + o], wk, fmaf(dg[(2 * Tk + t) * D + o], wv, acc[t])));
        }
#pragma unroll
        for (int t = 0; t < AE_MAXTK; ++t)
            if (t < Tk) {
                if (t == 0) {
                    if (d_sent) d_sent[(size_t)b * D + c] = acc[0];
                } else if (d_attrs) {
                    d_attrs[((size_t)b * (Tk - 1) + (t - 1)) * D + c] = acc[t];
                }
            }
    }
}

// dW_m[o][c] = sum over token rows (b, t) of g[b][m][t][o] * combine[b][t][c];  db_m[o] = sum of g[b][m][t][o]
// grid (D/32, D/8, 3): warp = output channel o, lane = input channel c; fixed summation order.
__global__ void __launch_bounds__(256) attr_enhance_dw_kernel(const float* __restrict__ g, const float* __restrict__ sent,
                                                              const float* __restrict__ attrs, int B, int D, int Tk,
                                                              float* __restrict__ dWq, float* __restrict__ dbq, float* __restrict__ dWk,
                                                              float* __restrict__ dbk, float* __restrict__ dWv, float* __restrict__ dbv) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), o = blockIdx.y * 8 + (threadIdx.x >> 5), m = blockIdx.z;
    if (o >= D) return;
    float acc = 0.f, accb = 0.f;
    for (int b = 0; b < B; ++b) {
        for (int t = 0; t < Tk; ++t) {
            const float gv = g[(((size_t)b * 3 + m) * Tk + t) * D + o];
            if (c < D) {
                const float x = t == 0 ? __ldg(sent + (size_t)b * D + c) : __ldg(attrs + ((size_t)b * (Tk - 1) + (t - 1)) * D + c);
                acc = fmaf(gv, x, acc);
            }
            accb += gv;
        }
    }
    float* dW = m == 0 ? dWq : m == 1 ? dWk : dWv;
    float* db = m == 0 ? dbq : m == 1 ? dbk : dbv;
    if (c < D && dW) dW[(size_t)o * D + c] = acc;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && db) db[o] = accb;
}

static int ae_check(int B, int D, int Tk) {
    EEGAN_REQUIRE(B > 0 && D > 0 && Tk >= 1 && Tk <= AE_MAXTK, "attr_enhance: B=%d D=%d tokens=%d (1..%d tokens)", B, D, Tk, AE_MAXTK);
    EEGAN_REQUIRE((size_t)(8 * Tk * D + 2 * Tk * Tk) * sizeof(float) <= 200 * 1024, "attr_enhance: D=%d too large for the shared-memory form", D);
    return EEGAN_OK;
}

}  // namespace eegan

using namespace eegan;

extern "C" int eegan_attr_enhance_fwd(const float* sent, const float* attrs, const float* Wq, const float* bq, const float* Wk,
                                      const float* bk, const float* Wv, const float* bv, int B, int D, int attr_num, float norm_fact,
                                      float* out, float* qkv, float* p, void* stream) {
    const int Tk = attr_num + 1;
    int rc = ae_check(B, D, Tk);
    if (rc) return rc;
    EEGAN_REQUIRE(sent && (attrs || attr_num == 0) && Wq && bq && Wk && bk && Wv && bv && out, "attr_enhance fwd: null pointer");
    const size_t smem = (size_t)(4 * Tk * D + Tk * Tk) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attr_enhance_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(attr_enhance_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    attr_enhance_fwd_kernel<<<B, AE_THREADS, smem, (cudaStream_t)stream>>>(sent, attrs, Wq, bq, Wk, bk, Wv, bv, D, Tk, norm_fact, out, qkv, p);
    return check_launch("attr_enhance fwd");
}

extern "C" int eegan_attr_enhance_bwd(const float* d_attn_sent, const float* d_attn_attrs, const float* sent, const float* attrs,
                                      const float* qkv, const float* p, const float* Wq, const float* Wk, const float* Wv, int B, int D,
                                      int attr_num, float norm_fact, float* g, float* d_sent, float* d_attrs, float* dWq, float* dbq,
                                      float* dWk, float* dbk, float* dWv, float* dbv, void* stream) {
    const int Tk = attr_num + 1;
    int rc = ae_check(B, D, Tk);
    if (rc) return rc;
    EEGAN_REQUIRE(sent && qkv && p && Wq && Wk && Wv && g, "attr_enhance bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(7 * Tk * D + 2 * Tk * Tk) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attr_enhance_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    attr_enhance_bwd_kernel<<<B, AE_THREADS, smem, st>>>(d_attn_sent, d_attn_attrs, qkv, p, Wq, Wk, Wv, D, Tk, norm_fact, g, d_sent, d_attrs);
    EEGAN_LAUNCH_CHECK("attr_enhance bwd");
    if (dWq || dWk || dWv || dbq || dbk || dbv) {
        attr_enhance_dw_kernel<<<dim3((D + 31) / 32, (D + 7) / 8, 3), 256, 0, st>>>(g, sent, attrs, B, D, Tk, dWq, dbq, dWk, dbk, dWv, dbv);
        EEGAN_LAUNCH_CHECK("attr_enhance dW");
    }
    return EEGAN_OK;
}
