// Two-way cross-entropy over a B x B score grid (tail of words_loss, DAMSM_losses.py:331-338,
// and of sent_loss :258-267) and the sentence cosine scores (:134-166, :248-258).
#include "common.cuh"

namespace eegan {

// blocks [0,B): row a -> scores_out[a][:] (scaled, class-masked) and row log-sum-exp
// blocks [B,2B): column b -> column log-sum-exp
__global__ void __launch_bounds__(256) pair_ce_lse_kernel(const float* __restrict__ in, float scale,
                                                          const int64_t* __restrict__ class_ids, int B,
                                                          float* __restrict__ out, float* __restrict__ lse) {
    __shared__ float red[32];
    pdl_trigger();
    pdl_wait();
    const bool is_row = blockIdx.x < B;
    const int fixed = is_row ? blockIdx.x : blockIdx.x - B;
    const int64_t cf = class_ids ? class_ids[fixed] : 0;
    float mx = -INFINITY;
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
        const int a = is_row ? fixed : k, b = is_row ? k : fixed;
        float v = in[(size_t)a * B + b] * scale;
        if (class_ids && k != fixed && class_ids[k] == cf) v = -INFINITY;
        if (is_row) out[(size_t)a * B + b] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    float s = 0.f;
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
        const int a = is_row ? fixed : k, b = is_row ? k : fixed;
        float v = in[(size_t)a * B + b] * scale;
        if (class_ids && k != fixed && class_ids[k] == cf) v = -INFINITY;
        s += expf(v - mx);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) lse[(is_row ? 0 : B) + fixed] = mx + logf(s);
}

// loss0 = mean_a (lse_row[a] - s[a][labels[a]]);  loss1 = mean_b (lse_col[b] - s[labels[b]][b])
__global__ void __launch_bounds__(256) pair_ce_loss_kernel(const float* __restrict__ s, const float* __restrict__ lse,
                                                           const int64_t* __restrict__ labels, int B,
                                                           float* __restrict__ loss01) {
    __shared__ float red[32];
    pdl_trigger();
    pdl_wait();
    float l0 = 0.f, l1 = 0.f;
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
        const long long lab64 = labels[k];
        if (lab64 < 0 || lab64 >= B) {  // torch's CrossEntropyLoss raises here; a kernel cannot, so the loss is poisoned instead
            l0 = l1 = NAN;
            continue;
        }
        const int lab = (int)lab64;
        l0 += lse[k] - s[(size_t)k * B + lab];
        l1 += lse[B + k] - s[(size_t)lab * B + k];
    }
    l0 = block_sum(l0, red);
    l1 = block_sum(l1, red);
    if (threadIdx.x == 0) {
        loss01[0] = l0 / (float)B;
        loss01[1] = l1 / (float)B;
    }
}

__global__ void __launch_bounds__(256) pair_ce_bwd_kernel(const float* __restrict__ s, const float* __restrict__ lse,
                                                          const int64_t* __restrict__ labels,
                                                          const float* __restrict__ g, float scale, int B,
                                                          float* __restrict__ ds) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= (size_t)B * B) return;
    const int a = (int)(idx / B), b = (int)(idx - (size_t)a * B);
    const float v = s[idx];
    float out = 0.f;
    if (v != -INFINITY) {
        const float g0 = g[0], g1 = g[1];
        const float pr = expf(v - lse[a]) - ((int)labels[a] == b ? 1.f : 0.f);
        const float pc = expf(v - lse[B + b]) - ((int)labels[b] == a ? 1.f : 0.f);
        out = scale * (g0 * pr + g1 * pc) / (float)B;
    }
    ds[idx] = out;
}

// ---------------------------------------------------------------------------------------
// sentence scores
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int D, int lane) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(a[d], b[d], s);
    return warp_sum(s);
}

// one CTA per image row i; warps loop over sentences j
__global__ void __launch_bounds__(256) sent_scores_fwd_kernel(const float* __restrict__ cnn, const float* __restrict__ rnn,
                                                              int B, int D, float g3, float eps,
                                                              float* __restrict__ scores, float* __restrict__ norms) {
    extern __shared__ float ci[];  // [D]
    __shared__ float red[32];
    const int i = blockIdx.x;
    float ss = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float v = cnn[(size_t)i * D + d];
        ci[d] = v;
        ss = fmaf(v, v, ss);
    }
    ss = block_sum(ss, red);
    const float nc = sqrtf(ss);
    if (threadIdx.x == 0) norms[i] = nc;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = w; j < B; j += nw) {
        const float* rj = rnn + (size_t)j * D;
        const float dot = warp_dot(ci, rj, D, lane);
        const float rr = warp_dot(rj, rj, D, lane);
        if (lane == 0) {
            const float nr = sqrtf(rr);
            scores[(size_t)i * B + j] = dot / fmaxf(nc * nr, eps) * g3;
            if (i == 0) norms[B + j] = nr;
        }
    }
}

// blocks [0,B): d_cnn row i;  blocks [B,2B): d_rnn row j.   smem: self[D], coef[B]
__global__ void __launch_bounds__(256) sent_scores_bwd_kernel(const float* __restrict__ cnn, const float* __restrict__ rnn,
                                                              const float* __restrict__ norms, const float* __restrict__ G,
                                                              int B, int D, float g3, float eps,
                                                              float* __restrict__ d_cnn, float* __restrict__ d_rnn) {
    extern __shared__ float sm[];
    float* self = sm;       // [D]
    float* coef = sm + D;   // [B]  d_dot for each partner
    __shared__ float red[32];
    __shared__ float selfc;
    const bool is_cnn = blockIdx.x < B;
    const int f = is_cnn ? blockIdx.x : blockIdx.x - B;
    const float* mine = (is_cnn ? cnn : rnn) + (size_t)f * D;
    const float* other = is_cnn ? rnn : cnn;
    const float nf = norms[(is_cnn ? 0 : B) + f];
    for (int d = threadIdx.x; d < D; d += blockDim.x) self[d] = mine[d];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float part = 0.f;  // lane-0 partial of sum_k d_norm_self[k]
    for (int k = w; k < B; k += nw) {
        const float dot = warp_dot(self, other + (size_t)k * D, D, lane);
        if (lane == 0) {
            const float no = norms[(is_cnn ? B : 0) + k];
            const float nn = nf * no;
            const float den = fmaxf(nn, eps);
            const float gv = G[is_cnn ? (size_t)f * B + k : (size_t)k * B + f] * g3;
            coef[k] = gv / den;
            if (nn > eps) part += -gv * dot / (den * den) * no;  // d(den)/d|self| = |other|
        }
    }
    part = block_sum(part, red);
    if (threadIdx.x == 0) selfc = (nf > 0.f) ? part / nf : 0.f;
    __syncthreads();
    float* dst = (is_cnn ? d_cnn : d_rnn) + (size_t)f * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = selfc * self[d];
        for (int k = 0; k < B; ++k) acc = fmaf(coef[k], other[(size_t)k * D + d], acc);
        dst[d] = acc;
    }
}

}  // namespace eegan

using namespace eegan;

extern "C" int eegan_pair_ce_fwd(const float* scores_in, float scale, const int64_t* class_ids, const int64_t* labels,
                                 int B, float* scores_out, float* loss01, float* lse, void* stream) {
    EEGAN_REQUIRE(B > 0, "pair_ce: B=%d", B);
    EEGAN_REQUIRE(scores_in && scores_out && lse, "pair_ce fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    launch_pdl(pair_ce_lse_kernel, dim3(2 * B), dim3(256), 0, st, scores_in, scale, class_ids, B, scores_out, lse);
    if (labels && loss01) launch_pdl(pair_ce_loss_kernel, dim3(1), dim3(256), 0, st, (const float*)scores_out, (const float*)lse, labels, B, loss01);
    EEGAN_LAUNCH_CHECK("pair_ce fwd");
    return EEGAN_OK;
}

extern "C" int eegan_pair_ce_bwd(const float* scores_out, const float* lse, const int64_t* labels,
                                 const float* g_loss01, float scale, int B, float* dscores_in, void* stream) {
    EEGAN_REQUIRE(B > 0, "pair_ce: B=%d", B);
    EEGAN_REQUIRE(scores_out && lse && labels && g_loss01 && dscores_in, "pair_ce bwd: null pointer");
    const size_t n = (size_t)B * B;
    launch_pdl(pair_ce_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, scores_out, lse, labels,
               g_loss01, scale, B, dscores_in);
    EEGAN_LAUNCH_CHECK("pair_ce bwd");
    return EEGAN_OK;
}

extern "C" int eegan_sent_scores_fwd(const float* cnn, const float* rnn, int B, int D, float gamma3, float eps,
                                     float* scores, float* norms, void* stream) {
    EEGAN_REQUIRE(B > 0 && D > 0 && D <= 8192, "sent_scores: B=%d D=%d", B, D);
    EEGAN_REQUIRE(cnn && rnn && scores && norms, "sent_scores fwd: null pointer");
    sent_scores_fwd_kernel<<<B, 256, D * sizeof(float), (cudaStream_t)stream>>>(cnn, rnn, B, D, gamma3, eps, scores, norms);
    EEGAN_LAUNCH_CHECK("sent_scores fwd");
    return EEGAN_OK;
}

extern "C" int eegan_sent_scores_bwd(const float* cnn, const float* rnn, const float* norms, const float* dscores,
                                     int B, int D, float gamma3, float eps, float* d_cnn, float* d_rnn, void* stream) {
    EEGAN_REQUIRE(B > 0 && D > 0 && (size_t)(B + D) * 4 <= 48 * 1024, "sent_scores bwd: B=%d D=%d too large", B, D);
    EEGAN_REQUIRE(cnn && rnn && norms && dscores && d_cnn && d_rnn, "sent_scores bwd: null pointer");
    sent_scores_bwd_kernel<<<2 * B, 256, (size_t)(B + D) * sizeof(float), (cudaStream_t)stream>>>(
        cnn, rnn, norms, dscores, B, D, gamma3, eps, d_cnn, d_rnn);
    EEGAN_LAUNCH_CHECK("sent_scores bwd");
    return EEGAN_OK;
}
