// R-precision scoring (test.py:306-336, Tester.cal_sim_one_by_one): for every generated image b,
// the cosine of its global code against R_val sentence codes — candidate 0 is the ground-truth
// caption, the others are mismatched captions — and whether candidate 0 wins the argmax:
//   scores0[b][k] = <cnn_b, rnn_bk> / max(|cnn_b| |rnn_bk|, 1e-8)          (test.py:323-327)
//   hit[b] = argmax_k scores0[b][k] == 0                                    (test.py:329)
// The reference walks the batch one sample at a time (a 1 x 100 torch.mm, two norms and a clamp
// per sample, with a host sync on the argmax); here the whole batch is one launch: CTA = image,
// warp = candidate, lanes over the embedding.  Ties resolve to the lowest index (torch.argmax).
#include "common.cuh"

namespace eegan {

__global__ void __launch_bounds__(256) rprecision_kernel(const float* __restrict__ cnn, const float* __restrict__ rnn, int Rv,
                                                         int D, float eps, float* __restrict__ scores,
                                                         int32_t* __restrict__ best, uint8_t* __restrict__ hit) {
    extern __shared__ float s_sc[];  // [Rv]
    __shared__ float s_cn;
    __shared__ float red[32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* c = cnn + (size_t)b * D;
    float cc = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) cc = fmaf(c[d], c[d], cc);
    cc = block_sum(cc, red);
    if (threadIdx.x == 0) s_cn = sqrtf(cc);
    __syncthreads();
    const float cn = s_cn;
    for (int k = w; k < Rv; k += nw) {
        const float* r = rnn + ((size_t)b * Rv + k) * D;
        float dot = 0.f, rr = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float rv = r[d];
            dot = fmaf(c[d], rv, dot);
            rr = fmaf(rv, rv, rr);
        }
        dot = warp_sum(dot);
        rr = warp_sum(rr);
        if (lane == 0) {
            const float s = dot / fmaxf(cn * sqrtf(rr), eps);
            s_sc[k] = s;
            if (scores) scores[(size_t)b * Rv + k] = s;
        }
    }
    __syncthreads();
    if (w == 0) {  // argmax, lowest index on ties
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int k = lane; k < Rv; k += 32) {
            const float v = s_sc[k];
            if (v > bv || (v == bv && k < bi)) { bv = v; bi = k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            if (bi == 0x7fffffff) bi = 0;  // all-NaN row: torch.argmax returns an index too; report 0
            if (best) best[b] = bi;
            if (hit) hit[b] = bi == 0 ? 1 : 0;
        }
    }
}

}  // namespace eegan

using namespace eegan;

extern "C" int eegan_rprecision(const float* cnn_code, const float* rnn_codes, int B, int R_val, int D, float eps,
                                float* scores, int32_t* best, uint8_t* hit, void* stream) {
    EEGAN_REQUIRE(cnn_code && rnn_codes && (scores || best || hit), "rprecision: null pointer");
    EEGAN_REQUIRE(B > 0 && R_val > 0 && D > 0, "rprecision: empty problem (B=%d R_val=%d D=%d)", B, R_val, D);
    EEGAN_REQUIRE(R_val <= 8192, "rprecision: R_val=%d exceeds 8192 candidates", R_val);
    rprecision_kernel<<<B, 256, R_val * sizeof(float), (cudaStream_t)stream>>>(cnn_code, rnn_codes, R_val, D, eps, scores, best, hit);
    EEGAN_LAUNCH_CHECK("rprecision");
    return EEGAN_OK;
}
