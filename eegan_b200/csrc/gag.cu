// GlobalAttentionGeneral (miscc/DAMSM_losses.py:65-132): per pixel q of sample b
//   s[t] = sum_d x[d,q] key[d,t];  masked rows -> -inf;  p = softmax_t(s);
//   out[d,q] = sum_t value[d,t] p[t];  attn[t,q] = p[t]
// HBM-bound streaming kernels: one thread owns PX pixels (coalesced along q for every d),
// key/value (and the mask) live in shared memory and are read as broadcast float4.
// Algorithmic bytes per pixel fwd+bwd: (5*idf + 2*T)*4.
#include "common.cuh"

namespace eegan {

constexpr int GAG_PX = 2;       // pixels per thread
constexpr int GAG_THREADS = 256;

// TP = T padded to a multiple of 4 (template => fully unrolled register arrays)
template <int TP>
__global__ void __launch_bounds__(GAG_THREADS) gag_fwd_kernel(const float* __restrict__ x, const float* __restrict__ key,
                                                              const float* __restrict__ value,
                                                              const uint8_t* __restrict__ mask, int mask_mode, int B,
                                                              int idf, int Q, int T, float* __restrict__ out,
                                                              float* __restrict__ attn) {
    extern __shared__ __align__(16) float sm[];
    float* ks = sm;               // [idf][TP]
    float* vs = sm + idf * TP;    // [idf][TP]
    const int b = blockIdx.y;
    for (int idx = threadIdx.x; idx < idf * TP; idx += blockDim.x) {
        const int d = idx / TP, t = idx - d * TP;
        ks[idx] = (t < T) ? key[((size_t)b * idf + d) * T + t] : 0.f;
        vs[idx] = (t < T) ? value[((size_t)b * idf + d) * T + t] : 0.f;
    }
    __syncthreads();
    const int q0 = (blockIdx.x * GAG_THREADS) * GAG_PX + threadIdx.x;
    const float* xb = x + (size_t)b * idf * Q;
    float s[GAG_PX][TP];
#pragma unroll
    for (int u = 0; u < GAG_PX; ++u)
#pragma unroll
        for (int t = 0; t < TP; ++t) s[u][t] = 0.f;
    bool ok[GAG_PX];
#pragma unroll
    for (int u = 0; u < GAG_PX; ++u) ok[u] = q0 + u * GAG_THREADS < Q;

#pragma unroll 4
    for (int d = 0; d < idf; ++d) {
        float xv[GAG_PX];
#pragma unroll
        for (int u = 0; u < GAG_PX; ++u) xv[u] = ok[u] ? __ldg(xb + (size_t)d * Q + q0 + u * GAG_THREADS) : 0.f;
#pragma unroll
        for (int t = 0; t < TP; t += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(ks + d * TP + t);
#pragma unroll
            for (int u = 0; u < GAG_PX; ++u) {
                s[u][t + 0] = fmaf(xv[u], k4.x, s[u][t + 0]);
                s[u][t + 1] = fmaf(xv[u], k4.y, s[u][t + 1]);
                s[u][t + 2] = fmaf(xv[u], k4.z, s[u][t + 2]);
                s[u][t + 3] = fmaf(xv[u], k4.w, s[u][t + 3]);
            }
        }
    }
    // masked softmax over words (:114-119)
#pragma unroll
    for (int u = 0; u < GAG_PX; ++u) {
        const int q = q0 + u * GAG_THREADS;
        if (!ok[u]) continue;
        const uint8_t* mrow = nullptr;
        if (mask) {
            const long long row = mask_mode == 0 ? ((long long)b * Q + q) % B : b;
            mrow = mask + row * T;
        }
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const bool dead = (t >= T) || (mrow && mrow[t < T ? t : 0]);
            s[u][t] = dead ? -INFINITY : s[u][t];
            mx = fmaxf(mx, s[u][t]);
        }
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            // a fully masked row gives exp(-inf - -inf) = NaN exactly like the reference's softmax
            const float e = (t < T) ? expf(s[u][t] - mx) : 0.f;
            s[u][t] = e;
            sum += e;
        }
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            s[u][t] = s[u][t] / sum;
            if (t < T) attn[((size_t)b * T + t) * Q + q] = s[u][t];
        }
    }
    // out[d,q] = sum_t value[d,t] p[t]  (:127)
    float* ob = out + (size_t)b * idf * Q;
#pragma unroll 2
    for (int d = 0; d < idf; ++d) {
        float acc[GAG_PX];
#pragma unroll
        for (int u = 0; u < GAG_PX; ++u) acc[u] = 0.f;
#pragma unroll
        for (int t = 0; t < TP; t += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vs + d * TP + t);
#pragma unroll
            for (int u = 0; u < GAG_PX; ++u) {
                acc[u] = fmaf(v4.x, s[u][t + 0], acc[u]);
                acc[u] = fmaf(v4.y, s[u][t + 1], acc[u]);
                acc[u] = fmaf(v4.z, s[u][t + 2], acc[u]);
                acc[u] = fmaf(v4.w, s[u][t + 3], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < GAG_PX; ++u)
            if (ok[u]) ob[(size_t)d * Q + q0 + u * GAG_THREADS] = acc[u];
    }
}

// Backward.  One CTA walks pixel tiles of 256 pixels of sample b (grid.x CTAs share a sample):
//   pass A (thread = pixel): dp[t] = sum_d dout[d,q] v[d,t] + dattn[t,q];  ds = p (dp - sum p dp)
//   pass B, per chunk of DC channels: dx[d,q] = sum_t ds[t] key[d,t]  (thread = pixel), then
//           dkey[d,t] += sum_q x[d,q] ds[q,t], dvalue[d,t] += sum_q dout[d,q] p[q,t]
//           (thread = (row, pixel quarter) over the x / dout chunk staged in shared memory),
//   accumulated per CTA in shared memory and flushed once with atomics.
constexpr int GAG_BT = 256;  // pixels per tile == threads
constexpr int GAG_DC = 32;   // channels per chunk

template <int TP>
__global__ void __launch_bounds__(GAG_BT) gag_bwd_kernel(const float* __restrict__ x, const float* __restrict__ key,
                                                         const float* __restrict__ value, const float* __restrict__ attn,
                                                         const float* __restrict__ d_out, const float* __restrict__ d_attn,
                                                         int idf, int Q, int T, float* __restrict__ d_x,
                                                         float* __restrict__ d_key, float* __restrict__ d_value) {
    extern __shared__ __align__(16) float sm[];
    float* ks = sm;                               // [idf][TP]
    float* vs = ks + idf * TP;                    // [idf][TP]
    float* dks = vs + idf * TP;                   // [idf][TP] accumulators
    float* dvs = dks + idf * TP;                  // [idf][TP]
    float* ds_s = dvs + idf * TP;                 // [BT][TP]
    float* p_s = ds_s + GAG_BT * TP;              // [BT][TP]
    float* xs = p_s + GAG_BT * TP;                // [DC][BT+1]
    float* gs = xs + GAG_DC * (GAG_BT + 1);       // [DC][BT+1]  d_out chunk
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < idf * TP; idx += GAG_BT) {
        const int d = idx / TP, t = idx - d * TP;
        ks[idx] = (t < T) ? key[((size_t)b * idf + d) * T + t] : 0.f;
        vs[idx] = (t < T) ? value[((size_t)b * idf + d) * T + t] : 0.f;
        dks[idx] = 0.f;
        dvs[idx] = 0.f;
    }
    __syncthreads();
    const float* xb = x + (size_t)b * idf * Q;
    const float* gb = d_out ? d_out + (size_t)b * idf * Q : nullptr;
    float* dxb = d_x + (size_t)b * idf * Q;
    const int ntiles = (Q + GAG_BT - 1) / GAG_BT;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int q = tile * GAG_BT + tid;
        const bool ok = q < Q;
        float dp[TP], p[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const bool on = ok && t < T;
            p[t] = on ? attn[((size_t)b * T + t) * Q + q] : 0.f;
            dp[t] = (on && d_attn) ? d_attn[((size_t)b * T + t) * Q + q] : 0.f;
        }
        if (gb) {
#pragma unroll 4
            for (int d = 0; d < idf; ++d) {
                const float gv = ok ? __ldg(gb + (size_t)d * Q + q) : 0.f;
#pragma unroll
                for (int t = 0; t < TP; t += 4) {
                    const float4 v4 = *reinterpret_cast<const float4*>(vs + d * TP + t);
                    dp[t + 0] = fmaf(gv, v4.x, dp[t + 0]);
                    dp[t + 1] = fmaf(gv, v4.y, dp[t + 1]);
                    dp[t + 2] = fmaf(gv, v4.z, dp[t + 2]);
                    dp[t + 3] = fmaf(gv, v4.w, dp[t + 3]);
                }
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t < TP; ++t) dot = fmaf(p[t], dp[t], dot);
        float ds[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            ds[t] = p[t] * (dp[t] - dot);
            ds_s[tid * TP + t] = ds[t];
            p_s[tid * TP + t] = p[t];
        }
        for (int d0 = 0; d0 < idf; d0 += GAG_DC) {
            __syncthreads();  // previous chunk's readers are done (and ds_s/p_s are visible)
#pragma unroll 4
            for (int dd = 0; dd < GAG_DC; ++dd) {
                const int d = d0 + dd;
                if (d >= idf) break;
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < TP; t += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(ks + d * TP + t);
                    acc = fmaf(ds[t + 0], k4.x, acc);
                    acc = fmaf(ds[t + 1], k4.y, acc);
                    acc = fmaf(ds[t + 2], k4.z, acc);
                    acc = fmaf(ds[t + 3], k4.w, acc);
                }
                if (ok) dxb[(size_t)d * Q + q] = acc;
                xs[dd * (GAG_BT + 1) + tid] = ok ? __ldg(xb + (size_t)d * Q + q) : 0.f;
                gs[dd * (GAG_BT + 1) + tid] = (ok && gb) ? __ldg(gb + (size_t)d * Q + q) : 0.f;
            }
            __syncthreads();
            // 2*DC rows (dkey rows then dvalue rows) x 4 pixel quarters = 256 threads
            const int row = tid % (2 * GAG_DC), quarter = tid / (2 * GAG_DC);
            const bool is_k = row < GAG_DC;
            const int dd = is_k ? row : row - GAG_DC;
            if (d0 + dd < idf) {
                const float* src = (is_k ? xs : gs) + dd * (GAG_BT + 1);
                const float* rhs = is_k ? ds_s : p_s;
                float acc[TP];
#pragma unroll
                for (int t = 0; t < TP; ++t) acc[t] = 0.f;
                const int qa = quarter * (GAG_BT / 4);
#pragma unroll 4
                for (int qq = qa; qq < qa + GAG_BT / 4; ++qq) {
                    const float sv = src[qq];
#pragma unroll
                    for (int t = 0; t < TP; t += 4) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rhs + qq * TP + t);
                        acc[t + 0] = fmaf(sv, r4.x, acc[t + 0]);
                        acc[t + 1] = fmaf(sv, r4.y, acc[t + 1]);
                        acc[t + 2] = fmaf(sv, r4.z, acc[t + 2]);
                        acc[t + 3] = fmaf(sv, r4.w, acc[t + 3]);
                    }
                }
                float* dst = (is_k ? dks : dvs) + (d0 + dd) * TP;
#pragma unroll
                for (int t = 0; t < TP; ++t) atomicAdd(dst + t, acc[t]);
            }
        }
        __syncthreads();  // ds_s / p_s are rewritten by the next tile
    }
    __syncthreads();
    for (int idx = tid; idx < idf * TP; idx += GAG_BT) {
        const int d = idx / TP, t = idx - d * TP;
        if (t < T) {
            atomicAdd(d_key + ((size_t)b * idf + d) * T + t, dks[idx]);
            atomicAdd(d_value + ((size_t)b * idf + d) * T + t, dvs[idx]);
        }
    }
}

template <int TP>
static int gag_fwd_launch(const float* x, const float* key, const float* value, const uint8_t* mask, int mask_mode,
                          int B, int idf, int Q, int T, float* out, float* attn, cudaStream_t st) {
    const size_t smem = (size_t)2 * idf * TP * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gag_fwd_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gag fwd smem: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
    }
    dim3 grid((Q + GAG_THREADS * GAG_PX - 1) / (GAG_THREADS * GAG_PX), B);
    gag_fwd_kernel<TP><<<grid, GAG_THREADS, smem, st>>>(x, key, value, mask, mask_mode, B, idf, Q, T, out, attn);
    return check_launch("gag fwd");
}

template <int TP>
static int gag_bwd_launch(const float* x, const float* key, const float* value, const float* attn, const float* d_out,
                          const float* d_attn, int B, int idf, int Q, int T, float* d_x, float* d_key, float* d_value,
                          cudaStream_t st) {
    const size_t smem = ((size_t)4 * idf * TP + 2 * GAG_BT * TP + 2 * GAG_DC * (GAG_BT + 1)) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gag_bwd_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gag bwd smem: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
    }
    cudaMemsetAsync(d_key, 0, (size_t)B * idf * T * sizeof(float), st);
    cudaMemsetAsync(d_value, 0, (size_t)B * idf * T * sizeof(float), st);
    const int ntiles = (Q + GAG_BT - 1) / GAG_BT;
    int per_sample = (2 * 148 + B - 1) / B;  // ~2 CTAs per SM over the whole grid
    if (per_sample > ntiles) per_sample = ntiles;
    if (per_sample < 1) per_sample = 1;
    gag_bwd_kernel<TP><<<dim3(per_sample, B), GAG_BT, smem, st>>>(x, key, value, attn, d_out, d_attn, idf, Q, T, d_x,
                                                                  d_key, d_value);
    return check_launch("gag bwd");
}

}  // namespace eegan

using namespace eegan;

#define GAG_DISPATCH(T, CALL)                                                   \
    ((T) <= 8 ? CALL<8> : (T) <= 12 ? CALL<12> : (T) <= 16 ? CALL<16> : (T) <= 20 ? CALL<20> : (T) <= 24 ? CALL<24> : CALL<32>)

extern "C" int eegan_gag_fwd(const float* x, const float* key, const float* value, const uint8_t* mask, int mask_mode,
                             int B, int idf, int Q, int T, float* out, float* attn, void* stream) {
    EEGAN_REQUIRE(B > 0 && idf > 0 && Q > 0 && T > 0, "gag: empty shape B=%d idf=%d Q=%d T=%d", B, idf, Q, T);
    EEGAN_REQUIRE(T <= 32 && idf <= 512, "gag: T=%d (<=32) idf=%d (<=512) unsupported", T, idf);
    EEGAN_REQUIRE(x && key && value && out && attn, "gag fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    return GAG_DISPATCH(T, gag_fwd_launch)(x, key, value, mask, mask_mode, B, idf, Q, T, out, attn, st);
}

extern "C" int eegan_gag_bwd(const float* x, const float* key, const float* value, const float* attn,
                             const float* d_out, const float* d_attn, int B, int idf, int Q, int T, float* d_x,
                             float* d_key, float* d_value, void* stream) {
    EEGAN_REQUIRE(B > 0 && idf > 0 && Q > 0 && T > 0, "gag: empty shape B=%d idf=%d Q=%d T=%d", B, idf, Q, T);
    EEGAN_REQUIRE(T <= 32 && idf <= 256, "gag bwd: T=%d (<=32) idf=%d (<=256) unsupported", T, idf);
    EEGAN_REQUIRE(x && key && value && attn && d_x && d_key && d_value, "gag bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    return GAG_DISPATCH(T, gag_bwd_launch)(x, key, value, attn, d_out, d_attn, B, idf, Q, T, d_x, d_key, d_value, st);
}
