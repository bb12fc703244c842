// GlobalAttentionGeneral (miscc/DAMSM_losses.py:65-132): per pixel q of sample b
//   s[t] = sum_d x[d,q] key[d,t];  masked rows -> -inf;  p = softmax_t(s);
//   out[d,q] = sum_t value[d,t] p[t];  attn[t,q] = p[t]
// HBM-bound streaming kernels: one thread owns PX pixels (coalesced along q for every d),
// key/value (and the mask) live in shared memory and are read as broadcast float4.
// Algorithmic bytes per pixel fwd+bwd: (5*idf + 2*T)*4.
#include <atomic>

#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

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

// ---------------------------------------------------------------------------------------
// TMA-fed forward: the x tile is streamed by the TMA engine in [8 channels][512 pixels] chunks
// (two 256-pixel boxes) through a 3-stage mbarrier ring, so the compute threads never wait on a
// global load; key / value stay in shared memory; a thread owns four consecutive pixels and both
// contractions run on packed fma.rn.f32x2.  Needs Q % 4 == 0, idf % 8 == 0, 16-byte aligned x / out / attn.
// ---------------------------------------------------------------------------------------
constexpr int GF_TILE = 512, GF_DC = 8, GF_NS = 3;
constexpr int GF_NT = 128, GF_PX = GF_TILE / GF_NT;  // a thread owns 4 pixels: every key / value float4 (a broadcast LDS.128 holds the
                                                     // shared-memory pipe for four cycles) feeds 16 FMA instead of 8
constexpr int GF_STAGE_BYTES = GF_DC * GF_TILE * 4;  // 16 KB

template <int TP>
__global__ void __launch_bounds__(GF_NT, 3) gag_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ key,
                                                             const float* __restrict__ value, const uint8_t* __restrict__ mask,
                                                             int mask_mode, int B, int idf, int Q, int T, float* __restrict__ out,
                                                             float* __restrict__ attn) {
    extern __shared__ uint8_t gsm[];
    const uint32_t base = (smem_u32(gsm) + 127u) & ~127u;
    float* ks = reinterpret_cast<float*>(gsm + (base - smem_u32(gsm)) + GF_NS * GF_STAGE_BYTES);  // [idf][TP]
    float* vs = ks + idf * TP;
    __shared__ __align__(8) unsigned long long bars[2 * GF_NS];
    __shared__ uint32_t mbits[1024];  // per mask row: bit t set = word t is padding
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (GF_NS + s); };
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int q0 = blockIdx.x * GF_TILE;
    for (int idx = tid; idx < idf * TP; idx += GF_NT) {
        const int d = idx / TP, t = idx - d * TP;
        ks[idx] = (t < T) ? key[((size_t)b * idf + d) * T + t] : 0.f;
        vs[idx] = (t < T) ? value[((size_t)b * idf + d) * T + t] : 0.f;
    }
    if (mask)
        for (int row = tid; row < B; row += GF_NT) {
            uint32_t bits = 0;
            for (int t = 0; t < T; ++t) bits |= (mask[(size_t)row * T + t] ? 1u : 0u) << t;
            mbits[row] = bits;
        }
    if (tid == 0) {
        for (int s = 0; s < GF_NS; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), GF_NT / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int nchunks = idf / GF_DC;
    auto issue = [&](int c) {
        const int s = c % GF_NS;
        mbar_arrive_expect_tx(full(s), GF_STAGE_BYTES);
        tma_load_2d(base + s * GF_STAGE_BYTES, &tmx, full(s), q0, b * idf + c * GF_DC);
        tma_load_2d(base + s * GF_STAGE_BYTES + GF_STAGE_BYTES / 2, &tmx, full(s), q0 + 256, b * idf + c * GF_DC);
    };
    if (tid == 0)
        for (int c = 0; c < GF_NS - 1 && c < nchunks; ++c) issue(c);

    // scores as PAIRS of adjacent words: every multiply-add of the two contractions is a packed fma.rn.f32x2 (FFMA2) — a
    // three-register FFMA issues every other cycle per scheduler on sm_100, the packed form does two per issue
    // (each half is an exact fma, so the scores are bit-identical to the scalar form)
    float2 s2[GF_PX][TP / 2];
#define GF_S(u, t) (((t) & 1) ? s2[u][(t) >> 1].y : s2[u][(t) >> 1].x)
#pragma unroll
    for (int u = 0; u < GF_PX; ++u)
#pragma unroll
        for (int t = 0; t < TP / 2; ++t) s2[u][t] = make_float2(0.f, 0.f);

    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0) {
            const int cc = c + GF_NS - 1;
            if (cc < nchunks) {
                if (cc >= GF_NS) mbar_wait(empty(cc % GF_NS), ((cc / GF_NS) - 1) & 1);
                issue(cc);
            }
        }
        const int st = c % GF_NS;
        mbar_wait(full(st), (c / GF_NS) & 1);
        const float* xs = reinterpret_cast<const float*>(gsm + (base - smem_u32(gsm)) + st * GF_STAGE_BYTES);
#pragma unroll 4
        for (int dd = 0; dd < GF_DC; ++dd) {
            // the thread's quad = pixels 4 tid .. 4 tid + 3 of the tile: one 16-byte read (the stage holds two 256-pixel TMA boxes)
            const float4 x4 = *reinterpret_cast<const float4*>(xs + (tid >> 6) * GF_DC * 256 + dd * 256 + 4 * (tid & 63));
            const float2 xd[GF_PX] = {make_float2(x4.x, x4.x), make_float2(x4.y, x4.y), make_float2(x4.z, x4.z), make_float2(x4.w, x4.w)};
            const float* kr = ks + (c * GF_DC + dd) * TP;
#pragma unroll
            for (int t = 0; t < TP; t += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kr + t);
                const float2 k01 = make_float2(k4.x, k4.y), k23 = make_float2(k4.z, k4.w);
#pragma unroll
                for (int u = 0; u < GF_PX; ++u) {
                    s2[u][t / 2] = __ffma2_rn(xd[u], k01, s2[u][t / 2]);
                    s2[u][t / 2 + 1] = __ffma2_rn(xd[u], k23, s2[u][t / 2 + 1]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty(st));
    }
    static_assert(GF_PX == 4, "the forward is written for pixel quads");
    const int qb = q0 + GF_PX * tid;  // Q % 4 == 0: a quad is entirely inside the row or entirely outside
    const bool ok = qb < Q;
    const uint32_t tail = (T < 32) ? (0xffffffffu << T) : 0u;  // columns t >= T never exist
    const uint32_t rowbase = (uint32_t)(((long long)b * Q) % B);
    if (ok) {
#pragma unroll
        for (int u = 0; u < GF_PX; ++u) {
            const int q = qb + u;
            uint32_t dead = tail;
            if (mask) dead |= mbits[mask_mode == 0 ? (rowbase + (uint32_t)q) % (uint32_t)B : (uint32_t)b];
            float mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < TP; ++t) {
                GF_S(u, t) = ((dead >> t) & 1u) ? -INFINITY : GF_S(u, t);
                mx = fmaxf(mx, GF_S(u, t));
            }
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < TP; ++t) {
                // a fully masked row gives exp(-inf - -inf) = NaN exactly like the reference's softmax
                const float e = (t < T) ? __expf(GF_S(u, t) - mx) : 0.f;
                GF_S(u, t) = e;
                sum += e;
            }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int t = 0; t < TP; ++t) GF_S(u, t) *= inv;
        }
#pragma unroll
        for (int t = 0; t < TP; ++t)
            if (t < T) *reinterpret_cast<float4*>(attn + ((size_t)b * T + t) * Q + qb) = make_float4(GF_S(0, t), GF_S(1, t), GF_S(2, t), GF_S(3, t));
    }
    float* ob = out + (size_t)b * idf * Q + qb;
#pragma unroll 2
    for (int d = 0; d < idf; ++d) {
        float2 acc[GF_PX];  // (even words, odd words) partial sums
#pragma unroll
        for (int u = 0; u < GF_PX; ++u) acc[u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < TP; t += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vs + d * TP + t);
            const float2 v01 = make_float2(v4.x, v4.y), v23 = make_float2(v4.z, v4.w);
#pragma unroll
            for (int u = 0; u < GF_PX; ++u) {
                acc[u] = __ffma2_rn(v01, s2[u][t / 2], acc[u]);
                acc[u] = __ffma2_rn(v23, s2[u][t / 2 + 1], acc[u]);
            }
        }
        if (ok) *reinterpret_cast<float4*>(ob + (size_t)d * Q) = make_float4(acc[0].x + acc[0].y, acc[1].x + acc[1].y, acc[2].x + acc[2].y, acc[3].x + acc[3].y);
    }
#undef GF_S
}

template <int TP>
static int gag_fwd_tma_launch(const float* x, const float* key, const float* value, const uint8_t* mask, int mask_mode,
                              int B, int idf, int Q, int T, float* out, float* attn, cudaStream_t st) {
    CUtensorMap tmx;
    int rc = make_tmap_2d(&tmx, x, (unsigned long long)B * idf, (unsigned long long)Q, (unsigned long long)Q, 256, GF_DC);
    if (rc) return rc;
    const size_t smem = (size_t)GF_NS * GF_STAGE_BYTES + (size_t)2 * idf * TP * sizeof(float) + 128;
    static SmemGrant grant;
    if (int rc2 = grant_dyn_smem(gag_fwd_tma_kernel<TP>, smem, grant, "gag fwd")) return rc2;
    gag_fwd_tma_kernel<TP><<<dim3((Q + GF_TILE - 1) / GF_TILE, B), GF_NT, smem, st>>>(tmx, key, value, mask, mask_mode, B, idf, Q,
                                                                                     T, out, attn);
    return check_launch("gag fwd (tma)");
}

// ---------------------------------------------------------------------------------------
// TMA-fed backward.  CTA = 256 threads walking 256-pixel tiles of sample b.  Per tile:
//   pass A  stream d_out in [8 ch][256 px] chunks: dp[t] += d_out[d,q] v[d,t]; then
//           ds = p (dp + d_attn - sum_t p (dp + d_attn)), kept in registers and in smem [px][TP]
//   pass B  stream x and d_out chunks again (L2 hits): dx[d,q] = sum_t ds[t] key[d,t] (thread = pixel)
//           and dkey[d,:] += sum_q x[d,q] ds[q,:], dvalue[d,:] += sum_q d_out[d,q] p[q,:]
//           with lanes = the 16 chunk rows (8 x + 8 d_out) and a broadcast read of ds / p rows.
// Chunks arrive as 128B-swizzled TMA boxes of [8 rows][32 px] so that BOTH access patterns
// (lanes = consecutive pixels, lanes = different rows of one pixel) are (nearly) conflict-free.
// ---------------------------------------------------------------------------------------
constexpr int GB_TILE = 256, GB_DC = 8, GB_NS = 3;
constexpr int GB_BOX_BYTES = GB_DC * 128;                  // one [8][32 px] box = 1 KB
constexpr int GB_OPER_BYTES = (GB_TILE / 32) * GB_BOX_BYTES;  // 8 KB per operand per chunk
constexpr int GB_STAGE_BYTES = 2 * GB_OPER_BYTES;          // x then d_out

__device__ __forceinline__ uint32_t gb_off(int row, int px) {  // byte offset of (row, px) inside an operand block
    return (uint32_t)((px >> 5) * GB_BOX_BYTES + row * 128 + ((((px & 31) >> 2) ^ (row & 7)) << 4) + ((px & 3) << 2));
}

template <int TP>
__global__ void __launch_bounds__(256, 2) gag_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                             const __grid_constant__ CUtensorMap tmg,
                                                             const float* __restrict__ key, const float* __restrict__ value,
                                                             const float* __restrict__ attn, const float* __restrict__ d_attn,
                                                             int idf, int Q, int T, float* __restrict__ d_x,
                                                             float* __restrict__ d_key, float* __restrict__ d_value) {
    extern __shared__ uint8_t gsm[];
    const uint32_t base = (smem_u32(gsm) + 1023u) & ~1023u;
    uint8_t* gbase = gsm + (base - smem_u32(gsm));
    float* ks = reinterpret_cast<float*>(gbase + GB_NS * GB_STAGE_BYTES);  // [idf][TP]
    float* vs = ks + idf * TP;
    float* dks = vs + idf * TP;   // accumulators
    float* dvs = dks + idf * TP;
    float* ds_s = dvs + idf * TP;  // [256][TP]
    float* p_s = ds_s + GB_TILE * TP;
    float* red_s = p_s + GB_TILE * TP;  // [2][8 warps][16 rows][TP]
    __shared__ __align__(8) unsigned long long bars[2 * GB_NS];
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (GB_NS + s); };
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < idf * TP; idx += 256) {
        const int d = idx / TP, t = idx - d * TP;
        ks[idx] = (t < T) ? key[((size_t)b * idf + d) * T + t] : 0.f;
        vs[idx] = (t < T) ? value[((size_t)b * idf + d) * T + t] : 0.f;
        dks[idx] = 0.f;
        dvs[idx] = 0.f;
    }
    if (tid == 0) {
        for (int s = 0; s < GB_NS; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 8);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int nch = idf / GB_DC;
    const int ntiles = (Q + GB_TILE - 1) / GB_TILE;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int per_tile = 2 * nch;             // pass A chunks then pass B chunks
    const int total = my_tiles * per_tile;    // items of this CTA, ring position = item index
    auto issue = [&](int it) {
        const int s = it % GB_NS;
        const int ti = it / per_tile, r = it - ti * per_tile;
        const int q0 = ((int)blockIdx.x + ti * (int)gridDim.x) * GB_TILE;
        const bool passB = r >= nch;
        const int row0 = b * idf + (passB ? r - nch : r) * GB_DC;
        const uint32_t dst = base + s * GB_STAGE_BYTES;
        mbar_arrive_expect_tx(full(s), passB ? GB_STAGE_BYTES : GB_OPER_BYTES);
#pragma unroll
        for (int j = 0; j < GB_TILE / 32; ++j) {
            tma_load_2d(dst + GB_OPER_BYTES + j * GB_BOX_BYTES, &tmg, full(s), q0 + 32 * j, row0);
            if (passB) tma_load_2d(dst + j * GB_BOX_BYTES, &tmx, full(s), q0 + 32 * j, row0);
        }
    };
    if (tid == 0)
        for (int it = 0; it < GB_NS - 1 && it < total; ++it) issue(it);

    // dk/dv phase mapping: lanes 0-7 = x rows, 8-15 = d_out rows; 16 pixel groups of 16 pixels
    const int crow = lane & 15, cgrp = (warp << 1) | (lane >> 4);
    const bool is_k = crow < GB_DC;
    const int cr = is_k ? crow : crow - GB_DC;

    int it = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int q = ((int)blockIdx.x + ti * (int)gridDim.x) * GB_TILE + tid;
        const bool ok = q < Q;
        float dp[TP], p[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const bool on = ok && t < T;
            p[t] = on ? attn[((size_t)b * T + t) * Q + q] : 0.f;
            dp[t] = (on && d_attn) ? d_attn[((size_t)b * T + t) * Q + q] : 0.f;
        }
        // ---------------- pass A ----------------
        for (int c = 0; c < nch; ++c, ++it) {
            if (tid == 0) {
                const int nx = it + GB_NS - 1;
                if (nx < total) {
                    if (nx >= GB_NS) mbar_wait(empty(nx % GB_NS), ((nx / GB_NS) - 1) & 1);
                    issue(nx);
                }
            }
            const int st = it % GB_NS;
            mbar_wait(full(st), (it / GB_NS) & 1);
            const uint8_t* gch = gbase + st * GB_STAGE_BYTES + GB_OPER_BYTES;
#pragma unroll
            for (int dd = 0; dd < GB_DC; ++dd) {
                const float gv = *reinterpret_cast<const float*>(gch + gb_off(dd, tid));
                const float* vr = vs + (c * GB_DC + dd) * TP;
#pragma unroll
                for (int t = 0; t < TP; t += 4) {
                    const float4 v4 = *reinterpret_cast<const float4*>(vr + t);
                    dp[t + 0] = fmaf(gv, v4.x, dp[t + 0]);
                    dp[t + 1] = fmaf(gv, v4.y, dp[t + 1]);
                    dp[t + 2] = fmaf(gv, v4.z, dp[t + 2]);
                    dp[t + 3] = fmaf(gv, v4.w, dp[t + 3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(st));
        }
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t < TP; ++t) dot = fmaf(p[t], dp[t], dot);
        float ds[TP];
        __syncthreads();  // the previous tile's dk/dv readers are done with ds_s / p_s
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            ds[t] = p[t] * (dp[t] - dot);
            ds_s[tid * TP + t] = ds[t];
            p_s[tid * TP + t] = p[t];
        }
        __syncthreads();
        // ---------------- pass B ----------------
        for (int c = 0; c < nch; ++c, ++it) {
            if (tid == 0) {
                const int nx = it + GB_NS - 1;
                if (nx < total) {
                    if (nx >= GB_NS) mbar_wait(empty(nx % GB_NS), ((nx / GB_NS) - 1) & 1);
                    issue(nx);
                }
            }
            // dx needs no streamed data: do it while the chunk is in flight
#pragma unroll
            for (int dd = 0; dd < GB_DC; ++dd) {
                const int d = c * GB_DC + dd;
                const float* kr = ks + d * TP;
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < TP; t += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(kr + t);
                    acc = fmaf(ds[t + 0], k4.x, acc);
                    acc = fmaf(ds[t + 1], k4.y, acc);
                    acc = fmaf(ds[t + 2], k4.z, acc);
                    acc = fmaf(ds[t + 3], k4.w, acc);
                }
                if (ok) d_x[((size_t)b * idf + d) * Q + q] = acc;
            }
            const int st = it % GB_NS;
            mbar_wait(full(st), (it / GB_NS) & 1);
            const uint8_t* src = gbase + st * GB_STAGE_BYTES + (is_k ? 0 : GB_OPER_BYTES);
            const float* rhs = is_k ? ds_s : p_s;
            float acc[TP];
#pragma unroll
            for (int t = 0; t < TP; ++t) acc[t] = 0.f;
#pragma unroll 4
            for (int k = 0; k < GB_TILE / 16; ++k) {
                const int px = cgrp * (GB_TILE / 16) + k;
                const float sv = *reinterpret_cast<const float*>(src + gb_off(cr, px));
                const float* rr = rhs + px * TP;
#pragma unroll
                for (int t = 0; t < TP; t += 4) {
                    const float4 r4 = *reinterpret_cast<const float4*>(rr + t);
                    acc[t + 0] = fmaf(sv, r4.x, acc[t + 0]);
                    acc[t + 1] = fmaf(sv, r4.y, acc[t + 1]);
                    acc[t + 2] = fmaf(sv, r4.z, acc[t + 2]);
                    acc[t + 3] = fmaf(sv, r4.w, acc[t + 3]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(st));
            // CTA-wide reduction over the 16 pixel groups without atomics: lane pairs by shuffle, the 8 warps through a
            // double-buffered scratch, then one owner thread per (row, t) adds into the per-CTA accumulators
#pragma unroll
            for (int t = 0; t < TP; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 16);
            float* red = red_s + (it & 1) * (8 * 16 * TP);
            if (lane < 16) {
#pragma unroll
                for (int t = 0; t < TP; t += 4)
                    *reinterpret_cast<float4*>(red + (warp * 16 + crow) * TP + t) = make_float4(acc[t], acc[t + 1], acc[t + 2], acc[t + 3]);
            }
            __syncthreads();
            for (int o = tid; o < 16 * TP; o += 256) {
                float sum = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) sum += red[w8 * 16 * TP + o];
                const int row = o / TP, t = o - row * TP;
                float* dst = (row < GB_DC ? dks : dvs) + (c * GB_DC + (row < GB_DC ? row : row - GB_DC)) * TP + t;
                *dst += sum;
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < idf * TP; idx += 256) {
        const int d = idx / TP, t = idx - d * TP;
        if (t < T) {
            atomicAdd(d_key + ((size_t)b * idf + d) * T + t, dks[idx]);
            atomicAdd(d_value + ((size_t)b * idf + d) * T + t, dvs[idx]);
        }
    }
}

template <int TP>
static int gag_bwd_tma_launch(const float* x, const float* key, const float* value, const float* attn, const float* d_out,
                              const float* d_attn, int B, int idf, int Q, int T, float* d_x, float* d_key, float* d_value,
                              cudaStream_t st) {
    CUtensorMap tmx, tmg;
    int rc = make_tmap_2d(&tmx, x, (unsigned long long)B * idf, (unsigned long long)Q, (unsigned long long)Q, 32, GB_DC, true);
    if (rc) return rc;
    rc = make_tmap_2d(&tmg, d_out, (unsigned long long)B * idf, (unsigned long long)Q, (unsigned long long)Q, 32, GB_DC, true);
    if (rc) return rc;
    const size_t smem = (size_t)GB_NS * GB_STAGE_BYTES + ((size_t)4 * idf * TP + 2 * GB_TILE * TP + 2 * 8 * 16 * TP) * sizeof(float) + 1024;
    static SmemGrant grant;
    if (int rc2 = grant_dyn_smem(gag_bwd_tma_kernel<TP>, smem, grant, "gag bwd")) return rc2;
    cudaMemsetAsync(d_key, 0, (size_t)B * idf * T * sizeof(float), st);
    cudaMemsetAsync(d_value, 0, (size_t)B * idf * T * sizeof(float), st);
    const int ntiles = (Q + GB_TILE - 1) / GB_TILE;
    int per_sample = (2 * 148 + B - 1) / B;
    if (per_sample > ntiles) per_sample = ntiles;
    if (per_sample < 1) per_sample = 1;
    gag_bwd_tma_kernel<TP><<<dim3(per_sample, B), 256, smem, st>>>(tmx, tmg, key, value, attn, d_attn, idf, Q, T, d_x, d_key,
                                                                   d_value);
    return check_launch("gag bwd (tma)");
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

namespace eegan {
bool gag_tc_fwd_supported(const float* x, int B, int idf, int Q, int T);
int gag_tc_fwd_launch(const float* x, const float* key, const float* value, const uint8_t* mask, int mask_mode, int B, int idf,
                      int Q, int T, float* out, float* attn, cudaStream_t st);
}  // namespace eegan

using namespace eegan;

// forward engine: 1 = tcgen05 kernel of gag_tc.cu where the shape allows (default: 54 / 101 / 207 us against 70 / 121 / 248 us
// for the CUDA-core kernels at 64^2 x 128 / 128^2 x 64 / 256^2 x 32, B = 48), 0 = CUDA-core kernels
static std::atomic<int> g_gag_engine{[] {
    const char* e = getenv("EEGAN_GAG_TC");
    return e ? atoi(e) : 1;
}()};
extern "C" int eegan_set_gag_engine(int engine) {
    EEGAN_REQUIRE(engine == 0 || engine == 1, "gag engine must be 0 (CUDA cores) or 1 (tensor-core forward)");
    g_gag_engine.store(engine);
    return EEGAN_OK;
}
extern "C" int eegan_get_gag_engine(void) { return g_gag_engine.load(); }

#define GAG_DISPATCH(T, CALL)                                                   \
    ((T) <= 8 ? CALL<8> : (T) <= 12 ? CALL<12> : (T) <= 16 ? CALL<16> : (T) <= 20 ? CALL<20> : (T) <= 24 ? CALL<24> : CALL<32>)

extern "C" int eegan_gag_fwd(const float* x, const float* key, const float* value, const uint8_t* mask, int mask_mode,
                             int B, int idf, int Q, int T, float* out, float* attn, void* stream) {
    EEGAN_REQUIRE(B > 0 && idf > 0 && Q > 0 && T > 0, "gag: empty shape B=%d idf=%d Q=%d T=%d", B, idf, Q, T);
    EEGAN_REQUIRE(T <= 32 && idf <= 512, "gag: T=%d (<=32) idf=%d (<=512) unsupported", T, idf);
    EEGAN_REQUIRE(x && key && value && out && attn, "gag fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (g_gag_engine.load() == 1 && gag_tc_fwd_supported(x, B, idf, Q, T))  // tensor-core kernel (gag_tc.cu)
        return gag_tc_fwd_launch(x, key, value, mask, mask_mode, B, idf, Q, T, out, attn, st);
    const size_t tma_smem = (size_t)GF_NS * GF_STAGE_BYTES + (size_t)2 * idf * 32 * sizeof(float) + 128;
    if (Q % 4 == 0 && idf % GF_DC == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(attn)) & 15) == 0 &&
        tma_smem <= 100 * 1024 && B <= 1024)
        return GAG_DISPATCH(T, gag_fwd_tma_launch)(x, key, value, mask, mask_mode, B, idf, Q, T, out, attn, st);
    return GAG_DISPATCH(T, gag_fwd_launch)(x, key, value, mask, mask_mode, B, idf, Q, T, out, attn, st);
}

extern "C" int eegan_gag_bwd(const float* x, const float* key, const float* value, const float* attn,
                             const float* d_out, const float* d_attn, int B, int idf, int Q, int T, float* d_x,
                             float* d_key, float* d_value, void* stream) {
    EEGAN_REQUIRE(B > 0 && idf > 0 && Q > 0 && T > 0, "gag: empty shape B=%d idf=%d Q=%d T=%d", B, idf, Q, T);
    EEGAN_REQUIRE(T <= 32 && idf <= 256, "gag bwd: T=%d (<=32) idf=%d (<=256) unsupported", T, idf);
    EEGAN_REQUIRE(x && key && value && attn && d_x && d_key && d_value, "gag bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t tma_smem = (size_t)GB_NS * GB_STAGE_BYTES + ((size_t)4 * idf * 32 + 2 * GB_TILE * 32) * sizeof(float) + 1024;
    if (d_out && Q % 4 == 0 && idf % GB_DC == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0 &&
        tma_smem <= 220 * 1024)
        return GAG_DISPATCH(T, gag_bwd_tma_launch)(x, key, value, attn, d_out, d_attn, B, idf, Q, T, d_x, d_key, d_value, st);
    return GAG_DISPATCH(T, gag_bwd_launch)(x, key, value, attn, d_out, d_attn, B, idf, Q, T, d_x, d_key, d_value, st);
}
