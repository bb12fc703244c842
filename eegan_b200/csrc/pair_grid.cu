// DAMSM pair grid: every caption against every image (miscc/DAMSM_losses.py:272-342 — the
// caption loop :281-321 over func_attention :25-63 and cosine_similarity :17-23), forward
// and backward, as a pipeline of GEMM-shaped contractions over *packed* word columns
// (n = (caption i, word t), only t < cap_lens[i]) with small fused row/column kernels
// between them.  Math: SURVEY.md App. A.
//
//   fwd:  S[j][n][r] = sum_d Wp[n][d] C[j][d][r]                (GEMM1, K = D)
//         P = softmax over the words of a caption; A = softmax over regions of g1*P
//         U[j][n][d] = sum_r A[j][n][r] C[j][d][r]              (GEMM2, K = R)
//         cos, m[j][i] = log sum_t exp(g2 cos)
//   bwd:  DU = a1 w - a2 u ; dA = DU . C (GEMM3) ; softmax backwards -> DS
//         dC[j] = DU^T A + Wp^T DS (GEMM4a/b, K = packed columns)
//         dWp   = sum_j DS C^T (GEMM5, split over j) + cosine term
//
// The forward stashes P, A, U, cos, |u| in the caller's workspace; the backward consumes
// them (no recompute).  Nothing here synchronises the device or allocates memory.
#include "common.cuh"
#include <atomic>

#include <stdlib.h>

#include "gemm_ffma.cuh"
#include "gemm_tc.cuh"
#include "pair_h.cuh"
#include "pair_v3.cuh"
#include "ptx.cuh"

namespace eegan {

// contraction engine: 3 = tcgen05 half-pair ("3xFP16") engine, operands pre-split as fp16 hi/lo, attention fused
// into the GEMM epilogues (pair_grid_h.cu, default), 2 = tcgen05 3xTF32 with the attention fused into the GEMM
// epilogues (pair_grid_v3.cu), 1 = tcgen05 3xTF32 GEMMs + separate row/column kernels,
// 0 = CUDA-core fp32 FFMA (exact-fp32 A/B reference).  Process-wide (the backward runs on
// autograd's thread); set via eegan_set_contraction_engine().
static int default_engine() {
    const char* e = getenv("EEGAN_ENGINE");  // A/B runs of the whole test suite / bench on another engine
    const int v = e ? atoi(e) : 3;
    return (v >= 0 && v <= 3) ? v : 3;
}
static std::atomic<int> g_engine{default_engine()};
static bool fused_ok(int D) { return D % 128 == 0; }  // the fused engine's dU kernel works in 128-float chunks

constexpr int PAIR_DU_JG = 8;  // images per CTA of the dU kernel

struct PairWs {
    int* col_start;  // [Bc+1] exclusive prefix of clamp(cap_lens)
    int* ntot;       // [1]    = col_start[Bc]
    int* col_cap;    // [NtM]  caption of packed column n
    float* Wp;       // [NtM][D] packed words, d contiguous
    float* wn;       // [NtM]  |w_n|
    float* SP;       // [Bi][NtM][Rp] S, then P in place
    float* A;        // [Bi][NtM][Rp]
    float* U;        // [Bi][NtM][D]  U, then DU in place (bwd)
    float* cosv;     // [Bi][NtM]
    float* un;       // [Bi][NtM]
    float* alpha;    // [3][Bi][NtM]
    float* DA;       // [Bi][NtM][Rp] dA, then DS in place
    float* dWpart;   // [nsplit][NtM][D]
    float* dwcos;    // [ceil(Bi/8)][NtM][D] per-image-group cosine part of dW
    float* Cp;       // [Bi][D][Rp] image features re-pitched for TMA (only if Rp != R)
    int Rp;          // stash pitch of the region index: R rounded up to 4 floats (16 B)
    int nsplit;
    size_t bytes;
};

// GEMM5 (dW) reduces over the images inside its K loop; the images are split into nsplit groups
// whose partial sums are added by the unpack kernel.  The persistent GEMM runs one CTA per SM, so
// the split is chosen to give about one full wave of LIVE tiles: the live row count sum(cap_lens)
// is only known on the device, so the expected fill of ragged captions (~65 % of T_max) is used —
// a wrong guess only costs load balance, never correctness.
static int pick_nsplit(int Bi, int NtM, int D) {
    const int live_m = ((int)(0.65f * NtM) + 127) / 128 > 0 ? ((int)(0.65f * NtM) + 127) / 128 : 1;
    const int tiles = live_m * ((D + 127) / 128);
    int ns = 148 / tiles;
    if (ns < 1) ns = 1;
    if (ns > Bi) ns = Bi;
    const int nred = (Bi + ns - 1) / ns;
    return (Bi + nred - 1) / nred;  // drop empty groups
}

static PairWs carve(void* base, int Bi, int Bc, int D, int R, int Tm) {
    PairWs w;
    const size_t NtM = (size_t)Bc * Tm;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    const size_t Rp = (size_t)(R + 3) / 4 * 4;
    w.Rp = (int)Rp;
    w.nsplit = pick_nsplit(Bi, (int)NtM, D);
    w.col_start = (int*)take((Bc + 1) * sizeof(int));
    w.ntot = (int*)take(sizeof(int));
    w.col_cap = (int*)take(NtM * sizeof(int));
    w.Wp = (float*)take(NtM * D * sizeof(float));
    w.wn = (float*)take(NtM * sizeof(float));
    w.SP = (float*)take((size_t)Bi * NtM * Rp * sizeof(float));
    w.A = (float*)take((size_t)Bi * NtM * Rp * sizeof(float));
    w.U = (float*)take((size_t)Bi * NtM * D * sizeof(float));
    w.cosv = (float*)take((size_t)Bi * NtM * sizeof(float));
    w.un = (float*)take((size_t)Bi * NtM * sizeof(float));
    w.alpha = (float*)take((size_t)3 * Bi * NtM * sizeof(float));
    w.DA = (float*)take((size_t)Bi * NtM * Rp * sizeof(float));
    w.dWpart = (float*)take((size_t)w.nsplit * NtM * D * sizeof(float));
    w.dwcos = (float*)take((size_t)((Bi + PAIR_DU_JG - 1) / PAIR_DU_JG) * NtM * D * sizeof(float));
    w.Cp = (float*)take(Rp != (size_t)R ? (size_t)Bi * D * Rp * sizeof(float) : 0);
    w.bytes = off;
    return w;
}

// ---------------------------------------------------------------------------------------
// prologue: column packing
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pair_scan_kernel(const int32_t* __restrict__ cap_lens, int Bc, int Tm,
                                                         int* __restrict__ col_start, int* __restrict__ ntot,
                                                         int* __restrict__ col_cap) {
    __shared__ int s[1024];
    __shared__ int carry;
    const int tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < Bc; base += 1024) {
        const int i = base + tid;
        int v = (i < Bc) ? min(max(cap_lens[i], 0), Tm) : 0;
        s[tid] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = (tid >= o) ? s[tid - o] : 0;
            __syncthreads();
            s[tid] += t;
            __syncthreads();
        }
        const int excl = carry + s[tid] - v;
        if (i < Bc) {
            col_start[i] = excl;
            for (int t = 0; t < v; ++t) col_cap[excl + t] = i;
        }
        __syncthreads();
        if (tid == 1023) carry += s[1023];
        __syncthreads();
    }
    if (tid == 0) {
        col_start[Bc] = carry;
        *ntot = carry;
    }
}

__global__ void __launch_bounds__(128) pair_pack_words_kernel(const float* __restrict__ words, const int* __restrict__ col_start,
                                                              const int* __restrict__ col_cap, const int* __restrict__ ntot,
                                                              int D, int Tm, float* __restrict__ Wp, float* __restrict__ wn) {
    __shared__ float red[32];
    const int n = blockIdx.x;
    if (n >= *ntot) return;
    const int i = col_cap[n];
    const int t = n - col_start[i];
    float ss = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float v = __ldg(words + ((size_t)i * D + d) * Tm + t);
        Wp[(size_t)n * D + d] = v;
        ss = fmaf(v, v, ss);
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) wn[n] = sqrtf(ss);
}

// ---------------------------------------------------------------------------------------
// forward row/column kernels
// ---------------------------------------------------------------------------------------
// One CTA per (caption i, image j): P = softmax_words(S) (:44-45), A = softmax_regions(g1 P)
// (:53-54).  P overwrites S.  Dynamic smem: T_max * R floats.
// TMA-staged version: the caption's [T][Rp] score block is ONE contiguous chunk, so a single
// cp.async.bulk pulls it into shared memory (all bytes in flight at once, no registers), both
// softmaxes run out of shared memory, and P / A leave through two bulk stores.
__global__ void __launch_bounds__(256) pair_attn_softmax_kernel(float* __restrict__ SP, float* __restrict__ A,
                                                                const int* __restrict__ col_start, int NtM, int R, int Rp,
                                                                float g1, float* __restrict__ att, int diag_offset,
                                                                int Tm) {
    extern __shared__ __align__(128) float smf[];
    float* ps = smf;                       // [T][Rp]  S -> P
    float* as = smf + (size_t)Tm * Rp;     // [T][Rp]  A
    __shared__ __align__(8) unsigned long long bar_storage;
    const uint32_t bar = smem_u32(&bar_storage);
    const int i = blockIdx.x, j = blockIdx.y;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    if (T <= 0) return;
    float* Sj = SP + ((size_t)j * NtM + cs) * Rp;
    float* Aj = A + ((size_t)j * NtM + cs) * Rp;
    const uint32_t bytes = (uint32_t)T * Rp * 4u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        bulk_load(smem_u32(ps), Sj, bytes, bar);
    }
    mbar_wait(bar, 0);
    // softmax over the caption's words, per region (:44-45)
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float mx = -INFINITY;
        for (int t = 0; t < T; ++t) mx = fmaxf(mx, ps[t * Rp + r]);
        float sum = 0.f;
        for (int t = 0; t < T; ++t) {
            const float e = __expf(ps[t * Rp + r] - mx);
            ps[t * Rp + r] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
        for (int t = 0; t < T; ++t) ps[t * Rp + r] *= inv;
    }
    __syncthreads();
    // softmax over regions of gamma1 * P, per word (:53-54)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {
        const float* pr = ps + t * Rp;
        float* ar = as + t * Rp;
        float mx = -INFINITY;
        for (int r = lane; r < R; r += 32) mx = fmaxf(mx, pr[r]);
        mx = warp_max(mx) * g1;
        float sum = 0.f;
        for (int r = lane; r < R; r += 32) {
            const float e = __expf(fmaf(g1, pr[r], -mx));
            ar[r] = e;
            sum += e;
        }
        const float inv = 1.0f / warp_sum(sum);
        for (int r = lane; r < Rp; r += 32) ar[r] = (r < R) ? ar[r] * inv : 0.f;
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the bulk-store engine
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store(Sj, smem_u32(ps), bytes);
        bulk_store(Aj, smem_u32(as), bytes);
        bulk_store_commit_wait();
    }
    if ((att != nullptr) && (j == i + diag_offset)) {  // the caption's own image: att_maps (:301)
        float* attp = att + (size_t)i * Tm * R;
        for (int idx = threadIdx.x; idx < Tm * R; idx += blockDim.x) {
            const int t = idx / R, r = idx - t * R;
            attp[idx] = t < T ? as[t * Rp + r] : 0.f;
        }
    }
}

// Pitch-agnostic fallback (func_attention's unpadded buffers): same math, plain loads/stores.
__global__ void __launch_bounds__(256) pair_attn_softmax_generic_kernel(float* __restrict__ SP, float* __restrict__ A,
                                                                        const int* __restrict__ col_start, int NtM, int R,
                                                                        int Rp, float g1) {
    extern __shared__ float p[];  // [T][R]
    const int i = blockIdx.x, j = blockIdx.y;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    float* Sj = SP + ((size_t)j * NtM + cs) * Rp;
    float* Aj = A + ((size_t)j * NtM + cs) * Rp;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float mx = -INFINITY;
        for (int t = 0; t < T; ++t) mx = fmaxf(mx, Sj[(size_t)t * Rp + r]);
        float sum = 0.f;
        for (int t = 0; t < T; ++t) {
            const float e = __expf(Sj[(size_t)t * Rp + r] - mx);
            p[t * R + r] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
        for (int t = 0; t < T; ++t) {
            const float pv = p[t * R + r] * inv;
            p[t * R + r] = pv;
            Sj[(size_t)t * Rp + r] = pv;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {
        float mx = -INFINITY;
        for (int r = lane; r < R; r += 32) mx = fmaxf(mx, p[t * R + r]);
        mx = warp_max(mx) * g1;
        float sum = 0.f;
        for (int r = lane; r < R; r += 32) {
            const float e = __expf(fmaf(g1, p[t * R + r], -mx));
            p[t * R + r] = e;
            sum += e;
        }
        const float inv = 1.0f / warp_sum(sum);
        for (int r = lane; r < R; r += 32) Aj[(size_t)t * Rp + r] = p[t * R + r] * inv;
    }
}

// One CTA per (caption i, image j): cos_t (:17-23) and m = log sum_t exp(g2 cos_t) (:315-317).
__global__ void __launch_bounds__(256) pair_cos_lse_kernel(const float* __restrict__ U, const float* __restrict__ Wp,
                                                           const float* __restrict__ wn, const int* __restrict__ col_start,
                                                           int NtM, int D, int Bc, float g2, float* __restrict__ cosv,
                                                           float* __restrict__ un, float* __restrict__ m) {
    __shared__ float ex[32];
    const int i = blockIdx.x, j = blockIdx.y;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {
        const int n = cs + t;
        const float4* u4 = reinterpret_cast<const float4*>(U + ((size_t)j * NtM + n) * D);
        const float4* w4 = reinterpret_cast<const float4*>(Wp + (size_t)n * D);
        float dot = 0.f, uu = 0.f;
        for (int q = lane; q < D / 4; q += 32) {
            const float4 a = u4[q], b = __ldg(w4 + q);
            dot = fmaf(a.x, b.x, dot); dot = fmaf(a.y, b.y, dot); dot = fmaf(a.z, b.z, dot); dot = fmaf(a.w, b.w, dot);
            uu = fmaf(a.x, a.x, uu); uu = fmaf(a.y, a.y, uu); uu = fmaf(a.z, a.z, uu); uu = fmaf(a.w, a.w, uu);
        }
        dot = warp_sum(dot);
        uu = warp_sum(uu);
        if (lane == 0) {
            const float unv = sqrtf(uu);
            const float c = dot / fmaxf(wn[n] * unv, 1e-8f);
            cosv[(size_t)j * NtM + n] = c;
            un[(size_t)j * NtM + n] = unv;
            ex[t] = expf(g2 * c);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int t = 0; t < T; ++t) s += ex[t];
        m[(size_t)j * Bc + i] = logf(s);
    }
}

// ---------------------------------------------------------------------------------------
// backward row/column kernels
// ---------------------------------------------------------------------------------------
// One warp per (caption i, image j): dcos_t and the three per-column coefficients
//   a1 = dcos / max(|w||u|, eps);  a2 = dcos cos / |u|^2;  a3 = dcos cos / |w|^2  (0 when clamped)
// so that dU = a1 w - a2 u and the cosine part of dW is a1 u - a3 w.
__global__ void __launch_bounds__(32) pair_bwd_scalars_kernel(const float* __restrict__ dm, const float* __restrict__ cosv,
                                                              const float* __restrict__ un, const float* __restrict__ wn,
                                                              const int* __restrict__ col_start, int NtM, int Bi, int Bc,
                                                              float g2, float* __restrict__ alpha) {
    const int i = blockIdx.x, j = blockIdx.y, t = threadIdx.x;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    const bool on = t < T;
    const size_t idx = (size_t)j * NtM + cs + t;
    const float c = on ? cosv[idx] : 0.f;
    const float e = on ? expf(g2 * c) : 0.f;
    const float s = warp_sum(e);
    if (!on) return;
    const float dcos = dm[(size_t)j * Bc + i] * g2 * e / s;
    const float wv = wn[cs + t], uv = un[idx];
    const float nn = wv * uv;
    const bool live = nn > 1e-8f;
    const size_t plane = (size_t)Bi * NtM;
    alpha[idx] = dcos / fmaxf(nn, 1e-8f);
    alpha[plane + idx] = live ? dcos * c / (uv * uv) : 0.f;
    alpha[2 * plane + idx] = live ? dcos * c / (wv * wv) : 0.f;
}

// Thread = (packed column n, 4 channels), blockIdx.y = image group jg (PAIR_DU_JG images):
// U -> DU in place and the group's share of the cosine part of dW:
//   dwcos[jg][n][d] = sum_{j in jg} a1 u - (sum_{j in jg} a3) w        (summed over jg by unpack)
__global__ void __launch_bounds__(256) pair_du_dwcos_kernel(float* __restrict__ U, const float* __restrict__ Wp,
                                                            const float* __restrict__ alpha, const int* __restrict__ ntot,
                                                            int NtM, int Bi, int D, float* __restrict__ dwcos) {
    const int d4 = D >> 2;                               // float4 per column
    const int per_cta = blockDim.x / d4;                 // columns per CTA (D <= 1024 -> >= 1)
    const int n = blockIdx.x * per_cta + threadIdx.x / d4, q4 = threadIdx.x % d4;
    if (threadIdx.x >= per_cta * d4 || n >= *ntot) return;
    const int jg = blockIdx.y, j0 = jg * PAIR_DU_JG;
    const size_t plane = (size_t)Bi * NtM;
    const float4 wv = reinterpret_cast<const float4*>(Wp + (size_t)n * D)[q4];
    float4 uv[PAIR_DU_JG];
#pragma unroll
    for (int q = 0; q < PAIR_DU_JG; ++q)  // all loads first: 8 independent 16-byte requests in flight
        if (j0 + q < Bi) uv[q] = reinterpret_cast<const float4*>(U + ((size_t)(j0 + q) * NtM + n) * D)[q4];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float a3s = 0.f;
#pragma unroll
    for (int q = 0; q < PAIR_DU_JG; ++q) {
        if (j0 + q < Bi) {
            const size_t idx = (size_t)(j0 + q) * NtM + n;
            const float a1 = alpha[idx], a2 = alpha[plane + idx];
            a3s += alpha[2 * plane + idx];
            acc.x = fmaf(a1, uv[q].x, acc.x); acc.y = fmaf(a1, uv[q].y, acc.y);
            acc.z = fmaf(a1, uv[q].z, acc.z); acc.w = fmaf(a1, uv[q].w, acc.w);
            float4 o;
            o.x = a1 * wv.x - a2 * uv[q].x; o.y = a1 * wv.y - a2 * uv[q].y;
            o.z = a1 * wv.z - a2 * uv[q].z; o.w = a1 * wv.w - a2 * uv[q].w;
            reinterpret_cast<float4*>(U + idx * D)[q4] = o;
        }
    }
    float4 o;
    o.x = acc.x - a3s * wv.x; o.y = acc.y - a3s * wv.y; o.z = acc.z - a3s * wv.z; o.w = acc.w - a3s * wv.w;
    reinterpret_cast<float4*>(dwcos + ((size_t)jg * NtM + n) * D)[q4] = o;
}

// One CTA per (caption i, image j): dA -> DS in place.
//   dz = a (dA - sum_r a dA);  v = g1 p dz;  ds = v - p sum_t v
__global__ void __launch_bounds__(256) pair_softmax_bwd_kernel(float* __restrict__ DA, const float* __restrict__ A,
                                                               const float* __restrict__ P, const int* __restrict__ col_start,
                                                               int NtM, int R, int Rp, float g1, int Tm) {
    extern __shared__ __align__(128) float smf[];
    float* g = smf;                          // [T][Rp] dA -> v -> dS
    float* a = smf + (size_t)Tm * Rp;        // [T][Rp]
    float* p = smf + (size_t)2 * Tm * Rp;    // [T][Rp]
    __shared__ float csum[32];
    __shared__ __align__(8) unsigned long long bar_storage;
    const uint32_t bar = smem_u32(&bar_storage);
    const int i = blockIdx.x, j = blockIdx.y;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    if (T <= 0) return;
    const size_t base = ((size_t)j * NtM + cs) * Rp;
    const uint32_t bytes = (uint32_t)T * Rp * 4u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the three [T][Rp] blocks are contiguous: three bulk copies, all in flight
        mbar_arrive_expect_tx(bar, 3 * bytes);
        bulk_load(smem_u32(g), DA + base, bytes, bar);
        bulk_load(smem_u32(a), A + base, bytes, bar);
        bulk_load(smem_u32(p), P + base, bytes, bar);
    }
    mbar_wait(bar, 0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {  // csum[t] = sum_r a dA
        float s = 0.f;
        for (int r = lane; r < R; r += 32) s = fmaf(a[t * Rp + r], g[t * Rp + r], s);
        s = warp_sum(s);
        if (lane == 0) csum[t] = s;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float q = 0.f;
        for (int t = 0; t < T; ++t) {
            const float v = g1 * p[t * Rp + r] * a[t * Rp + r] * (g[t * Rp + r] - csum[t]);  // g1 p dz
            g[t * Rp + r] = v;
            q += v;
        }
        for (int t = 0; t < T; ++t) g[t * Rp + r] -= p[t * Rp + r] * q;
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store(DA + base, smem_u32(g), bytes);
        bulk_store_commit_wait();
    }
}

// Pitch-agnostic fallback (func_attention's unpadded buffers).
__global__ void __launch_bounds__(256) pair_softmax_bwd_generic_kernel(float* __restrict__ DA, const float* __restrict__ A,
                                                                       const float* __restrict__ P,
                                                                       const int* __restrict__ col_start, int NtM, int R,
                                                                       int Rp, float g1) {
    __shared__ float csum[32];
    const int i = blockIdx.x, j = blockIdx.y;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    const size_t base = ((size_t)j * NtM + cs) * Rp;
    float* dA = DA + base;
    const float* a = A + base;
    const float* p = P + base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {
        float s = 0.f;
        for (int r = lane; r < R; r += 32) s = fmaf(a[(size_t)t * Rp + r], dA[(size_t)t * Rp + r], s);
        s = warp_sum(s);
        if (lane == 0) csum[t] = s;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float q = 0.f;
        for (int t = 0; t < T; ++t) {
            const size_t k = (size_t)t * Rp + r;
            q += g1 * p[k] * a[k] * (dA[k] - csum[t]);
        }
        for (int t = 0; t < T; ++t) {
            const size_t k = (size_t)t * Rp + r;
            const float pv = p[k];
            dA[k] = g1 * pv * a[k] * (dA[k] - csum[t]) - pv * q;
        }
    }
}

// d_words[i][d][t] = dwcos + sum of the split-j partials; zero for padded words.
// grid (Bc, D/32): a 32(d) x T_max tile goes through shared memory so that both the packed
// reads (d contiguous) and the d_words writes (t contiguous) are coalesced.
__global__ void __launch_bounds__(256) pair_unpack_dw_kernel(const float* __restrict__ dWpart, const float* __restrict__ dwcos,
                                                             const int* __restrict__ col_start, int nsplit, int ngroups,
                                                             int NtM, int D, int Tm, float* __restrict__ d_words) {
    __shared__ float tile[32][33];
    const int i = blockIdx.x, d0 = blockIdx.y * 32;
    const int cs = col_start[i], T = col_start[i + 1] - cs;
    for (int idx = threadIdx.x; idx < 32 * Tm; idx += blockDim.x) {
        const int t = idx / 32, dd = idx % 32;
        float v = 0.f;
        if (t < T && d0 + dd < D) {
            const size_t k = (size_t)(cs + t) * D + d0 + dd;
            for (int s = 0; s < ngroups; ++s) v += dwcos[(size_t)s * NtM * D + k];
            for (int s = 0; s < nsplit; ++s) v += dWpart[(size_t)s * NtM * D + k];
        }
        tile[dd][t] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * Tm; idx += blockDim.x) {
        const int dd = idx / Tm, t = idx % Tm;
        if (d0 + dd < D) d_words[((size_t)i * D + d0 + dd) * Tm + t] = tile[dd][t];
    }
}

// img [Bi*D][R] -> Cp [Bi*D][Rp] (TMA needs 16-byte row pitches; R = 289 is odd)
__global__ void __launch_bounds__(256) pair_repitch_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           long long rows, int R, int Rp) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const float* s = src + row * R;
        float* d = dst + row * Rp;
        for (int r = lane; r < Rp; r += 32) d[r] = r < R ? __ldg(s + r) : 0.f;
    }
}

// The K = packed-column contraction (dC) runs in blocks of 32 rows: rows [ntot, roundup32(ntot)) of
// the four operands must be zero, not stale.  grid (32, Bi, 4).
__global__ void __launch_bounds__(128) pair_zero_tail_kernel(float* __restrict__ DU, float* __restrict__ A,
                                                             float* __restrict__ DS, float* __restrict__ Wp,
                                                             const int* __restrict__ ntot, int NtM, int D, int Rp) {
    const int n = *ntot + blockIdx.x;
    if (n >= NtM || n >= (*ntot + 31) / 32 * 32) return;
    const int j = blockIdx.y, which = blockIdx.z;
    float* row;
    int len;
    if (which == 0) { row = DU + ((size_t)j * NtM + n) * D; len = D; }
    else if (which == 1) { row = A + ((size_t)j * NtM + n) * Rp; len = Rp; }
    else if (which == 2) { row = DS + ((size_t)j * NtM + n) * Rp; len = Rp; }
    else { if (j) return; row = Wp + (size_t)n * D; len = D; }
    for (int k = threadIdx.x; k < len; k += blockDim.x) row[k] = 0.f;
}

static bool bulk_ok(const void* p, int Rp) { return (Rp % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0); }

static int softmax_fwd_launch(dim3 grid, cudaStream_t st, float* SP, float* A, const int* col_start, int NtM, int R, int Rp,
                              float g1, float* att, int diag_offset, int Tm) {
    if (bulk_ok(SP, Rp) && bulk_ok(A, Rp)) {
        const size_t smem = (size_t)2 * Tm * Rp * sizeof(float);
        static SmemGrant grant;
        if (int rc = grant_dyn_smem(pair_attn_softmax_kernel, smem, grant, "pair softmax")) return rc;
        pair_attn_softmax_kernel<<<grid, 256, smem, st>>>(SP, A, col_start, NtM, R, Rp, g1, att, diag_offset, Tm);
    } else {
        EEGAN_REQUIRE(att == nullptr, "softmax: att output needs the padded pitch");
        const size_t smem = (size_t)Tm * R * sizeof(float);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(pair_attn_softmax_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("smem attr: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
        }
        pair_attn_softmax_generic_kernel<<<grid, 256, smem, st>>>(SP, A, col_start, NtM, R, Rp, g1);
    }
    return check_launch("pair softmax");
}
static int softmax_bwd_launch(dim3 grid, cudaStream_t st, float* DA, const float* A, const float* P, const int* col_start,
                              int NtM, int R, int Rp, float g1, int Tm) {
    if (bulk_ok(DA, Rp) && bulk_ok(A, Rp) && bulk_ok(P, Rp)) {
        const size_t smem = (size_t)3 * Tm * Rp * sizeof(float);
        static SmemGrant grant;
        if (int rc = grant_dyn_smem(pair_softmax_bwd_kernel, smem, grant, "pair softmax bwd")) return rc;
        pair_softmax_bwd_kernel<<<grid, 256, smem, st>>>(DA, A, P, col_start, NtM, R, Rp, g1, Tm);
    } else {
        pair_softmax_bwd_generic_kernel<<<grid, 256, 0, st>>>(DA, A, P, col_start, NtM, R, Rp, g1);
    }
    return check_launch("pair softmax bwd");
}

static int validate(int Bi, int Bc, int D, int R, int Tm) {
    EEGAN_REQUIRE(Bi > 0 && Bc > 0, "pair grid: empty batch (B_img=%d B_cap=%d)", Bi, Bc);
    EEGAN_REQUIRE(D > 0 && D % 4 == 0 && D <= 1024, "pair grid: D=%d must be a multiple of 4 and <= 1024", D);
    EEGAN_REQUIRE(R > 0 && R <= 1024, "pair grid: R=%d must be in [1,1024]", R);
    EEGAN_REQUIRE(Tm > 0 && Tm <= 32, "pair grid: T_max=%d must be in [1,32]", Tm);
    EEGAN_REQUIRE((size_t)Tm * R * sizeof(float) <= 200 * 1024, "pair grid: T_max*R too large for shared memory");
    return EEGAN_OK;
}

}  // namespace eegan

using namespace eegan;

extern "C" size_t eegan_damsm_pair_workspace_bytes(int B_img, int B_cap, int D, int R, int T_max) {
    if (B_img <= 0 || B_cap <= 0 || D <= 0 || R <= 0 || T_max <= 0 || T_max > 32) return 0;
    const size_t a = carve(nullptr, B_img, B_cap, D, R, T_max).bytes;
    const size_t b = pair_v3_workspace_bytes(B_img, B_cap, D, R, T_max);
    const size_t c = pair_h_workspace_bytes(B_img, B_cap, D, R, T_max);
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

// ---- the six contractions, on either engine -------------------------------------------
static const float* image_operand(const PairWs& w, const float* img) { return w.Cp ? w.Cp : img; }

// S / dA [j][n][Rp] = X[n][d] . img[j][d][r]   (X = Wp unbatched, or DU batched)
static int gemm_nd_dr(const PairWs& w, const float* X, bool x_batched, const float* img, float* out, int Bi, int NtM, int D,
                      int R, cudaStream_t st) {
    if (g_engine.load()) {
        TcGemm g{};
        g.nseg = 1;
        g.A[0] = TcOperand{X, nullptr, 1, D, x_batched ? (long long)NtM * D : 0, x_batched ? Bi : 1, NtM, D};
        g.B[0] = TcOperand{image_operand(w, img), nullptr, 0, w.Rp, (long long)D * w.Rp, Bi, R, D};
        g.C = out; g.ldc = w.Rp; g.bC = (long long)NtM * w.Rp; g.M = NtM; g.N = R; g.dynM = w.ntot; g.batch = Bi; g.nred = 1;
        return tc_gemm_launch(g, st);
    }
    GemmArgs g{};
    g.A = X; g.B = img; g.C = out;
    g.M = NtM; g.N = R; g.K = D; g.dynM = w.ntot;
    g.sAm = D; g.sAk = 1; g.sBk = R; g.sBn = 1; g.sCm = w.Rp; g.sCn = 1;
    g.bA = x_batched ? (long long)NtM * D : 0; g.bB = (long long)D * R; g.bC = (long long)NtM * w.Rp;
    g.nred = 1;
    launch_gemm_ffma<128, 64, 8, 4, true, true>(g, Bi, st);
    return check_launch("pair gemm (n,d)x(d,r)");
}

// U [j][n][D] / dWpart[s][n][D] = X[j][n][r] . img[j][d][r], optionally reduced over j in nsplit groups
static int gemm_nr_dr(const PairWs& w, const float* X, const float* img, float* out, int Bi, int NtM, int D, int R,
                      int nsplit, cudaStream_t st) {
    const int nred = nsplit ? (Bi + nsplit - 1) / nsplit : 1;
    const int batch = nsplit ? nsplit : Bi;
    if (g_engine.load()) {
        TcGemm g{};
        g.nseg = 1;
        g.A[0] = TcOperand{X, nullptr, 1, w.Rp, (long long)NtM * w.Rp, Bi, NtM, R};
        g.B[0] = TcOperand{image_operand(w, img), nullptr, 1, w.Rp, (long long)D * w.Rp, Bi, D, R};
        g.C = out; g.ldc = D; g.bC = (long long)NtM * D; g.M = NtM; g.N = D; g.dynM = w.ntot; g.batch = batch;
        g.nred = nred; g.red_total = nsplit ? Bi : 0;
        return tc_gemm_launch(g, st);
    }
    GemmArgs g{};
    g.A = X; g.B = img; g.C = out;
    g.M = NtM; g.N = D; g.K = R; g.dynM = w.ntot;
    g.sAm = w.Rp; g.sAk = 1; g.sBk = 1; g.sBn = R; g.sCm = D; g.sCn = 1;
    g.rA = (long long)NtM * w.Rp; g.rB = (long long)D * R;
    g.bA = g.rA * nred; g.bB = g.rB * nred; g.bC = (long long)NtM * D;
    g.nred = nred; g.red_total = nsplit ? Bi : 0;
    launch_gemm_ffma<128, 128, 8, 8, true, false>(g, batch, st);
    return check_launch("pair gemm (n,r)x(d,r)");
}

// d_img[j][d][r] = sum_n DU[j][n][d] A[j][n][r] + Wp[n][d] DS[j][n][r]
static int gemm_dc(const PairWs& w, float* d_img, int Bi, int NtM, int D, int R, cudaStream_t st) {
    if (g_engine.load()) {
        pair_zero_tail_kernel<<<dim3(32, Bi, 4), 128, 0, st>>>(w.U, w.A, w.DA, w.Wp, w.ntot, NtM, D, w.Rp);
        TcGemm g{};
        g.nseg = 2;
        g.A[0] = TcOperand{w.U, nullptr, 0, D, (long long)NtM * D, Bi, D, NtM};
        g.B[0] = TcOperand{w.A, nullptr, 0, w.Rp, (long long)NtM * w.Rp, Bi, R, NtM};
        g.A[1] = TcOperand{w.Wp, nullptr, 0, D, 0, 1, D, NtM};
        g.B[1] = TcOperand{w.DA, nullptr, 0, w.Rp, (long long)NtM * w.Rp, Bi, R, NtM};
        g.C = d_img; g.ldc = R; g.bC = (long long)D * R; g.M = D; g.N = R; g.dynK = w.ntot; g.batch = Bi; g.nred = 1;
        return tc_gemm_launch(g, st);
    }
    GemmArgs g{};
    g.A = w.U; g.B = w.A; g.C = d_img;
    g.M = D; g.N = R; g.K = NtM; g.dynK = w.ntot;
    g.sAm = 1; g.sAk = D; g.sBk = w.Rp; g.sBn = 1; g.sCm = R; g.sCn = 1;
    g.bA = (long long)NtM * D; g.bB = (long long)NtM * w.Rp; g.bC = (long long)D * R;
    g.nred = 1;
    launch_gemm_ffma<128, 64, 8, 4, false, true>(g, Bi, st);
    g.A = w.Wp; g.B = w.DA; g.bA = 0; g.accumulate = 1;
    launch_gemm_ffma<128, 64, 8, 4, false, true>(g, Bi, st);
    return check_launch("pair gemm dC");
}

extern "C" int eegan_set_contraction_engine(int engine) {
    EEGAN_REQUIRE(engine >= 0 && engine <= 3,
                  "contraction engine must be 0 (fp32 FFMA), 1 (tcgen05 3xTF32), 2 (tcgen05 3xTF32, fused attention) or 3 (tcgen05 half pairs, fused attention)");
    g_engine.store(engine);
    return EEGAN_OK;
}
extern "C" int eegan_get_contraction_engine(void) { return g_engine.load(); }

extern "C" int eegan_damsm_pair_fwd(const float* img, const float* words, const int32_t* cap_lens, int Bi, int Bc,
                                    int D, int R, int Tm, float g1, float g2, float* m, float* att, int diag_offset,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate(Bi, Bc, D, R, Tm);
    if (rc) return rc;
    EEGAN_REQUIRE(img && words && cap_lens && m && workspace, "pair fwd: null pointer");
    if (g_engine.load() == 3 && fused_ok(D))
        return pair_h_fwd(img, words, cap_lens, Bi, Bc, D, R, Tm, g1, g2, m, att, diag_offset, workspace, workspace_bytes,
                          (cudaStream_t)stream);
    if (g_engine.load() >= 2 && fused_ok(D))
        return pair_v3_fwd(img, words, cap_lens, Bi, Bc, D, R, Tm, g1, g2, m, att, diag_offset, workspace, workspace_bytes,
                           (cudaStream_t)stream);
    PairWs w = carve(workspace, Bi, Bc, D, R, Tm);
    if (workspace_bytes < w.bytes) {
        set_error("pair fwd: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return EEGAN_ERR_WORKSPACE;
    }
    if (w.Rp == R) w.Cp = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    const int NtM = Bc * Tm;

    prof_mark(-1, st);
    pair_scan_kernel<<<1, 1024, 0, st>>>(cap_lens, Bc, Tm, w.col_start, w.ntot, w.col_cap);
    pair_pack_words_kernel<<<NtM, 128, 0, st>>>(words, w.col_start, w.col_cap, w.ntot, D, Tm, w.Wp, w.wn);
    if (w.Cp && g_engine.load())
        pair_repitch_kernel<<<148 * 8, 256, 0, st>>>(img, w.Cp, (long long)Bi * D, R, w.Rp);
    EEGAN_LAUNCH_CHECK("pair prologue");
    prof_mark(0, st);

    rc = gemm_nd_dr(w, w.Wp, false, img, w.SP, Bi, NtM, D, R, st);  // GEMM1: S = Wp . C
    if (rc) return rc;
    prof_mark(1, st);

    rc = softmax_fwd_launch(dim3(Bc, Bi), st, w.SP, w.A, w.col_start, NtM, R, w.Rp, g1, att, diag_offset, Tm);
    if (rc) return rc;
    prof_mark(2, st);

    rc = gemm_nr_dr(w, w.A, img, w.U, Bi, NtM, D, R, 0, st);  // GEMM2: U = A . C^T
    if (rc) return rc;
    prof_mark(3, st);

    pair_cos_lse_kernel<<<dim3(Bc, Bi), 256, 0, st>>>(w.U, w.Wp, w.wn, w.col_start, NtM, D, Bc, g2, w.cosv, w.un, m);
    EEGAN_LAUNCH_CHECK("pair cos/lse");
    prof_mark(4, st);
    return EEGAN_OK;
}

extern "C" int eegan_damsm_pair_bwd_phased(const float* img, const float* words, const int32_t* cap_lens, int Bi, int Bc,
                                           int D, int R, int Tm, float g1, float g2, const float* dm, float* d_img,
                                           float* d_words, int phases, void* workspace, size_t workspace_bytes, void* stream) {
    EEGAN_REQUIRE(phases >= 1 && phases <= 3, "pair bwd: phases=%d (1 = image part, 2 = words part, 3 = both)", phases);
    if (phases == 3) return eegan_damsm_pair_bwd(img, words, cap_lens, Bi, Bc, D, R, Tm, g1, g2, dm, d_img, d_words, workspace, workspace_bytes, stream);
    int rc = validate(Bi, Bc, D, R, Tm);
    if (rc) return rc;
    EEGAN_REQUIRE(img && dm && workspace, "pair bwd: null pointer");
    EEGAN_REQUIRE(g_engine.load() == 3 && fused_ok(D), "pair bwd: split phases need the default contraction engine (3) and D %% 128 == 0");
    return pair_h_bwd(img, Bi, Bc, D, R, Tm, g1, g2, dm, d_img, d_words, workspace, workspace_bytes, (cudaStream_t)stream, phases);
}

extern "C" int eegan_damsm_pair_bwd(const float* img, const float* words, const int32_t* cap_lens, int Bi, int Bc,
                                    int D, int R, int Tm, float g1, float g2, const float* dm, float* d_img,
                                    float* d_words, void* workspace, size_t workspace_bytes, void* stream) {
    (void)words; (void)cap_lens;
    int rc = validate(Bi, Bc, D, R, Tm);
    if (rc) return rc;
    EEGAN_REQUIRE(img && dm && workspace, "pair bwd: null pointer");
    if (g_engine.load() == 3 && fused_ok(D))
        return pair_h_bwd(img, Bi, Bc, D, R, Tm, g1, g2, dm, d_img, d_words, workspace, workspace_bytes, (cudaStream_t)stream);
    if (g_engine.load() >= 2 && fused_ok(D))
        return pair_v3_bwd(img, Bi, Bc, D, R, Tm, g1, g2, dm, d_img, d_words, workspace, workspace_bytes, (cudaStream_t)stream);
    PairWs w = carve(workspace, Bi, Bc, D, R, Tm);
    if (workspace_bytes < w.bytes) {
        set_error("pair bwd: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return EEGAN_ERR_WORKSPACE;
    }
    if (w.Rp == R) w.Cp = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    const int NtM = Bc * Tm;

    prof_mark(-1, st);
    pair_bwd_scalars_kernel<<<dim3(Bc, Bi), 32, 0, st>>>(dm, w.cosv, w.un, w.wn, w.col_start, NtM, Bi, Bc, g2, w.alpha);
    {
        const int per_cta = 256 / (D / 4) > 0 ? 256 / (D / 4) : 1;
        pair_du_dwcos_kernel<<<dim3((NtM + per_cta - 1) / per_cta, (Bi + PAIR_DU_JG - 1) / PAIR_DU_JG), D / 4 > 256 ? D / 4 : 256, 0, st>>>(
            w.U, w.Wp, w.alpha, w.ntot, NtM, Bi, D, w.dwcos);
    }
    EEGAN_LAUNCH_CHECK("pair bwd scalars");
    prof_mark(5, st);

    rc = gemm_nd_dr(w, w.U, true, img, w.DA, Bi, NtM, D, R, st);  // GEMM3: dA = DU . C
    if (rc) return rc;
    prof_mark(6, st);

    rc = softmax_bwd_launch(dim3(Bc, Bi), st, w.DA, w.A, w.SP, w.col_start, NtM, R, w.Rp, g1, Tm);
    if (rc) return rc;
    prof_mark(7, st);

    if (d_img) {
        rc = gemm_dc(w, d_img, Bi, NtM, D, R, st);  // GEMM4: dC = DU^T A + Wp^T DS
        if (rc) return rc;
        prof_mark(8, st);
    }
    if (d_words) {
        rc = gemm_nr_dr(w, w.DA, img, w.dWpart, Bi, NtM, D, R, w.nsplit, st);  // GEMM5: dWp = sum_j DS . C^T
        if (rc) return rc;
        prof_mark(9, st);
        pair_unpack_dw_kernel<<<dim3(Bc, (D + 31) / 32), 256, 0, st>>>(w.dWpart, w.dwcos, w.col_start, w.nsplit, (Bi + PAIR_DU_JG - 1) / PAIR_DU_JG, NtM, D, Tm, d_words);
        EEGAN_LAUNCH_CHECK("pair GEMM5");
        prof_mark(10, st);
    }
    return EEGAN_OK;
}

// =======================================================================================
// func_attention (miscc/DAMSM_losses.py:25-63) as a stand-alone op: sample b's query against
// sample b's context (the pair grid above is the all-pairs form of the same attention).
// Re-uses the grid's softmax kernels with a single "caption" of T words per sample.
// =======================================================================================
namespace eegan {

struct FaWs {
    int* col_start;  // {0, T}
    float* P;        // [B][T][R]
    float* DA;       // [B][T][R]
    size_t bytes;
};
static FaWs carve_fa(void* base, int B, int R, int T) {
    FaWs w;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    w.col_start = (int*)take(2 * sizeof(int));
    w.P = (float*)take((size_t)B * T * R * sizeof(float));
    w.DA = (float*)take((size_t)B * T * R * sizeof(float));
    w.bytes = off;
    return w;
}
__global__ void fa_init_kernel(int* col_start, int T) {
    col_start[0] = 0;
    col_start[1] = T;
}

// one warp per row: cos = <a,b> / max(|a||b|, eps)   (DAMSM_losses.py:17-23)
__global__ void __launch_bounds__(256) cosine_rows_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                              long long rows, int D, float eps, float* __restrict__ out,
                                                              float* __restrict__ norms) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* a = x1 + row * D;
    const float* b = x2 + row * D;
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float av = a[d], bv = b[d];
        ab = fmaf(av, bv, ab); aa = fmaf(av, av, aa); bb = fmaf(bv, bv, bb);
    }
    ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
    if (lane == 0) {
        const float na = sqrtf(aa), nb = sqrtf(bb);
        out[row] = ab / fmaxf(na * nb, eps);
        norms[2 * row] = na;
        norms[2 * row + 1] = nb;
    }
}
__global__ void __launch_bounds__(256) cosine_rows_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                              const float* __restrict__ out, const float* __restrict__ norms,
                                                              const float* __restrict__ g, long long rows, int D, float eps,
                                                              float* __restrict__ d1, float* __restrict__ d2) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float na = norms[2 * row], nb = norms[2 * row + 1], c = out[row], gv = g[row];
    const float nn = na * nb;
    const bool live = nn > eps;
    const float k = gv / fmaxf(nn, eps);
    const float ka = live ? gv * c / (na * na) : 0.f, kb = live ? gv * c / (nb * nb) : 0.f;
    for (int d = lane; d < D; d += 32) {
        const float av = x1[row * D + d], bv = x2[row * D + d];
        d1[row * D + d] = k * bv - ka * av;
        d2[row * D + d] = k * av - kb * bv;
    }
}

}  // namespace eegan

extern "C" size_t eegan_func_attention_workspace_bytes(int B, int D, int R, int T) {
    (void)D;
    if (B <= 0 || R <= 0 || T <= 0) return 0;
    return carve_fa(nullptr, B, R, T).bytes;
}

extern "C" int eegan_func_attention_fwd(const float* query, const float* context, int B, int D, int R, int T,
                                        float gamma1, float* u, float* attn, void* workspace, size_t workspace_bytes,
                                        void* stream) {
    int rc = validate(B, B, D, R, T);
    if (rc) return rc;
    EEGAN_REQUIRE(query && context && u && attn && workspace, "func_attention fwd: null pointer");
    FaWs w = carve_fa(workspace, B, R, T);
    if (workspace_bytes < w.bytes) { set_error("func_attention: workspace too small"); return EEGAN_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    fa_init_kernel<<<1, 1, 0, st>>>(w.col_start, T);
    GemmArgs g{};
    // S[b][t][r] = sum_d q[b][d][t] c[b][d][r]      (:42)
    g.A = query; g.B = context; g.C = w.P;
    g.M = T; g.N = R; g.K = D;
    g.sAm = 1; g.sAk = T; g.sBk = R; g.sBn = 1; g.sCm = R; g.sCn = 1;
    g.bA = (long long)D * T; g.bB = (long long)D * R; g.bC = (long long)T * R;
    g.nred = 1;
    launch_gemm_ffma<128, 64, 8, 4, false, true>(g, B, st);
    rc = softmax_fwd_launch(dim3(1, B), st, w.P, attn, w.col_start, T, R, R, gamma1, nullptr, 0, T);
    if (rc) return rc;
    // u[b][d][t] = sum_r c[b][d][r] a[b][t][r]      (:61)
    g = GemmArgs{};
    g.A = context; g.B = attn; g.C = u;
    g.M = D; g.N = T; g.K = R;
    g.sAm = R; g.sAk = 1; g.sBk = 1; g.sBn = R; g.sCm = T; g.sCn = 1;
    g.bA = (long long)D * R; g.bB = (long long)T * R; g.bC = (long long)D * T;
    g.nred = 1;
    launch_gemm_ffma<128, 64, 8, 4, true, false>(g, B, st);
    EEGAN_LAUNCH_CHECK("func_attention fwd");
    return EEGAN_OK;
}

extern "C" int eegan_func_attention_bwd(const float* query, const float* context, const float* attn, const float* d_u,
                                        const float* d_attn, int B, int D, int R, int T, float gamma1, float* d_query,
                                        float* d_context, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = validate(B, B, D, R, T);
    if (rc) return rc;
    EEGAN_REQUIRE(query && context && attn && d_query && d_context && workspace, "func_attention bwd: null pointer");
    FaWs w = carve_fa(workspace, B, R, T);
    if (workspace_bytes < w.bytes) { set_error("func_attention: workspace too small"); return EEGAN_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nTR = (size_t)B * T * R * sizeof(float);
    if (d_attn) cudaMemcpyAsync(w.DA, d_attn, nTR, cudaMemcpyDeviceToDevice, st);
    else cudaMemsetAsync(w.DA, 0, nTR, st);
    GemmArgs g{};
    if (d_u) {  // dA[b][t][r] += sum_d d_u[b][d][t] c[b][d][r]
        g.A = d_u; g.B = context; g.C = w.DA;
        g.M = T; g.N = R; g.K = D;
        g.sAm = 1; g.sAk = T; g.sBk = R; g.sBn = 1; g.sCm = R; g.sCn = 1;
        g.bA = (long long)D * T; g.bB = (long long)D * R; g.bC = (long long)T * R;
        g.nred = 1; g.accumulate = 1;
        launch_gemm_ffma<128, 64, 8, 4, false, true>(g, B, st);
    }
    rc = softmax_bwd_launch(dim3(1, B), st, w.DA, attn, w.P, w.col_start, T, R, R, gamma1, T);
    if (rc) return rc;
    // d_context[b][d][r] = sum_t d_u[b][d][t] a[b][t][r] + q[b][d][t] ds[b][t][r]
    g = GemmArgs{};
    g.B = attn; g.C = d_context;
    g.M = D; g.N = R; g.K = T;
    g.sAm = T; g.sAk = 1; g.sBk = R; g.sBn = 1; g.sCm = R; g.sCn = 1;
    g.bA = (long long)D * T; g.bB = (long long)T * R; g.bC = (long long)D * R;
    g.nred = 1;
    if (d_u) {
        g.A = d_u;
        launch_gemm_ffma<128, 64, 8, 4, true, true>(g, B, st);
    }
    g.A = query; g.B = w.DA; g.accumulate = d_u ? 1 : 0;
    launch_gemm_ffma<128, 64, 8, 4, true, true>(g, B, st);
    // d_query[b][d][t] = sum_r c[b][d][r] ds[b][t][r]
    g = GemmArgs{};
    g.A = context; g.B = w.DA; g.C = d_query;
    g.M = D; g.N = T; g.K = R;
    g.sAm = R; g.sAk = 1; g.sBk = 1; g.sBn = R; g.sCm = T; g.sCn = 1;
    g.bA = (long long)D * R; g.bB = (long long)T * R; g.bC = (long long)D * T;
    g.nred = 1;
    launch_gemm_ffma<128, 64, 8, 4, true, false>(g, B, st);
    EEGAN_LAUNCH_CHECK("func_attention bwd");
    return EEGAN_OK;
}

extern "C" int eegan_cosine_rows_fwd(const float* x1, const float* x2, long long rows, int D, float eps, float* out,
                                     float* norms, void* stream) {
    EEGAN_REQUIRE(rows > 0 && D > 0 && x1 && x2 && out && norms, "cosine_rows fwd: bad arguments");
    cosine_rows_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x1, x2, rows, D, eps, out, norms);
    EEGAN_LAUNCH_CHECK("cosine_rows fwd");
    return EEGAN_OK;
}
extern "C" int eegan_cosine_rows_bwd(const float* x1, const float* x2, const float* out, const float* norms,
                                     const float* g, long long rows, int D, float eps, float* d_x1, float* d_x2,
                                     void* stream) {
    EEGAN_REQUIRE(rows > 0 && D > 0 && x1 && x2 && out && norms && g && d_x1 && d_x2, "cosine_rows bwd: bad arguments");
    cosine_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x1, x2, out, norms, g, rows, D,
                                                                                         eps, d_x1, d_x2);
    EEGAN_LAUNCH_CHECK("cosine_rows bwd");
    return EEGAN_OK;
}
