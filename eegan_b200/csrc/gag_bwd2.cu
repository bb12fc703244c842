// GlobalAttentionGeneral backward, second form (miscc/DAMSM_losses.py:96-132; math in SURVEY.md App. A).
//
// The first backward (gag.cu) does everything in one kernel and is bound by shared-memory operand fetches: every FMA of
// its four contractions takes one operand from shared memory, and the d_key / d_value sums over pixels need a CTA-wide
// reduction per 8-channel chunk.  Here the work is split so that every inner loop is FMA-bound out of registers:
//
//   kernel 1 (thread = 4 consecutive pixels)   dp[t] = sum_d d_out[d,q] v[d,t] + d_attn[t,q];  ds = p (dp - sum_t p dp)
//                                   d_x[d,q] = sum_t ds[t] key[d,t];  ds[t,q] -> workspace
//                                   key / value rows are broadcast float4 reads shared by the thread's four pixels; every
//                                   global access is a 16-byte one.
//   kernel 2, launched twice        d_key[d,:] = sum_q x[d,q] ds[:,q]   and   d_value[d,:] = sum_q d_out[d,q] p[:,q]
//   (gag_bwd_rowsum_px_kernel)      over one chunk of pixels, a warp = 8 channels (4 once T > 20), lanes along the pixels:
//                                   both operands arrive by 16-byte cp.async through a 4-deep shared-memory ring as
//                                   contiguous 512-byte segments, the 8 x TP sums stay in registers for the whole chunk.
//                                   (gag_bwd_rowsum_kernel, the first mapping — lanes on 32 channel groups, i.e. 32 rows
//                                   per 16-byte load — is kept behind EEGAN_GAG_ROWSUM=0: 1.8x slower.)
//   kernel 3                        fixed-order sum of the per-chunk partials (no atomics, nothing to zero).
//
// Extra HBM traffic against the one-kernel form: ds written and read once (2 T floats per pixel), d_out and attn read twice.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace eegan {

constexpr int G2_THREADS = 128;
constexpr int G2_DC = 8;   // channels per streamed d_out chunk (pixel kernel)
constexpr int G2_NS = 3;   // ring depth (3 x 16 KB + key/value: three CTAs per SM at idf = 128)

template <int TP, int PX>
__global__ void __launch_bounds__(G2_THREADS, 3) gag_bwd_px_kernel(const float* __restrict__ key, const float* __restrict__ value,
                                                                const float* __restrict__ attn, const float* __restrict__ d_out,
                                                                const float* __restrict__ d_attn, int idf, int Q, int T,
                                                                float* __restrict__ d_x, float* __restrict__ ds_out) {
    extern __shared__ __align__(16) float g2sm[];
    float* ks = g2sm;             // [idf][TP]
    float* vs = ks + idf * TP;    // [idf][TP]
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int idx = tid; idx < idf * TP; idx += G2_THREADS) {
        const int d = idx / TP, t = idx - d * TP;
        ks[idx] = t < T ? key[((size_t)b * idf + d) * T + t] : 0.f;
        vs[idx] = t < T ? value[((size_t)b * idf + d) * T + t] : 0.f;
    }
    __syncthreads();
    // a thread owns PX = 4 CONSECUTIVE pixels: every global access of the kernel is a 16-byte one (Q % 4 == 0, checked by
    // the caller: a quad is entirely inside the row or entirely outside)
    static_assert(PX == 4, "the pixel kernel is written for quads");
    int qb = (blockIdx.x * G2_THREADS + tid) * PX;
    const bool ok = qb < Q;
    if (!ok) qb = Q - PX;  // clamped loads, guarded stores
    float dp[PX][TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) {
        const float4 v = (d_attn && t < T) ? __ldg(reinterpret_cast<const float4*>(d_attn + ((size_t)b * T + t) * Q + qb))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        dp[0][t] = v.x; dp[1][t] = v.y; dp[2][t] = v.z; dp[3][t] = v.w;
    }
    if (d_out) {
        // d_out streams through a ring of [G2_DC channels][PX * 128 pixels] chunks filled by 16-byte cp.async, G2_NS - 1 chunks
        // ahead of the FMAs, so that the HBM latency never stalls the (few, register-heavy) warps; a thread's quad is one
        // conflict-free 16-byte read of a chunk row.
        constexpr int CH = PX * G2_THREADS;  // pixels per chunk row
        float* ring = vs + idf * TP;         // [G2_NS][G2_DC][CH]
        const int q0 = blockIdx.x * CH;
        const float* gbase = d_out + (size_t)b * idf * Q + q0;
        const int nch = idf / G2_DC;
        auto issue = [&](int c) {
            if (c < nch) {
                float* dst = ring + (size_t)(c % G2_NS) * G2_DC * CH;
#pragma unroll
                for (int i = 0; i < G2_DC * CH / 4 / G2_THREADS; ++i) {
                    const int f = tid + i * G2_THREADS, dd = f / (CH / 4), c4 = f - dd * (CH / 4);
                    const int src_bytes = q0 + 4 * c4 < Q ? 16 : 0;  // beyond the row: zero-fill (Q % 4 == 0)
                    const float* src = gbase + (size_t)(c * G2_DC + dd) * Q + (src_bytes ? 4 * c4 : 0);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + dd * CH + 4 * c4)),
                                 "l"(src), "r"(src_bytes) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");  // one group per chunk slot, empty past the end: uniform counting
        };
        for (int c = 0; c < G2_NS - 1; ++c) issue(c);
        for (int c = 0; c < nch; ++c) {
            asm volatile("cp.async.wait_group %0;" ::"n"(G2_NS - 2) : "memory");  // chunk c has landed (this thread's copies)
            __syncthreads();                                                     // ... and everyone's; stage (c-1) % NS is free
            issue(c + G2_NS - 1);
            const float* tile = ring + (size_t)(c % G2_NS) * G2_DC * CH;
#pragma unroll
            for (int dd = 0; dd < G2_DC; ++dd) {
                const float4 g4 = *reinterpret_cast<const float4*>(tile + dd * CH + PX * tid);
                const float g[PX] = {g4.x, g4.y, g4.z, g4.w};
                const float* vr = vs + (c * G2_DC + dd) * TP;
#pragma unroll
                for (int t = 0; t < TP; t += 4) {
                    const float4 v4 = *reinterpret_cast<const float4*>(vr + t);
#pragma unroll
                    for (int p = 0; p < PX; ++p) {
                        dp[p][t + 0] = fmaf(g[p], v4.x, dp[p][t + 0]);
                        dp[p][t + 1] = fmaf(g[p], v4.y, dp[p][t + 1]);
                        dp[p][t + 2] = fmaf(g[p], v4.z, dp[p][t + 2]);
                        dp[p][t + 3] = fmaf(g[p], v4.w, dp[p][t + 3]);
                    }
                }
            }
        }
    }
    {  // ds = p (dp - sum_t p dp), in place
        float dot[PX] = {0.f, 0.f, 0.f, 0.f};
        float pr[PX][TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const float4 v = t < T ? __ldg(reinterpret_cast<const float4*>(attn + ((size_t)b * T + t) * Q + qb)) : make_float4(0.f, 0.f, 0.f, 0.f);
            pr[0][t] = v.x; pr[1][t] = v.y; pr[2][t] = v.z; pr[3][t] = v.w;
#pragma unroll
            for (int p = 0; p < PX; ++p) dot[p] = fmaf(pr[p][t], dp[p][t], dot[p]);
        }
#pragma unroll
        for (int t = 0; t < TP; ++t) {
#pragma unroll
            for (int p = 0; p < PX; ++p) dp[p][t] = pr[p][t] * (dp[p][t] - dot[p]);
            if (t < T && ok)
                *reinterpret_cast<float4*>(ds_out + ((size_t)b * T + t) * Q + qb) = make_float4(dp[0][t], dp[1][t], dp[2][t], dp[3][t]);
        }
    }
    float* xbase = d_x + (size_t)b * idf * Q;
#pragma unroll 4
    for (int d = 0; d < idf; ++d) {
        const float* kr = ks + d * TP;
        float acc[PX];
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[p] = 0.f;
#pragma unroll
        for (int t = 0; t < TP; t += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kr + t);
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                acc[p] = fmaf(dp[p][t + 0], k4.x, acc[p]);
                acc[p] = fmaf(dp[p][t + 1], k4.y, acc[p]);
                acc[p] = fmaf(dp[p][t + 2], k4.z, acc[p]);
                acc[p] = fmaf(dp[p][t + 3], k4.w, acc[p]);
            }
        }
        if (ok) *reinterpret_cast<float4*>(xbase + (size_t)d * Q + qb) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// out[d,:] = sum over the chunk's pixels q of rows[d,q] * w[:,q]   (rows = x, w = ds -> d_key;  rows = d_out, w = p -> d_value)
// grid (chunks, B); blockDim = (idf / 4) * sub: thread = (row group of 4 channels, pixel subgroup sg); a round covers
// sub * 8 pixels; the subgroups meet in shared memory at the end (fixed order).
constexpr int G2_RPT = 4;   // rows (channels) per thread
constexpr int G2_PPR = 8;   // pixels per thread and round
template <int TP>
__global__ void __launch_bounds__(128, 3) gag_bwd_rowsum_kernel(const float* __restrict__ rows, const float* __restrict__ w, int idf,
                                                             int sub, int Q, int T, int Qc, float* __restrict__ part) {
    extern __shared__ __align__(16) float g2sm[];
    const int nrg = idf / G2_RPT, round_px = sub * G2_PPR;
    const int b = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
    const int rg = tid % nrg, sg = tid / nrg;
    const int q_beg = c * Qc, q_end = min(Q, q_beg + Qc);
    float acc[G2_RPT][TP];
#pragma unroll
    for (int r = 0; r < G2_RPT; ++r)
#pragma unroll
        for (int t = 0; t < TP; ++t) acc[r][t] = 0.f;
    const float* rbase = rows ? rows + ((size_t)b * idf + rg * G2_RPT) * Q : nullptr;
    // Software pipeline over the rounds: the loads of round i+1 — this thread's 4 x 8 row values (registers, ping-pong)
    // and its share of the w rows (4-byte cp.async straight into the other half of the double-buffered staging) — are
    // issued right before the FMAs of round i, so their latency hides behind ~640 FMA per thread; one barrier per round.
    float* w_s = g2sm;  // two buffers of [round_px][TP]
    const int wbuf = round_px * TP;
    auto load_x = [&](int q0, float (&xd)[G2_RPT][G2_PPR]) {
        const int qs = q0 + sg * G2_PPR;
        if (rbase && qs + G2_PPR <= q_end) {
#pragma unroll
            for (int r = 0; r < G2_RPT; ++r)
#pragma unroll
                for (int k = 0; k < G2_PPR / 4; ++k) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(rbase + (size_t)r * Q + qs) + k);
                    xd[r][4 * k] = a.x; xd[r][4 * k + 1] = a.y; xd[r][4 * k + 2] = a.z; xd[r][4 * k + 3] = a.w;
                }
        } else {
#pragma unroll
            for (int r = 0; r < G2_RPT; ++r)
#pragma unroll
                for (int k = 0; k < G2_PPR; ++k) xd[r][k] = (rbase && qs + k < q_end) ? __ldg(rbase + (size_t)r * Q + qs + k) : 0.f;
        }
    };
    auto issue_w = [&](int q0, float* dst) {
        for (int idx = tid; idx < TP * round_px; idx += blockDim.x) {
            const int t = idx / round_px, qq = idx - t * round_px;  // consecutive threads -> consecutive pixels: coalesced
            const int qg = q0 + qq;
            float* d = dst + qq * TP + t;
            if (t < T && qg < q_end) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(d)),
                             "l"(w + ((size_t)b * T + t) * Q + qg) : "memory");
            } else {
                *d = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto do_round = [&](float (&xc)[G2_RPT][G2_PPR], float (&xnext)[G2_RPT][G2_PPR], int q0, int cur) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // w_s[cur] has landed for every thread; every thread is done reading w_s[cur ^ 1]
        if (q0 + round_px < q_end) {
            load_x(q0 + round_px, xnext);
            issue_w(q0 + round_px, w_s + (cur ^ 1) * wbuf);
        }
        const float* wr = w_s + cur * wbuf + sg * G2_PPR * TP;
#pragma unroll
        for (int k = 0; k < G2_PPR; ++k) {
#pragma unroll
            for (int t = 0; t < TP; t += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(wr + k * TP + t);
#pragma unroll
                for (int r = 0; r < G2_RPT; ++r) {
                    acc[r][t + 0] = fmaf(xc[r][k], w4.x, acc[r][t + 0]);
                    acc[r][t + 1] = fmaf(xc[r][k], w4.y, acc[r][t + 1]);
                    acc[r][t + 2] = fmaf(xc[r][k], w4.z, acc[r][t + 2]);
                    acc[r][t + 3] = fmaf(xc[r][k], w4.w, acc[r][t + 3]);
                }
            }
        }
    };
    float xa[G2_RPT][G2_PPR], xb[G2_RPT][G2_PPR];
    load_x(q_beg, xa);
    issue_w(q_beg, w_s);
    for (int q0 = q_beg; q0 < q_end; q0 += 2 * round_px) {
        do_round(xa, xb, q0, 0);
        if (q0 + round_px < q_end) do_round(xb, xa, q0 + round_px, 1);
    }
    __syncthreads();
    // subgroups meet in shared memory, laid out [row r][t][subgroup][row group] so that a warp's lanes (consecutive row
    // groups) hit consecutive words
    float* red = g2sm;
    if (sub > 1) {
#pragma unroll
        for (int r = 0; r < G2_RPT; ++r)
#pragma unroll
            for (int t = 0; t < TP; ++t) red[((size_t)(r * TP + t) * sub + sg) * nrg + rg] = acc[r][t];
        __syncthreads();
        if (sg == 0)
            for (int s2 = 1; s2 < sub; ++s2)
#pragma unroll
                for (int r = 0; r < G2_RPT; ++r)
#pragma unroll
                    for (int t = 0; t < TP; ++t) acc[r][t] += red[((size_t)(r * TP + t) * sub + s2) * nrg + rg];
    }
    if (sg == 0) {
#pragma unroll
        for (int r = 0; r < G2_RPT; ++r) {
            float* pp = part + (((size_t)b * gridDim.x + c) * idf + rg * G2_RPT + r) * TP;
#pragma unroll
            for (int t = 0; t < TP; t += 4) *reinterpret_cast<float4*>(pp + t) = make_float4(acc[r][t], acc[r][t + 1], acc[r][t + 2], acc[r][t + 3]);
        }
    }
}

// Row sums, second mapping (default): the same contraction as gag_bwd_rowsum_kernel, with the warp's lanes along the PIXELS.
// The first mapping puts a warp's lanes on 32 different channel groups, so one 16-byte load per lane touches 32 different
// rows Q floats apart — 32 DRAM pages for 1 KB, and 32 L1 tag look-ups per instruction.  Here a warp owns RPT channels and
// a round of 128 pixels (lane l: pixels 4l .. 4l+3).  Both operands of a round arrive in shared memory by 16-byte cp.async
// through an NS-deep ring filled NS - 1 rounds ahead of the FMAs: the warp's own RPT rows as contiguous 512-byte segments,
// the w tile [TP][128] (shared by the CTA's 4 warps) likewise; every shared-memory read of the FMA loop is a conflict-free
// LDS.128 (RPT x 4 FMA per w load).  The RPT x TP sums stay in registers over the CTA's whole pixel chunk and meet once,
// through shared memory, at the end.  grid (chunks, idf / (4 RPT), B); 4 warps per CTA.
constexpr int G3_RPX = 128;  // pixels per round
template <int TP, int RPT, int NS>
__global__ void __launch_bounds__(128, (RPT > 4 || TP > 20) ? 2 : 3)
gag_bwd_rowsum_px_kernel(const float* __restrict__ rows, const float* __restrict__ w, int idf, int Q, int T, int Qc,
                         float* __restrict__ part) {
    extern __shared__ __align__(16) float g2sm[];
    constexpr int WT = TP * G3_RPX;        // floats per w tile
    constexpr int XT = 4 * RPT * G3_RPX;   // floats per x tile (4 warps x RPT rows)
    constexpr int ST = WT + XT;            // floats per stage: [TP][128] then [4 warps][RPT][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, b = blockIdx.z;
    const int ch0 = (blockIdx.y * 4 + warp) * RPT;
    const int q_beg = c * Qc, q_end = min(Q, q_beg + Qc);
    float acc[RPT][TP];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int t = 0; t < TP; ++t) acc[r][t] = 0.f;
    const float* rbase = rows ? rows + ((size_t)b * idf + ch0) * Q : nullptr;
    const float* wbase = w + (size_t)b * T * Q;
    auto issue = [&](int q0, int stage) {  // one commit group per call, empty past the chunk's end: uniform counting
        if (q0 < q_end) {
            float* wd = g2sm + stage * ST;
            float* xd = wd + WT + warp * RPT * G3_RPX;
            const int q = q0 + 4 * lane;  // q_end is a multiple of 4: a float4 group is all in or all out
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                float* d = xd + r * G3_RPX + 4 * lane;
                if (rbase && q < q_end) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d)),
                                 "l"(rbase + (size_t)r * Q + q) : "memory");
                } else {
                    *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int k = 0; k < TP / 4; ++k) {  // TP * 32 float4 groups over 128 threads: warp `warp` takes rows warp, warp + 4, ...
                const int t = warp + 4 * k;
                float* d = wd + t * G3_RPX + 4 * lane;
                if (t < T && q < q_end) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(d)),
                                 "l"(wbase + (size_t)t * Q + q) : "memory");
                } else {
                    *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s0 = 0; s0 < NS - 1; ++s0) issue(q_beg + s0 * G3_RPX, s0);
    int i = 0;
    for (int q0 = q_beg; q0 < q_end; q0 += G3_RPX, ++i) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NS - 2) : "memory");  // round i has landed (this thread's copies)
        __syncthreads();  // ... and everyone's; every thread is done reading round i - 1, whose stage is refilled next
        issue(q0 + (NS - 1) * G3_RPX, (i + NS - 1) % NS);
        const float* wt = g2sm + (i % NS) * ST + 4 * lane;
        const float* xt = wt + WT + warp * RPT * G3_RPX;
        float4 xv[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) xv[r] = *reinterpret_cast<const float4*>(xt + r * G3_RPX);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + t * G3_RPX);
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                float a0 = acc[r][t];
                a0 = fmaf(xv[r].x, wv.x, a0); a0 = fmaf(xv[r].y, wv.y, a0);
                a0 = fmaf(xv[r].z, wv.z, a0); a0 = fmaf(xv[r].w, wv.w, a0);
                acc[r][t] = a0;
            }
        }
    }
    // the lanes' partial sums meet through shared memory (the ring is free now): each warp parks its RPT x TP values as
    // [value][lane] with pitch 33 (conflict-free both ways), then lane l adds up values l, l + 32, ... and stores them —
    // RPT * TP / 32 coalesced rows instead of RPT * TP five-step shuffle reductions
    __syncthreads();
    float* red = g2sm + warp * (RPT * TP * 33);
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int t = 0; t < TP; ++t) red[(r * TP + t) * 33 + lane] = acc[r][t];
    __syncwarp();
    float* pp = part + (((size_t)b * gridDim.x + c) * idf + ch0) * TP;  // the warp's RPT rows are contiguous: [RPT][TP]
    for (int v = lane; v < RPT * TP; v += 32) {
        const float* rv = red + v * 33;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 32; ++k) s4[k & 3] += rv[k];
        pp[v] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
}

__global__ void __launch_bounds__(256) gag_bwd_kv_reduce_kernel(const float* __restrict__ part_k, const float* __restrict__ part_v,
                                                                int B, int idf, int T, int TP, int S, float* __restrict__ d_key,
                                                                float* __restrict__ d_value) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * idf * T) return;
    const int t = (int)(i % T);
    const long long bd = i / T;
    const int d = (int)(bd % idf), b = (int)(bd / idf);
    float sk = 0.f, sv = 0.f;
    const size_t o0 = (((size_t)b * S) * idf + d) * TP + t, step = (size_t)idf * TP;
    for (int c = 0; c < S; c += 4) {  // four partials of each array in flight; the order of the additions stays 0, 1, 2, ...
        float k4[4], v4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool on = c + q < S;
            k4[q] = on ? __ldg(part_k + o0 + (size_t)(c + q) * step) : 0.f;
            v4[q] = on ? __ldg(part_v + o0 + (size_t)(c + q) * step) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sk += k4[q];
            sv += v4[q];
        }
    }
    d_key[i] = sk;
    d_value[i] = sv;
}

struct G2Plan {
    int TP, sub, threads, S, Qc;
    int px_map;  // 8 / 4: gag_bwd_rowsum_px_kernel with that many channels per warp (default 8; EEGAN_GAG_RPT=4); 0: gag_bwd_rowsum_kernel (EEGAN_GAG_ROWSUM=0)
    size_t ds_bytes, part_bytes;
};

static bool g2_plan(int B, int idf, int Q, int T, G2Plan* pl) {
    if (idf % 32 != 0 || idf > 512 || T > 32) return false;
    pl->TP = T <= 8 ? 8 : T <= 12 ? 12 : T <= 16 ? 16 : T <= 20 ? 20 : T <= 24 ? 24 : 32;
    const int nrg = idf / G2_RPT;              // 8 .. 128 row groups
    pl->sub = 128 / nrg > 0 ? 128 / nrg : 1;   // pixel subgroups per CTA
    pl->threads = nrg * pl->sub;               // <= 128
    static const int px_map = [] { const char* e = getenv("EEGAN_GAG_ROWSUM"); return e ? atoi(e) != 0 : 1; }();
    static const int px_rpt = [] { const char* e = getenv("EEGAN_GAG_RPT"); return (e && atoi(e) == 4) ? 4 : 8; }();
    // 8 channels per warp need 8 x TP sums in registers: TP <= 20 (beyond that the kernel spills)
    pl->px_map = px_map ? ((px_rpt == 8 && pl->TP <= 20 && idf % 32 == 0) ? 8 : (idf % 16 == 0 ? 4 : 0)) : 0;
    if (pl->px_map) {
        // pixel chunks of whole 256-pixel rounds; the chunk count that fills the CTA slots (2 or 3 per SM) best, at most 4 waves
        const int rpt = pl->px_map;
        const int slots = 148 * ((rpt > 4 || pl->TP > 20) ? 2 : 3);
        const int per_chunk = B * (idf / (4 * rpt));
        const int max_s = (Q + G3_RPX - 1) / G3_RPX;
        int best = 1;
        double best_eff = 0.0;
        for (int S = 1; S <= max_s && (long long)S * per_chunk <= 4LL * slots; ++S) {
            const long long n = (long long)S * per_chunk, waves = (n + slots - 1) / slots;
            const double eff = (double)n / (double)(waves * slots);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = S; }
        }
        pl->Qc = ((Q + best - 1) / best + G3_RPX - 1) / G3_RPX * G3_RPX;
        pl->S = (Q + pl->Qc - 1) / pl->Qc;
        pl->ds_bytes = align_up((size_t)B * T * Q * sizeof(float), 256);
        pl->part_bytes = align_up((size_t)B * pl->S * idf * pl->TP * sizeof(float), 256);
        return true;
    }
    const int round_px = pl->sub * G2_PPR;
    // about nine CTAs per SM over the grid (EEGAN_GAG_CPS probe: 3 / 4 / 6 / 9 / 12 -> backward 1.45 / 1.37 / 1.31 / 1.17 / 1.20 ms at 256^2 x 32, flat elsewhere)
    static const int cps = [] { const char* e = getenv("EEGAN_GAG_CPS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 9; }();
    int S = (cps * 148 + B - 1) / B;
    const int max_s = (Q + round_px - 1) / round_px;
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    pl->Qc = ((Q + S - 1) / S + round_px - 1) / round_px * round_px;
    pl->S = (Q + pl->Qc - 1) / pl->Qc;
    pl->ds_bytes = align_up((size_t)B * T * Q * sizeof(float), 256);
    pl->part_bytes = align_up((size_t)B * pl->S * idf * pl->TP * sizeof(float), 256);
    return true;
}

template <int TP>
static int g2_launch(const G2Plan& pl, const float* x, const float* key, const float* value, const float* attn, const float* d_out,
                     const float* d_attn, int B, int idf, int Q, int T, float* d_x, float* d_key, float* d_value, float* dsw,
                     float* pk, float* pv, cudaStream_t st) {
    constexpr int PX = 4;
    const size_t smem1 = ((size_t)2 * idf * TP + (size_t)G2_NS * G2_DC * PX * G2_THREADS) * sizeof(float);
    if (smem1 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gag_bwd_px_kernel<TP, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) { set_error("gag bwd2 smem: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
    }
    gag_bwd_px_kernel<TP, PX><<<dim3((Q + G2_THREADS * PX - 1) / (G2_THREADS * PX), B), G2_THREADS, smem1, st>>>(
        key, value, attn, d_out, d_attn, idf, Q, T, d_x, dsw);
    EEGAN_LAUNCH_CHECK("gag bwd2 (pixels)");
    const size_t stage2 = (size_t)2 * pl.sub * G2_PPR * TP * sizeof(float), red2 = pl.sub > 1 ? (size_t)pl.sub * idf * TP * sizeof(float) : 0;
    const size_t smem2 = stage2 > red2 ? stage2 : red2;
    if (smem2 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gag_bwd_rowsum_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) { set_error("gag bwd2 smem: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
    }
    if (pl.px_map) {
        const int rpt = pl.px_map;
        static const int ns_env = [] { const char* e = getenv("EEGAN_GAG_NS"); return e ? atoi(e) : 0; }();  // ring depth probe
        const int ns = ns_env == 3 ? 3 : 4;
        const size_t ring3 = (size_t)ns * (TP + 4 * rpt) * G3_RPX * sizeof(float), red3 = (size_t)4 * rpt * TP * 33 * sizeof(float);
        const size_t smem3 = ring3 > red3 ? ring3 : red3;
        const dim3 grid3(pl.S, idf / (4 * rpt), B);
        auto run = [&](auto kern) -> int {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
            if (e != cudaSuccess) { set_error("gag bwd2 smem: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
            kern<<<grid3, 128, smem3, st>>>(x, dsw, idf, Q, T, pl.Qc, pk);
            kern<<<grid3, 128, smem3, st>>>(d_out, attn, idf, Q, T, pl.Qc, pv);
            return EEGAN_OK;
        };
        int rc;
        if (rpt == 8) rc = ns == 3 ? run(gag_bwd_rowsum_px_kernel<TP, 8, 3>) : run(gag_bwd_rowsum_px_kernel<TP, 8, 4>);
        else rc = ns == 3 ? run(gag_bwd_rowsum_px_kernel<TP, 4, 3>) : run(gag_bwd_rowsum_px_kernel<TP, 4, 4>);
        if (rc) return rc;
    } else {
        gag_bwd_rowsum_kernel<TP><<<dim3(pl.S, B), pl.threads, smem2, st>>>(x, dsw, idf, pl.sub, Q, T, pl.Qc, pk);
        gag_bwd_rowsum_kernel<TP><<<dim3(pl.S, B), pl.threads, smem2, st>>>(d_out, attn, idf, pl.sub, Q, T, pl.Qc, pv);
    }
    EEGAN_LAUNCH_CHECK("gag bwd2 (key/value)");
    const long long n = (long long)B * idf * T;
    gag_bwd_kv_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pk, pv, B, idf, T, TP, pl.S, d_key, d_value);
    return check_launch("gag bwd2 (reduce)");
}

}  // namespace eegan

using namespace eegan;

// gag_tc_bwd.cu: the one-pass tensor-core backward (idf = 32 / 64 / 128)
namespace eegan {
int gag_tc_bwd_chunks(int B, int Q);
bool gag_tc_bwd_supported(const float* x, const float* attn, const float* d_out, const float* d_attn, const float* d_x, int B, int idf, int Q, int T);
size_t gag_tc_bwd_part_floats(int B, int idf, int Q);
int gag_tc_bwd_launch(const float* x, const float* key, const float* value, const float* attn, const float* d_out, const float* d_attn,
                      int B, int idf, int Q, int T, float* d_x, float* part_k, float* part_v, int* chunks, cudaStream_t st);
}  // namespace eegan

// backward engine of eegan_gag_bwd_ws: 1 = tcgen05 one-pass kernel where the shape allows (default), 0 = CUDA-core kernels
static std::atomic<int> g_gag_bwd_engine{[] {
    const char* e = getenv("EEGAN_GAG_TC_BWD");
    return e ? atoi(e) : 1;
}()};
extern "C" int eegan_set_gag_bwd_engine(int engine) {
    EEGAN_REQUIRE(engine == 0 || engine == 1, "gag backward engine must be 0 (CUDA cores) or 1 (tensor cores)");
    g_gag_bwd_engine.store(engine);
    return EEGAN_OK;
}
extern "C" int eegan_get_gag_bwd_engine(void) { return g_gag_bwd_engine.load(); }

static size_t g2_tc_bytes(int B, int idf, int Q, int T) {
    if (!(idf == 32 || idf == 64 || idf == 128) || T > 32 || Q % 4 != 0) return 0;
    return 2 * align_up(gag_tc_bwd_part_floats(B, idf, Q) * sizeof(float), 256);
}

extern "C" size_t eegan_gag_bwd_workspace_bytes(int B, int idf, int Q, int T) {
    G2Plan pl;
    if (B <= 0 || idf <= 0 || Q <= 0 || T <= 0 || !g2_plan(B, idf, Q, T, &pl)) return 0;
    const size_t cc = pl.ds_bytes + 2 * pl.part_bytes, tc = g2_tc_bytes(B, idf, Q, T);
    return cc > tc ? cc : tc;
}

extern "C" int eegan_gag_bwd(const float* x, const float* key, const float* value, const float* attn, const float* d_out,
                             const float* d_attn, int B, int idf, int Q, int T, float* d_x, float* d_key, float* d_value,
                             void* stream);

// Backward with a caller-owned workspace (eegan_gag_bwd_workspace_bytes): the two-kernel register-tiled form above.
// Shapes it does not cover (idf not a multiple of 32, rows not 16-byte aligned) and workspace == NULL go to eegan_gag_bwd.
extern "C" int eegan_gag_bwd_ws(const float* x, const float* key, const float* value, const float* attn, const float* d_out,
                                const float* d_attn, int B, int idf, int Q, int T, float* d_x, float* d_key, float* d_value,
                                void* workspace, size_t workspace_bytes, void* stream) {
    EEGAN_REQUIRE(B > 0 && idf > 0 && Q > 0 && T > 0, "gag: empty shape B=%d idf=%d Q=%d T=%d", B, idf, Q, T);
    EEGAN_REQUIRE(x && key && value && attn && d_x && d_key && d_value, "gag bwd: null pointer");
    G2Plan pl;
    const bool aligned = Q % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(attn) |
                                         reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(d_attn) | reinterpret_cast<uintptr_t>(d_x)) & 15) == 0;
    if (!workspace || !aligned || !g2_plan(B, idf, Q, T, &pl) || B > 65535)
        return eegan_gag_bwd(x, key, value, attn, d_out, d_attn, B, idf, Q, T, d_x, d_key, d_value, stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (g_gag_bwd_engine.load() == 1 && gag_tc_bwd_supported(x, attn, d_out, d_attn, d_x, B, idf, Q, T)) {
        const size_t half = g2_tc_bytes(B, idf, Q, T) / 2;
        EEGAN_REQUIRE(workspace_bytes >= 2 * half, "gag bwd: workspace %zu < %zu bytes", workspace_bytes, 2 * half);
        float* tk = (float*)workspace;
        float* tv = (float*)((char*)workspace + half);
        int S = 0;
        if (int rc = gag_tc_bwd_launch(x, key, value, attn, d_out, d_attn, B, idf, Q, T, d_x, tk, tv, &S, st)) return rc;
        const long long n = (long long)B * idf * T;
        gag_bwd_kv_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tk, tv, B, idf, T, 32, S, d_key, d_value);
        return check_launch("gag tc bwd (reduce)");
    }
    EEGAN_REQUIRE(workspace_bytes >= pl.ds_bytes + 2 * pl.part_bytes, "gag bwd: workspace %zu < %zu bytes", workspace_bytes,
                  pl.ds_bytes + 2 * pl.part_bytes);
    float* dsw = (float*)workspace;
    float* pk = (float*)((char*)workspace + pl.ds_bytes);
    float* pv = (float*)((char*)workspace + pl.ds_bytes + pl.part_bytes);
#define G2_CALL(TPV) g2_launch<TPV>(pl, x, key, value, attn, d_out, d_attn, B, idf, Q, T, d_x, d_key, d_value, dsw, pk, pv, st)
    switch (pl.TP) {
        case 8: return G2_CALL(8);
        case 12: return G2_CALL(12);
        case 16: return G2_CALL(16);
        case 20: return G2_CALL(20);
        case 24: return G2_CALL(24);
        default: return G2_CALL(32);
    }
#undef G2_CALL
}
