// Kernels shared by the two fused pair-grid pipelines (pair_grid_v3.cu: 3xTF32 engine; pair_grid_h.cu: half-pair engine):
// caption packing, cosine + log-sum-exp, diagonal attention maps, d_words unpack.  Included by both translation units
// (static linkage: no relocatable device code needed).
#pragma once
#include "common.cuh"

namespace eegan {

constexpr int V3_BIN = 64;     // packed columns per bin (captions never straddle a bin)
constexpr int V3_DU_JG = 64;   // images per CTA of the dU kernel (8 warps x 8 images)

static int v3_nsplit(int Bi, int NtP, int D) {
    // GEMM5 reduces over images inside its K loop; split so that about one wave of live tiles exists
    const int live_m = ((int)(0.6f * NtP) + 127) / 128 > 0 ? ((int)(0.6f * NtP) + 127) / 128 : 1;
    const int tiles = live_m * ((D + 127) / 128);
    int ns = 148 / tiles;
    if (ns < 1) ns = 1;
    if (ns > Bi) ns = Bi;
    const int nred = (Bi + ns - 1) / ns;
    return (Bi + nred - 1) / nred;
}

// ---------------------------------------------------------------------------------------
// prologue
// ---------------------------------------------------------------------------------------
// Greedy, order-preserving packing of whole captions into 64-column bins, then the column ->
// caption map.  The lengths are staged in shared memory so that the one sequential pass (thread
// 0; B is at most a few thousand) never waits on global memory.
__device__ __forceinline__ void v3_scan_body(int* s_buf /* shared: [2 Bc] */, const int32_t* __restrict__ cap_lens, int Bc, int Tm,
                                             int maxbins, int* __restrict__ col_start, int* __restrict__ cap_len,
                                             int* __restrict__ bin_cap, int* __restrict__ bin_used, int* __restrict__ meta,
                                             int* __restrict__ col_cap) {
    int* s_len = s_buf;  // [Bc] lengths, then [Bc] first columns
    int* s_cs = s_buf + Bc;
    __shared__ int s_nbins;
    for (int i = threadIdx.x; i < Bc; i += blockDim.x) s_len[i] = min(max(cap_lens[i], 0), Tm);
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0, fill = 0;
        bin_cap[0] = 0;
        for (int i = 0; i < Bc; ++i) {
            const int T = s_len[i];
            if (fill + T > V3_BIN) {
                bin_used[b] = fill;
                ++b;
                bin_cap[b] = i;
                fill = 0;
            }
            s_cs[i] = b * V3_BIN + fill;
            fill += T;
        }
        bin_used[b] = fill;
        const int nbins = b + 1;  // <= maxbins by construction
        bin_cap[nbins] = Bc;
        for (int q = nbins + 1; q <= maxbins; ++q) bin_cap[q] = Bc;
        for (int q = nbins; q < maxbins; ++q) bin_used[q] = 0;
        col_start[Bc] = nbins * V3_BIN;
        meta[0] = nbins;
        meta[1] = nbins * V3_BIN;
        s_nbins = nbins;
    }
    __syncthreads();
    const int ntotp = s_nbins * V3_BIN;
    for (int n = threadIdx.x; n < ntotp; n += blockDim.x) col_cap[n] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < Bc; i += blockDim.x) {
        const int cs = s_cs[i], T = s_len[i];
        col_start[i] = cs;
        cap_len[i] = T;
        for (int t = 0; t < T; ++t) col_cap[cs + t] = i;
    }
}

static __global__ void __launch_bounds__(256) v3_scan_kernel(const int32_t* __restrict__ cap_lens, int Bc, int Tm, int maxbins,
                                                             int* __restrict__ col_start, int* __restrict__ cap_len,
                                                             int* __restrict__ bin_cap, int* __restrict__ bin_used,
                                                             int* __restrict__ meta, int* __restrict__ col_cap) {
    extern __shared__ int s_scan_buf[];
    pdl_trigger();
    pdl_wait();
    v3_scan_body(s_scan_buf, cap_lens, Bc, Tm, maxbins, col_start, cap_len, bin_cap, bin_used, meta, col_cap);
}

// ---------------------------------------------------------------------------------------
// forward: cosine + log-sum-exp   (DAMSM_losses.py:17-23, :315-317)
// ---------------------------------------------------------------------------------------
// One CTA per (64-column bin, image j), 8 warps x 8 packed columns, two columns in flight per warp:
//   Z = sum of the per-32-region partials (fixed order: deterministic), u = U'/Z,
//   cos = <w,u> / max(|w||u|, 1e-8);   then one thread per caption of the bin: m[j][i] = log sum_t exp(g2 cos_t).
static __global__ void __launch_bounds__(256) v3_cos_lse_kernel(const float* __restrict__ U, const float* __restrict__ Wp,
                                                         const float* __restrict__ wn, const float* __restrict__ Zpart,
                                                         const int* __restrict__ col_start, const int* __restrict__ cap_len,
                                                         const int* __restrict__ bin_cap, const int* __restrict__ bin_used,
                                                         const int* __restrict__ meta, int NtP, int D, int Bc, int nz, float g2,
                                                         float* __restrict__ Z, float* __restrict__ cosv, float* __restrict__ un,
                                                         float* __restrict__ m, float* __restrict__ mst) {
    __shared__ float s_cos[V3_BIN];
    const int b = blockIdx.x, j = blockIdx.y;
    pdl_trigger();
    pdl_wait();
    if (b >= meta[0]) return;
    const int used = bin_used[b];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int d4 = D >> 2;
    constexpr int NCW = 8;  // packed columns per warp, all in flight at once (one round of loads per warp)
    {
        const int c0 = w * 8;
        if (c0 < used) {
        float dot[NCW], uu[NCW], zz[NCW];
        size_t nn[NCW];
        const float4* up[NCW];
        const float4* wp[NCW];
#pragma unroll
        for (int k = 0; k < NCW; ++k) {
            nn[k] = (size_t)b * V3_BIN + min(c0 + k, used - 1);
            up[k] = reinterpret_cast<const float4*>(U + ((size_t)j * NtP + nn[k]) * D);
            wp[k] = reinterpret_cast<const float4*>(Wp + nn[k] * D);
            zz[k] = lane < nz ? Zpart[((size_t)j * nz + lane) * NtP + nn[k]] : 0.f;
            dot[k] = 0.f;
            uu[k] = 0.f;
        }
        for (int q = lane; q < d4; q += 32) {
            float4 a[NCW], bw[NCW];
#pragma unroll
            for (int k = 0; k < NCW; ++k) { a[k] = up[k][q]; bw[k] = __ldg(wp[k] + q); }
#pragma unroll
            for (int k = 0; k < NCW; ++k) {
                dot[k] = fmaf(a[k].x, bw[k].x, dot[k]); dot[k] = fmaf(a[k].y, bw[k].y, dot[k]);
                dot[k] = fmaf(a[k].z, bw[k].z, dot[k]); dot[k] = fmaf(a[k].w, bw[k].w, dot[k]);
                uu[k] = fmaf(a[k].x, a[k].x, uu[k]); uu[k] = fmaf(a[k].y, a[k].y, uu[k]);
                uu[k] = fmaf(a[k].z, a[k].z, uu[k]); uu[k] = fmaf(a[k].w, a[k].w, uu[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < NCW; ++k) { dot[k] = warp_sum(dot[k]); uu[k] = warp_sum(uu[k]); zz[k] = warp_sum(zz[k]); }
        if (lane < NCW && c0 + lane < used) {
            float z = zz[0], d = dot[0], u2 = uu[0];
            size_t n = nn[0];
#pragma unroll
            for (int k = 1; k < NCW; ++k)
                if (lane == k) { z = zz[k]; d = dot[k]; u2 = uu[k]; n = nn[k]; }
            const float unv = sqrtf(u2) / z;  // |u|, u = U' / Z
            const float c = (d / z) / fmaxf(wn[n] * unv, 1e-8f);
            Z[(size_t)j * NtP + n] = z;
            cosv[(size_t)j * NtP + n] = c;
            un[(size_t)j * NtP + n] = unv;
            s_cos[c0 + lane] = c;
        }
        }
    }
    __syncthreads();
    const int i0 = bin_cap[b], i1 = bin_cap[b + 1];
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const int cl = col_start[i] - b * V3_BIN, T = cap_len[i];
        float sum = 0.f;
        for (int t = 0; t < T; ++t) sum += expf(g2 * s_cos[cl + t]);
        const float v = logf(sum);
        m[(size_t)j * Bc + i] = v;
        mst[(size_t)j * Bc + i] = v;
    }
}

// att_maps[i][t][r] = E^T[j][r][cs+t] / Z[j][cs+t] for the caption's own image j = i + diag_offset (:301).
// from_p: the half-pair engine keeps E only as fp16 pairs, so E is recomputed from the fp32 P stash.
// grid (captions, 32-region slabs): the slab goes through shared memory so that both the E^T reads
// (words contiguous) and the att writes (regions contiguous) are coalesced.
static __global__ void __launch_bounds__(256) v3_att_diag_kernel(const float* __restrict__ E, const float* __restrict__ Z,
                                                          const int* __restrict__ col_start, const int* __restrict__ cap_len,
                                                          int NtP, int R, int Tm, int Bi, int diag_offset, float* __restrict__ att,
                                                          int from_p, float g1) {
    __shared__ float tile[32][33];
    const int i = blockIdx.x, j = i + diag_offset, r0 = blockIdx.y * 32;
    pdl_trigger();
    pdl_wait();
    float* out = att + (size_t)i * Tm * R;
    const bool have = j >= 0 && j < Bi;
    const int cs = col_start[i], T = have ? cap_len[i] : 0;
    for (int idx = threadIdx.x; idx < 32 * Tm; idx += blockDim.x) {
        const int rr = idx / Tm, t = idx - rr * Tm;
        float v = 0.f;
        if (t < T && r0 + rr < R) {
            float e = E[((size_t)j * R + r0 + rr) * NtP + cs + t];
            if (from_p) {  // the array holds P: E = exp(g1 (P - 1)), same ex2.approx form as the GEMM epilogue that summed Z
                float y;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(g1 * (e - 1.0f) * 1.4426950408889634f));
                e = y;
            }
            v = e / Z[(size_t)j * NtP + cs + t];
        }
        tile[t][rr] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * Tm; idx += blockDim.x) {
        const int t = idx / 32, rr = idx - t * 32;
        if (r0 + rr < R) out[(size_t)t * R + r0 + rr] = tile[t][rr];
    }
}

// d_words[i][d][t] = dwcos + sum of the split-j partials; zero for padded words.
static __global__ void __launch_bounds__(256) v3_unpack_dw_kernel(const float* __restrict__ dWpart, const float* __restrict__ dwcos,
                                                           const int* __restrict__ col_start, const int* __restrict__ cap_len,
                                                           int nsplit, int ngroups, int NtP, int D, int Tm,
                                                           float* __restrict__ d_words) {
    __shared__ float tile[32][33];
    const int i = blockIdx.x, d0 = blockIdx.y * 32;
    pdl_trigger();
    pdl_wait();
    const int cs = col_start[i], T = cap_len[i];
    for (int idx = threadIdx.x; idx < 32 * Tm; idx += blockDim.x) {
        const int t = idx / 32, dd = idx % 32;
        float v = 0.f;
        if (t < T && d0 + dd < D) {
            const size_t k = (size_t)(cs + t) * D + d0 + dd, plane = (size_t)NtP * D;
            // the cosine part and the first sixteen split partials leave in ONE batch of loads (one memory round trip for the
            // usual nsplit <= 16 instead of three); the summation order stays fixed
            float c0 = ngroups > 0 ? dwcos[k] : 0.f;
            float p16[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) p16[q] = (q < nsplit) ? dWpart[(size_t)q * plane + k] : 0.f;
            for (int s = 1; s < ngroups; ++s) c0 += dwcos[(size_t)s * plane + k];
            v = c0;
            v += ((p16[0] + p16[1]) + (p16[2] + p16[3])) + ((p16[4] + p16[5]) + (p16[6] + p16[7]));
            v += ((p16[8] + p16[9]) + (p16[10] + p16[11])) + ((p16[12] + p16[13]) + (p16[14] + p16[15]));
            for (int s0 = 16; s0 < nsplit; s0 += 8) {  // eight partials in flight per further round
                float p8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) p8[q] = (s0 + q < nsplit) ? dWpart[(size_t)(s0 + q) * plane + k] : 0.f;
                v += ((p8[0] + p8[1]) + (p8[2] + p8[3])) + ((p8[4] + p8[5]) + (p8[6] + p8[7]));
            }
        }
        tile[dd][t] = v;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * Tm; idx += blockDim.x) {
        const int dd = idx / Tm, t = idx % Tm;
        if (d0 + dd < D) d_words[((size_t)i * D + d0 + dd) * Tm + t] = tile[dd][t];
    }
}

}  // namespace eegan
