// DAMSM pair grid, region-major fused pipeline (contraction engine 2, the default).
//
// Same math as pair_grid.cu (miscc/DAMSM_losses.py:272-342) but the word-region
// attention runs inside the tcgen05 GEMM epilogues, so the only kernels between the contractions
// are two small per-column passes:
//
//   fwd  scan/pack                 packed word columns n' in 64-column bins of whole captions
//        GEMM1 + attention fwd     S^T[j][r][n'] = C_j^T Wp^T in TMEM; per region thread: P = softmax_words(S),
//                                  E = exp(g1 (P - 1)); stores P^T, E^T and per-32-region partial sums of E
//        GEMM2                     U'[j][n'][d] = sum_r E^T[r][n'] C_j[d][r]      (u = U' / Z, Z = sum_r E)
//        cos/lse                   cos, |u|, m[j][i] = log sum_t exp(g2 cos); att_maps from the diagonal
//   bwd  dU                        DUz = (a1 w - a2 u) / Z, csz = <DU, u> / Z, cosine part of dW
//        GEMM3 + attention bwd     acc = C_j^T DUz^T = dA / Z in TMEM; per region thread:
//                                  v = g1 P E (acc - csz), dS = v - P sum_t v; stores dS^T
//        GEMM4                     dC_j[d][r] = sum_n' DUz[n'][d] E^T[r][n'] + Wp[n'][d] dS^T[r][n']
//        GEMM5 + unpack            dWp[n'][d] = sum_j sum_r dS^T[j][r][n'] C_j[d][r]
//
// sum_r A dA = <DU, u> (because u = A C^T), which is why the softmax backward needs no reduction
// over regions.  All stash arrays are [image j][region r][packed column n'] with n' contiguous, so
// every GEMM reads them through a legal K-major / MN-major tensor map without a transposed copy.
#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "pair_v3.cuh"
#include "pair_v3_kernels.cuh"

namespace eegan {


struct V3Ws {
    int* col_start;  // [Bc+1] first packed column of caption i (col_start[Bc] = live padded extent)
    int* cap_len;    // [Bc]   clamped caption lengths
    int* bin_cap;    // [maxbins+1]
    int* bin_used;   // [maxbins]
    int* meta;       // [0] = nbins, [1] = ntotp = nbins * 64
    int* col_cap;    // [NtP] caption of packed column n' (-1 = padding)
    float* Wp;       // [NtP][D]  packed words (zero rows for padding columns)
    float* wn;       // [NtP]
    float* Wlo;      // [NtP][D]  tf32 residual of Wp (pre-split B operand of GEMM1)
    float* Cp;       // [Bi][D][Rp] image features re-pitched for TMA (only if Rp != R)
    float* Clo;      // [Bi][D][Rp] tf32 residual of the image features (pre-split B operand of GEMM2 / GEMM5)
    float* P;        // [Bi][R][NtP]
    float* E;        // [Bi][R][NtP]
    float* Zpart;    // [Bi][ceil(R/32)][NtP]
    float* Z;        // [Bi][NtP]
    float* U;        // [Bi][NtP][D]  unnormalised word contexts U'
    float* DUz;      // [Bi][NtP][D]
    float* DUzlo;    // [Bi][NtP][D]  tf32 residual of DUz (pre-split B operand of GEMM3)
    float* cosv;     // [Bi][NtP]
    float* un;       // [Bi][NtP]  |u|
    float* csz;      // [Bi][NtP]
    float* mst;      // [Bi][Bc]   forward m (log-sum-exp), needed by the backward
    float* dS;       // [Bi][R][NtP]
    float* dWpart;   // [nsplit][NtP][D]
    float* dwcos;    // [ngroups][NtP][D]
    int Rp, NtP, maxbins, nsplit, ngroups, nz;
    size_t bytes;
};

// Pre-split (hi = raw, lo array) B operands take the B-splitter warps off the critical path at the price of 16 KB
// more L2->SM traffic per k-block.  Measured at B=48 (scratch notes in profiles/README.md): Wp for GEMM1 -5 us (on:
// the operand is tiny and L2-resident); the image features for GEMM2/GEMM5 and dU for GEMM3 +-0 (off).
static int presplit_c() {
    static int v = [] {
        const char* e = getenv("EEGAN_V3_PRESPLIT_C");
        return e ? atoi(e) : 0;
    }();
    return v;
}
static int presplit_du() {
    static int v = [] {
        const char* e = getenv("EEGAN_V3_PRESPLIT_DU");
        return e ? atoi(e) : 0;
    }();
    return v;
}

// extra (never live) bins in the column pitch of the stash arrays: keeps region rows off power-of-two strides
static int pad_bins() {
    static int v = [] {
        const char* e = getenv("EEGAN_V3_PADBINS");
        return e ? atoi(e) : 0;
    }();
    return v;
}


static V3Ws v3_carve(void* base, int Bi, int Bc, int D, int R, int Tm) {
    V3Ws w;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    const int per_bin = V3_BIN / Tm;  // >= 2 because Tm <= 32
    w.maxbins = (Bc + per_bin - 1) / per_bin;
    w.NtP = (w.maxbins + pad_bins()) * V3_BIN;
    w.Rp = (R + 3) / 4 * 4;
    w.nz = (R + 31) / 32;
    w.nsplit = v3_nsplit(Bi, w.NtP, D);
    w.ngroups = (Bi + V3_DU_JG - 1) / V3_DU_JG;
    const size_t NtP = (size_t)w.NtP;
    w.col_start = (int*)take((Bc + 1) * sizeof(int));
    w.cap_len = (int*)take(Bc * sizeof(int));
    w.bin_cap = (int*)take((w.maxbins + 1) * sizeof(int));
    w.bin_used = (int*)take(w.maxbins * sizeof(int));
    w.meta = (int*)take(4 * sizeof(int));
    w.col_cap = (int*)take(NtP * sizeof(int));
    w.Wp = (float*)take(NtP * D * sizeof(float));
    w.wn = (float*)take(NtP * sizeof(float));
    w.Wlo = (float*)take(NtP * D * sizeof(float));
    w.Cp = (float*)take(w.Rp != R ? (size_t)Bi * D * w.Rp * sizeof(float) : 0);
    w.Clo = (float*)take(presplit_c() ? (size_t)Bi * D * w.Rp * sizeof(float) : 0);
    w.P = (float*)take((size_t)Bi * R * NtP * sizeof(float));
    w.E = (float*)take((size_t)Bi * R * NtP * sizeof(float));
    w.Zpart = (float*)take((size_t)Bi * w.nz * NtP * sizeof(float));
    w.Z = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.U = (float*)take((size_t)Bi * NtP * D * sizeof(float));
    w.DUz = (float*)take((size_t)Bi * NtP * D * sizeof(float));
    w.DUzlo = (float*)take(presplit_du() ? (size_t)Bi * NtP * D * sizeof(float) : 0);
    w.cosv = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.un = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.csz = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.mst = (float*)take((size_t)Bi * Bc * sizeof(float));
    w.dS = (float*)take((size_t)Bi * R * NtP * sizeof(float));
    w.dWpart = (float*)take((size_t)w.nsplit * NtP * D * sizeof(float));
    w.dwcos = (float*)take((size_t)w.ngroups * NtP * D * sizeof(float));
    w.bytes = off;
    if (w.Rp == R) w.Cp = nullptr;
    if (!presplit_c()) w.Clo = nullptr;
    return w;
}


// One launch for the two independent input re-layouts:
//   CTAs [0, NtP):   one packed column each: gather the word vector (words is [i][d][t]) and its norm
//   CTAs [NtP, ...): img [Bi*D][R] -> Cp [Bi*D][Rp] (TMA needs 16-byte row pitches; R = 289 is odd)
__global__ void __launch_bounds__(128) v3_pack_kernel(const float* __restrict__ words, const int* __restrict__ col_start,
                                                      const int* __restrict__ col_cap, const int* __restrict__ meta, int D,
                                                      int Tm, int NtP, float* __restrict__ Wp, float* __restrict__ Wlo,
                                                      float* __restrict__ wn, const float* __restrict__ img, float* __restrict__ Cp,
                                                      float* __restrict__ Clo, long long rows, int R, int Rp) {
    __shared__ float red[32];
    if ((int)blockIdx.x >= NtP) {
        const int lane = threadIdx.x & 31;
        const long long warps = (long long)(gridDim.x - NtP) * (blockDim.x >> 5);
        for (long long row = (long long)(blockIdx.x - NtP) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
            const float* sp = img + row * R;
            float* dp = Cp ? Cp + row * Rp : nullptr;
            for (int r0 = 0; r0 < Rp; r0 += 32 * 10) {  // up to ten loads per lane in flight (R = 289 -> one round)
                float v[10];
#pragma unroll
                for (int q = 0; q < 10; ++q) {
                    const int r = r0 + 32 * q + lane;
                    v[q] = r < R ? __ldg(sp + r) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 10; ++q) {
                    const int r = r0 + 32 * q + lane;
                    if (r < Rp) {
                        if (Cp) dp[r] = v[q];
                        if (Clo) {
                            const float h = __uint_as_float(__float_as_uint(v[q]) & 0xffffe000u);
                            Clo[row * Rp + r] = __uint_as_float((__float_as_uint(v[q] - h) + 0x1000u) & 0xffffe000u);
                        }
                    }
                }
            }
        }
        return;
    }
    const int n = blockIdx.x;
    if (n >= meta[1]) return;
    const int i = col_cap[n];
    float ss = 0.f;
    if (i < 0) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            Wp[(size_t)n * D + d] = 0.f;
            Wlo[(size_t)n * D + d] = 0.f;
        }
    } else {
        const int t = n - col_start[i];
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const float v = __ldg(words + ((size_t)i * D + d) * Tm + t);
            Wp[(size_t)n * D + d] = v;
            // lo = tf32(v - trunc_tf32(v)): what the B splitters of the GEMM would compute in shared memory
            const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            Wlo[(size_t)n * D + d] = __uint_as_float((__float_as_uint(v - h) + 0x1000u) & 0xffffe000u);
            ss = fmaf(v, v, ss);
        }
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) wn[n] = sqrtf(ss);
}


// ---------------------------------------------------------------------------------------
// backward: per packed column
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float v3_lo(float x) {
    const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    return __uint_as_float((__float_as_uint(x - h) + 0x1000u) & 0xffffe000u);
}

// One CTA per (packed column n', group of V3_DU_JG = 64 images); warp w takes images j0 + 8 w ... + 7, four per round
// with all their loads issued before the math:
//   dcos = dm g2 exp(g2 cos - m);  a1 = dcos / max(|w||u|, eps);  a2 = dcos cos / |u|^2;  a3 = dcos cos / |w|^2
//   DU = a1 w - a2 u  (u = U'/Z);  DUz = DU / Z;  csz = <DU, u> / Z;  dwcos[g][n'] = sum_j a1 u - (sum_j a3) w
// (the per-warp parts of dwcos meet in shared memory and are added in a fixed order).
template <int NQ>  // float4 per lane: D = 128 * NQ
__global__ void __launch_bounds__(256) v3_du_kernel(const float* __restrict__ U, const float* __restrict__ Wp,
                                                    const float* __restrict__ wn, const float* __restrict__ Z,
                                                    const float* __restrict__ cosv, const float* __restrict__ un,
                                                    const float* __restrict__ dm, const float* __restrict__ mst,
                                                    const int* __restrict__ col_cap, const int* __restrict__ meta, int NtP,
                                                    int Bi, int Bc, int D, float g2, float* __restrict__ DUz,
                                                    float* __restrict__ DUzlo, float* __restrict__ csz,
                                                    float* __restrict__ dwcos) {
    __shared__ float4 s_acc[8][32 * NQ];
    __shared__ float s_a3[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x;
    if (n >= meta[1]) return;
    const int g = blockIdx.y, j0 = g * V3_DU_JG + warp * 8;
    const int i = col_cap[n];
    float4* dwc = reinterpret_cast<float4*>(dwcos + ((size_t)g * NtP + n) * D);
    if (i < 0) {  // padding column: zero operand rows so that GEMM4's K loop adds nothing
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + q;
            if (j >= Bi) break;
            float4* o = reinterpret_cast<float4*>(DUz + ((size_t)j * NtP + n) * D);
            float4* ol = DUzlo ? reinterpret_cast<float4*>(DUzlo + ((size_t)j * NtP + n) * D) : nullptr;
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                o[lane + 32 * c] = zero;
                if (ol) ol[lane + 32 * c] = zero;
            }
            if (lane == 0) csz[(size_t)j * NtP + n] = 0.f;
        }
        if (warp == 0) {
#pragma unroll
            for (int c = 0; c < NQ; ++c) dwc[lane + 32 * c] = zero;
        }
        return;
    }
    float4 wv[NQ], acc[NQ];
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
        wv[c] = __ldg(reinterpret_cast<const float4*>(Wp + (size_t)n * D) + lane + 32 * c);
        acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float wnv = wn[n];
    float a3s = 0.f;
#pragma unroll 1
    for (int q0 = 0; q0 < 8 && j0 + q0 < Bi; q0 += 4) {
        float zs[4], cs4[4], us[4], dms[4], ms[4];
        float4 uv[4][NQ];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = min(j0 + q0 + q, Bi - 1);
            const size_t idx = (size_t)j * NtP + n;
            zs[q] = Z[idx]; cs4[q] = cosv[idx]; us[q] = un[idx];
            dms[q] = dm[(size_t)j * Bc + i]; ms[q] = mst[(size_t)j * Bc + i];
            const float4* u4 = reinterpret_cast<const float4*>(U + idx * D);
#pragma unroll
            for (int k = 0; k < NQ; ++k) uv[q][k] = u4[lane + 32 * k];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + q0 + q;
            if (j >= Bi) break;
            const size_t idx = (size_t)j * NtP + n;
            const float c = cs4[q], unv = us[q];
            const float dcos = dms[q] * g2 * expf(g2 * c - ms[q]);
            const float nn = wnv * unv;
            const bool live = nn > 1e-8f;
            const float a1 = dcos / fmaxf(nn, 1e-8f);
            const float a2 = live ? dcos * c / (unv * unv) : 0.f;
            a3s += live ? dcos * c / (wnv * wnv) : 0.f;
            const float iz = 1.0f / zs[q];
            const float a2z = a2 * iz, a1z = a1 * iz;
            float4* o4 = reinterpret_cast<float4*>(DUz + idx * D);
            float cs = 0.f;
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                const float4 u = uv[q][k];
                float4 du;
                du.x = a1 * wv[k].x - a2z * u.x; du.y = a1 * wv[k].y - a2z * u.y;
                du.z = a1 * wv[k].z - a2z * u.z; du.w = a1 * wv[k].w - a2z * u.w;
                cs = fmaf(du.x, u.x, cs); cs = fmaf(du.y, u.y, cs); cs = fmaf(du.z, u.z, cs); cs = fmaf(du.w, u.w, cs);
                acc[k].x = fmaf(a1z, u.x, acc[k].x); acc[k].y = fmaf(a1z, u.y, acc[k].y);
                acc[k].z = fmaf(a1z, u.z, acc[k].z); acc[k].w = fmaf(a1z, u.w, acc[k].w);
                du.x *= iz; du.y *= iz; du.z *= iz; du.w *= iz;
                o4[lane + 32 * k] = du;
                if (DUzlo) {  // lo = tf32(x - trunc_tf32(x)), what the B splitters of GEMM3 would compute
                    float4 l;
                    l.x = v3_lo(du.x); l.y = v3_lo(du.y); l.z = v3_lo(du.z); l.w = v3_lo(du.w);
                    reinterpret_cast<float4*>(DUzlo + idx * D)[lane + 32 * k] = l;
                }
            }
            cs = warp_sum(cs);  // <DU, U'>
            if (lane == 0) csz[idx] = cs * iz * iz;
        }
    }
#pragma unroll
    for (int k = 0; k < NQ; ++k) s_acc[warp][lane + 32 * k] = acc[k];
    if (lane == 0) s_a3[warp] = a3s;
    __syncthreads();
    for (int q = threadIdx.x; q < 32 * NQ; q += blockDim.x) {
        float4 t = s_acc[0][q];
        float a3 = s_a3[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            const float4 o = s_acc[w][q];
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
            a3 += s_a3[w];
        }
        const float4 wq = __ldg(reinterpret_cast<const float4*>(Wp + (size_t)n * D) + q);
        t.x -= a3 * wq.x; t.y -= a3 * wq.y; t.z -= a3 * wq.z; t.w -= a3 * wq.w;
        dwc[q] = t;
    }
}


// A operand staged in tensor memory (gemm_ts.cu) unless EEGAN_V3_TS=0 asks for the all-shared-memory kernel (A/B timing)
static int use_ts() {
    static int v = [] {
        const char* e = getenv("EEGAN_V3_TS");
        return e ? atoi(e) : 1;
    }();
    return v;
}

static int presplit_w() {
    static int v = [] {
        const char* e = getenv("EEGAN_V3_PRESPLIT_W");
        return e ? atoi(e) : 1;
    }();
    return v;
}

static TcAttnEpi attn_args(const V3Ws& w, float g1) {
    TcAttnEpi a{};
    a.nbins = w.meta;
    a.bin_cap = w.bin_cap;
    a.bin_used = w.bin_used;
    a.col_start = w.col_start;
    a.cap_len = w.cap_len;
    a.P = w.P;
    a.Zpart = w.Zpart;
    a.csz = w.csz;
    a.g1 = g1;
    return a;
}

size_t pair_v3_workspace_bytes(int Bi, int Bc, int D, int R, int Tm) { return v3_carve(nullptr, Bi, Bc, D, R, Tm).bytes; }

int pair_v3_fwd(const float* img, const float* words, const int32_t* cap_lens, int Bi, int Bc, int D, int R, int Tm, float g1,
                float g2, float* m, float* att, int diag_offset, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    EEGAN_REQUIRE(D % 128 == 0 && D <= 1024, "pair grid (fused engine): D=%d must be a multiple of 128 and <= 1024", D);
    EEGAN_REQUIRE(R <= 1024, "pair grid (fused engine): R=%d must be <= 1024", R);  // cos/lse sums <= 32 region partials per column
    EEGAN_REQUIRE(Bc <= 4096, "pair grid (fused engine): at most 4096 captions per call (got %d)", Bc);
    V3Ws w = v3_carve(workspace, Bi, Bc, D, R, Tm);
    if (workspace_bytes < w.bytes) {
        set_error("pair fwd: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return EEGAN_ERR_WORKSPACE;
    }
    const float* C = w.Cp ? w.Cp : img;
    const int NtP = w.NtP;

    prof_mark(-1, st);
    v3_scan_kernel<<<1, 256, 2 * Bc * sizeof(int), st>>>(cap_lens, Bc, Tm, w.maxbins, w.col_start, w.cap_len, w.bin_cap, w.bin_used, w.meta, w.col_cap);
    v3_pack_kernel<<<NtP + ((w.Cp || w.Clo) ? 148 * 8 : 0), 128, 0, st>>>(words, w.col_start, w.col_cap, w.meta, D, Tm, NtP, w.Wp,
                                                                         w.Wlo, w.wn, img, w.Cp, w.Clo, (long long)Bi * D, R, w.Rp);
    EEGAN_LAUNCH_CHECK("pair prologue");
    prof_mark(0, st);

    {  // GEMM1 + attention forward: S^T[j][r][n'] -> P^T, E^T, Zpart
        TcGemm g{};
        g.nseg = 1;
        g.ts = use_ts();
        g.A[0] = TcOperand{C, nullptr, 0, w.Rp, (long long)D * w.Rp, Bi, R, D};
        g.B[0] = TcOperand{w.Wp, presplit_w() ? w.Wlo : nullptr, 1, D, 0, 1, NtP, D};
        g.C = w.E; g.ldc = NtP; g.bC = (long long)R * NtP; g.M = R; g.N = NtP; g.dynN = w.meta + 1; g.batch = Bi; g.nred = 1;
        g.epi = TC_EPI_ATTN_FWD;
        g.attn = attn_args(w, g1);
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
    }
    prof_mark(1, st);

    {  // GEMM2: U'[j][n'][d] = sum_r E^T[j][r][n'] C[j][d][r]
        TcGemm g{};
        g.nseg = 1;
        g.ts = use_ts();
        g.A[0] = TcOperand{w.E, nullptr, 0, NtP, (long long)R * NtP, Bi, NtP, R};
        g.B[0] = TcOperand{C, w.Clo, 1, w.Rp, (long long)D * w.Rp, Bi, D, R};
        g.C = w.U; g.ldc = D; g.bC = (long long)NtP * D; g.M = NtP; g.N = D; g.dynM = w.meta + 1; g.batch = Bi; g.nred = 1;
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
    }
    prof_mark(3, st);

    v3_cos_lse_kernel<<<dim3(w.maxbins, Bi), 256, 0, st>>>(w.U, w.Wp, w.wn, w.Zpart, w.col_start, w.cap_len, w.bin_cap, w.bin_used,
                                                           w.meta, NtP, D, Bc, w.nz, g2, w.Z, w.cosv, w.un, m, w.mst);
    if (att) v3_att_diag_kernel<<<dim3(Bc, (R + 31) / 32), 256, 0, st>>>(w.E, w.Z, w.col_start, w.cap_len, NtP, R, Tm, Bi, diag_offset, att, 0, g1);
    EEGAN_LAUNCH_CHECK("pair cos/lse");
    prof_mark(4, st);
    return EEGAN_OK;
}

int pair_v3_bwd(const float* img, int Bi, int Bc, int D, int R, int Tm, float g1, float g2, const float* dm, float* d_img,
                float* d_words, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    V3Ws w = v3_carve(workspace, Bi, Bc, D, R, Tm);
    if (workspace_bytes < w.bytes) {
        set_error("pair bwd: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return EEGAN_ERR_WORKSPACE;
    }
    const float* C = w.Cp ? w.Cp : img;
    const int NtP = w.NtP;

    prof_mark(-1, st);
    {
        dim3 grid(NtP, w.ngroups);
#define V3_DU(NQ)                                                                                                            \
    v3_du_kernel<NQ><<<grid, 256, 0, st>>>(w.U, w.Wp, w.wn, w.Z, w.cosv, w.un, dm, w.mst, w.col_cap, w.meta, NtP, Bi, Bc, D, g2, \
                                           w.DUz, presplit_du() ? w.DUzlo : nullptr, w.csz, w.dwcos)
        switch (D / 128) {
            case 1: V3_DU(1); break;
            case 2: V3_DU(2); break;
            case 3: V3_DU(3); break;
            case 4: V3_DU(4); break;
            case 5: V3_DU(5); break;
            case 6: V3_DU(6); break;
            case 7: V3_DU(7); break;
            default: V3_DU(8); break;
        }
#undef V3_DU
    }
    EEGAN_LAUNCH_CHECK("pair dU");
    prof_mark(5, st);

    {  // GEMM3 + attention backward: acc = C^T DUz^T -> dS^T
        TcGemm g{};
        g.nseg = 1;
        g.ts = use_ts();
        g.A[0] = TcOperand{C, nullptr, 0, w.Rp, (long long)D * w.Rp, Bi, R, D};
        g.B[0] = TcOperand{w.DUz, presplit_du() ? w.DUzlo : nullptr, 1, D, (long long)NtP * D, Bi, NtP, D};
        g.C = w.dS; g.ldc = NtP; g.bC = (long long)R * NtP; g.M = R; g.N = NtP; g.dynN = w.meta + 1; g.batch = Bi; g.nred = 1;
        g.epi = TC_EPI_ATTN_BWD;
        g.attn = attn_args(w, g1);
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
    }
    prof_mark(6, st);

    if (d_img) {  // GEMM4: dC[j][d][r] = sum_n' DUz[j][n'][d] E^T[j][r][n'] + Wp[n'][d] dS^T[j][r][n']
        TcGemm g{};
        g.nseg = 2;
        g.ts = use_ts();
        g.A[0] = TcOperand{w.DUz, nullptr, 0, D, (long long)NtP * D, Bi, D, NtP};
        g.B[0] = TcOperand{w.E, nullptr, 1, NtP, (long long)R * NtP, Bi, R, NtP};
        g.A[1] = TcOperand{w.Wp, nullptr, 0, D, 0, 1, D, NtP};
        g.B[1] = TcOperand{w.dS, nullptr, 1, NtP, (long long)R * NtP, Bi, R, NtP};
        g.C = d_img; g.ldc = R; g.bC = (long long)D * R; g.M = D; g.N = R; g.dynK = w.meta + 1; g.batch = Bi; g.nred = 1;
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
        prof_mark(8, st);
    }
    if (d_words) {  // GEMM5: dWp[n'][d] = sum_j sum_r dS^T[j][r][n'] C[j][d][r], images split in nsplit groups
        const int nred = (Bi + w.nsplit - 1) / w.nsplit;
        TcGemm g{};
        g.nseg = 1;
        g.ts = use_ts();
        g.A[0] = TcOperand{w.dS, nullptr, 0, NtP, (long long)R * NtP, Bi, NtP, R};
        g.B[0] = TcOperand{C, w.Clo, 1, w.Rp, (long long)D * w.Rp, Bi, D, R};
        g.C = w.dWpart; g.ldc = D; g.bC = (long long)NtP * D; g.M = NtP; g.N = D; g.dynM = w.meta + 1; g.batch = w.nsplit;
        g.nred = nred; g.red_total = Bi;
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
        prof_mark(9, st);
        v3_unpack_dw_kernel<<<dim3(Bc, (D + 31) / 32), 256, 0, st>>>(w.dWpart, w.dwcos, w.col_start, w.cap_len, w.nsplit, w.ngroups,
                                                                     NtP, D, Tm, d_words);
        EEGAN_LAUNCH_CHECK("pair GEMM5");
        prof_mark(10, st);
    }
    return EEGAN_OK;
}

}  // namespace eegan
