// DAMSM pair grid on the half-pair ("3xFP16") contraction engine (contraction engine 3).
//
// Same region-major fused pipeline as pair_grid_v3.cu (miscc/DAMSM_losses.py:272-342):
// five tcgen05 contractions with the word-region attention inside the GEMM1 / GEMM3 epilogues.  What changes
// is how the GEMM operands live in HBM: every tensor a contraction reads — image features C, packed words W,
// E = exp(g1 (P - 1)), DUz, dS — is written ONCE, by the kernel that produces it, as two fp16 arrays
// (hi, lo) of x * 2^e with one power-of-two scale per tensor (gemm_h.cuh).  That is the same 4 bytes per element as
// fp32 and the same 22 mantissa bits the 3xTF32 engine used, but the GEMM main loop becomes pure TMA -> MMA at the
// fp16 rate with no operand split in the kernel.
//
// Scales (all on the device, no host sync; `scal` block of the workspace):
//   C, W   exact max |x| (h_pre_kernel partials)        -> max scaled into [2^7, 2^8)
//   E      in (0, 1]: constant 2^12
//   DUz    exact row 2-norms |dcos| sqrt(1 - cos^2) / (|u| Z) from the per-column scalars (h_duscale_kernel)
//          -> largest row norm scaled into [2^11, 2^12): elements <= 2^12
//   dS     rigorous bound |dS| <= 4 g1 max(E) max_r|C[:,r]| max|DUz row|  -> bound scaled into [2^13, 2^14)
// so nothing can overflow fp16; a loose bound only costs low-order bits of the smallest elements.
#include <stdlib.h>

#include "common.cuh"
#include "gemm_h.cuh"
#include "pair_h.cuh"
#include "pair_v3_kernels.cuh"

namespace eegan {

enum HScal {
    HS_MAXC = 0, HS_MAXW = 1, HS_MAXCN2 = 2, HS_EMAX = 3, HS_MAXDUN = 4,
    HS_SC = 8, HS_IC = 9, HS_SW = 10, HS_IW = 11, HS_SE = 12, HS_IE = 13, HS_SDU = 14, HS_IDU = 15, HS_SDS = 16, HS_IDS = 17,
    HS_COUNT = 32
};

struct HWs {
    int* col_start;
    int* cap_len;
    int* bin_cap;
    int* bin_used;
    int* meta;
    int* col_cap;
    float* scal;     // [HS_COUNT] maxima and scales
    float* pmax;     // [3][npart] per-CTA partial maxima of the prologue: |c|, column norm^2 of C, |w|
    float* pdun;     // [Bi] per-image partial maxima of the DUz row norms
    float* Wp;       // [NtP][D] packed words, fp32 (cos/lse, dU)
    float* wn;       // [NtP]
    __half* Wh;      // [NtP][D]
    __half* Wl;
    __half* Ch;      // [Bi][D][Rp]
    __half* Cl;
    float* P;        // [Bi][R][NtP]
    __half* Eh;      // [Bi][R][NtP]
    __half* El;
    float* Zpart;    // [Bi][ceil(R/32)][NtP]
    float* Z;        // [Bi][NtP]
    float* U;        // [Bi][NtP][D]
    __half* DUh;     // [Bi][NtP][D]
    __half* DUl;
    float* cosv;
    float* un;
    float* csz;
    float* mst;      // [Bi][Bc]
    __half* dSh;     // [Bi][R][NtP]
    __half* dSl;
    float* dWpart;   // [nsplit][NtP][D]
    float* dwcos;    // [ngroups][NtP][D]
    int Rp, NtP, maxbins, nsplit, ngroups, nz, npart;
    size_t bytes;
};

static HWs h_carve(void* base, int Bi, int Bc, int D, int R, int Tm) {
    HWs w;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* q = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return q;
    };
    const int per_bin = V3_BIN / Tm;
    w.maxbins = (Bc + per_bin - 1) / per_bin;
    w.NtP = w.maxbins * V3_BIN;
    w.Rp = (R + 7) / 8 * 8;  // 16-byte row pitch for the half arrays
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
    w.scal = (float*)take(HS_COUNT * sizeof(float));
    w.npart = Bi * w.nz > Bc ? Bi * w.nz : Bc;
    w.pmax = (float*)take((size_t)3 * w.npart * sizeof(float));
    w.pdun = (float*)take((size_t)Bi * w.maxbins * sizeof(float));
    w.Wp = (float*)take(NtP * D * sizeof(float));
    w.wn = (float*)take(NtP * sizeof(float));
    w.Wh = (__half*)take(NtP * D * sizeof(__half));
    w.Wl = (__half*)take(NtP * D * sizeof(__half));
    w.Ch = (__half*)take((size_t)Bi * D * w.Rp * sizeof(__half));
    w.Cl = (__half*)take((size_t)Bi * D * w.Rp * sizeof(__half));
    w.P = (float*)take((size_t)Bi * R * NtP * sizeof(float));
    w.Eh = (__half*)take((size_t)Bi * R * NtP * sizeof(__half));
    w.El = (__half*)take((size_t)Bi * R * NtP * sizeof(__half));
    w.Zpart = (float*)take((size_t)Bi * w.nz * NtP * sizeof(float));
    w.Z = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.U = (float*)take((size_t)Bi * NtP * D * sizeof(float));
    w.DUh = (__half*)take((size_t)Bi * NtP * D * sizeof(__half));
    w.DUl = (__half*)take((size_t)Bi * NtP * D * sizeof(__half));
    w.cosv = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.un = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.csz = (float*)take((size_t)Bi * NtP * sizeof(float));
    w.mst = (float*)take((size_t)Bi * Bc * sizeof(float));
    w.dSh = (__half*)take((size_t)Bi * R * NtP * sizeof(__half));
    w.dSl = (__half*)take((size_t)Bi * R * NtP * sizeof(__half));
    w.dWpart = (float*)take((size_t)w.nsplit * NtP * D * sizeof(float));
    w.dwcos = (float*)take((size_t)w.ngroups * NtP * D * sizeof(float));
    w.bytes = off;
    return w;
}

// largest power of two s with maxv * s < 2^target_exp (1 for a zero or non-finite maximum)
__device__ __forceinline__ float h_pow2_scale(float maxv, int target_exp) {
    if (!(maxv > 0.f) || !(maxv < INFINITY)) return 1.0f;
    int e;
    frexpf(maxv, &e);  // maxv = f 2^e, f in [0.5, 1)
    const int k = max(-120, min(120, target_exp - e));
    return ldexpf(1.0f, k);
}
__device__ __forceinline__ void h_atomic_max_pos(float* addr, float v) {  // v >= 0: the bit patterns order like the values
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
__device__ __forceinline__ uint32_t h_pack2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void h_split4(const float4 v, float s, uint2& hi, uint2& lo) {
    __half h0, h1, h2, h3, l0, l1, l2, l3;
    h_split(v.x * s, h0, l0);
    h_split(v.y * s, h1, l1);
    h_split(v.z * s, h2, l2);
    h_split(v.w * s, h3, l3);
    hi = make_uint2(h_pack2(h0, h1), h_pack2(h2, h3));
    lo = make_uint2(h_pack2(l0, l1), h_pack2(l2, l3));
}

// ---------------------------------------------------------------------------------------
// prologue: two launches
// ---------------------------------------------------------------------------------------
// Launch 1, three independent roles (none needs another's result):
//   CTA 0                      caption packing (v3_scan_body)
//   CTAs 1 .. Bi*nslab         one (image, 32-region slab) each: max |c| and the largest column norm |C[j][:, r]|^2
//   the last Bc CTAs           one caption each: max |w| over its live words
// The maxima leave as per-CTA partials (no atomics, nothing to zero); launch 2 reduces them.
__global__ void __launch_bounds__(256) h_pre_kernel(const int32_t* __restrict__ cap_lens, int Bc, int Tm, int maxbins,
                                                    int* __restrict__ col_start, int* __restrict__ cap_len, int* __restrict__ bin_cap,
                                                    int* __restrict__ bin_used, int* __restrict__ meta, int* __restrict__ col_cap,
                                                    const float* __restrict__ img, int Bi, int D, int R, int nslab,
                                                    const float* __restrict__ words, float* __restrict__ pmax, int npart) {
    extern __shared__ int s_pre_buf[];
    __shared__ float red[32];
    __shared__ float s_part[8][32];
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x == 0) {
        v3_scan_body(s_pre_buf, cap_lens, Bc, Tm, maxbins, col_start, cap_len, bin_cap, bin_used, meta, col_cap);
        return;
    }
    const int idx = blockIdx.x - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (idx < Bi * nslab) {
        const int j = idx / nslab, r = (idx % nslab) * 32 + lane;
        const float* src = img + (size_t)j * D * R + r;
        float ss = 0.f, am = 0.f;
        if (r < R) {
            for (int d0 = warp; d0 < D; d0 += 32) {  // four loads in flight per thread
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = d0 + 8 * q < D ? __ldg(src + (size_t)(d0 + 8 * q) * R) : 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ss = fmaf(v[q], v[q], ss);
                    am = fmaxf(am, fabsf(v[q]));
                }
            }
        }
        s_part[warp][lane] = ss;
        am = block_max(am, red);  // its __syncthreads also publishes s_part
        if (warp == 0) {
            float n2 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) n2 += s_part[q][lane];
            n2 = warp_max(n2);
            if (lane == 0) {
                pmax[idx] = am;
                pmax[npart + idx] = n2;
            }
        }
        return;
    }
    const int i = idx - Bi * nslab;
    const int T = min(max(cap_lens[i], 0), Tm);
    const float* wsrc = words + (size_t)i * D * Tm;
    float am = 0.f;
    for (int k = threadIdx.x; k < D * Tm; k += blockDim.x) {
        const float v = __ldg(wsrc + k);
        if (k % Tm < T) am = fmaxf(am, fabsf(v));
    }
    am = block_max(am, red);
    if (threadIdx.x == 0) pmax[2 * npart + i] = am;
}

// Launch 2: every CTA reduces the partial maxima to the scales of C and W (CTA 0 publishes them), then
//   CTAs [0, NtP):   one packed column each: gather the word vector (words is [i][d][t]) -> Wp (fp32), |w|, Wh / Wl
//   CTAs [NtP, ...): img rows [Bi*D][R] -> Ch, Cl [Bi*D][Rp] (pad columns zero)
__global__ void __launch_bounds__(256) h_pack_kernel(const float* __restrict__ words, const int* __restrict__ col_start,
                                                     const int* __restrict__ col_cap, const int* __restrict__ meta, int D, int Tm,
                                                     int NtP, float* __restrict__ Wp, float* __restrict__ wn, __half* __restrict__ Wh,
                                                     __half* __restrict__ Wl, const float* __restrict__ img, long long rows, int R,
                                                     int Rp, __half* __restrict__ Ch, __half* __restrict__ Cl,
                                                     const float* __restrict__ pmax, int nC, int nW, int npart,
                                                     float* __restrict__ scal) {
    __shared__ float red[32];
    pdl_trigger();
    pdl_wait();
    float mc = 0.f, mn = 0.f, mw = 0.f;
    for (int k = threadIdx.x; k < nC; k += blockDim.x) {
        mc = fmaxf(mc, pmax[k]);
        mn = fmaxf(mn, pmax[npart + k]);
    }
    for (int k = threadIdx.x; k < nW; k += blockDim.x) mw = fmaxf(mw, pmax[2 * npart + k]);
    mc = block_max(mc, red);
    mw = block_max(mw, red);
    const float sC = h_pow2_scale(mc, 8), sW = h_pow2_scale(mw, 8);
    if (blockIdx.x == 0) {
        mn = block_max(mn, red);
        if (threadIdx.x == 0) {
            scal[HS_MAXC] = mc; scal[HS_MAXW] = mw; scal[HS_MAXCN2] = mn;
            scal[HS_EMAX] = 0.f;  // GEMM1's epilogue collects max E with atomicMax
            scal[HS_SC] = sC; scal[HS_IC] = 1.0f / sC;
            scal[HS_SW] = sW; scal[HS_IW] = 1.0f / sW;
            scal[HS_SE] = H_E_SCALE; scal[HS_IE] = 1.0f / H_E_SCALE;
        }
    }
    if ((int)blockIdx.x >= NtP) {
        const int lane = threadIdx.x & 31;
        const long long warps = (long long)(gridDim.x - NtP) * (blockDim.x >> 5);
        for (long long row = (long long)(blockIdx.x - NtP) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
            const float* sp = img + row * R;
            uint32_t* dh = reinterpret_cast<uint32_t*>(Ch + row * Rp);  // Rp is even: a lane owns regions 2l, 2l + 1
            uint32_t* dl = reinterpret_cast<uint32_t*>(Cl + row * Rp);
            for (int r0 = 0; r0 < Rp; r0 += 64 * 5) {  // ten loads per lane in flight (R = 289 -> one round)
                float v[10];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int r = r0 + 64 * q + 2 * lane;
                    v[2 * q] = r < R ? __ldg(sp + r) : 0.f;
                    v[2 * q + 1] = r + 1 < R ? __ldg(sp + r + 1) : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int r = r0 + 64 * q + 2 * lane;
                    if (r < Rp) {
                        __half h0, l0, h1, l1;
                        h_split(v[2 * q] * sC, h0, l0);
                        h_split(v[2 * q + 1] * sC, h1, l1);
                        dh[r >> 1] = h_pack2(h0, h1);
                        dl[r >> 1] = h_pack2(l0, l1);
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
        const __half zero = __float2half_rn(0.f);
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            Wp[(size_t)n * D + d] = 0.f;
            Wh[(size_t)n * D + d] = zero;
            Wl[(size_t)n * D + d] = zero;
        }
    } else {
        const int t = n - col_start[i];
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const float v = __ldg(words + ((size_t)i * D + d) * Tm + t);
            Wp[(size_t)n * D + d] = v;
            __half h, l;
            h_split(v * sW, h, l);
            Wh[(size_t)n * D + d] = h;
            Wl[(size_t)n * D + d] = l;
            ss = fmaf(v, v, ss);
        }
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) wn[n] = sqrtf(ss);
}

// ---------------------------------------------------------------------------------------
// backward: per packed column
// ---------------------------------------------------------------------------------------
// Largest row 2-norm of DUz over all (image j, packed column n'), from the per-column scalars alone:
//   du = dcos (w / (|w||u|) - cos u / |u|^2)  =>  |du| = |dcos| sqrt(1 - cos^2) / |u|,  DUz = du / Z
// (clamped columns, |w||u| <= 1e-8: du = dcos w / 1e-8).  The 1e-6 floor under 1 - cos^2 keeps rounding noise of
// near-parallel columns inside the bound.  One CTA per image walks that image's live columns and leaves ONE partial
// maximum (pdun[j]): the dU kernel's CTAs each reduce the partials again, so their number must grow with B, not B^2
// (with one partial per (image, bin) every dU CTA re-read B * bins floats: 350 KB at B = 512, 13 ms of the step).
__global__ void __launch_bounds__(256) h_duscale_kernel(const float* __restrict__ wn, const float* __restrict__ Z,
                                                        const float* __restrict__ cosv, const float* __restrict__ un,
                                                        const float* __restrict__ dm, const float* __restrict__ mst,
                                                        const int* __restrict__ col_cap, const int* __restrict__ meta, int NtP, int Bc,
                                                        float g2, float* __restrict__ pdun) {
    __shared__ float red[32];
    const int j = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    float nz = 0.f;
    const int nlive = meta[1];
    for (int n = threadIdx.x; n < nlive; n += blockDim.x) {
        const int i = col_cap[n];
        if (i < 0) continue;
        const size_t idx = (size_t)j * NtP + n;
        const float c = cosv[idx], unv = un[idx];
        const float dcos = fabsf(dm[(size_t)j * Bc + i] * g2 * expf(g2 * c - mst[(size_t)j * Bc + i]));
        const float nn = wn[n] * unv;
        const float nrm = nn > 1e-8f ? dcos * sqrtf(fmaxf(1.0f - c * c, 1e-6f)) / unv : dcos * 1e8f * wn[n];
        float v = nrm / Z[idx];
        if (!(v >= 0.f)) v = INFINITY;  // NaN upstream: poison the maximum so that the scale falls back to 1
        nz = fmaxf(nz, v);
    }
    nz = block_max(nz, red);
    if (threadIdx.x == 0) pdun[j] = nz;  // every CTA writes: nothing to zero beforehand
}

// One CTA per (packed column n', group of 64 images): as v3_du_kernel (pair_grid_v3.cu), with DUz written as half pairs.
//   dcos = dm g2 exp(g2 cos - m);  a1 = dcos / max(|w||u|, eps);  a2 = dcos cos / |u|^2;  a3 = dcos cos / |w|^2
//   DU = a1 w - a2 u  (u = U'/Z);  DUz = DU / Z;  csz = <DU, u> / Z;  dwcos[g][n'] = sum_j a1 u - (sum_j a3) w
template <int NQ>  // float4 per lane: D = 128 * NQ
__global__ void __launch_bounds__(256) h_du_kernel(const float* __restrict__ U, const float* __restrict__ Wp,
                                                   const float* __restrict__ wn, const float* __restrict__ Z,
                                                   const float* __restrict__ cosv, const float* __restrict__ un,
                                                   const float* __restrict__ dm, const float* __restrict__ mst,
                                                   const int* __restrict__ col_cap, const int* __restrict__ meta, int NtP, int Bi,
                                                   int Bc, int D, float g1, float g2, __half* __restrict__ DUh,
                                                   __half* __restrict__ DUl, float* __restrict__ csz, float* __restrict__ dwcos,
                                                   const float* __restrict__ pdun, int npdun, float* __restrict__ scal) {
    __shared__ float4 s_acc[8][32 * NQ];
    __shared__ float s_a3[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    if (n >= meta[1]) return;
    __shared__ float red[32];
    float maxdun = 0.f;
    for (int k = threadIdx.x; k < npdun; k += blockDim.x) maxdun = fmaxf(maxdun, pdun[k]);
    maxdun = block_max(maxdun, red);
    const float sDU = h_pow2_scale(maxdun, 12);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        scal[HS_MAXDUN] = maxdun;
        scal[HS_SDU] = sDU; scal[HS_IDU] = 1.0f / sDU;
        const float bound = 4.0f * g1 * fmaxf(scal[HS_EMAX], 1e-30f) * sqrtf(scal[HS_MAXCN2]) * maxdun;
        const float sDS = h_pow2_scale(bound, 14);
        scal[HS_SDS] = sDS; scal[HS_IDS] = 1.0f / sDS;
    }
    // warp w takes images g*64 + w, w + 8, w + 16, ... (strided, so that all 8 warps share the work at any Bi), four per round
    const int g = blockIdx.y, j0 = g * V3_DU_JG + warp;
    const int i = col_cap[n];
    float4* dwc = reinterpret_cast<float4*>(dwcos + ((size_t)g * NtP + n) * D);
    if (i < 0) {  // padding column: zero operand rows so that GEMM4's K loop adds nothing
        const uint2 zero2 = make_uint2(0u, 0u);
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + 8 * q;
            if (j >= Bi) break;
            uint2* oh = reinterpret_cast<uint2*>(DUh + ((size_t)j * NtP + n) * D);
            uint2* ol = reinterpret_cast<uint2*>(DUl + ((size_t)j * NtP + n) * D);
#pragma unroll
            for (int c = 0; c < NQ; ++c) {
                oh[lane + 32 * c] = zero2;
                ol[lane + 32 * c] = zero2;
            }
            if (lane == 0) csz[(size_t)j * NtP + n] = 0.f;
        }
        if (warp == 0) {
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
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
    for (int q0 = 0; q0 < 8 && j0 + 8 * q0 < Bi; q0 += 4) {
        float zs[4], cs4[4], us[4], dms[4], ms[4];
        float4 uv[4][NQ];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = min(j0 + 8 * (q0 + q), Bi - 1);
            const size_t idx = (size_t)j * NtP + n;
            zs[q] = Z[idx]; cs4[q] = cosv[idx]; us[q] = un[idx];
            dms[q] = dm[(size_t)j * Bc + i]; ms[q] = mst[(size_t)j * Bc + i];
            const float4* u4 = reinterpret_cast<const float4*>(U + idx * D);
#pragma unroll
            for (int k = 0; k < NQ; ++k) uv[q][k] = u4[lane + 32 * k];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + 8 * (q0 + q);
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
            const float izs = iz * sDU;
            uint2* oh = reinterpret_cast<uint2*>(DUh + idx * D);
            uint2* ol = reinterpret_cast<uint2*>(DUl + idx * D);
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
                uint2 hi, lo;
                h_split4(du, izs, hi, lo);
                oh[lane + 32 * k] = hi;
                ol[lane + 32 * k] = lo;
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

// ---------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------
// EEGAN_FUSED_FWD=1 (read once) selects the one-launch forward (hf_fwd_kernel, gemm_h.cu: S + attention + U' per (image, word
// tile)).  Parity-green (the whole GPU suite passes on it) and 35 MB less DRAM traffic per step at B = 48, but not faster yet
// (68.9 us against 39.4 + 30.6 us for the two launches; profiles/README.md "r2 fused forward" has the ablation): off by default.
static bool h_fused_fwd_on() {
    static const bool on = [] { const char* e = getenv("EEGAN_FUSED_FWD"); return e ? atoi(e) != 0 : false; }();
    return on;
}

static HAttnEpi h_attn_args(const HWs& w, float g1) {
    HAttnEpi a{};
    a.base.nbins = w.meta;
    a.base.bin_cap = w.bin_cap;
    a.base.bin_used = w.bin_used;
    a.base.col_start = w.col_start;
    a.base.cap_len = w.cap_len;
    a.base.P = w.P;
    a.base.Zpart = w.Zpart;
    a.base.csz = w.csz;
    a.base.g1 = g1;
    return a;
}

size_t pair_h_workspace_bytes(int Bi, int Bc, int D, int R, int Tm) { return h_carve(nullptr, Bi, Bc, D, R, Tm).bytes; }

int pair_h_fwd(const float* img, const float* words, const int32_t* cap_lens, int Bi, int Bc, int D, int R, int Tm, float g1,
               float g2, float* m, float* att, int diag_offset, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    EEGAN_REQUIRE(D % 128 == 0 && D <= 1024, "pair grid (half-pair engine): D=%d must be a multiple of 128 and <= 1024", D);
    EEGAN_REQUIRE(R <= 1024, "pair grid (half-pair engine): R=%d must be <= 1024", R);
    EEGAN_REQUIRE(Bc <= 4096, "pair grid (half-pair engine): at most 4096 captions per call (got %d)", Bc);
    HWs w = h_carve(workspace, Bi, Bc, D, R, Tm);
    if (workspace_bytes < w.bytes) {
        set_error("pair fwd: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return EEGAN_ERR_WORKSPACE;
    }
    const int NtP = w.NtP;
    const int nslab = (R + 31) / 32;

    prof_mark(-1, st);
    launch_pdl(h_pre_kernel, dim3(1 + Bi * nslab + Bc), dim3(256), 2 * Bc * sizeof(int), st, cap_lens, Bc, Tm, w.maxbins, w.col_start,
               w.cap_len, w.bin_cap, w.bin_used, w.meta, w.col_cap, img, Bi, D, R, nslab, words, w.pmax, w.npart);
    launch_pdl(h_pack_kernel, dim3(NtP + 148 * 4), dim3(256), 0, st, words, (const int*)w.col_start, (const int*)w.col_cap,
               (const int*)w.meta, D, Tm, NtP, w.Wp, w.wn, w.Wh, w.Wl, img, (long long)Bi * D, R, w.Rp, w.Ch, w.Cl,
               (const float*)w.pmax, Bi * nslab, Bc, w.npart, w.scal);
    EEGAN_LAUNCH_CHECK("pair prologue");
    prof_mark(0, st);

    const HOperand opC_mn{w.Ch, w.Cl, w.Rp, (long long)D * w.Rp, Bi, R, D, w.scal + HS_IC};  // A: rows = regions, K = channels
    const HOperand opC_k{w.Ch, w.Cl, w.Rp, (long long)D * w.Rp, Bi, D, R, w.scal + HS_IC};   // B: rows = channels, K = regions
    if (h_fused_fwd_on() && D == 256) {  // one launch: S + attention forward + U' per (image, word tile)
        HFusedFwd f{};
        f.Ch = w.Ch; f.Cl = w.Cl; f.Wh = w.Wh; f.Wl = w.Wl;
        f.Bi = Bi; f.D = D; f.R = R; f.Rp = w.Rp; f.NtP = NtP;
        f.nlive = w.meta + 1;
        f.inv_c = w.scal + HS_IC; f.inv_w = w.scal + HS_IW; f.inv_e = w.scal + HS_IE;
        f.U = w.U;
        f.attn = h_attn_args(w, g1);
        f.attn.out_hi = w.Eh; f.attn.out_lo = w.El; f.attn.emax = w.scal + HS_EMAX;
        int rc = h_fused_fwd_launch(f, st);
        if (rc) return rc;
        prof_mark(1, st);
    } else {
    {  // GEMM1 + attention forward: S^T[j][r][n'] -> P^T, E^T (half pairs), Zpart
        HGemm g{};
        g.nseg = 1;
        g.A[0] = opC_mn;
        g.B[0] = HOperand{w.Wh, w.Wl, D, 0, 1, NtP, D, w.scal + HS_IW};
        g.ldc = NtP; g.bC = (long long)R * NtP; g.M = R; g.N = NtP; g.dynN = w.meta + 1; g.batch = Bi; g.nred = 1;
        g.epi = TC_EPI_ATTN_FWD;
        g.attn = h_attn_args(w, g1);
        g.attn.out_hi = w.Eh; g.attn.out_lo = w.El; g.attn.emax = w.scal + HS_EMAX;
        int rc = h_gemm_launch(g, st);
        if (rc) return rc;
    }
    prof_mark(1, st);

    {  // GEMM2: U'[j][n'][d] = sum_r E^T[j][r][n'] C[j][d][r]
        HGemm g{};
        g.nseg = 1;
        g.A[0] = HOperand{w.Eh, w.El, NtP, (long long)R * NtP, Bi, NtP, R, w.scal + HS_IE};
        g.B[0] = opC_k;
        g.C = w.U; g.ldc = D; g.bC = (long long)NtP * D; g.M = NtP; g.N = D; g.dynM = w.meta + 1; g.batch = Bi; g.nred = 1;
        g.epi = TC_EPI_PLAIN;
        int rc = h_gemm_launch(g, st);
        if (rc) return rc;
    }
    prof_mark(3, st);
    }

    launch_pdl(v3_cos_lse_kernel, dim3(w.maxbins, Bi), dim3(256), 0, st, (const float*)w.U, (const float*)w.Wp, (const float*)w.wn,
               (const float*)w.Zpart, (const int*)w.col_start, (const int*)w.cap_len, (const int*)w.bin_cap, (const int*)w.bin_used,
               (const int*)w.meta, NtP, D, Bc, w.nz, g2, w.Z, w.cosv, w.un, m, w.mst);
    if (att)
        launch_pdl(v3_att_diag_kernel, dim3(Bc, (R + 31) / 32), dim3(256), 0, st, (const float*)w.P, (const float*)w.Z,
                   (const int*)w.col_start, (const int*)w.cap_len, NtP, R, Tm, Bi, diag_offset, att, 1, g1);
    EEGAN_LAUNCH_CHECK("pair cos/lse");
    prof_mark(4, st);
    return EEGAN_OK;
}

int pair_h_bwd(const float* img, int Bi, int Bc, int D, int R, int Tm, float g1, float g2, const float* dm, float* d_img,
               float* d_words, void* workspace, size_t workspace_bytes, cudaStream_t st, int phases) {
    (void)img;
    HWs w = h_carve(workspace, Bi, Bc, D, R, Tm);
    if (workspace_bytes < w.bytes) {
        set_error("pair bwd: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
        return EEGAN_ERR_WORKSPACE;
    }
    const int NtP = w.NtP;

    prof_mark(-1, st);
    const HOperand opC_mn{w.Ch, w.Cl, w.Rp, (long long)D * w.Rp, Bi, R, D, w.scal + HS_IC};
    const HOperand opC_k{w.Ch, w.Cl, w.Rp, (long long)D * w.Rp, Bi, D, R, w.scal + HS_IC};
    if (phases & 1) {
    launch_pdl(h_duscale_kernel, dim3(Bi), dim3(256), 0, st, (const float*)w.wn, (const float*)w.Z, (const float*)w.cosv,
               (const float*)w.un, dm, (const float*)w.mst, (const int*)w.col_cap, (const int*)w.meta, NtP, Bc, g2, w.pdun);
    {
        dim3 grid(NtP, w.ngroups);
#define H_DU(NQ)                                                                                                                  \
    launch_pdl(h_du_kernel<NQ>, grid, dim3(256), 0, st, (const float*)w.U, (const float*)w.Wp, (const float*)w.wn, (const float*)w.Z, \
               (const float*)w.cosv, (const float*)w.un, dm, (const float*)w.mst, (const int*)w.col_cap, (const int*)w.meta, NtP, Bi, \
               Bc, D, g1, g2, w.DUh, w.DUl, w.csz, w.dwcos, (const float*)w.pdun, Bi, w.scal)
        switch (D / 128) {
            case 1: H_DU(1); break;
            case 2: H_DU(2); break;
            case 3: H_DU(3); break;
            case 4: H_DU(4); break;
            case 5: H_DU(5); break;
            case 6: H_DU(6); break;
            case 7: H_DU(7); break;
            default: H_DU(8); break;
        }
#undef H_DU
    }
    EEGAN_LAUNCH_CHECK("pair dU");
    prof_mark(5, st);

    {  // GEMM3 + attention backward: acc = C^T DUz^T = dA / Z -> dS^T (half pairs)
        HGemm g{};
        g.nseg = 1;
        g.A[0] = opC_mn;
        g.B[0] = HOperand{w.DUh, w.DUl, D, (long long)NtP * D, Bi, NtP, D, w.scal + HS_IDU};
        g.ldc = NtP; g.bC = (long long)R * NtP; g.M = R; g.N = NtP; g.dynN = w.meta + 1; g.batch = Bi; g.nred = 1;
        g.epi = TC_EPI_ATTN_BWD;
        g.attn = h_attn_args(w, g1);
        g.attn.out_hi = w.dSh; g.attn.out_lo = w.dSl; g.attn.out_scale = w.scal + HS_SDS;
        int rc = h_gemm_launch(g, st);
        if (rc) return rc;
    }
    prof_mark(6, st);

    if (d_img) {  // GEMM4: dC[j][d][r] = sum_n' DUz[j][n'][d] E^T[j][r][n'] + Wp[n'][d] dS^T[j][r][n']  (one accumulator per term)
        HGemm g{};
        g.nseg = 2;
        g.A[0] = HOperand{w.DUh, w.DUl, D, (long long)NtP * D, Bi, D, NtP, w.scal + HS_IDU};
        g.B[0] = HOperand{w.Eh, w.El, NtP, (long long)R * NtP, Bi, R, NtP, w.scal + HS_IE};
        g.A[1] = HOperand{w.Wh, w.Wl, D, 0, 1, D, NtP, w.scal + HS_IW};
        g.B[1] = HOperand{w.dSh, w.dSl, NtP, (long long)R * NtP, Bi, R, NtP, w.scal + HS_IDS};
        g.C = d_img; g.ldc = R; g.bC = (long long)D * R; g.M = D; g.N = R; g.dynK = w.meta + 1; g.batch = Bi; g.nred = 1;
        g.epi = TC_EPI_PLAIN;
        int rc = h_gemm_launch(g, st);
        if (rc) return rc;
        prof_mark(8, st);
    }
    }  // phases & 1
    if (d_words && (phases & 2)) {  // GEMM5: dWp[n'][d] = sum_j sum_r dS^T[j][r][n'] C[j][d][r], images split in nsplit groups
        const int nred = (Bi + w.nsplit - 1) / w.nsplit;
        HGemm g{};
        g.nseg = 1;
        g.A[0] = HOperand{w.dSh, w.dSl, NtP, (long long)R * NtP, Bi, NtP, R, w.scal + HS_IDS};
        g.B[0] = opC_k;
        g.C = w.dWpart; g.ldc = D; g.bC = (long long)NtP * D; g.M = NtP; g.N = D; g.dynM = w.meta + 1; g.batch = w.nsplit;
        g.nred = nred; g.red_total = Bi;
        g.epi = TC_EPI_PLAIN;
        int rc = h_gemm_launch(g, st);
        if (rc) return rc;
        prof_mark(9, st);
        launch_pdl(v3_unpack_dw_kernel, dim3(Bc, (D + 31) / 32), dim3(256), 0, st, (const float*)w.dWpart, (const float*)w.dwcos,
                   (const int*)w.col_start, (const int*)w.cap_len, w.nsplit, w.ngroups, NtP, D, Tm, d_words);
        EEGAN_LAUNCH_CHECK("pair GEMM5");
        prof_mark(10, st);
    }
    return EEGAN_OK;
}

}  // namespace eegan
