// tcgen05 3xFP16 GEMM on pre-split half operands — see gemm_h.cuh for the design.
#include <stdlib.h>

#include "gemm_h.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"

// Work-skipping timing switches (EEGAN_H_DBG) exist only in builds made with -DEEGAN_DEBUG_SWITCHES; the shipped
// library compiles them out (the epilogue code below sees the constant 0) and reads no environment variable per launch.
#ifdef EEGAN_DEBUG_SWITCHES
#define H_DBG(p) ((p).dbg)
#else
#define H_DBG(p) 0
#endif

namespace eegan {

// staging row pitch of the attention epilogues (floats): even, so that the read-out takes a lane's two adjacent
// columns with one 8-byte load; the thread-per-row writes of the caption phase pay a 2-way bank conflict for it
constexpr int H_EPI_PITCH = 66;

struct HMaps {
    CUtensorMap m[2][2][2];  // [segment][0 = A, 1 = B][0 = hi, 1 = lo]
};

struct HArgs {
    float* C;
    long long ldc, bC;
    int M, N;
    const int* dynM;
    const int* dynN;
    const int* dynK;
    int K[2];
    int nseg, nred, red_total, batch;
    int a_batched[2], b_batched[2];
    const float* inv_a[2];
    const float* inv_b[2];
    HAttnEpi attn;
    int dbg;  // timing experiments only (-DEEGAN_DEBUG_SWITCHES builds, EEGAN_H_DBG): 1 read-out without global stores, 2 no read-out, 4 no caption phase
};

template <int EPI, bool DUAL, int BN = H_BN, int NSO = 0>
struct HCfg {
    static constexpr bool kAttn = EPI != TC_EPI_PLAIN;
    static constexpr int kBTile = BN * H_BK * 2;                      // one of B hi / lo
    static constexpr int kStageBytes = 2 * H_A_TILE + 2 * kBTile;     // A_hi A_lo B_hi B_lo
    static constexpr int kEWarps = 8;  // epilogue warps: two per TMEM lane quarter, one 64-column half of the tile each
    static constexpr int kStages = NSO ? NSO : (kAttn ? 4 : (BN > 128 ? 3 : 5));  // NSO: stage-count probe (EEGAN_H_STAGES)
    static constexpr int kThreads = 32 * (2 + kEWarps);
    static constexpr int kEpiPitch = kAttn ? H_EPI_PITCH : 33;
    static constexpr int kEpiWarpBytes = 32 * kEpiPitch * 4;
    static constexpr int kSmem = kStages * kStageBytes + kEWarps * (kEpiWarpBytes + 256) + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int kAccStride = DUAL ? 2 * BN : BN;  // TMEM columns between the two accumulator buffers
};
static_assert(HCfg<TC_EPI_PLAIN, false>::kSmem <= 232448 && HCfg<TC_EPI_ATTN_FWD, false>::kSmem <= 232448 &&
                  HCfg<TC_EPI_PLAIN, false, 256>::kSmem <= 232448, "shared memory budget");

__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void h_mma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory descriptors (version 1 = Blackwell) of the two canonical layouts the TMA boxes land in:
//   A, MN-major, SWIZZLE_128B: box = [32 k][64 rows] halves = 32 rows of 128 B; 8-k groups 1024 B apart (SBO),
//      the second 64-row chunk of the 128-row tile is the next box, 4096 B on (LBO); a K-step of 16 = 2048 B.
//   B, K-major, SWIZZLE_64B: box = [128 rows][32 k] halves = rows of 64 B; 8-row groups 512 B apart (SBO);
//      a K-step of 16 halves advances the start address by 32 B inside the swizzle row.
__device__ __forceinline__ uint64_t h_desc_a(uint32_t tile, int kstep) {
    const uint32_t start = tile + (uint32_t)kstep * 2048u;
    uint64_t d = (uint64_t)((start & 0x3FFFF) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint64_t h_desc_b(uint32_t tile, int kstep) {
    const uint32_t start = tile + (uint32_t)kstep * 32u;
    uint64_t d = (uint64_t)((start & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;  // SWIZZLE_64B
    return d;
}

// ---------------------------------------------------------------------------------------
// epilogues (one call = one output tile for one epilogue warp); same structure as tc_device.cuh,
// plus the power-of-two descale of the accumulator and half-pair outputs
// ---------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void h_attn_fwd_caption(const uint32_t (&vr)[32], int T, uint32_t cap_row, float inv) {
    float x[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) x[q] = (q < NV - 3 || q < T) ? __uint_as_float(vr[q]) * inv : -INFINITY;
    float m4[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
    for (int q = 4; q < NV; ++q) m4[q & 3] = fmaxf(m4[q & 3], x[q]);
    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        x[q] = fast_exp(x[q] - mx);
        s4[q & 3] += x[q];
    }
    const float invs = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (q < NV - 3 || q < T) sts_f32(cap_row + (uint32_t)q * 4u, x[q] * invs);
}

// dS = v - P sum_t v,  v = g1 P E (acc - csz),  E = exp(g1 (P - 1))
template <int NV>
__device__ __forceinline__ void h_attn_bwd_caption(const uint32_t (&vr)[32], int T, uint32_t cap_row, uint32_t cz_row, float g1,
                                                   float inv) {
    float pv[NV], vv[NV];
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const bool ok = (q < NV - 3 || q < T);
        pv[q] = ok ? lds_f32(cap_row + (uint32_t)q * 4u) : 0.f;
        const float cz = lds_f32(cz_row + (uint32_t)q * 4u);
        const float ev = fast_exp(g1 * (pv[q] - 1.0f));
        const float t = g1 * pv[q] * ev * (__uint_as_float(vr[q]) * inv - cz);
        vv[q] = ok ? t : 0.f;
        s4[q & 3] += vv[q];
    }
    const float qs = (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (q < NV - 3 || q < T) sts_f32(cap_row + (uint32_t)q * 4u, vv[q] - pv[q] * qs);
}

template <int EPI>
__device__ __forceinline__ void h_attn_caption(uint32_t taddr, int T, uint32_t cap_row, uint32_t cz_row, float g1, float inv) {
    uint32_t vr[32];
    const int bucket = (T + 3) >> 2;  // 1..8, warp-uniform
    if (bucket <= 2) tmem_ld8(taddr, vr);
    else if (bucket <= 4) tmem_ld16(taddr, vr);
    else tmem_ld32(taddr, vr);
    tmem_ld_wait();
#define H_CAP(NV)                                                                   \
    case (NV) / 4:                                                                  \
        if (EPI == TC_EPI_ATTN_FWD) h_attn_fwd_caption<NV>(vr, T, cap_row, inv);    \
        else h_attn_bwd_caption<NV>(vr, T, cap_row, cz_row, g1, inv);               \
        break;
    switch (bucket) {
        H_CAP(4) H_CAP(8) H_CAP(12) H_CAP(16) H_CAP(20) H_CAP(24) H_CAP(28) H_CAP(32)
        default: break;
    }
#undef H_CAP
}

template <int EPI, bool DUAL, int BN>
__device__ __forceinline__ void h_epilogue_tile(const HArgs& p, const EpiTile& t, int lane, float inv0, float inv1) {
    using Cfg = HCfg<EPI, DUAL, BN>;  // (the epilogue does not depend on the stage count)
    float* Cz = p.C + (long long)t.z * p.bC;
    const int row0 = t.m0 + t.quarter * 32;
    const int rows_live = max(0, min(32, t.Mlive - row0));
    if (EPI == TC_EPI_PLAIN) {
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_after();
        const int c_lo = t.half * (BN / 64), c_hi = c_lo + BN / 64;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            float v[32];
            if (t.total > 0) {
                uint32_t r0[32];
                tmem_ld32(t.tacc + (uint32_t)(c * 32), r0);
                if (DUAL) {
                    uint32_t r1[32];
                    tmem_ld32(t.tacc + (uint32_t)(BN + c * 32), r1);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = fmaf(__uint_as_float(r1[q]), inv1, __uint_as_float(r0[q]) * inv0);
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r0[q]) * inv0;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0.f;
            }
            if (c == c_hi - 1) {  // this warp's share of the accumulator is read: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t.empty_bar);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 32; ++q) sts_f32(t.stage + (uint32_t)(lane * 33 + q) * 4u, v[q]);
            __syncwarp();
            const int gn = t.n0 + c * 32 + lane;
            if (gn < t.Nlive) {
                float* dst = Cz + (long long)row0 * p.ldc + gn;
#pragma unroll 1
                for (int r8 = 0; r8 < rows_live; r8 += 8) {
                    float a[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) a[q] = lds_f32(t.stage + (uint32_t)((r8 + q) * 33 + lane) * 4u);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (r8 + q < rows_live) dst[(long long)(r8 + q) * p.ldc] = a[q];
                }
            }
        }
        return;
    }
    // ---- word-region attention on the accumulator: thread = region (TMEM lane), columns = packed words ----
    const TcAttnEpi& e = p.attn.base;
    const int nbins = *e.nbins;
    const uint32_t my_row = t.stage + (uint32_t)(lane * H_EPI_PITCH) * 4u;
    float* Pz = e.P + (long long)t.z * p.bC;
    __half* Hz = p.attn.out_hi + (long long)t.z * p.bC;
    __half* Lz = p.attn.out_lo + (long long)t.z * p.bC;
    const float oscale = EPI == TC_EPI_ATTN_FWD ? H_E_SCALE : *p.attn.out_scale;
    const int b0 = t.n0 >> 6;
    const int bc_l = e.bin_cap[min(b0 + lane, nbins)];
    const int bu_l = (lane < 2 && b0 + lane < nbins) ? e.bin_used[b0 + lane] : 0;
    // P rows of this warp's (32 regions x 64 columns) block -> staging, as 8-byte asynchronous copies (global -> shared without
    // a register stop-over): all 32 rows are in flight at once and nothing waits until cp.async.wait_all in front of the caption
    // phase.  (The register form — 8 rows of loads, then their shared stores — stalled the epilogue on the L2 latency four
    // times per tile: 17 % of this kernel's stall samples sat on those stores, profiles/README.md "r2".)
    auto load_p_bin = [&](int col0) {
        const float* src = Pz + (long long)row0 * p.ldc + col0 + 2 * lane;
        const uint32_t dst = t.stage + (uint32_t)(2 * lane) * 4u;
#pragma unroll 4
        for (int r = 0; r < rows_live; ++r)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (uint32_t)(r * H_EPI_PITCH) * 4u), "l"(src + (long long)r * p.ldc) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncwarp();
    const int h = t.half;  // the 64-column bin of the tile this warp handles
    const int b = b0 + h;
    if (b >= nbins) {  // no live bin for this warp in the tile: only keep the accumulator hand-shake going
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t.empty_bar);
        return;
    }
    const int i0 = __shfl_sync(0xffffffffu, bc_l, h), i1 = __shfl_sync(0xffffffffu, bc_l, h + 1);
    const int used = __shfl_sync(0xffffffffu, bu_l, h);
    const int col0 = t.n0 + 64 * h;
    if (EPI == TC_EPI_ATTN_BWD) {
        load_p_bin(col0);
        sts_f32(t.czs + (uint32_t)lane * 4u, __ldg(e.csz + (long long)t.z * p.ldc + col0 + lane));
        sts_f32(t.czs + (uint32_t)(lane + 32) * 4u, __ldg(e.csz + (long long)t.z * p.ldc + col0 + lane + 32));
    }
    __syncwarp();
    bool waited = false;
#pragma unroll 1
    for (int ic = i0; ic < i1; ic += 32) {
        const int nc = min(32, i1 - ic);
        const int myT = lane < nc ? e.cap_len[ic + lane] : 0;
        const int myC = lane < nc ? e.col_start[ic + lane] - col0 : 0;
        if (!waited) {
            mbar_wait(t.full_bar, t.full_parity);
            tc_fence_after();
            if (EPI == TC_EPI_ATTN_BWD) {  // the P rows copied asynchronously into staging have landed, for every lane of the warp
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
            }
            waited = true;
        }
#pragma unroll 1
        for (int ci = 0; ci < nc; ++ci) {
            const int T = __shfl_sync(0xffffffffu, myT, ci);
            if (T <= 0) continue;
            const int cl = __shfl_sync(0xffffffffu, myC, ci);
            if (!(H_DBG(p) & 4)) h_attn_caption<EPI>(t.tacc + (uint32_t)(64 * h + cl), T, my_row + (uint32_t)cl * 4u, t.czs + (uint32_t)cl * 4u, e.g1, inv0);
        }
    }
    if (!waited) {
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_after();
        if (EPI == TC_EPI_ATTN_BWD) asm volatile("cp.async.wait_all;" ::: "memory");
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t.empty_bar);
    for (int c = used; c < 64; ++c) sts_f32(my_row + (uint32_t)c * 4u, 0.f);  // padding columns of the bin
    __syncwarp();
    // read-out: region rows, coalesced; a lane owns the adjacent columns 2 lane, 2 lane + 1 of the bin: one 8-byte P store and
    // one 4-byte store per half array and row.  Padding columns of the bin hold P = 0, hence finite E' / dS' = 0 — every
    // consumer multiplies them by zero rows (GEMM4) or never reads the outputs they feed (GEMM2 / GEMM5 rows, Zpart).
    if (H_DBG(p) & 2) return;
    const bool nostore = H_DBG(p) & 1;
    const uint32_t rd0 = t.stage + (uint32_t)(2 * lane) * 4u;
    const long long o0 = (long long)row0 * p.ldc + col0 + 2 * lane;
    uint32_t* hp = reinterpret_cast<uint32_t*>(Hz + o0);
    uint32_t* lp = reinterpret_cast<uint32_t*>(Lz + o0);
    const long long hstep = p.ldc >> 1;  // row pitch in 4-byte units of the half arrays
    const int full = rows_live & ~3;
    if (EPI == TC_EPI_ATTN_FWD) {
        // E' = 2^12 exp(g1 (P - 1)) = ex2(P a + b)
        const float a = e.g1 * 1.4426950408889634f, bb = 12.0f - a;
        float2* pp = reinterpret_cast<float2*>(Pz + o0);
        const long long pstep = p.ldc >> 1;
        float z0 = 0.f, z1 = 0.f, em = 0.f;
        auto row_out = [&](float2 pv) {
            float e0, e1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(pv.x, a, bb)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(pv.y, a, bb)));
            const __half2 h = __floats2half2_rn(e0, e1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(e0 - hf.x, e1 - hf.y);
            if (!nostore) {
                *pp = pv;
                *hp = *reinterpret_cast<const uint32_t*>(&h);
                *lp = *reinterpret_cast<const uint32_t*>(&l);
            } else {
                z0 += hf.x + __half22float2(l).x;
            }
            pp += pstep; hp += hstep; lp += hstep;
            z0 += e0;
            z1 += e1;
            em = fmaxf(em, fmaxf(e0, e1));
        };
#pragma unroll 1
        for (int r4 = 0; r4 < full; r4 += 4) {
            float2 pv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) pv[q] = lds_f32x2(rd0 + (uint32_t)((r4 + q) * H_EPI_PITCH) * 4u);
#pragma unroll
            for (int q = 0; q < 4; ++q) row_out(pv[q]);
        }
#pragma unroll 1
        for (int r = full; r < rows_live; ++r) row_out(lds_f32x2(rd0 + (uint32_t)(r * H_EPI_PITCH) * 4u));
        if (rows_live > 0) {
            float* zp = e.Zpart + ((long long)t.z * ((p.M + 31) / 32) + (row0 >> 5)) * p.ldc + col0 + 2 * lane;
            *reinterpret_cast<float2*>(zp) = make_float2(z0 * (1.0f / H_E_SCALE), z1 * (1.0f / H_E_SCALE));
        }
        em = warp_max(em) * (1.0f / H_E_SCALE);
        if (lane == 0 && em > 0.f) atomicMax(reinterpret_cast<int*>(p.attn.emax), __float_as_int(em));
    } else {
        auto row_out = [&](float2 dv) {
            const float x0 = dv.x * oscale, x1 = dv.y * oscale;  // |dS'| < 2^14 by construction of the scale (pair_grid_h.cu)
            const __half2 h = __floats2half2_rn(x0, x1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            if (!nostore) {
                *hp = *reinterpret_cast<const uint32_t*>(&h);
                *lp = *reinterpret_cast<const uint32_t*>(&l);
            } else if (hf.x + __half22float2(l).y == 123.456f) {
                *hp = 0u;
            }
            hp += hstep; lp += hstep;
        };
#pragma unroll 1
        for (int r4 = 0; r4 < full; r4 += 4) {
            float2 dv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) dv[q] = lds_f32x2(rd0 + (uint32_t)((r4 + q) * H_EPI_PITCH) * 4u);
#pragma unroll
            for (int q = 0; q < 4; ++q) row_out(dv[q]);
        }
#pragma unroll 1
        for (int r = full; r < rows_live; ++r) row_out(lds_f32x2(rd0 + (uint32_t)(r * H_EPI_PITCH) * 4u));
    }
}

// ---------------------------------------------------------------------------------------
// kernel: persistent CTAs (one per SM) walk 128 x 128 output tiles (n fastest, then m, then batch)
// ---------------------------------------------------------------------------------------
template <int EPI, bool DUAL, int BN, int NSO = 0>
__global__ void __launch_bounds__(HCfg<EPI, DUAL, BN, NSO>::kThreads, 1)
h_gemm_kernel(const __grid_constant__ HMaps tm, const HArgs p) {
    using Cfg = HCfg<EPI, DUAL, BN, NSO>;
    static_assert(BN == 128 || (BN == 256 && EPI == TC_EPI_PLAIN && !DUAL), "256-wide tiles: plain single-accumulator epilogue only");
    constexpr int NS = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_stage = base + NS * Cfg::kStageBytes;
    const uint32_t epi_czs = epi_stage + Cfg::kEWarps * Cfg::kEpiWarpBytes;
    const uint32_t bars = epi_czs + Cfg::kEWarps * 256u;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (NS + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (2 * NS + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (2 * NS + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * NS + 4);

    // ---- set-up that depends on nothing the previous kernel wrote: overlaps its tail under PDL ----
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), Cfg::kEWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; ++q) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.m[q >> 2][(q >> 1) & 1][q & 1]) : "memory");
    }
    if (warp == 1) {  // all 512 columns: the attention epilogues read 32-column windows that may overhang an accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    pdl_wait();  // from here on: global memory written by the previous kernels

    const int Mlive = p.dynM ? min(*p.dynM, p.M) : p.M;
    const int Nlive = p.dynN ? min(*p.dynN, p.N) : p.N;
    const int mt = (Mlive + H_BM - 1) / H_BM, nt = (Nlive + BN - 1) / BN;
    const int ntiles = mt * nt * p.batch;  // CTAs beyond the live tiles fall through the role loops

    int kb0 = 0, kb1 = 0;
    {
        const int K0 = p.dynK ? min(*p.dynK, p.K[0]) : p.K[0];
        kb0 = (K0 + H_BK - 1) / H_BK;
        if (p.nseg > 1) {
            const int K1 = p.dynK ? min(*p.dynK, p.K[1]) : p.K[1];
            kb1 = (K1 + H_BK - 1) / H_BK;
        }
    }
    const int kbt = kb0 + kb1;
    auto tile_total = [&](int z) {
        int nred = p.nred;
        if (p.red_total > 0) nred = max(0, min(p.nred, p.red_total - z * p.nred));
        return nred * kbt;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int z = t / (mt * nt), rem_t = t - z * (mt * nt);
                const int m0 = (rem_t / nt) * H_BM, n0 = (rem_t % nt) * BN;
                const int total = tile_total(z);
                for (int k = 0; k < total; ++k, ++it) {
                    const int s = it % NS, ph = (it / NS) & 1;
                    const int red = k / kbt, rem = k - red * kbt;
                    const int seg = rem >= kb0 ? 1 : 0;
                    const int k0 = (seg ? rem - kb0 : rem) * H_BK;
                    const int zr = z * p.nred + red;
                    const int zA = p.a_batched[seg] ? zr : 0, zB = p.b_batched[seg] ? zr : 0;
                    mbar_wait(empty(s), ph ^ 1);
                    mbar_arrive_expect_tx(full(s), (uint32_t)Cfg::kStageBytes);
                    const uint32_t sA = base + s * Cfg::kStageBytes, sB = sA + 2 * H_A_TILE;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const CUtensorMap* ta = &tm.m[seg][0][h];
                        tma_load_3d(sA + h * H_A_TILE, ta, full(s), m0, k0, zA);
                        tma_load_3d(sA + h * H_A_TILE + 4096, ta, full(s), m0 + 64, k0, zA);
                        tma_load_3d(sB + h * Cfg::kBTile, &tm.m[seg][1][h], full(s), k0, n0, zB);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) /*D=f32*/ | (0u << 7) /*A=f16*/ | (0u << 10) /*B=f16*/ | (1u << 15) /*A MN-major*/ |
                                       (0u << 16) /*B K-major*/ | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(H_BM >> 4) << 24);
            int it = 0, ti = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
                const int z = t / (mt * nt);
                const int total = tile_total(z);
                const int acc = ti & 1;
                const uint32_t tmem_d0 = tmem_base + (uint32_t)(acc * Cfg::kAccStride);
                mbar_wait(tmem_empty(acc), ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int k = 0; k < total; ++k, ++it) {
                    const int s = it % NS, ph = (it / NS) & 1;
                    const int rem = k % kbt;
                    const int seg = rem >= kb0 ? 1 : 0;
                    const bool first = DUAL ? (k < kbt && (seg ? rem == kb0 : rem == 0)) : (k == 0);
                    const uint32_t tmem_d = tmem_d0 + (DUAL ? (uint32_t)(seg * BN) : 0u);
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * Cfg::kStageBytes, a_lo = a_hi + H_A_TILE;
                    const uint32_t b_hi = a_hi + 2 * H_A_TILE, b_lo = b_hi + Cfg::kBTile;
#pragma unroll
                    for (int ks = 0; ks < H_BK / 16; ++ks) {
                        const uint64_t dah = h_desc_a(a_hi, ks), dal = h_desc_a(a_lo, ks);
                        const uint64_t dbh = h_desc_b(b_hi, ks), dbl = h_desc_b(b_lo, ks);
                        h_mma_f16(tmem_d, dal, dbh, idesc, (!first || ks > 0) ? 1u : 0u);
                        h_mma_f16(tmem_d, dah, dbl, idesc, 1u);
                        h_mma_f16(tmem_d, dah, dbh, idesc, 1u);
                    }
                    tc_commit(empty(s));
                }
                tc_commit(tmem_full(acc));
            }
        }
    } else {
        // ===== epilogue =====
        const int ew = warp - 2;
        float inv0 = __ldg(p.inv_a[0]) * __ldg(p.inv_b[0]);
        float inv1 = (DUAL && p.nseg > 1) ? __ldg(p.inv_a[1]) * __ldg(p.inv_b[1]) : 0.f;
        EpiTile et;
        et.quarter = warp & 3;
        et.half = ew >> 2;
        et.stage = epi_stage + (uint32_t)ew * Cfg::kEpiWarpBytes;
        et.czs = epi_czs + (uint32_t)ew * 256u;
        et.Mlive = Mlive;
        et.Nlive = Nlive;
        int ti = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
            const int rem_t = t % (mt * nt);
            const int acc = ti & 1;
            et.z = t / (mt * nt);
            et.m0 = (rem_t / nt) * H_BM;
            et.n0 = (rem_t % nt) * BN;
            et.total = tile_total(et.z);
            et.tacc = tmem_base + ((uint32_t)(et.quarter * 32) << 16) + (uint32_t)(acc * Cfg::kAccStride);
            et.full_bar = tmem_full(acc);
            et.full_parity = (uint32_t)((ti >> 1) & 1);
            et.empty_bar = tmem_empty(acc);
            h_epilogue_tile<EPI, DUAL, BN>(p, et, lane, inv0, inv1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn h_get_encode() {
    static thread_local bool bound = false;  // the driver entry point needs the primary context bound to this thread
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct HMapKey {
    const __half* ptr;
    long long ld, bstride;
    int is_b, nbatch, rows, K;
    bool operator==(const HMapKey& o) const {
        return ptr == o.ptr && ld == o.ld && bstride == o.bstride && is_b == o.is_b && nbatch == o.nbatch && rows == o.rows && K == o.K;
    }
};
struct HMapSlot {
    HMapKey key;
    CUtensorMap map;
    bool used;
};
static thread_local HMapSlot g_hmap_cache[48];
static thread_local int g_hmap_next = 0;

static int h_make_map(CUtensorMap* m, const __half* ptr, const HOperand& o, bool is_b, int bn) {
    const HMapKey key{ptr, o.ld, o.bstride, is_b ? bn : 0, o.nbatch, o.rows, o.K};
    for (int i = 0; i < 48; ++i)
        if (g_hmap_cache[i].used && g_hmap_cache[i].key == key) {
            *m = g_hmap_cache[i].map;
            return EEGAN_OK;
        }
    EncodeTiledFn enc = h_get_encode();
    if (!enc) { set_error("h gemm: cuTensorMapEncodeTiled unavailable"); return EEGAN_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (o.ld % 8) || (o.bstride % 8)) {
        set_error("h gemm: operand base/pitch must be 16-byte aligned (ptr=%p ld=%lld bstride=%lld)", (const void*)ptr, o.ld, o.bstride);
        return EEGAN_ERR_INVALID;
    }
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (is_b) {  // K-major [rows][K]
        dims[0] = (cuuint64_t)o.K; dims[1] = (cuuint64_t)o.rows;
        box[0] = H_BK; box[1] = (cuuint32_t)bn;
    } else {     // MN-major [K][rows]
        dims[0] = (cuuint64_t)o.rows; dims[1] = (cuuint64_t)o.K;
        box[0] = 64; box[1] = H_BK;
    }
    dims[2] = (cuuint64_t)(o.nbatch > 0 ? o.nbatch : 1);
    box[2] = 1;
    strides[0] = (cuuint64_t)o.ld * 2;
    strides[1] = (cuuint64_t)(o.bstride > 0 ? o.bstride : (long long)dims[1] * o.ld) * 2;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, is_b ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("h gemm: cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu ld=%lld", (int)r, (unsigned long long)dims[0],
                  (unsigned long long)dims[1], (unsigned long long)dims[2], o.ld);
        return EEGAN_ERR_CUDA;
    }
    HMapSlot& slot = g_hmap_cache[g_hmap_next];
    g_hmap_next = (g_hmap_next + 1) % 48;
    slot.key = key;
    slot.map = *m;
    slot.used = true;
    return EEGAN_OK;
}

static int h_num_sms() { return num_sms_current(); }

template <int EPI, bool DUAL, int BN = H_BN, int NSO = 0>
static int h_launch_t(const HMaps& maps, const HArgs& a, unsigned grid, cudaStream_t st) {
    using Cfg = HCfg<EPI, DUAL, BN, NSO>;
    static SmemGrant grant;
    if (int rc = grant_dyn_smem(h_gemm_kernel<EPI, DUAL, BN, NSO>, (size_t)Cfg::kSmem, grant, "h gemm")) return rc;
    cudaError_t e = launch_pdl(h_gemm_kernel<EPI, DUAL, BN, NSO>, dim3(grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmem, st, maps, a);
    if (e != cudaSuccess) { set_error("h gemm launch: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
    return check_launch("h gemm");
}

int h_gemm_launch(const HGemm& g, cudaStream_t st) {
    EEGAN_REQUIRE(g.nseg == 1 || g.nseg == 2, "h gemm: nseg=%d", g.nseg);
    EEGAN_REQUIRE(g.M > 0 && g.N > 0 && g.batch > 0, "h gemm: empty problem");
    // 256-wide tiles (EEGAN_H_WIDE=1; off by default) for the plain single-segment GEMM whose N is a multiple of 256
    // (U': N = D): one A tile feeds twice the output, 48 KB instead of 64 KB of operands per 128 x 256 x 32 of work.
    // Measured neutral at B = 48 (GEMM2 30.8 vs 30.7 us): at 10 k-blocks per tile the launch is bound by its ramp and
    // tile-boundary latencies, not by L2 -> SM bandwidth; kept for larger D / R.
    static const bool wide_ok = [] { const char* e = getenv("EEGAN_H_WIDE"); return e ? atoi(e) != 0 : false; }();
    const int bn = (wide_ok && g.epi == TC_EPI_PLAIN && g.nseg == 1 && g.N % 256 == 0 && !g.dynN && g.red_total == 0) ? 256 : H_BN;
    HMaps maps;
    HArgs a{};
    for (int s = 0; s < 2; ++s) {
        const int src = s < g.nseg ? s : 0;
        const HOperand* ops[2] = {&g.A[src], &g.B[src]};
        for (int o = 0; o < 2; ++o) {
            EEGAN_REQUIRE(ops[o]->hi && ops[o]->lo && ops[o]->inv_scale, "h gemm: operand arrays missing");
            int rc = h_make_map(&maps.m[s][o][0], ops[o]->hi, *ops[o], o == 1, bn);
            if (rc) return rc;
            rc = h_make_map(&maps.m[s][o][1], ops[o]->lo, *ops[o], o == 1, bn);
            if (rc) return rc;
        }
        a.K[s] = s < g.nseg ? g.A[src].K : 0;
        a.a_batched[s] = g.A[src].bstride > 0;
        a.b_batched[s] = g.B[src].bstride > 0;
        a.inv_a[s] = g.A[src].inv_scale;
        a.inv_b[s] = g.B[src].inv_scale;
    }
    a.C = g.C; a.ldc = g.ldc; a.bC = g.bC; a.M = g.M; a.N = g.N; a.dynM = g.dynM; a.dynN = g.dynN; a.dynK = g.dynK;
    a.attn = g.attn;
    a.nseg = g.nseg; a.nred = g.nred > 0 ? g.nred : 1; a.red_total = g.red_total; a.batch = g.batch;
#ifdef EEGAN_DEBUG_SWITCHES
    if (const char* e = getenv("EEGAN_H_DBG")) a.dbg = atoi(e);
#endif
    const long long tiles = (long long)((g.N + bn - 1) / bn) * ((g.M + H_BM - 1) / H_BM) * g.batch;
    const unsigned grid = (unsigned)(tiles < h_num_sms() ? tiles : h_num_sms());
    if (g.epi != TC_EPI_PLAIN) {
        const TcAttnEpi& e = g.attn.base;
        EEGAN_REQUIRE(g.nseg == 1, "h gemm: the attention epilogues take one segment");
        EEGAN_REQUIRE(e.nbins && e.bin_cap && e.bin_used && e.col_start && e.cap_len && e.P && g.attn.out_hi && g.attn.out_lo,
                      "h gemm: attention epilogue arguments missing");
        EEGAN_REQUIRE(g.ldc % 64 == 0, "h gemm: attention epilogue needs a column pitch that is a multiple of 64");
        if (g.epi == TC_EPI_ATTN_FWD) {
            EEGAN_REQUIRE(e.Zpart && g.attn.emax, "h gemm: attention forward epilogue needs Zpart and emax");
            return h_launch_t<TC_EPI_ATTN_FWD, false>(maps, a, grid, st);
        }
        EEGAN_REQUIRE(g.epi == TC_EPI_ATTN_BWD && e.csz && g.attn.out_scale, "h gemm: attention backward epilogue needs csz and the dS scale");
        return h_launch_t<TC_EPI_ATTN_BWD, false>(maps, a, grid, st);
    }
    EEGAN_REQUIRE(g.C, "h gemm: no output");
    if (g.nseg == 2) return h_launch_t<TC_EPI_PLAIN, true>(maps, a, grid, st);
    if (bn == 256) return h_launch_t<TC_EPI_PLAIN, false, 256>(maps, a, grid, st);
#ifdef EEGAN_DEBUG_SWITCHES
    static const int nso = [] { const char* e = getenv("EEGAN_H_STAGES"); return e ? atoi(e) : 0; }();  // stage-count probe
    if (nso == 2) return h_launch_t<TC_EPI_PLAIN, false, H_BN, 2>(maps, a, grid, st);
    if (nso == 3) return h_launch_t<TC_EPI_PLAIN, false, H_BN, 3>(maps, a, grid, st);
    if (nso == 4) return h_launch_t<TC_EPI_PLAIN, false, H_BN, 4>(maps, a, grid, st);
#endif
    return h_launch_t<TC_EPI_PLAIN, false>(maps, a, grid, st);
}

// =======================================================================================
// Fused forward of the pair grid: per (image j, 128-column word tile), for every 128-region slab
//     S_slab = C[j][:, slab]^T W_tile          (tcgen05, K = D)           -> TMEM accumulator (double-buffered)
//     P = softmax_words(S),  E' = 2^12 exp(g1 (P - 1))                    (attention epilogue, thread = region)
//     U'[n'][d] += sum_{r in slab} E'[r][n'] C[j][d][r]   (tcgen05, K = 128 regions, N = D = 256) -> second TMEM accumulator
// E' goes from the epilogue's registers into a shared-memory tile laid out as the MN-major SWIZZLE_128B A operand of the
// second contraction (and, as before, to the HBM stash the backward reads) — GEMM2 no longer re-reads E from L2 / HBM, the
// U' contraction rides behind the next slab's S contraction on the same tensor pipe, and one launch ramp disappears.
// Roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-9 epilogue (two per TMEM lane quarter: one 64-column
// bin of the tile each).  MMA order per item (ns slabs): M1(0); for s >= 1: M1(s), M2(s-1); M2(ns-1) — the producer issues
// its loads in the same order through one 3-stage ring of 32 KB blocks (S: C boxes + W boxes; U': one [256 d][32 r] box pair).
// TMEM: S0 [0,128) S1 [128,256) U' [256,512).  Needs D == 256.
// =======================================================================================
constexpr int HF_NS = 3;
constexpr int HF_STAGE = 32768;
constexpr int HF_EHALF = 32768;             // one of E hi / lo: [4 k-blocks of 32 regions][2 chunks of 64 columns][32 rows][128 B]
constexpr int HF_EWARPS = 8;
constexpr int HF_EPI_WARP_BYTES = 32 * H_EPI_PITCH * 4;
constexpr int HF_THREADS = 32 * (2 + HF_EWARPS);
constexpr int HF_SMEM = HF_NS * HF_STAGE + 2 * HF_EHALF + HF_EWARPS * HF_EPI_WARP_BYTES + 128;
static_assert(HF_SMEM <= 232448, "fused forward: shared memory budget");

struct HFMaps {
    CUtensorMap c_mn[2];  // C as A of S:  MN-major [d][r], box [32 d][64 r]           (hi, lo)
    CUtensorMap w_k[2];   // W as B of S:  K-major [n'][d], box [128 n'][32 d]
    CUtensorMap c_k[2];   // C as B of U': K-major [d][r],  box [256 d][32 r]
};
struct HFArgs {
    float* U;              // [Bi][NtP][D]
    long long ldc, bC;     // stash pitch (NtP) and image stride (R * NtP) of P / E / (Zpart uses ldc)
    int R, D, NtP, Bi;
    const int* nlive;      // live packed columns (bins * 64)
    const float* inv_c;
    const float* inv_w;
    const float* inv_e;
    HAttnEpi attn;
    int dbg;  // -DEEGAN_DEBUG_SWITCHES builds only (EEGAN_HF_DBG): 1 no stash stores, 2 no read-out, 4 no caption phase, 8 no U' MMAs, 16 no U' drain stores
};
#ifdef EEGAN_DEBUG_SWITCHES
#define HF_DBG(p) ((p).dbg)
#else
#define HF_DBG(p) 0
#endif

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// attention forward on one S tile (one epilogue warp: 32 regions x one 64-column bin), E' also into the shared-memory operand tile
__device__ __forceinline__ void hf_attn_fwd_tile(const HFArgs& p, const EpiTile& t, int lane, float inv0, uint32_t e_hi, uint32_t e_lo,
                                                 uint32_t e_free_bar, int e_free_wait, uint32_t e_free_parity, uint32_t e_ready_bar) {
    const TcAttnEpi& e = p.attn.base;
    const int nbins = *e.nbins;
    const int row0 = t.m0 + t.quarter * 32;
    const int rows_live = max(0, min(32, t.Mlive - row0));
    const uint32_t my_row = t.stage + (uint32_t)(lane * H_EPI_PITCH) * 4u;
    const int h = t.half;
    const int b0 = t.n0 >> 6;
    const int b = b0 + h;
    if (b >= nbins) {  // no live bin for this warp in the tile: keep the hand-shakes going (in step with the live warps:
        // an arrival for slab g may only follow the completion of slab g - 1's phase, hence the same e_free wait)
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t.empty_bar);
        if (e_free_wait) mbar_wait(e_free_bar, e_free_parity);
        __syncwarp();
        if (lane == 0) mbar_arrive(e_ready_bar);
        return;
    }
    const int bc_l = e.bin_cap[min(b0 + lane, nbins)];
    const int bu_l = (lane < 2 && b0 + lane < nbins) ? e.bin_used[b0 + lane] : 0;
    const int i0 = __shfl_sync(0xffffffffu, bc_l, h), i1 = __shfl_sync(0xffffffffu, bc_l, h + 1);
    const int used = __shfl_sync(0xffffffffu, bu_l, h);
    const int col0 = t.n0 + 64 * h;
    bool waited = false;
#pragma unroll 1
    for (int ic = i0; ic < i1; ic += 32) {
        const int nc = min(32, i1 - ic);
        const int myT = lane < nc ? e.cap_len[ic + lane] : 0;
        const int myC = lane < nc ? e.col_start[ic + lane] - col0 : 0;
        if (!waited) {
            mbar_wait(t.full_bar, t.full_parity);
            tc_fence_after();
            waited = true;
        }
#pragma unroll 1
        for (int ci = 0; ci < nc; ++ci) {
            const int T = __shfl_sync(0xffffffffu, myT, ci);
            if (T <= 0) continue;
            const int cl = __shfl_sync(0xffffffffu, myC, ci);
            if (!(HF_DBG(p) & 4)) h_attn_caption<TC_EPI_ATTN_FWD>(t.tacc + (uint32_t)(64 * h + cl), T, my_row + (uint32_t)cl * 4u, 0u, e.g1, inv0);
        }
    }
    if (!waited) {
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_after();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(t.empty_bar);
    for (int c = used; c < 64; ++c) sts_f32(my_row + (uint32_t)c * 4u, 0.f);  // padding columns of the bin
    __syncwarp();
    if (e_free_wait) mbar_wait(e_free_bar, e_free_parity);  // the previous slab's U' contraction has read the operand tile
    // read-out: region rows, coalesced; a lane owns the adjacent columns 2 lane, 2 lane + 1 of the bin
    float* Pz = e.P + (long long)t.z * p.bC;
    __half* Hz = p.attn.out_hi + (long long)t.z * p.bC;
    __half* Lz = p.attn.out_lo + (long long)t.z * p.bC;
    const uint32_t rd0 = t.stage + (uint32_t)(2 * lane) * 4u;
    const long long o0 = (long long)row0 * p.ldc + col0 + 2 * lane;
    uint32_t* hp = reinterpret_cast<uint32_t*>(Hz + o0);
    uint32_t* lp = reinterpret_cast<uint32_t*>(Lz + o0);
    float2* pp = reinterpret_cast<float2*>(Pz + o0);
    const long long hstep = p.ldc >> 1;
    // operand tile: k-block = lane quarter, chunk = bin h, row r of the quarter: 128-byte rows, 16-byte units XOR (r & 7)
    const uint32_t eoff = (uint32_t)t.quarter * 8192u + (uint32_t)h * 4096u;
    const uint32_t eunit = (uint32_t)(lane >> 2), ein = (uint32_t)(lane & 3) << 2;
    const float a = e.g1 * 1.4426950408889634f, bb = 12.0f - a;
    float z0 = 0.f, z1 = 0.f, em = 0.f;
    const int rows_do = (HF_DBG(p) & 2) ? 0 : rows_live;
#pragma unroll 4
    for (int r = 0; r < rows_do; ++r) {
        const float2 pv = lds_f32x2(rd0 + (uint32_t)(r * H_EPI_PITCH) * 4u);
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(pv.x, a, bb)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(pv.y, a, bb)));
        const __half2 hh = __floats2half2_rn(e0, e1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(e0 - hf.x, e1 - hf.y);
        const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hh), lu = *reinterpret_cast<const uint32_t*>(&ll);
        if (!(HF_DBG(p) & 1)) {
            *pp = pv;
            *hp = hu;
            *lp = lu;
        }
        pp += hstep; hp += hstep; lp += hstep;
        const uint32_t ea = eoff + (uint32_t)r * 128u + ((eunit ^ (uint32_t)(r & 7)) << 4) + ein;
        sts_u32(e_hi + ea, hu);
        sts_u32(e_lo + ea, lu);
        z0 += e0;
        z1 += e1;
        em = fmaxf(em, fmaxf(e0, e1));
    }
    if (rows_live > 0) {
        float* zp = e.Zpart + ((long long)t.z * ((p.R + 31) / 32) + (row0 >> 5)) * p.ldc + col0 + 2 * lane;
        *reinterpret_cast<float2*>(zp) = make_float2(z0 * (1.0f / H_E_SCALE), z1 * (1.0f / H_E_SCALE));
    }
    em = warp_max(em) * (1.0f / H_E_SCALE);
    if (lane == 0 && em > 0.f) atomicMax(reinterpret_cast<int*>(p.attn.emax), __float_as_int(em));
    fence_proxy_async();  // the tensor core reads the tile through the async proxy
    __syncwarp();
    if (lane == 0) mbar_arrive(e_ready_bar);
}

__global__ void __launch_bounds__(HF_THREADS, 1) hf_fwd_kernel(const __grid_constant__ HFMaps tm, const HFArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t e_hi = base + HF_NS * HF_STAGE, e_lo = e_hi + HF_EHALF;
    const uint32_t epi_stage = e_lo + HF_EHALF;
    const uint32_t bars = epi_stage + HF_EWARPS * HF_EPI_WARP_BYTES;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (HF_NS + s); };
    auto s_full = [&](int a) { return bars + 8u * (2 * HF_NS + a); };
    auto s_empty = [&](int a) { return bars + 8u * (2 * HF_NS + 2 + a); };
    const uint32_t e_ready = bars + 8u * (2 * HF_NS + 4), e_free = e_ready + 8u, u_full = e_ready + 16u, u_empty = e_ready + 24u;
    const uint32_t tmem_slot = e_ready + 32u;

    pdl_trigger();
    if ((base & 1023u) != 0u) __trap();  // SWIZZLE_128B tiles need the 1024-byte alignment the declaration asks for
    if (threadIdx.x == 0) {
        for (int s = 0; s < HF_NS; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(s_full(a), 1);
            mbar_init(s_empty(a), HF_EWARPS);
        }
        mbar_init(e_ready, HF_EWARPS);
        mbar_init(e_free, 1);
        mbar_init(u_full, 1);
        mbar_init(u_empty, HF_EWARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.c_mn[q]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.w_k[q]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.c_k[q]) : "memory");
        }
    }
    // the operand tile starts as zeros: rows beyond the live regions of a slab are never written and must stay finite
    for (uint32_t o = threadIdx.x * 16u; o < 2u * HF_EHALF; o += HF_THREADS * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(e_hi + o), "r"(0u) : "memory");
    fence_proxy_async();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    pdl_wait();  // from here on: global memory written by the previous kernels

    const int nlive = min(*p.nlive, p.NtP);
    const int ntl = (nlive + H_BN - 1) / H_BN;
    const int nitems = ntl * p.Bi;
    const int ns = (p.R + H_BM - 1) / H_BM;
    const int kb1 = p.D / H_BK;
    auto kb2_of = [&](int s) { return (min(H_BM, p.R - s * H_BM) + H_BK - 1) / H_BK; };

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            auto acquire = [&]() {
                const int s = it % HF_NS, ph = (it / HF_NS) & 1;
                mbar_wait(empty(s), ph ^ 1);
                mbar_arrive_expect_tx(full(s), (uint32_t)HF_STAGE);
                ++it;
                return s;
            };
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int j = item / ntl, n0 = (item - j * ntl) * H_BN;
                auto load1 = [&](int sl) {
                    for (int kb = 0; kb < kb1; ++kb) {
                        const int s = acquire();
                        const uint32_t sA = base + s * HF_STAGE, sB = sA + 2 * H_A_TILE;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            tma_load_3d(sA + h * H_A_TILE, &tm.c_mn[h], full(s), sl * H_BM, kb * H_BK, j);
                            tma_load_3d(sA + h * H_A_TILE + 4096, &tm.c_mn[h], full(s), sl * H_BM + 64, kb * H_BK, j);
                            tma_load_3d(sB + h * H_B_TILE, &tm.w_k[h], full(s), kb * H_BK, n0, 0);
                        }
                    }
                };
                auto load2 = [&](int sl) {
                    const int nkb = kb2_of(sl);
                    for (int kb = 0; kb < nkb; ++kb) {
                        const int s = acquire();
                        const uint32_t sB = base + s * HF_STAGE;
                        tma_load_3d(sB, &tm.c_k[0], full(s), sl * H_BM + kb * H_BK, 0, j);
                        tma_load_3d(sB + 16384, &tm.c_k[1], full(s), sl * H_BM + kb * H_BK, 0, j);
                    }
                };
                load1(0);
                for (int sl = 1; sl < ns; ++sl) {
                    load1(sl);
                    load2(sl - 1);
                }
                load2(ns - 1);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_common = (1u << 4) /*D=f32*/ | (0u << 7) /*A=f16*/ | (0u << 10) /*B=f16*/ | (1u << 15) /*A MN-major*/ |
                                              (0u << 16) /*B K-major*/ | ((uint32_t)(H_BM >> 4) << 24);
            constexpr uint32_t idesc1 = idesc_common | ((uint32_t)(H_BN >> 3) << 17);
            constexpr uint32_t idesc2 = idesc_common | ((uint32_t)(256 >> 3) << 17);
            const uint32_t tmem_u = tmem_base + 256u;
            int it = 0, g1c = 0 /*S slabs issued*/, g2c = 0 /*U' slabs issued*/, items_done = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++items_done) {
                auto mma1 = [&]() {
                    const int acc = g1c & 1;
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * H_BN);
                    mbar_wait(s_empty(acc), ((g1c >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < kb1; ++kb, ++it) {
                        const int s = it % HF_NS, ph = (it / HF_NS) & 1;
                        mbar_wait(full(s), ph);
                        tc_fence_after();
                        const uint32_t a_hi = base + s * HF_STAGE, a_lo = a_hi + H_A_TILE;
                        const uint32_t b_hi = a_hi + 2 * H_A_TILE, b_lo = b_hi + H_B_TILE;
#pragma unroll
                        for (int ks = 0; ks < H_BK / 16; ++ks) {
                            const uint64_t dah = h_desc_a(a_hi, ks), dal = h_desc_a(a_lo, ks);
                            const uint64_t dbh = h_desc_b(b_hi, ks), dbl = h_desc_b(b_lo, ks);
                            h_mma_f16(tmem_d, dal, dbh, idesc1, (kb > 0 || ks > 0) ? 1u : 0u);
                            h_mma_f16(tmem_d, dah, dbl, idesc1, 1u);
                            h_mma_f16(tmem_d, dah, dbh, idesc1, 1u);
                        }
                        tc_commit(empty(s));
                    }
                    tc_commit(s_full(acc));
                    ++g1c;
                };
                auto mma2 = [&](int sl) {
                    mbar_wait(e_ready, (uint32_t)(g2c & 1));
                    if (sl == 0) mbar_wait(u_empty, (uint32_t)((items_done & 1) ^ 1));  // the previous item's U' has been drained
                    tc_fence_after();
                    const int nkb = kb2_of(sl);
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % HF_NS, ph = (it / HF_NS) & 1;
                        mbar_wait(full(s), ph);
                        tc_fence_after();
                        const uint32_t b_hi = base + s * HF_STAGE, b_lo = b_hi + 16384;
                        const uint32_t a_hi = e_hi + (uint32_t)kb * 8192u, a_lo = e_lo + (uint32_t)kb * 8192u;
#pragma unroll
                        for (int ks = 0; ks < ((HF_DBG(p) & 8) ? 0 : H_BK / 16); ++ks) {
                            const uint64_t dah = h_desc_a(a_hi, ks), dal = h_desc_a(a_lo, ks);
                            const uint64_t dbh = h_desc_b(b_hi, ks), dbl = h_desc_b(b_lo, ks);
                            h_mma_f16(tmem_u, dal, dbh, idesc2, (sl > 0 || kb > 0 || ks > 0) ? 1u : 0u);
                            h_mma_f16(tmem_u, dah, dbl, idesc2, 1u);
                            h_mma_f16(tmem_u, dah, dbh, idesc2, 1u);
                        }
                        tc_commit(empty(s));
                    }
                    tc_commit(e_free);
                    if (sl == ns - 1) tc_commit(u_full);
                    ++g2c;
                };
                mma1();
                for (int sl = 1; sl < ns; ++sl) {
                    mma1();
                    mma2(sl - 1);
                }
                mma2(ns - 1);
            }
        }
    } else {
        const int ew = warp - 2;
        const float inv_s = __ldg(p.inv_c) * __ldg(p.inv_w);
        const float inv_u = __ldg(p.inv_e) * __ldg(p.inv_c);
        HArgs pu{};  // the U' drain reuses the plain epilogue
        pu.C = p.U; pu.ldc = p.D; pu.bC = (long long)p.NtP * p.D; pu.M = p.NtP; pu.N = p.D;
        EpiTile et;
        et.quarter = warp & 3;
        et.half = ew >> 2;
        et.stage = epi_stage + (uint32_t)ew * HF_EPI_WARP_BYTES;
        et.czs = 0;
        int g = 0, idone = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++idone) {
            const int j = item / ntl, n0 = (item - j * ntl) * H_BN;
            for (int sl = 0; sl < ns; ++sl, ++g) {
                const int acc = g & 1;
                et.z = j; et.m0 = sl * H_BM; et.n0 = n0; et.total = kb1; et.Mlive = p.R; et.Nlive = nlive;
                et.tacc = tmem_base + ((uint32_t)(et.quarter * 32) << 16) + (uint32_t)(acc * H_BN);
                et.full_bar = s_full(acc);
                et.full_parity = (uint32_t)((g >> 1) & 1);
                et.empty_bar = s_empty(acc);
                hf_attn_fwd_tile(p, et, lane, inv_s, e_hi, e_lo, e_free, g > 0, (uint32_t)((g - 1) & 1), e_ready);
            }
            EpiTile eu = et;
            eu.z = j; eu.m0 = n0; eu.n0 = 0; eu.total = 1; eu.Mlive = nlive; eu.Nlive = p.D;
            eu.tacc = tmem_base + ((uint32_t)(et.quarter * 32) << 16) + 256u;
            eu.full_bar = u_full;
            eu.full_parity = (uint32_t)(idone & 1);
            eu.empty_bar = u_empty;
            h_epilogue_tile<TC_EPI_PLAIN, false, 256>(pu, eu, lane, inv_u, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int h_fused_fwd_launch(const HFusedFwd& f, cudaStream_t st) {
    EEGAN_REQUIRE(f.D == 256, "fused forward: D=%d (the U' accumulator takes exactly 256 tensor-memory columns)", f.D);
    EEGAN_REQUIRE(f.NtP % 64 == 0 && f.R > 0 && f.Bi > 0, "fused forward: bad shape");
    HFMaps maps;
    const HOperand c_mn{f.Ch, f.Cl, f.Rp, (long long)f.D * f.Rp, f.Bi, f.R, f.D, f.inv_c};
    const HOperand c_k{f.Ch, f.Cl, f.Rp, (long long)f.D * f.Rp, f.Bi, f.D, f.R, f.inv_c};
    const HOperand w_k{f.Wh, f.Wl, f.D, 0, 1, f.NtP, f.D, f.inv_w};
    int rc;
    if ((rc = h_make_map(&maps.c_mn[0], f.Ch, c_mn, false, H_BN))) return rc;
    if ((rc = h_make_map(&maps.c_mn[1], f.Cl, c_mn, false, H_BN))) return rc;
    if ((rc = h_make_map(&maps.w_k[0], f.Wh, w_k, true, H_BN))) return rc;
    if ((rc = h_make_map(&maps.w_k[1], f.Wl, w_k, true, H_BN))) return rc;
    if ((rc = h_make_map(&maps.c_k[0], f.Ch, c_k, true, 256))) return rc;
    if ((rc = h_make_map(&maps.c_k[1], f.Cl, c_k, true, 256))) return rc;
    HFArgs a{};
    a.U = f.U; a.ldc = f.NtP; a.bC = (long long)f.R * f.NtP; a.R = f.R; a.D = f.D; a.NtP = f.NtP; a.Bi = f.Bi;
    a.nlive = f.nlive; a.inv_c = f.inv_c; a.inv_w = f.inv_w; a.inv_e = f.inv_e; a.attn = f.attn;
#ifdef EEGAN_DEBUG_SWITCHES
    if (const char* e = getenv("EEGAN_HF_DBG")) a.dbg = atoi(e);
#endif
    const TcAttnEpi& e = f.attn.base;
    EEGAN_REQUIRE(e.nbins && e.bin_cap && e.bin_used && e.col_start && e.cap_len && e.P && e.Zpart && f.attn.out_hi && f.attn.out_lo &&
                      f.attn.emax && f.U && f.nlive, "fused forward: arguments missing");
    static SmemGrant grant;
    if ((rc = grant_dyn_smem(hf_fwd_kernel, (size_t)HF_SMEM, grant, "fused forward"))) return rc;
    const long long items = (long long)(f.NtP / H_BN + (f.NtP % H_BN ? 1 : 0)) * f.Bi;
    const unsigned grid = (unsigned)(items < h_num_sms() ? items : h_num_sms());
    cudaError_t err = launch_pdl(hf_fwd_kernel, dim3(grid), dim3(HF_THREADS), (size_t)HF_SMEM, st, maps, a);
    if (err != cudaSuccess) { set_error("fused forward launch: %s", cudaGetErrorString(err)); return EEGAN_ERR_CUDA; }
    return check_launch("fused forward");
}

// elementwise split of an fp32 array into the half pair of x * s (test entry point below)
__global__ void h_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long long n, float s) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        __half h, l;
        h_split(x[i] * s, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}
__global__ void h_set2_kernel(float* p, float a, float b, float c, float d) {
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
}

}  // namespace eegan

using namespace eegan;

// Stand-alone entry point (tests / microbench): C[z] = A[z] * B[z]^T through the half-pair engine.
//   A is MN-major: [K][lda] with the M index contiguous;  B is K-major: [N][ldb].  lda, ldb multiples of 8.
//   sa, sb: power-of-two scales the operands are stored with.  dual != 0 runs the two-accumulator form with the
//   same operands in both segments (the second stored with scales 4 sa, sb / 8): the result is 2 A B^T.
//   workspace: 2 * (elements of A + elements of B) halves (x2 when dual) + 16 bytes.
extern "C" int eegan_gemm_f16x3(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb,
                                long long ldc, long long bsA, long long bsB, long long bsC, int batch, float sa, float sb, int dual,
                                void* workspace, size_t workspace_bytes, void* stream) {
    EEGAN_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0 && workspace, "gemm_f16x3: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nA = bsA > 0 ? bsA * batch : (long long)K * lda, nB = bsB > 0 ? bsB * batch : (long long)N * ldb;
    const int nset = (dual & 1) ? 2 : 1;
    const size_t need = (size_t)nset * 2 * (align_up(nA * 2, 256) + align_up(nB * 2, 256)) + 256;
    EEGAN_REQUIRE(workspace_bytes >= need, "gemm_f16x3: workspace %zu < %zu", workspace_bytes, need);
    char* p = (char*)workspace;
    float* scales = (float*)p;
    p += 256;
    HGemm g{};
    g.nseg = nset;
    h_set2_kernel<<<1, 1, 0, st>>>(scales, 1.0f / sa, 1.0f / sb, 1.0f / (4.0f * sa), 8.0f / sb);
    for (int s = 0; s < nset; ++s) {
        __half* ah = (__half*)p; p += align_up(nA * 2, 256);
        __half* al = (__half*)p; p += align_up(nA * 2, 256);
        __half* bh = (__half*)p; p += align_up(nB * 2, 256);
        __half* bl = (__half*)p; p += align_up(nB * 2, 256);
        if (!(dual & 4)) {  // dual & 4 (microbenchmarks): the workspace already holds the split operands of an earlier call
            h_split_kernel<<<296, 256, 0, st>>>(A, ah, al, nA, s ? 4.0f * sa : sa);
            h_split_kernel<<<296, 256, 0, st>>>(B, bh, bl, nB, s ? sb / 8.0f : sb);
        }
        g.A[s] = HOperand{ah, al, lda, bsA, batch, M, K, scales + 2 * s};
        g.B[s] = HOperand{bh, bl, ldb, bsB, batch, N, K, scales + 2 * s + 1};
    }
    EEGAN_LAUNCH_CHECK("gemm_f16x3 split");
    g.C = C; g.ldc = ldc; g.bC = bsC; g.M = M; g.N = N; g.batch = batch; g.nred = 1; g.red_total = 0;
    g.epi = TC_EPI_PLAIN;
    return h_gemm_launch(g, st);
}
