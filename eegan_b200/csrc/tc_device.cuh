// Device-side pieces shared by the two tcgen05 3xTF32 GEMM kernels (gemm_tc.cu: both operands
// from shared memory; gemm_ts.cu: A operand staged in tensor memory): PTX wrappers, kernel
// argument block, and the epilogues (plain store / word-region attention forward / backward).
#pragma once
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace eegan {

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem: lane = row m, 8 consecutive 32-bit columns = K] * B[smem desc]
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the tensor core reads only the top 19 bits of an fp32 operand (hardware truncation to tf32)
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// round-to-nearest (ties away) to the 10-bit tf32 mantissa with two integer ops; same result as
// cvt.rna.tf32.f32 for finite inputs (inf/nan inputs poison the output either way)
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

// Canonical 128B-swizzle UMMA shared-memory descriptor (version 1 = Blackwell).
//   K-major : rows at 128 B, 8-row groups at SBO = 1024 B; a K-step of 8 fp32 advances the start by 32 B.
//   MN-major: 32-bit operands only exist in the "128B swizzle, 32B atomicity" layout (descriptor
//             layout type 1, TMA SWIZZLE_128B_ATOM_32B): [k][32 fp32] rows of 128 B, atoms of 4
//             k-rows (SBO = 512 B between 4-row groups), 32-wide MN chunks LBO = 4096 B apart
//             (one TMA box of 32 k-rows each); a K-step of 8 advances the start by 1024 B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t tile, bool kmajor, int kstep, uint32_t mn_lbo, uint32_t mn_sbo) {
    const uint32_t start = tile + (kmajor ? kstep * 32 : kstep * 1024);
    const uint64_t lbo = kmajor ? 1 : mn_lbo;
    const uint64_t sbo = kmajor ? (1024 >> 4) : mn_sbo;
    uint64_t d = (uint64_t)((start & 0x3FFFF) >> 4);
    d |= lbo << 16;
    d |= sbo << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    d |= (uint64_t)(kmajor ? 2 : 1) << 61;  // SWIZZLE_128B (K-major) / SWIZZLE_128B_BASE32B (MN-major tf32)
    return d;
}

// [segment][0 = A, 1 = B][0 = raw / hi, 1 = pre-split lo]
struct TcMaps {
    CUtensorMap m[2][2][2];
};

struct TcArgs {
    float* C;
    long long ldc, bC;
    int M, N;
    const int* dynM;
    const int* dynN;
    const int* dynK;
    int K[2];
    int nseg, nred, red_total, batch;
    int a_batched[2], b_batched[2];
    int a_pre[2], b_pre[2];   // operand arrives pre-split (raw + lo arrays): no in-kernel split for it
    int dbg;                  // timing experiments only (EEGAN_TS_DBG): 1 one MMA per K-step, 2 skip the A split, 4 skip the B split
    int trunc_hi;             // 1: leave the raw operand as hi (hardware truncation), write lo only
    uint32_t mn_lbo, mn_sbo;  // debug-overridable descriptor fields of MN-major tiles (16-byte units)
    TcAttnEpi attn;
};

#define TC_LD_REGS8(v, o) \
    "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define TC_ST_REGS8(v, o) \
    "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
// this warp's 32 TMEM lanes x NV consecutive fp32 columns starting at taddr -> v[0..NV-1]
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : TC_LD_REGS8(v, 0), TC_LD_REGS8(v, 8), TC_LD_REGS8(v, 16), TC_LD_REGS8(v, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TC_LD_REGS8(v, 0), TC_LD_REGS8(v, 8)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : TC_LD_REGS8(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// v[0..31] -> this warp's 32 TMEM lanes x 32 consecutive columns starting at taddr
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        TC_ST_REGS8(v, 0), TC_ST_REGS8(v, 8), TC_ST_REGS8(v, 16), TC_ST_REGS8(v, 24)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// exp(x) = ex2(x log2 e), flush-to-zero: two instructions (the default __expf adds a denormal-range fix-up)
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------
// epilogues: one call handles one output tile for one of the four epilogue warps
// ---------------------------------------------------------------------------------------
struct EpiTile {
    int z, m0, n0;        // batch index and tile origin
    int total;            // k-blocks accumulated into the tile (0 = nothing: store zeros)
    int Mlive, Nlive;
    uint32_t tacc;        // TMEM address of the accumulator, this warp's lane quarter
    uint32_t stage;       // this warp's staging buffer: [32][TC_EPI_PITCH] floats
    uint32_t czs;         // this warp's 64-float scratch (attention backward)
    uint32_t full_bar;    // mbarrier: accumulator complete
    uint32_t full_parity;
    uint32_t empty_bar;   // mbarrier: accumulator drained (one arrival per epilogue warp)
    int quarter;          // TMEM lane quarter of this warp
    int half;             // attention epilogues: the 64-column bin of the tile this warp handles (0 / 1), -1 = both
};

// one caption of NV <= 32 live-or-padding columns (T in (NV-4, NV]), forward: P = softmax over its words
template <int NV>
__device__ __forceinline__ void attn_fwd_caption(const uint32_t (&vr)[32], int T, uint32_t cap_row) {
    float x[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) x[q] = (q < NV - 3 || q < T) ? __uint_as_float(vr[q]) : -INFINITY;
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
    const float inv = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (q < NV - 3 || q < T) sts_f32(cap_row + (uint32_t)q * 4u, x[q] * inv);
}

// backward: dS = v - P sum_t v,  v = g1 P E (acc - csz),  E = exp(g1 (P - 1))   (acc = dA / Z, csz = (sum_r A dA) / Z)
template <int NV>
__device__ __forceinline__ void attn_bwd_caption(const uint32_t (&vr)[32], int T, uint32_t cap_row, uint32_t cz_row, float g1) {
    float pv[NV], vv[NV];
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const bool ok = (q < NV - 3 || q < T);
        pv[q] = ok ? lds_f32(cap_row + (uint32_t)q * 4u) : 0.f;
        const float cz = lds_f32(cz_row + (uint32_t)q * 4u);  // same address in every lane: broadcast
        const float ev = fast_exp(g1 * (pv[q] - 1.0f));
        const float t = g1 * pv[q] * ev * (__uint_as_float(vr[q]) - cz);
        vv[q] = ok ? t : 0.f;
        s4[q & 3] += vv[q];
    }
    const float qs = (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (q < NV - 3 || q < T) sts_f32(cap_row + (uint32_t)q * 4u, vv[q] - pv[q] * qs);
}

template <int EPI>
__device__ __forceinline__ void attn_caption(uint32_t taddr, int T, uint32_t cap_row, uint32_t cz_row, float g1) {
    uint32_t vr[32];
    const int bucket = (T + 3) >> 2;  // 1..8, warp-uniform
    if (bucket <= 2) tmem_ld8(taddr, vr);
    else if (bucket <= 4) tmem_ld16(taddr, vr);
    else tmem_ld32(taddr, vr);
    tmem_ld_wait();
#define TC_CAP(NV)                                                                  \
    case (NV) / 4:                                                                  \
        if (EPI == TC_EPI_ATTN_FWD) attn_fwd_caption<NV>(vr, T, cap_row);           \
        else attn_bwd_caption<NV>(vr, T, cap_row, cz_row, g1);                      \
        break;
    switch (bucket) {
        TC_CAP(4) TC_CAP(8) TC_CAP(12) TC_CAP(16) TC_CAP(20) TC_CAP(24) TC_CAP(28) TC_CAP(32)
        default: break;
    }
#undef TC_CAP
}

template <int EPI>
__device__ __forceinline__ void tc_epilogue_tile(const TcArgs& p, const EpiTile& t, int lane) {
    float* Cz = p.C + (long long)t.z * p.bC;
    const int row0 = t.m0 + t.quarter * 32;
    const int rows_live = max(0, min(32, t.Mlive - row0));  // rows of this warp that exist
    if (EPI == TC_EPI_PLAIN) {
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_after();
        // TMEM -> registers -> per-warp smem transpose -> coalesced global rows
        // with two warps per lane quarter (half = 0 / 1) each takes two of the four 32-column chunks
        const int c_lo = t.half < 0 ? 0 : t.half * (TC_BN / 64), c_hi = t.half < 0 ? TC_BN / 32 : c_lo + TC_BN / 64;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            uint32_t v[32];
            if (t.total > 0) {
                tmem_ld32(t.tacc + (uint32_t)(c * 32), v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = 0u;
            }
            if (c == c_hi - 1) {  // this warp's share of the accumulator is read: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t.empty_bar);
            }
            __syncwarp();  // previous chunk's readers are done with the staging buffer
#pragma unroll
            for (int q = 0; q < 32; ++q) sts_f32(t.stage + (uint32_t)(lane * 33 + q) * 4u, __uint_as_float(v[q]));
            __syncwarp();
            const int gn = t.n0 + c * 32 + lane;
            if (gn < t.Nlive) {
                float* dst = Cz + (long long)row0 * p.ldc + gn;
#pragma unroll 1
                for (int r8 = 0; r8 < rows_live; r8 += 8) {  // 8 rows of shared loads in flight, then 8 row stores
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
    const TcAttnEpi& e = p.attn;
    const int nbins = *e.nbins;
    const uint32_t my_row = t.stage + (uint32_t)(lane * TC_EPI_PITCH) * 4u;
    float* Pz = e.P + (long long)t.z * p.bC;
    const int b0 = t.n0 >> 6;
    // bin metadata for both bins of the tile, one coalesced load, before the accumulator is needed
    const int bc_l = e.bin_cap[min(b0 + lane, nbins)];
    const int bu_l = (lane < 2 && b0 + lane < nbins) ? e.bin_used[b0 + lane] : 0;
    // P rows of this warp's regions -> staging (coalesced 128-byte rows, 8 rows of loads in flight)
    auto load_p_bin = [&](int col0) {
#pragma unroll 1
        for (int r8 = 0; r8 < rows_live; r8 += 8) {
            float a0[8], a1[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int rr = min(r8 + q, rows_live - 1);
                const float* src = Pz + (long long)(row0 + rr) * p.ldc + col0 + lane;
                a0[q] = __ldg(src);
                a1[q] = __ldg(src + 32);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                sts_f32(t.stage + (uint32_t)((r8 + q) * TC_EPI_PITCH + lane) * 4u, a0[q]);
                sts_f32(t.stage + (uint32_t)((r8 + q) * TC_EPI_PITCH + lane + 32) * 4u, a1[q]);
            }
        }
    };
    __syncwarp();  // the previous tile's read-out is done with the staging buffer
    const int h_lo = t.half < 0 ? 0 : t.half, h_hi = t.half < 0 ? 2 : t.half + 1;
    if (EPI == TC_EPI_ATTN_BWD && b0 + h_lo < nbins) load_p_bin(t.n0 + 64 * h_lo);
    bool waited = false;
    if (b0 + h_lo >= nbins) {  // no live bin for this warp in the tile: only keep the accumulator hand-shake going
        mbar_wait(t.full_bar, t.full_parity);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t.empty_bar);
        return;
    }
#pragma unroll 1
    for (int h = h_lo; h < h_hi; ++h) {
        const int b = b0 + h;
        if (b >= nbins) break;  // warp-uniform
        const int i0 = __shfl_sync(0xffffffffu, bc_l, h), i1 = __shfl_sync(0xffffffffu, bc_l, h + 1);
        const int used = __shfl_sync(0xffffffffu, bu_l, h);
        const int col0 = t.n0 + 64 * h;  // first packed column of the bin
        const bool last = (h == h_hi - 1) || (b + 1 >= nbins);
        const bool c0ok = lane < used, c1ok = lane + 32 < used;
        if (EPI == TC_EPI_ATTN_BWD) {
            if (h > h_lo) {
                __syncwarp();  // bin 0's read-out is done with the staging buffer
                load_p_bin(col0);
            }
            sts_f32(t.czs + (uint32_t)lane * 4u, __ldg(e.csz + (long long)t.z * p.ldc + col0 + lane));
            sts_f32(t.czs + (uint32_t)(lane + 32) * 4u, __ldg(e.csz + (long long)t.z * p.ldc + col0 + lane + 32));
        }
        __syncwarp();
#pragma unroll 1
        for (int ic = i0; ic < i1; ic += 32) {
            // caption metadata for up to 32 captions: one coalesced load, broadcast by shuffle below
            const int nc = min(32, i1 - ic);
            const int myT = lane < nc ? e.cap_len[ic + lane] : 0;
            const int myC = lane < nc ? e.col_start[ic + lane] - col0 : 0;
            if (!waited) {  // first use of the accumulator in this tile
                mbar_wait(t.full_bar, t.full_parity);
                tc_fence_after();
                waited = true;
            }
#pragma unroll 1
            for (int ci = 0; ci < nc; ++ci) {
                const int T = __shfl_sync(0xffffffffu, myT, ci);
                if (T <= 0) continue;
                const int cl = __shfl_sync(0xffffffffu, myC, ci);  // bin-local first column of the caption
                attn_caption<EPI>(t.tacc + (uint32_t)(64 * h + cl), T, my_row + (uint32_t)cl * 4u, t.czs + (uint32_t)cl * 4u, e.g1);
            }
        }
        if (!waited) {  // live bins without a caption cannot happen with the packer; keeps the barrier protocol sound
            mbar_wait(t.full_bar, t.full_parity);
            tc_fence_after();
            waited = true;
        }
        if (last) {  // accumulator fully read: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t.empty_bar);
        }
        for (int c = used; c < 64; ++c) sts_f32(my_row + (uint32_t)c * 4u, 0.f);  // padding columns of the bin
        __syncwarp();
        // read-out: region rows (2 x 128 bytes each), coalesced, 4 rows per round
        if (EPI == TC_EPI_ATTN_FWD) {
            float z0 = 0.f, z1 = 0.f;
#pragma unroll 1
            for (int r4 = 0; r4 < rows_live; r4 += 4) {
                float p0[4], p1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    p0[q] = lds_f32(t.stage + (uint32_t)((r4 + q) * TC_EPI_PITCH + lane) * 4u);
                    p1[q] = lds_f32(t.stage + (uint32_t)((r4 + q) * TC_EPI_PITCH + lane + 32) * 4u);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (r4 + q < rows_live) {
                        // E = exp(g1 (P - 1)): the region softmax of g1 P (:53-54) without its normaliser
                        const float e0 = c0ok ? fast_exp(e.g1 * (p0[q] - 1.0f)) : 0.f;
                        const float e1 = c1ok ? fast_exp(e.g1 * (p1[q] - 1.0f)) : 0.f;
                        const long long o = (long long)(row0 + r4 + q) * p.ldc + col0 + lane;
                        Pz[o] = p0[q];
                        Pz[o + 32] = p1[q];
                        Cz[o] = e0;
                        Cz[o + 32] = e1;
                        z0 += e0;
                        z1 += e1;
                    }
                }
            }
            float* zp = e.Zpart + ((long long)t.z * ((p.M + 31) / 32) + (row0 >> 5)) * p.ldc + col0 + lane;
            if (rows_live > 0) {
                zp[0] = z0;
                zp[32] = z1;
            }
        } else {
#pragma unroll 1
            for (int r4 = 0; r4 < rows_live; r4 += 4) {
                float d0[4], d1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    d0[q] = lds_f32(t.stage + (uint32_t)((r4 + q) * TC_EPI_PITCH + lane) * 4u);
                    d1[q] = lds_f32(t.stage + (uint32_t)((r4 + q) * TC_EPI_PITCH + lane + 32) * 4u);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (r4 + q < rows_live) {
                        const long long o = (long long)(row0 + r4 + q) * p.ldc + col0 + lane;
                        Cz[o] = d0[q];
                        Cz[o + 32] = d1[q];
                    }
            }
        }
    }
}

}  // namespace eegan
