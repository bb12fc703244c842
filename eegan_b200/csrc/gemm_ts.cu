// tcgen05 3xTF32 GEMM with the A operand staged in tensor memory ("TS" form of tcgen05.mma).
//
// Why: with both operands in shared memory, the three MMAs of a 3xTF32 K-step re-read A and B
// tiles six times, and the hi/lo split adds a read and a write of both tiles — at 128x128 tiles the
// MMA operand fetch alone saturates the 128 B/clk shared-memory port, so the kernel ran at ~1/3 of
// the tensor pipe.  Here the A tile (always MN-major in the pair-grid pipeline: regions / packed
// columns / channels contiguous) lands un-swizzled in shared memory, four splitter warps read it
// once — thread = row m, 32 k values — and write hi = trunc_tf32(x) and lo = tf32(x - hi) straight
// into TMEM with tcgen05.st; the MMAs then take A from TMEM and only B (K-major, 128B swizzle) from
// shared memory.  Shared-memory traffic per 128x128x32 block drops from ~192 KB to ~128 KB, and
// the freed stage space pays for a 4-deep TMA ring.
//
//   warp 0      TMA producer (A box [32 k][128 m] plain, B box [128 n][32 k] SWIZZLE_128B)
//   warp 1      TMEM allocator + single-thread MMA issuer
//   warps 2-5   A splitters: smem -> registers -> TMEM (lane quarter = warp & 3)
//   warps 6-9   B splitters: lo tile in shared memory (hi = the raw tile, hardware truncation)
//   warps 10-13 epilogue (tc_device.cuh)
// TMEM columns: [0,256) two accumulators, [256,512) four A stages of 32 hi + 32 lo columns.
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"

#ifdef EEGAN_DEBUG_SWITCHES  // work-skipping timing switches: compiled out of the shipped library
#define TS_DBG(p) ((p).dbg)
#else
#define TS_DBG(p) 0
#endif

namespace eegan {

// The attention epilogues are the long pole of the K = D contractions (exp-heavy per-region work on a
// 128 x 128 tile against only 8 k-blocks of MMA), so those instantiations run 8 epilogue warps — two per
// TMEM lane quarter, one 64-column bin each — paid for with a 3-deep ring and 2 B-splitter warps.
template <int EPI>
struct TsCfg {
    static constexpr bool kAttn = EPI != TC_EPI_PLAIN;
    static constexpr int kStages = kAttn ? 3 : 4;
    static constexpr int kBWarps = kAttn ? 2 : 4;   // B-splitter warps
    static constexpr int kEWarps = kAttn ? 8 : 4;   // epilogue warps
    static constexpr int kEpi0 = 6 + kBWarps;       // first epilogue warp
    static constexpr int kThreads = 32 * (kEpi0 + kEWarps);
    static constexpr int kSmem = kStages * 3 * TC_TILE_BYTES + kEWarps * (32 * TC_EPI_PITCH * 4 + 256) + 1024 /*align*/ + 256 /*barriers*/;
};
constexpr int TS_STAGE_BYTES = 3 * TC_TILE_BYTES;  // A raw, B raw (= hi), B lo
constexpr int TS_A_COL0 = 2 * TC_BN;                // first TMEM column of the A stages
static_assert(TS_A_COL0 + 4 * 2 * TC_BK <= TC_TMEM_COLS, "TMEM budget");
static_assert(TsCfg<TC_EPI_PLAIN>::kSmem <= 232448 && TsCfg<TC_EPI_ATTN_FWD>::kSmem <= 232448, "shared memory budget");
static_assert((TsCfg<TC_EPI_ATTN_FWD>::kEpi0 & 3) == 0 && (TsCfg<TC_EPI_PLAIN>::kEpi0 & 3) == 2, "epilogue warps must cover the four lane quarters");

template <int EPI>
__global__ void __launch_bounds__(TsCfg<EPI>::kThreads, 1)
ts_gemm_kernel(const __grid_constant__ TcMaps tm, const TcArgs p) {
    using Cfg = TsCfg<EPI>;
    constexpr int TS_STAGES = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Mlive = p.dynM ? min(*p.dynM, p.M) : p.M;
    const int Nlive = p.dynN ? min(*p.dynN, p.N) : p.N;
    const int mt = (Mlive + TC_BM - 1) / TC_BM, nt = (Nlive + TC_BN - 1) / TC_BN;
    const int ntiles = mt * nt * p.batch;
    if ((int)blockIdx.x >= ntiles) return;  // uniform: before any barrier / TMEM state exists

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_stage = base + TS_STAGES * TS_STAGE_BYTES;
    const uint32_t epi_czs = epi_stage + Cfg::kEWarps * (32 * TC_EPI_PITCH * 4);
    const uint32_t bars = epi_czs + Cfg::kEWarps * 256u;
    auto full = [&](int s) { return bars + 8u * s; };
    auto conv = [&](int s) { return bars + 8u * (TS_STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * TS_STAGES + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (3 * TS_STAGES + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (3 * TS_STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * TS_STAGES + 4);

    int kb0 = 0, kb1 = 0;
    {
        const int K0 = p.dynK ? min(*p.dynK, p.K[0]) : p.K[0];
        kb0 = (K0 + TC_BK - 1) / TC_BK;
        if (p.nseg > 1) {
            const int K1 = p.dynK ? min(*p.dynK, p.K[1]) : p.K[1];
            kb1 = (K1 + TC_BK - 1) / TC_BK;
        }
    }
    const int kbt = kb0 + kb1;
    auto tile_total = [&](int z) {  // k-blocks of the tile whose batch index is z
        int nred = p.nred;
        if (p.red_total > 0) nred = max(0, min(p.nred, p.red_total - z * p.nred));
        return nred * kbt;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_STAGES; ++s) {
            mbar_init(full(s), 1);
            mbar_init(conv(s), 4 + Cfg::kBWarps);  // A-splitter + B-splitter warps
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), Cfg::kEWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;  // running k-block counter across tiles: stage ring position
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int z = t / (mt * nt), rem_t = t - z * (mt * nt);
                const int m0 = (rem_t / nt) * TC_BM, n0 = (rem_t % nt) * TC_BN;
                const int total = tile_total(z);
                for (int k = 0; k < total; ++k, ++it) {
                    const int s = it % TS_STAGES, ph = (it / TS_STAGES) & 1;
                    const int red = k / kbt, rem = k - red * kbt;
                    const int seg = rem >= kb0 ? 1 : 0;
                    const int k0 = (seg ? rem - kb0 : rem) * TC_BK;
                    const int zr = z * p.nred + red;
                    const int zA = p.a_batched[seg] ? zr : 0, zB = p.b_batched[seg] ? zr : 0;
                    const bool bpre = p.b_pre[seg];
                    mbar_wait(empty(s), ph ^ 1);
                    mbar_arrive_expect_tx(full(s), (uint32_t)((2 + (bpre ? 1 : 0)) * TC_TILE_BYTES));
                    const uint32_t sA = base + s * TS_STAGE_BYTES, sB = sA + TC_TILE_BYTES;
                    tma_load_3d(sA, &tm.m[seg][0][0], full(s), m0, k0, zA);
                    tma_load_3d(sB, &tm.m[seg][1][0], full(s), k0, n0, zB);
                    if (bpre) tma_load_3d(sB + TC_TILE_BYTES, &tm.m[seg][1][1], full(s), k0, n0, zB);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) /*D=f32*/ | (2u << 7) /*A=tf32*/ | (2u << 10) /*B=tf32*/ |
                                       ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);  // A, B both K-major
            int it = 0, ti = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
                const int z = t / (mt * nt);
                const int total = tile_total(z);
                const int acc = ti & 1;
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_BN);
                mbar_wait(tmem_empty(acc), ((ti >> 1) & 1) ^ 1);  // the epilogue drained this accumulator
                tc_fence_after();
                for (int k = 0; k < total; ++k, ++it) {
                    const int s = it % TS_STAGES, ph = (it / TS_STAGES) & 1;
                    mbar_wait(conv(s), ph);
                    tc_fence_after();
                    const uint32_t b_hi = base + s * TS_STAGE_BYTES + TC_TILE_BYTES, b_lo = b_hi + TC_TILE_BYTES;
                    const uint32_t a_hi = tmem_base + (uint32_t)(TS_A_COL0 + s * 2 * TC_BK), a_lo = a_hi + TC_BK;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        const uint64_t dbh = umma_desc(b_hi, true, ks, 0, 0), dbl = umma_desc(b_lo, true, ks, 0, 0);
                        tc_mma_tf32_ts(tmem_d, a_lo + ks * 8, dbh, idesc, (k > 0 || ks > 0) ? 1u : 0u);
                        if (TS_DBG(p) & 1) continue;
                        tc_mma_tf32_ts(tmem_d, a_hi + ks * 8, dbl, idesc, 1u);
                        tc_mma_tf32_ts(tmem_d, a_hi + ks * 8, dbh, idesc, 1u);
                    }
                    tc_commit(empty(s));  // implies tcgen05.fence::before_thread_sync; also frees the A columns
                }
                tc_commit(tmem_full(acc));
            }
        }
    } else if (warp < 6) {
        // ===== A splitters: shared memory [32 k][128 m] -> TMEM [lane m][32 hi | 32 lo columns] =====
        const int quarter = warp & 3;
        const uint32_t my_m = (uint32_t)(quarter * 32 + lane) * 4u;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int total = tile_total(t / (mt * nt));
            for (int k = 0; k < total; ++k, ++it) {
                const int s = it % TS_STAGES, ph = (it / TS_STAGES) & 1;
                mbar_wait(full(s), ph);
                const uint32_t sA = base + s * TS_STAGE_BYTES + my_m;
                if (TS_DBG(p) & 2) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(conv(s));
                    continue;
                }
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float x = lds_f32(sA + (uint32_t)q * (TC_BM * 4u));
                    const float h = trunc_tf32(x);
                    hi[q] = __float_as_uint(h);
                    lo[q] = __float_as_uint(to_tf32(x - h));
                }
                const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(TS_A_COL0 + s * 2 * TC_BK);
                tmem_st32(ta, hi);
                tmem_st32(ta + TC_BK, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(conv(s));
            }
        }
    } else if (warp < Cfg::kEpi0) {
        // ===== B splitters: lo tile, position-preserving (swizzle-agnostic) =====
        const int ctid = threadIdx.x - 6 * 32;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int total = tile_total(t / (mt * nt));
            for (int k = 0; k < total; ++k, ++it) {
                const int s = it % TS_STAGES, ph = (it / TS_STAGES) & 1;
                const int seg = (k % kbt) >= kb0 ? 1 : 0;
                mbar_wait(full(s), ph);
                if (!p.b_pre[seg] && !(TS_DBG(p) & 4)) {
                    const uint32_t hi = base + s * TS_STAGE_BYTES + TC_TILE_BYTES, lo = hi + TC_TILE_BYTES;
                    constexpr int NF = (TC_TILE_BYTES / 16) / (32 * Cfg::kBWarps);  // float4 per thread
                    float4 v[NF];
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        const uint32_t off = (uint32_t)(i * 32 * Cfg::kBWarps + ctid) * 16u;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(hi + off));
                    }
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        const uint32_t off = (uint32_t)(i * 32 * Cfg::kBWarps + ctid) * 16u;
                        float4 l;
                        l.x = to_tf32(v[i].x - trunc_tf32(v[i].x)); l.y = to_tf32(v[i].y - trunc_tf32(v[i].y));
                        l.z = to_tf32(v[i].z - trunc_tf32(v[i].z)); l.w = to_tf32(v[i].w - trunc_tf32(v[i].w));
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(conv(s));
            }
        }
    } else {
        // ===== epilogue (last four warps) =====
        EpiTile et;
        et.quarter = warp & 3;  // TMEM lane quarter this warp may access
        et.half = Cfg::kEWarps == 8 ? (warp - Cfg::kEpi0) >> 2 : -1;
        et.stage = epi_stage + (uint32_t)(warp - Cfg::kEpi0) * (32 * TC_EPI_PITCH * 4);
        et.czs = epi_czs + (uint32_t)(warp - Cfg::kEpi0) * 256u;
        et.Mlive = Mlive;
        et.Nlive = Nlive;
        int ti = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
            const int rem_t = t % (mt * nt);
            const int acc = ti & 1;
            et.z = t / (mt * nt);
            et.m0 = (rem_t / nt) * TC_BM;
            et.n0 = (rem_t % nt) * TC_BN;
            et.total = tile_total(et.z);
            et.tacc = tmem_base + ((uint32_t)(et.quarter * 32) << 16) + (uint32_t)(acc * TC_BN);
            et.full_bar = tmem_full(acc);
            et.full_parity = (uint32_t)((ti >> 1) & 1);
            et.empty_bar = tmem_empty(acc);
            tc_epilogue_tile<EPI>(p, et, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

template <int EPI>
static int ts_launch(const TcMaps& maps, const TcArgs& a, unsigned grid, cudaStream_t st) {
    static SmemGrant grant;
    if (int rc = grant_dyn_smem(ts_gemm_kernel<EPI>, (size_t)TsCfg<EPI>::kSmem, grant, "ts gemm")) return rc;
    ts_gemm_kernel<EPI><<<grid, TsCfg<EPI>::kThreads, TsCfg<EPI>::kSmem, st>>>(maps, a);
    return check_launch("ts gemm");
}

int ts_gemm_dispatch(const TcMaps& maps, const TcArgs& args, unsigned grid, int epi, cudaStream_t st) {
    EEGAN_REQUIRE(!args.a_pre[0] && !args.a_pre[1], "ts gemm: the A operand is split on its way into TMEM; pre-split A is not taken");
    if (epi == TC_EPI_ATTN_FWD) return ts_launch<TC_EPI_ATTN_FWD>(maps, args, grid, st);
    if (epi == TC_EPI_ATTN_BWD) return ts_launch<TC_EPI_ATTN_BWD>(maps, args, grid, st);
    return ts_launch<TC_EPI_PLAIN>(maps, args, grid, st);
}

}  // namespace eegan
