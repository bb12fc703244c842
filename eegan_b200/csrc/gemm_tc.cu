// tcgen05 3xTF32 GEMM engine — see gemm_tc.cuh for the design.
#include <stdlib.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"

namespace eegan {

// Persistent, warp-specialised kernel: grid = min(#live tiles, #SMs); every CTA walks tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, then m, then batch z).
//   warp 0      TMA producer            warp 1      TMEM alloc + MMA issuer
//   warps 2-9   hi/lo splitters         warps 10-13 epilogue (TMEM lane quarter = warp & 3)
// The smem stage ring runs continuously across tiles; two TMEM accumulators (2 x 128 columns)
// let the epilogue of tile i overlap the main loop of tile i+1.
template <bool A_K, bool B_K, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ TcMaps tm, const TcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Mlive = p.dynM ? min(*p.dynM, p.M) : p.M;
    const int Nlive = p.dynN ? min(*p.dynN, p.N) : p.N;
    const int mt = (Mlive + TC_BM - 1) / TC_BM, nt = (Nlive + TC_BN - 1) / TC_BN;
    const int ntiles = mt * nt * p.batch;
    if ((int)blockIdx.x >= ntiles) return;  // uniform: before any barrier / TMEM state exists

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_stage = base + TC_STAGES * TC_STAGE_BYTES;       // 4 warps x [32][TC_EPI_PITCH] floats
    const uint32_t epi_czs = epi_stage + TC_EPI_STAGE_BYTES;            // 4 warps x 64 floats
    const uint32_t bars = epi_czs + 1024u;
    auto full = [&](int s) { return bars + 8u * s; };
    auto conv = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };
    auto tmem_full = [&](int a) { return bars + 8u * (3 * TC_STAGES + a); };
    auto tmem_empty = [&](int a) { return bars + 8u * (3 * TC_STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * TC_STAGES + 4);

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
    // k-blocks of the tile whose batch index is z
    auto tile_total = [&](int z) {
        int nred = p.nred;
        if (p.red_total > 0) nred = max(0, min(p.nred, p.red_total - z * p.nred));
        return nred * kbt;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full(s), 1);
            mbar_init(conv(s), TC_SPLIT_WARPS);
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full(a), 1);
            mbar_init(tmem_empty(a), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // all 512 columns: the attention epilogues read 32-column windows that may overhang an accumulator
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
                    const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    const int red = k / kbt, rem = k - red * kbt;
                    const int seg = rem >= kb0 ? 1 : 0;
                    const int k0 = (seg ? rem - kb0 : rem) * TC_BK;
                    const int zr = z * p.nred + red;
                    const int zA = p.a_batched[seg] ? zr : 0, zB = p.b_batched[seg] ? zr : 0;
                    const bool apre = p.a_pre[seg], bpre = p.b_pre[seg];
                    mbar_wait(empty(s), ph ^ 1);
                    mbar_arrive_expect_tx(full(s), (uint32_t)((2 + (apre ? 1 : 0) + (bpre ? 1 : 0)) * TC_TILE_BYTES));
                    const uint32_t sA = base + s * TC_STAGE_BYTES, sB = sA + TC_TILE_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {  // h = 1: the pre-split lo arrays land directly in the lo tiles
                        if (h == 0 || apre) {
                            const CUtensorMap* ta = &tm.m[seg][0][h];
                            const uint32_t dst = sA + h * 2 * TC_TILE_BYTES;
                            if (A_K) {
                                tma_load_3d(dst, ta, full(s), k0, m0, zA);
                            } else {
#pragma unroll
                                for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * 4096, ta, full(s), m0 + 32 * c, k0, zA);
                            }
                        }
                        if (h == 0 || bpre) {
                            const CUtensorMap* tb = &tm.m[seg][1][h];
                            const uint32_t dst = sB + h * 2 * TC_TILE_BYTES;
                            if (B_K) {
                                tma_load_3d(dst, tb, full(s), k0, n0, zB);
                            } else {
#pragma unroll
                                for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * 4096, tb, full(s), n0 + 32 * c, k0, zB);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) /*D=f32*/ | (2u << 7) /*A=tf32*/ | (2u << 10) /*B=tf32*/ |
                                       ((A_K ? 0u : 1u) << 15) | ((B_K ? 0u : 1u) << 16) |
                                       ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int it = 0, ti = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
                const int z = t / (mt * nt);
                const int total = tile_total(z);
                const int acc = ti & 1;
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_BN);
                mbar_wait(tmem_empty(acc), ((ti >> 1) & 1) ^ 1);  // the epilogue drained this accumulator
                tc_fence_after();
                for (int k = 0; k < total; ++k, ++it) {
                    const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    mbar_wait(conv(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * TC_STAGE_BYTES, b_hi = a_hi + TC_TILE_BYTES;
                    const uint32_t a_lo = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        const uint64_t dah = umma_desc(a_hi, A_K, ks, p.mn_lbo, p.mn_sbo), dal = umma_desc(a_lo, A_K, ks, p.mn_lbo, p.mn_sbo);
                        const uint64_t dbh = umma_desc(b_hi, B_K, ks, p.mn_lbo, p.mn_sbo), dbl = umma_desc(b_lo, B_K, ks, p.mn_lbo, p.mn_sbo);
                        tc_mma_tf32(tmem_d, dal, dbh, idesc, (k > 0 || ks > 0) ? 1u : 0u);
                        tc_mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                        tc_mma_tf32(tmem_d, dah, dbh, idesc, 1u);
                    }
                    tc_commit(empty(s));  // implies tcgen05.fence::before_thread_sync
                }
                tc_commit(tmem_full(acc));
            }
        }
    } else if (warp < 2 + TC_SPLIT_WARPS) {
        // ===== hi/lo splitters =====
        const int ctid = threadIdx.x - 64;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int total = tile_total(t / (mt * nt));
            for (int k = 0; k < total; ++k, ++it) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                const int seg = (k % kbt) >= kb0 ? 1 : 0;
                // float4 range of the stage that still needs splitting: A tile = [0,1024), B tile = [1024,2048)
                const int f0 = p.a_pre[seg] ? TC_TILE_BYTES / 16 : 0;
                const int f1 = p.b_pre[seg] ? TC_TILE_BYTES / 16 : 2 * TC_TILE_BYTES / 16;
                mbar_wait(full(s), ph);
                const uint32_t hi = base + s * TC_STAGE_BYTES, lo = hi + 2 * TC_TILE_BYTES;
#pragma unroll 4
                for (int f = f0 + ctid; f < f1; f += 32 * TC_SPLIT_WARPS) {
                    const uint32_t off = (uint32_t)f * 16u;
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(hi + off));
                    float4 h, l;
                    if (p.trunc_hi) {
                        // the tensor core reads only the top 19 bits of an fp32 operand: the raw tile IS hi = trunc(x);
                        // only lo = tf32(x - trunc(x)) is written
                        h.x = trunc_tf32(v.x); h.y = trunc_tf32(v.y); h.z = trunc_tf32(v.z); h.w = trunc_tf32(v.w);
                    } else {
                        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(hi + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                    }
                    l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                }
                if (f0 < f1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
                __syncwarp();
                if (lane == 0) mbar_arrive(conv(s));
            }
        }
    } else {
        // ===== epilogue (last four warps) =====
        EpiTile et;
        et.quarter = warp & 3;  // TMEM lane quarter this warp may access
        et.half = -1;
        et.stage = epi_stage + (uint32_t)(warp - 2 - TC_SPLIT_WARPS) * (32 * TC_EPI_PITCH * 4);
        et.czs = epi_czs + (uint32_t)(warp - 2 - TC_SPLIT_WARPS) * 256u;
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

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is a driver entry point: it needs the primary context bound to the calling
// thread.  A backward can be the first CUDA work of autograd's thread (no runtime call has bound it
// yet), so bind it once per thread through the runtime.
static void bind_context_once() {
    static thread_local bool bound = false;
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
}

static EncodeTiledFn get_encode() {
    bind_context_once();
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Encoding a tensor map is pure host work (a few microseconds); the same handful of operands
// recurs every step (the workspace comes back at the same address from the caching allocator),
// so the last encodes are memoised per host thread.
struct MapKey {
    const float* ptr;
    long long ld, bstride;
    int kmajor, nbatch, rows, K, box_rows, plain;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && ld == o.ld && bstride == o.bstride && kmajor == o.kmajor && nbatch == o.nbatch &&
               rows == o.rows && K == o.K && box_rows == o.box_rows && plain == o.plain;
    }
};
struct MapSlot {
    MapKey key;
    CUtensorMap map;
    bool used;
};
static thread_local MapSlot g_map_cache[32];
static thread_local int g_map_next = 0;

// plain = true: un-swizzled [K][128 rows] box of an MN-major operand (read by threads, not by the tensor core: gemm_ts.cu)
static int make_map(CUtensorMap* m, const TcOperand& o, int box_rows_kmajor, bool plain = false) {
    const MapKey key{o.ptr, o.ld, o.bstride, o.kmajor, o.nbatch, o.rows, o.K, box_rows_kmajor, plain ? 1 : 0};
    for (int i = 0; i < 32; ++i)
        if (g_map_cache[i].used && g_map_cache[i].key == key) {
            *m = g_map_cache[i].map;
            return EEGAN_OK;
        }
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("tc gemm: cuTensorMapEncodeTiled unavailable"); return EEGAN_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(o.ptr) & 15) || (o.ld % 4) || (o.bstride % 4)) {
        set_error("tc gemm: operand base/pitch must be 16-byte aligned (ptr=%p ld=%lld bstride=%lld)", (const void*)o.ptr, o.ld, o.bstride);
        return EEGAN_ERR_INVALID;
    }
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    if (o.kmajor) {  // [rows][K]
        dims[0] = (cuuint64_t)o.K; dims[1] = (cuuint64_t)o.rows;
        box[0] = TC_BK; box[1] = (cuuint32_t)box_rows_kmajor;
    } else {         // [K][rows]
        dims[0] = (cuuint64_t)o.rows; dims[1] = (cuuint64_t)o.K;
        box[0] = plain ? TC_BM : 32; box[1] = TC_BK;
    }
    dims[2] = (cuuint64_t)(o.nbatch > 0 ? o.nbatch : 1);
    box[2] = 1;
    strides[0] = (cuuint64_t)o.ld * 4;
    strides[1] = (cuuint64_t)(o.bstride > 0 ? o.bstride : (long long)dims[1] * o.ld) * 4;
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(o.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     plain ? CU_TENSOR_MAP_SWIZZLE_NONE : (o.kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tc gemm: cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu ld=%lld", (int)r, (unsigned long long)dims[0],
                  (unsigned long long)dims[1], (unsigned long long)dims[2], o.ld);
        return EEGAN_ERR_CUDA;
    }
    MapSlot& slot = g_map_cache[g_map_next];
    g_map_next = (g_map_next + 1) % 32;
    slot.key = key;
    slot.map = *m;
    slot.used = true;
    return EEGAN_OK;
}

int tc_make_map_plain(CUtensorMap* m, const TcOperand& o) { return make_map(m, o, TC_BM, true); }

int make_tmap_2d(CUtensorMap* m, const float* ptr, unsigned long long rows, unsigned long long cols, unsigned long long pitch,
                 unsigned box_cols, unsigned box_rows, bool swizzle128) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("tmap2d: cuTensorMapEncodeTiled unavailable"); return EEGAN_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (pitch % 4)) {
        set_error("tmap2d: base/pitch must be 16-byte aligned");
        return EEGAN_ERR_INVALID;
    }
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {pitch * 4};
    cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tmap2d: cuTensorMapEncodeTiled failed (%d)", (int)r); return EEGAN_ERR_CUDA; }
    return EEGAN_OK;
}

int make_tmap_3d(CUtensorMap* m, const float* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                 unsigned long long pitch1, unsigned long long pitch2, unsigned box0, unsigned box1) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("tmap3d: cuTensorMapEncodeTiled unavailable"); return EEGAN_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (pitch1 % 4) || (pitch2 % 4)) {
        set_error("tmap3d: base/pitches must be 16-byte aligned");
        return EEGAN_ERR_INVALID;
    }
    cuuint64_t dims[3] = {d0, d1, d2}, strides[2] = {pitch1 * 4, pitch2 * 4};
    cuuint32_t box[3] = {box0, box1, 1}, estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tmap3d: cuTensorMapEncodeTiled failed (%d)", (int)r); return EEGAN_ERR_CUDA; }
    return EEGAN_OK;
}

static int num_sms() { return num_sms_current(); }

template <bool A_K, bool B_K, int EPI>
static int launch_t(const TcMaps& maps, const TcArgs& a, dim3 grid, cudaStream_t st) {
    static SmemGrant grant;
    if (int rc = grant_dyn_smem(tc_gemm_kernel<A_K, B_K, EPI>, (size_t)TC_SMEM_BYTES, grant, "tc gemm")) return rc;
    tc_gemm_kernel<A_K, B_K, EPI><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(maps, a);
    return check_launch("tc gemm");
}

int tc_gemm_launch(const TcGemm& g, cudaStream_t st) {
    EEGAN_REQUIRE(g.nseg == 1 || g.nseg == 2, "tc gemm: nseg=%d", g.nseg);
    EEGAN_REQUIRE(g.M > 0 && g.N > 0 && g.batch > 0 && g.C, "tc gemm: empty problem");
    for (int s = 1; s < g.nseg; ++s)
        EEGAN_REQUIRE(g.A[s].kmajor == g.A[0].kmajor && g.B[s].kmajor == g.B[0].kmajor, "tc gemm: segments must share majorness");
    if (g.ts) EEGAN_REQUIRE(!g.A[0].kmajor && g.B[0].kmajor, "tc gemm: the TMEM-staged kernel takes an MN-major A and a K-major B");

    TcMaps maps;
    TcArgs a{};
    for (int s = 0; s < 2; ++s) {
        const int src = s < g.nseg ? s : 0;
        const TcOperand* ops[2] = {&g.A[src], &g.B[src]};
        for (int o = 0; o < 2; ++o) {
            const bool plain = g.ts && o == 0;
            int rc = make_map(&maps.m[s][o][0], *ops[o], o == 0 ? TC_BM : TC_BN, plain);
            if (rc) return rc;
            if (ops[o]->lo) {
                TcOperand lo = *ops[o];
                lo.ptr = ops[o]->lo;
                rc = make_map(&maps.m[s][o][1], lo, o == 0 ? TC_BM : TC_BN, plain);
                if (rc) return rc;
            } else {
                maps.m[s][o][1] = maps.m[s][o][0];
            }
        }
        a.K[s] = s < g.nseg ? g.A[src].K : 0;
        a.a_batched[s] = g.A[src].bstride > 0;
        a.b_batched[s] = g.B[src].bstride > 0;
        a.a_pre[s] = g.A[src].lo != nullptr;
        a.b_pre[s] = g.B[src].lo != nullptr;
    }
    a.C = g.C; a.ldc = g.ldc; a.bC = g.bC; a.M = g.M; a.N = g.N; a.dynM = g.dynM; a.dynN = g.dynN; a.dynK = g.dynK;
    a.attn = g.attn;
    a.nseg = g.nseg; a.nred = g.nred > 0 ? g.nred : 1; a.red_total = g.red_total; a.batch = g.batch;
    a.mn_lbo = 4096 >> 4; a.mn_sbo = 512 >> 4;
    a.trunc_hi = 1;
#ifdef EEGAN_DEBUG_SWITCHES  // descriptor / work-skipping probes: never in the shipped library
    if (const char* e = getenv("EEGAN_TS_DBG")) a.dbg = atoi(e);
    if (const char* e = getenv("EEGAN_TC_TRUNC_HI")) a.trunc_hi = atoi(e);
    if (const char* e = getenv("EEGAN_TC_MN_LBO")) a.mn_lbo = (uint32_t)atoi(e);
    if (const char* e = getenv("EEGAN_TC_MN_SBO")) a.mn_sbo = (uint32_t)atoi(e);
#endif
    const long long tiles = (long long)((g.N + TC_BN - 1) / TC_BN) * ((g.M + TC_BM - 1) / TC_BM) * g.batch;
    dim3 grid((unsigned)(tiles < num_sms() ? tiles : num_sms()));
    const bool ak = g.A[0].kmajor, bk = g.B[0].kmajor;
    if (g.epi != TC_EPI_PLAIN) {
        EEGAN_REQUIRE(!ak && bk && g.nseg == 1, "tc gemm: the attention epilogues take an MN-major A (regions) and a K-major B (packed words)");
        EEGAN_REQUIRE(g.attn.nbins && g.attn.bin_cap && g.attn.bin_used && g.attn.col_start && g.attn.cap_len && g.attn.P,
                      "tc gemm: attention epilogue arguments missing");
        EEGAN_REQUIRE(g.ldc % 64 == 0, "tc gemm: attention epilogue needs a column pitch that is a multiple of 64");
        if (g.epi == TC_EPI_ATTN_FWD) {
            EEGAN_REQUIRE(g.attn.Zpart, "tc gemm: attention forward epilogue needs Zpart");
            if (g.ts) return ts_gemm_dispatch(maps, a, grid.x, g.epi, st);
            return launch_t<false, true, TC_EPI_ATTN_FWD>(maps, a, grid, st);
        }
        EEGAN_REQUIRE(g.epi == TC_EPI_ATTN_BWD && g.attn.csz, "tc gemm: attention backward epilogue needs csz");
        if (g.ts) return ts_gemm_dispatch(maps, a, grid.x, g.epi, st);
        return launch_t<false, true, TC_EPI_ATTN_BWD>(maps, a, grid, st);
    }
    if (g.ts) return ts_gemm_dispatch(maps, a, grid.x, g.epi, st);
    if (ak && bk) return launch_t<true, true, TC_EPI_PLAIN>(maps, a, grid, st);
    if (ak && !bk) return launch_t<true, false, TC_EPI_PLAIN>(maps, a, grid, st);
    if (!ak && bk) return launch_t<false, true, TC_EPI_PLAIN>(maps, a, grid, st);
    return launch_t<false, false, TC_EPI_PLAIN>(maps, a, grid, st);
}

}  // namespace eegan

using namespace eegan;

// Stand-alone entry point (tests / microbench): C[z] = A[z] * B[z]^T in 3xTF32.
//   a_kmajor: A is [M][K] (ld = lda) else [K][M];  b_kmajor: B is [N][K] else [K][N].
//   staging: 0 = both operands read from shared memory; 1 = A staged in tensor memory (needs a_kmajor = 0, b_kmajor = 1).
extern "C" int eegan_gemm_tf32x3(const float* A, const float* B, float* C, int M, int N, int K, int a_kmajor, int b_kmajor,
                                 long long lda, long long ldb, long long ldc, long long bsA, long long bsB, long long bsC,
                                 int batch, int staging, void* stream) {
    EEGAN_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "gemm_tf32x3: bad arguments");
    TcGemm g{};
    g.nseg = 1;
    g.ts = staging;
    g.A[0] = TcOperand{A, nullptr, a_kmajor, lda, bsA, batch, M, K};
    g.B[0] = TcOperand{B, nullptr, b_kmajor, ldb, bsB, batch, N, K};
    g.C = C; g.ldc = ldc; g.bC = bsC; g.M = M; g.N = N; g.batch = batch; g.nred = 1; g.red_total = 0;
    return tc_gemm_launch(g, (cudaStream_t)stream);
}
