// GlobalAttentionGeneral backward with all four contractions on the tensor cores (miscc/DAMSM_losses.py:96-132;
// math in SURVEY.md App. A), sm_100a.  One pass over HBM: d_out, x, attn, d_attn are read once, d_x is written once,
// ds never leaves the SM (the three-kernel form of gag_bwd2.cu moves 1.3-1.5x these bytes and needs 4 idf T FMA per
// pixel on the CUDA cores).
//
// Per 128-pixel tile of sample b (q = pixel, c = channel, t = word):
//   (1) dP[q][t]  = sum_c d_out[c][q] value[c][t]         M = 128 q, N = 32,  K = idf
//   (2) dV[c][t] += sum_q d_out[c][q] p[t][q]             M = 64 c,  N = 32,  K = 128 q     (one accumulator per 64 channels)
//       ds[q][t]  = p (dP + d_attn - sum_t p (dP + d_attn))     registers of the thread that owns pixel q (TMEM lane q)
//   (3) dX[q][c]  = sum_t ds[t][q] key[c][t]              M = 128 q, N = idf, K = 32
//   (4) dK[c][t] += sum_q x[c][q] ds[t][q]                M = 64 c,  N = 32,  K = 128 q
// d_out is contracted over c in (1) and over q in (2), i.e. it is an MN-major operand once and a K-major operand once.
// 32-bit (tf32) operands have no shared-memory layout that serves both (MN-major tf32 exists only in the 32-byte-atom
// swizzle, which K-major operands cannot use), 16-bit operands do: the plain 128-byte swizzle.  So every operand is a
// PAIR of bf16 arrays hi = bf16(v), lo = bf16(v - hi) (16 significant bits, the fp32 exponent range: no scaling), and a
// contraction is three tcgen05.mma.kind::f16 per K-step (lo*hi, hi*lo, hi*hi) into an fp32 accumulator in tensor memory:
// relative error ~2^-16 per product, against the 1e-4 relative-to-max the gradients are held to.
//
// Streams: d_out and x arrive as [UC channels][128 pixels] fp32 units (UC = min(idf, 64)) by TMA into a ring of slots;
// four converter warps turn a unit IN PLACE into its bf16 pair (the fp32 unit and the pair have the same size), laid
// out as two [UC][64 q] panels of 128-byte rows, 128B-swizzled — the layout that is K-major for (2)/(4) and MN-major
// for (1).  p and ds are written by the threads that own the pixels as [32 t][64 q] panels: K-major B operand of
// (2)/(4), MN-major A operand of (3).  value^T and key are converted once per CTA.
//
//   warp 0       TMA producer: unit order  d_out(0) | d_out(1) x(0) | d_out(2) x(1) | ... | x(last)   (per accumulation group)
//   warp 1       TMEM allocator + MMA issuer of the d_out side: (1) + (2) of tile i, issued while the pixel warps still work on
//                tile i-1 (dP and the p panels are double-buffered)
//   warp 2       MMA issuer of the ds / x side: (3) + (4) of tile i-1.  Both issuers walk the whole unit sequence and skip the
//                other's units; each is warp-uniform with one elected lane (uniform-register descriptors, no broadcast loops)
//   warp 3       spare (registers are allocated per four warps)
//   warps 4-7    converters (fp32 unit -> bf16 hi/lo panels, in place)
//   warps 8-11   pixel warps: d_attn from global one tile ahead, dP -> ds (p re-read from the panels), ds panels; at the end of a
//                group dV / dK -> partials
//   warps 12-15  p / output warps: attn from global -> p panels of tile i+1 (loaded one iteration earlier), then the dX
//                accumulator (double-buffered) -> d_x, coalesced 128-byte rows
// TMEM columns: dP 0-63 (2 x 32), dV 64-127 (2 x 32), dK 128-191, dX 192-447 (2 x 128).
// d_key / d_value: the tensor core adds into its fp32 accumulator with truncation, so a long accumulation chain drifts; every CTA
// writes its accumulators out as a partial [b][chunk][c][32] at most every GB_GROUP tiles, and gag_bwd_kv_reduce_kernel
// (gag_bwd2.cu) sums the partials in fp32 in a fixed order.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"

namespace eegan {

constexpr int GB_THREADS = 32 * 16;
constexpr int GB_TILE = 128;       // pixels per tile
constexpr int GB_NS_MAX = 10;      // most ring slots
constexpr uint32_t GB_PT = 8192;   // bytes of one [32 t][128 q] bf16 tile: 2 panels x 32 rows x 128 B

struct GagTcBwdArgs {
    const float* key;     // [B][idf][T]
    const float* value;   // [B][idf][T]
    const float* attn;    // [B][T][Q]
    const float* d_attn;  // [B][T][Q] or null
    float* d_x;           // [B][idf][Q]
    float* part_k;        // [B][S * ngr][idf][32]
    float* part_v;        // [B][S * ngr][idf][32]
    int B, idf, Q, T;
    int uc, nu;           // channels per unit (32 or 64), units per tile and stream
    int ns;               // ring slots
    int ngr;              // accumulation groups per CTA (dV / dK partials are written out once per group)
};

// Shared-memory matrix descriptors (version 1 = Blackwell) as two 32-bit words, so that the single issuing thread builds
// one with an integer add: low word = start address >> 4 (14 bits) | LBO >> 4 << 16, high word = SBO >> 4 | version | layout.
//   K-major, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart (SBO); a K-step of 16 halves advances the start by 32 B
//   MN-major, SWIZZLE_128B: rows = k, 64 MN elements per 128-byte row, 8-k groups 1024 B apart (SBO), the next 64 MN
//             elements LBO bytes on; a K-step of 16 advances the start by 2048 B
//   K-major, 64-byte rows, SWIZZLE_64B: 8-row groups 512 B apart; a K-step of 16 halves advances the start by 32 B
constexpr uint32_t GB_HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t GB_HI64 = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint32_t gb_lo_k(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t gb_lo_mn(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }

__device__ __forceinline__ bool gb_elect() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// issued by the elected lane only (`leader`); every lane of the warp executes the call
__device__ __forceinline__ void gb_mma(bool leader, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                       uint32_t idesc, uint32_t accumulate) {
    if (leader)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            "}" ::"r"(d_tmem),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void gb_commit(bool leader, uint32_t bar) {
    if (leader) tc_commit(bar);
}

// (a, b) -> packed bf16 pairs: hi = rn(a), rn(b) (a in the low half = lower address), lo = rn(a - hi_a), rn(b - hi_b)
__device__ __forceinline__ void gb_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void gb_split1(float a, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(a - __uint_as_float((uint32_t)hi << 16)));
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ uint16_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
// predicated read-only load, zero when off: nothing is executed on the loaded value afterwards, so the loads of a whole tile
// stay in flight until their first real use (a select after the load would make the warp wait for each of them at once)
__device__ __forceinline__ float ldg_pred(const float* ptr, bool pred) {
    float v;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "mov.f32 %0, 0f00000000;\n\t"
        "@p ld.global.nc.f32 %0, [%1];\n\t"
        "}"
        : "=f"(v)
        : "l"(ptr), "r"((int)pred));
    return v;
}
// the two halves of a packed pair to two addresses
__device__ __forceinline__ void sts_u16x2(uint32_t addr_lo, uint32_t addr_hi, uint32_t packed) {
    asm volatile(
        "{\n\t"
        ".reg .b16 l, h;\n\t"
        "mov.b32 {l, h}, %2;\n\t"
        "st.shared.u16 [%0], l;\n\t"
        "st.shared.u16 [%1], h;\n\t"
        "}" ::"r"(addr_lo),
        "r"(addr_hi), "r"(packed)
        : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int TP, int UC>
__global__ void __launch_bounds__(GB_THREADS, 1)
gag_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_x, const GagTcBwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    // the warp index through a shuffle: provably warp-uniform, so the role branches are uniform branches and the MMA issuer's
    // descriptors live in uniform registers (without it every MMA is wrapped in an elect / broadcast loop)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int tiles_b = (p.Q + GB_TILE - 1) / GB_TILE;
    const int my_tiles = (tiles_b - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // >= 1: the launch keeps gridDim.x <= tiles_b
    const int NU = p.nu, NS = p.ns, T = p.T, idf = p.idf, NGR = p.ngr;
    // the CTA's tiles in NGR groups (all non-empty: my_tiles >= NGR, see gag_tc_bwd_groups); group g = local tiles [gr0(g), gr0(g+1))
    auto gr0 = [&](int g) { return (int)(((long long)g * my_tiles) / NGR); };

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    // layout: ring | p hi, p lo (x 2 buffers) | ds hi | ds lo | value^T hi | value^T lo | key hi | key lo | barriers
    const uint32_t slot_bytes = (uint32_t)UC * 512u, half_bytes = (uint32_t)UC * 256u, panel_bytes = (uint32_t)UC * 128u;
    const uint32_t p_hi = base + (uint32_t)NS * slot_bytes, p_lo = p_hi + GB_PT;  // two buffers (tile i -> buffer i & 1), 2 GB_PT apart
    const uint32_t ds_hi = p_hi + 4 * GB_PT, ds_lo = ds_hi + GB_PT;
    const uint32_t vt_hi = ds_lo + GB_PT, vt_lo = vt_hi + (uint32_t)NU * 4096u;
    const uint32_t key_hi = vt_lo + (uint32_t)NU * 4096u, key_lo = key_hi + (uint32_t)idf * 64u;
    const uint32_t bars = key_lo + (uint32_t)idf * 64u;
    auto full = [&](int s) { return bars + 8u * s; };                       // TMA landed
    // "bf16 panels written": one barrier array PER ISSUER (d_out units / x units).  With one array an issuer would see only every
    // other completion of a slot's barrier — the other issuer's units pass it by — and a parity wait tells completion n from
    // completion n - 1 only (tests/test_gag_tc_bwd_protocol.py: wrong unit contracted on a two-slot ring).
    auto conv_do = [&](int s) { return bars + 8u * (GB_NS_MAX + s); };
    auto conv_x = [&](int s) { return bars + 8u * (2 * GB_NS_MAX + s); };
    auto sfree = [&](int s) { return bars + 8u * (3 * GB_NS_MAX + s); };    // MMAs done with the slot
    const uint32_t b0 = bars + 8u * (4 * GB_NS_MAX);
    const uint32_t ds_full = b0 + 16, ds_empty = b0 + 24, acc_full = b0 + 32, acc_empty = b0 + 40;
    // one barrier pair PER p buffer: with a single pair a waiter could be lapped by two completions (the phase parity aliases)
    auto p_full = [&](int a) { return b0 + 112u + 8u * a; };
    auto p_empty = [&](int a) { return b0 + 128u + 8u * a; };
    auto dp_full = [&](int a) { return b0 + 48u + 8u * a; };
    auto dp_empty = [&](int a) { return b0 + 64u + 8u * a; };
    auto dx_full = [&](int a) { return b0 + 80u + 8u * a; };
    auto dx_empty = [&](int a) { return b0 + 96u + 8u * a; };
    const uint32_t tmem_slot = b0 + 144u;

    // ---- one-time: zero the operand tiles (word padding rows / columns), value^T and key as bf16 pairs ----
    {
        float4* z4 = reinterpret_cast<float4*>(gbase + (p_hi - base));
        const int nz4 = (int)((bars - p_hi) >> 4);
        for (int idx = threadIdx.x; idx < nz4; idx += blockDim.x) z4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        const float* kb_ = p.key + (size_t)b * idf * T;
        const float* vb_ = p.value + (size_t)b * idf * T;
        // all loads of a thread first (idf T <= 128 x 32 elements over 448 threads), then the conversions: one load latency
        constexpr int PER = (128 * 32 + GB_THREADS - 1) / GB_THREADS;
        float kr[PER], vr[PER];
        const int n = idf * T;
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int idx = (int)threadIdx.x + r * GB_THREADS;
            kr[r] = idx < n ? __ldg(kb_ + idx) : 0.f;
            vr[r] = idx < n ? __ldg(vb_ + idx) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int idx = (int)threadIdx.x + r * GB_THREADS;
            if (idx < n) {
                const int c = idx / T, t = idx - c * T;
                uint16_t h, l;
                // value^T: per unit a K-major tile [32 t rows][64 c] of 128-byte rows, SWIZZLE_128B
                gb_split1(vr[r], h, l);
                const int u = c / UC, cc = c - u * UC;
                const uint32_t vo = (uint32_t)u * 4096u + (uint32_t)t * 128u + (uint32_t)(((cc >> 3) ^ (t & 7)) << 4) + (uint32_t)(cc & 7) * 2u;
                sts_u16(vt_hi + vo, h);
                sts_u16(vt_lo + vo, l);
                // key: K-major [idf rows][32 t] of 64-byte rows, SWIZZLE_64B (16-byte chunk ^= bits 7-8 of the offset)
                gb_split1(kr[r], h, l);
                const uint32_t ko = (uint32_t)c * 64u + (uint32_t)(((t >> 3) ^ ((c >> 1) & 3)) << 4) + (uint32_t)(t & 7) * 2u;
                sts_u16(key_hi + ko, h);
                sts_u16(key_lo + ko, l);
            }
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < GB_NS_MAX; ++s) {
            mbar_init(full(s), 1);
            mbar_init(conv_do(s), 4);
            mbar_init(conv_x(s), 4);
            mbar_init(sfree(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(p_full(a), 4);
            mbar_init(p_empty(a), 1);
        }
        mbar_init(ds_full, 4);
        mbar_init(ds_empty, 1);
        mbar_init(acc_full, 2);
        mbar_init(acc_empty, 4);
        for (int a = 0; a < 2; ++a) {
            mbar_init(dp_full(a), 1);
            mbar_init(dp_empty(a), 4);
            mbar_init(dx_full(a), 1);
            mbar_init(dx_empty(a), 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();  // operand tiles written by the generic proxy -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    auto tile_q0 = [&](int i) { return ((int)blockIdx.x + i * (int)gridDim.x) * GB_TILE; };
    constexpr uint32_t COL_DP = 0, COL_DV = 64, COL_DK = 128, COL_DX = 192;

    // position of a unit in the ring, carried along instead of a division per unit
    struct Ring {
        int s = 0, ph = 0;
    };
    auto ring_adv = [&](Ring& r, int n) {  // n <= NU <= 2 <= NS
        r.s += n;
        if (r.s >= NS) {
            r.s -= NS;
            r.ph ^= 1;
        }
    };

    if (warp == 0) {
        // ===== TMA producer: per group  d_out(t0) | d_out(t0+1) x(t0) | ... | x(t1-1) =====
        if (lane == 0) {
            Ring r;
            auto load = [&](const CUtensorMap* map, int q0, int u) {
                mbar_wait(sfree(r.s), r.ph ^ 1);
                mbar_arrive_expect_tx(full(r.s), slot_bytes);
                tma_load_3d(base + (uint32_t)r.s * slot_bytes, map, full(r.s), q0, u * UC, b);
                ring_adv(r, 1);
            };
            for (int g = 0; g < NGR; ++g) {
                const int t0 = gr0(g), t1 = gr0(g + 1);
                for (int i = t0; i <= t1; ++i) {
                    if (i < t1)
                        for (int u = 0; u < NU; ++u) load(&tm_do, tile_q0(i), u);
                    if (i > t0)
                        for (int u = 0; u < NU; ++u) load(&tm_x, tile_q0(i - 1), u);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===== MMA issuers: warp 1 takes the d_out units ((1) + (2)), warp 2 the ds / x side ((3) + (4)); each walks the whole
        // unit sequence (uniform control flow, one elected lane issues) and skips the other's units.  One issuer for all four
        // contractions was the bottleneck at idf = 32 (60 MMAs and their waits per 2 us tile in one instruction stream).
        const bool leader = gb_elect();
        const bool side_a = warp == 1;
        constexpr uint32_t ID_BF16 = (1u << 4) /*D = f32*/ | (1u << 7) /*A = bf16*/ | (1u << 10) /*B = bf16*/;
        const uint32_t idesc1 = ID_BF16 | (1u << 15) /*A MN-major*/ | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc24 = ID_BF16 | ((32u >> 3) << 17) | ((64u >> 4) << 24);
        const uint32_t idesc3 = ID_BF16 | (1u << 15) | (((uint32_t)idf >> 3) << 17) | ((128u >> 4) << 24);
        constexpr int ks1 = UC / 16;  // K-steps of (1) per unit
        const uint32_t pan16 = panel_bytes >> 4, half16 = half_bytes >> 4;
        const uint32_t ph_k = gb_lo_k(p_hi), pl_k = gb_lo_k(p_lo), dsh_k = gb_lo_k(ds_hi), dsl_k = gb_lo_k(ds_lo);
        const uint32_t dsh_mn = gb_lo_mn(ds_hi, 4096u), dsl_mn = gb_lo_mn(ds_lo, 4096u);
        const uint32_t kh_k = gb_lo_k(key_hi), kl_k = gb_lo_k(key_lo);
        // (2) / (4): acc[c][t] (+)= sum_q unit[c][q] w[t][q] : A = unit, K-major (M = 64 channels), B = p / ds panels
        auto rowsum = [&](uint32_t d_acc, uint32_t a_k, uint32_t wh, uint32_t wl, bool fresh) {
#pragma unroll
            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ao = (uint32_t)pp * pan16 + (uint32_t)ks * 2u, bo = (uint32_t)pp * 256u + (uint32_t)ks * 2u;
                    gb_mma(leader, d_acc, a_k + half16 + ao, GB_HI128, wh + bo, GB_HI128, idesc24, (!fresh || pp > 0 || ks > 0) ? 1u : 0u);
                    gb_mma(leader, d_acc, a_k + ao, GB_HI128, wl + bo, GB_HI128, idesc24, 1u);
                    gb_mma(leader, d_acc, a_k + ao, GB_HI128, wh + bo, GB_HI128, idesc24, 1u);
                }
        };
        Ring r;
        uint32_t used = 0;  // bit s: parity of THIS issuer's units that went through slot s = the phase of its barrier of the slot
        auto wait_converted = [&]() {
            mbar_wait(side_a ? conv_do(r.s) : conv_x(r.s), (used >> r.s) & 1u);
            used ^= 1u << r.s;
        };
        for (int g = 0; g < NGR; ++g) {
            const int t0 = gr0(g), t1 = gr0(g + 1);
            if (g > 0) {  // the pixel warps have read the dV / dK accumulators of the previous group
                mbar_wait(acc_empty, (g - 1) & 1);
                tc_fence_after();
            }
            for (int i = t0; i <= t1; ++i) {
                if (i < t1) {
                    if (side_a) {
                        const int a = i & 1, k = i >> 1;
                        mbar_wait(p_full(a), k & 1);
                        mbar_wait(dp_empty(a), (k & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d_dp = tmem_base + COL_DP + (uint32_t)a * 32u;
                        for (int u = 0; u < NU; ++u) {
                            wait_converted();
                            tc_fence_after();
                            const uint32_t sa = base + (uint32_t)r.s * slot_bytes;
                            const uint32_t a_k = gb_lo_k(sa), a_mn = gb_lo_mn(sa, panel_bytes);
                            const uint32_t vh = gb_lo_k(vt_hi + (uint32_t)u * 4096u), vl = gb_lo_k(vt_lo + (uint32_t)u * 4096u);
                            // (1) dP += d_out^T value : A = unit, MN-major (M = pixels), B = value^T chunk u
#pragma unroll
                            for (int ks = 0; ks < ks1; ++ks) {
                                const uint32_t ao = (uint32_t)ks * 128u, bo = (uint32_t)ks * 2u;
                                gb_mma(leader, d_dp, a_mn + half16 + ao, GB_HI128, vh + bo, GB_HI128, idesc1, (u > 0 || ks > 0) ? 1u : 0u);
                                gb_mma(leader, d_dp, a_mn + ao, GB_HI128, vl + bo, GB_HI128, idesc1, 1u);
                                gb_mma(leader, d_dp, a_mn + ao, GB_HI128, vh + bo, GB_HI128, idesc1, 1u);
                            }
                            // (2) dV_u += d_out p^T
                            rowsum(tmem_base + COL_DV + (uint32_t)u * 32u, a_k, ph_k + (uint32_t)a * (2 * GB_PT >> 4), pl_k + (uint32_t)a * (2 * GB_PT >> 4), i == t0);
                            gb_commit(leader, sfree(r.s));
                            ring_adv(r, 1);
                        }
                        gb_commit(leader, dp_full(a));
                        gb_commit(leader, p_empty(a));
                    } else {
                        ring_adv(r, NU);
                    }
                }
                if (i > t0) {
                    if (!side_a) {
                        const int j = i - 1, a = j & 1, k = j >> 1;
                        mbar_wait(ds_full, j & 1);
                        mbar_wait(dx_empty(a), (k & 1) ^ 1);
                        tc_fence_after();
                        // (3) dX = ds key^T : A = ds panels, MN-major (M = pixels), B = key (K-major, 64-byte rows)
                        const uint32_t d_dx = tmem_base + COL_DX + (uint32_t)a * 128u;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t ao = (uint32_t)ks * 128u, bo = (uint32_t)ks * 2u;
                            gb_mma(leader, d_dx, dsl_mn + ao, GB_HI128, kh_k + bo, GB_HI64, idesc3, ks > 0 ? 1u : 0u);
                            gb_mma(leader, d_dx, dsh_mn + ao, GB_HI128, kl_k + bo, GB_HI64, idesc3, 1u);
                            gb_mma(leader, d_dx, dsh_mn + ao, GB_HI128, kh_k + bo, GB_HI64, idesc3, 1u);
                        }
                        gb_commit(leader, dx_full(a));
                        // (4) dK_u += x ds^T
                        for (int u = 0; u < NU; ++u) {
                            wait_converted();
                            tc_fence_after();
                            rowsum(tmem_base + COL_DK + (uint32_t)u * 32u, gb_lo_k(base + (uint32_t)r.s * slot_bytes), dsh_k, dsl_k, j == t0);
                            gb_commit(leader, sfree(r.s));
                            ring_adv(r, 1);
                        }
                        gb_commit(leader, ds_empty);
                    } else {
                        ring_adv(r, NU);
                    }
                }
            }
            gb_commit(leader, acc_full);  // one arrival per issuer
        }
    } else if (warp == 3) {
        // spare warp (the register file is allocated in units of four warps: 16 warps cost what 14 did)
    } else if (warp < 8) {
        // ===== converters: fp32 unit [UC][128 q] -> bf16 hi / lo, two [UC][64 q] panels each, in place =====
        const int ctid = threadIdx.x - 128;
        constexpr int nchunk = UC / 4;  // 16-byte chunks per thread
        Ring r;
        // the same unit sequence as the producer's, to know whose unit a slot holds (which issuer's barrier to arrive on)
        int g = 0, i = gr0(0), t0 = gr0(0), t1 = gr0(1), sub = 0;  // sub: 0 .. NU-1 d_out units of tile i, NU .. 2NU-1 x units of tile i-1
        auto next_unit = [&](bool& is_x) -> bool {
            for (;;) {
                if (g >= NGR) return false;
                const bool has_do = i < t1, has_x = i > t0;
                if (sub < NU && has_do) { is_x = false; ++sub; return true; }
                if (sub < NU) sub = NU;
                if (sub < 2 * NU && has_x) { is_x = true; ++sub; return true; }
                sub = 0;
                if (++i > t1) {
                    ++g;
                    if (g < NGR) { t0 = gr0(g); t1 = gr0(g + 1); i = t0; }
                }
            }
        };
        bool is_x = false;
        while (next_unit(is_x)) {
            const int s = r.s, ph = r.ph;
            ring_adv(r, 1);
            mbar_wait(full(s), ph);
            const uint32_t sb = base + (uint32_t)s * slot_bytes;
            float4 v[nchunk];
#pragma unroll
            for (int k = 0; k < nchunk; ++k) v[k] = lds_v4(sb + (uint32_t)(ctid + 128 * k) * 16u);
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the whole unit is in registers: the slot may be overwritten
#pragma unroll
            for (int k = 0; k < nchunk; ++k) {
                {
                    const int idx = ctid + 128 * k, row = idx >> 5, ch = idx & 31;
                    uint32_t h0, l0, h1, l1;
                    gb_split2(v[k].x, v[k].y, h0, l0);
                    gb_split2(v[k].z, v[k].w, h1, l1);
                    const uint32_t off = (uint32_t)(ch >> 4) * panel_bytes + (uint32_t)row * 128u + (uint32_t)((((ch & 15) >> 1) ^ (row & 7)) << 4) +
                                         (uint32_t)(ch & 1) * 8u;
                    sts_v2(sb + off, h0, h1);
                    sts_v2(sb + half_bytes + off, l0, l1);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(is_x ? conv_x(s) : conv_do(s));
        }
    } else {
        // ===== pixel warps (8-11) and p / output warps (12-15): a thread owns pixel pl of the tile = TMEM lane pl =====
        const int quarter = warp & 3;
        const int pl = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // byte offset of (row t, pixel pl) in a [32 t][128 q] bf16 tile = offs[t & 7] + t * 128 (the 128-byte swizzle XORs the
        // 16-byte chunk index with t & 7): eight registers, every store / load below takes t * 128 as an immediate
        uint32_t offs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) offs[k] = (uint32_t)(pl >> 6) * 4096u + (uint32_t)(pl & 7) * 2u + ((((uint32_t)(pl & 63) >> 3) ^ (uint32_t)k) << 4);
        // val[0..TP) -> the tile's hi / lo panels (rows >= T carry zeros: val is zero there)
        auto write_tile = [&](uint32_t hi_base, uint32_t lo_base, const float (&val)[TP]) {
#pragma unroll
            for (int t = 0; t < TP; t += 2) {
                uint32_t h2, l2;
                gb_split2(val[t], val[t + 1], h2, l2);
                const uint32_t o0 = offs[t & 7] + (uint32_t)t * 128u, o1 = offs[(t + 1) & 7] + (uint32_t)(t + 1) * 128u;
                sts_u16x2(hi_base + o0, hi_base + o1, h2);
                sts_u16x2(lo_base + o0, lo_base + o1, l2);
            }
            fence_proxy_async();
            __syncwarp();
        };
        // rows of a [B][T][Q] array at this thread's pixel of tile i: predicated loads, zeros beyond T / Q
        auto load_rows = [&](const float* arr_b, int i, float (&dst)[TP]) {
            const int q = tile_q0(i) + pl;
            const bool ok = q < p.Q;
            const float* src = arr_b + q;
#pragma unroll
            for (int t = 0; t < TP; ++t) dst[t] = ldg_pred(src + (size_t)t * p.Q, ok && t < T);
        };
        if (warp < 12) {
            // ---- pixel warps: dP -> ds = p (dp - sum p dp) -> ds panels; at the end of a group dV / dK -> partials.
            // The fp32 p of the tile is re-read as hi + lo from the thread's own column of the p panels (2^-17 relative).
            const float* dattn_b = p.d_attn ? p.d_attn + (size_t)b * T * p.Q : nullptr;
            float da[TP], dn[TP];
#pragma unroll
            for (int t = 0; t < TP; ++t) dn[t] = 0.f;
            if (dattn_b) load_rows(dattn_b, 0, dn);
            int g = 0, g_end = gr0(1);
            for (int i = 0; i < my_tiles; ++i) {
                const int a = i & 1, k = i >> 1;
#pragma unroll
                for (int t = 0; t < TP; ++t) da[t] = dn[t];
                if (dattn_b && i + 1 < my_tiles) load_rows(dattn_b, i + 1, dn);  // one tile ahead: the latency hides behind this tile
                mbar_wait(dp_full(a), k & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(lane_base + COL_DP + (uint32_t)a * 32u, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dp_empty(a));
                float pc[TP];
                const uint32_t cb = (uint32_t)a * 2u * GB_PT;
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    const uint32_t off = cb + offs[t & 7] + (uint32_t)t * 128u;
                    pc[t] = __uint_as_float((uint32_t)lds_u16(p_hi + off) << 16) + __uint_as_float((uint32_t)lds_u16(p_lo + off) << 16);
                }
                float dot = 0.f;
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    da[t] += __uint_as_float(v[t]);  // dp = dP + d_attn
                    dot = fmaf(pc[t], da[t], dot);
                }
#pragma unroll
                for (int t = 0; t < TP; ++t) da[t] = pc[t] * (da[t] - dot);  // ds
                mbar_wait(ds_empty, (i & 1) ^ 1);
                write_tile(ds_hi, ds_lo, da);
                if (lane == 0) mbar_arrive(ds_full);
                if (i + 1 == g_end) {
                    // ---- end of a group: dV / dK accumulators (M = 64: row m sits in lane (m % 16) + 32 (m / 16)) -> the group's partial.
                    // The tensor core accumulates with truncation, so a long chain of accumulations drifts (measured: 6e-5 of the
                    // maximum after 170 tiles); a group is at most GB_GROUP tiles and the partials are summed in fp32 by the reduce kernel.
                    mbar_wait(acc_full, g & 1);
                    tc_fence_after();
                    const int cl = quarter * 16 + lane;  // channel of the unit held by this lane (lanes 0-15)
                    const size_t pbase = (((size_t)b * gridDim.x + blockIdx.x) * NGR + g) * idf;
                    for (int u = 0; u < NU; ++u) {
                        uint32_t vk[32], vv[32];
                        tmem_ld32(lane_base + COL_DK + (uint32_t)u * 32u, vk);
                        tmem_ld32(lane_base + COL_DV + (uint32_t)u * 32u, vv);
                        tmem_ld_wait();
                        if (lane < 16 && cl < UC) {
                            float4* ok = reinterpret_cast<float4*>(p.part_k + (pbase + (size_t)u * UC + cl) * 32);
                            float4* ov = reinterpret_cast<float4*>(p.part_v + (pbase + (size_t)u * UC + cl) * 32);
#pragma unroll
                            for (int t = 0; t < 32; t += 4) {
                                ok[t >> 2] = make_float4(__uint_as_float(vk[t]), __uint_as_float(vk[t + 1]), __uint_as_float(vk[t + 2]), __uint_as_float(vk[t + 3]));
                                ov[t >> 2] = make_float4(__uint_as_float(vv[t]), __uint_as_float(vv[t + 1]), __uint_as_float(vv[t + 2]), __uint_as_float(vv[t + 3]));
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty);
                    ++g;
                    g_end = gr0(g + 1);
                }
            }
        } else {
            // ---- p / output warps: the p panels of tile i+1 (two buffers) at the top of iteration i, from registers loaded one
            // iteration earlier, then dX[q][c] of tile i -> d_x[b][c][q] (coalesced 128-byte rows per channel)
            const float* attn_b = p.attn + (size_t)b * T * p.Q;
            float pn[TP];
            load_rows(attn_b, 0, pn);
            write_tile(p_hi, p_lo, pn);
            if (lane == 0) mbar_arrive(p_full(0));
            if (my_tiles > 1) load_rows(attn_b, 1, pn);
            for (int i = 0; i < my_tiles; ++i) {
                const int a = i & 1, k = i >> 1;
                if (i + 1 < my_tiles) {
                    // the contractions (2) of tile i-1 are done with buffer (i+1) & 1: its completion number (i-1) >> 1
                    mbar_wait(p_empty(a ^ 1), (((i + 1) >> 1) & 1) ^ 1);
                    const uint32_t nb = (uint32_t)(a ^ 1) * 2u * GB_PT;
                    write_tile(p_hi + nb, p_lo + nb, pn);
                    if (lane == 0) mbar_arrive(p_full(a ^ 1));
                    if (i + 2 < my_tiles) load_rows(attn_b, i + 2, pn);
                }
                const int q = tile_q0(i) + pl;
                mbar_wait(dx_full(a), k & 1);
                tc_fence_after();
                float* orow = p.d_x + (size_t)b * idf * p.Q + q;
                for (int c = 0; c < idf; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + COL_DX + (uint32_t)a * 128u + (uint32_t)c, v);
                    tmem_ld_wait();
                    if (c + 32 >= idf) {  // accumulator fully read
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(dx_empty(a));
                    }
                    if (q < p.Q) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) orow[(size_t)(c + jj) * p.Q] = __uint_as_float(v[jj]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// Pixel chunks (CTAs) per sample: about one CTA per SM over the grid, every CTA at least one tile.
int gag_tc_bwd_chunks(int B, int Q) {
    const int tiles_b = (Q + GB_TILE - 1) / GB_TILE;
    int S = num_sms_current() / (B > 0 ? B : 1);
    if (S < 1) S = 1;
    if (S > tiles_b) S = tiles_b;
    return S;
}
// Accumulation groups per CTA: at most GB_GROUP tiles are accumulated in tensor memory before the partial is written out.
// Every CTA has at least floor(tiles / S) >= groups tiles, so no group is empty.
constexpr int GB_GROUP = 32;
static int gag_tc_bwd_groups(int B, int Q) {
    const int tiles_b = (Q + GB_TILE - 1) / GB_TILE, S = gag_tc_bwd_chunks(B, Q);
    const int n = (tiles_b + S - 1) / S;
    return (n + GB_GROUP - 1) / GB_GROUP;
}

bool gag_tc_bwd_supported(const float* x, const float* attn, const float* d_out, const float* d_attn, const float* d_x, int B, int idf, int Q, int T) {
    return d_out && (idf == 32 || idf == 64 || idf == 128) && T >= 1 && T <= 32 && Q % 4 == 0 && B <= 65535 &&
           ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(attn) |
             reinterpret_cast<uintptr_t>(d_attn) | reinterpret_cast<uintptr_t>(d_x)) & 15) == 0;
}

// Floats of ONE partial array ([B][S * groups][idf][32]).
size_t gag_tc_bwd_part_floats(int B, int idf, int Q) { return (size_t)B * gag_tc_bwd_chunks(B, Q) * gag_tc_bwd_groups(B, Q) * idf * 32; }

// Enqueue the fused backward; part_k / part_v receive the partials of d_key / d_value ([B][*chunks][idf][32]), to be summed by
// gag_bwd_kv_reduce_kernel.
int gag_tc_bwd_launch(const float* x, const float* key, const float* value, const float* attn, const float* d_out, const float* d_attn,
                      int B, int idf, int Q, int T, float* d_x, float* part_k, float* part_v, int* chunks, cudaStream_t st) {
    CUtensorMap tm_do, tm_x;
    const int UC = idf < 64 ? idf : 64;
    int rc = make_tmap_3d(&tm_do, d_out, (unsigned long long)Q, (unsigned long long)idf, (unsigned long long)B, (unsigned long long)Q,
                          (unsigned long long)idf * Q, GB_TILE, (unsigned)UC);
    if (rc) return rc;
    rc = make_tmap_3d(&tm_x, x, (unsigned long long)Q, (unsigned long long)idf, (unsigned long long)B, (unsigned long long)Q,
                      (unsigned long long)idf * Q, GB_TILE, (unsigned)UC);
    if (rc) return rc;
    GagTcBwdArgs a{};
    a.key = key; a.value = value; a.attn = attn; a.d_attn = d_attn; a.d_x = d_x; a.part_k = part_k; a.part_v = part_v;
    a.B = B; a.idf = idf; a.Q = Q; a.T = T;
    a.uc = UC; a.nu = idf / UC;
    a.ngr = gag_tc_bwd_groups(B, Q);
    const size_t fixed = 6 * (size_t)GB_PT + 2 * (size_t)a.nu * 4096 + 2 * (size_t)idf * 64 + 640 /*barriers*/ + 1024 /*align*/;
    const size_t slot = (size_t)UC * 512;
    a.ns = (int)((232448 - fixed) / slot);
    if (a.ns > GB_NS_MAX) a.ns = GB_NS_MAX;
    EEGAN_REQUIRE(a.ns >= 2, "gag tc bwd: no room for the unit ring");
    const size_t smem = (size_t)a.ns * slot + fixed;
    const int S = gag_tc_bwd_chunks(B, Q);
    *chunks = S * a.ngr;
    auto run = [&](auto kern, SmemGrant& grant) -> int {
        if (int r = grant_dyn_smem(kern, smem, grant, "gag tc bwd")) return r;
        kern<<<dim3(S, B), GB_THREADS, smem, st>>>(tm_do, tm_x, a);
        return EEGAN_OK;
    };
    static SmemGrant g0, g1, g2, g3;
    if (T <= 20) rc = UC == 64 ? run(gag_tc_bwd_kernel<20, 64>, g0) : run(gag_tc_bwd_kernel<20, 32>, g1);
    else rc = UC == 64 ? run(gag_tc_bwd_kernel<32, 64>, g2) : run(gag_tc_bwd_kernel<32, 32>, g3);
    if (rc) return rc;
    return check_launch("gag tc bwd");
}

}  // namespace eegan
