// GlobalAttentionGeneral forward on the tensor cores (miscc/DAMSM_losses.py:96-132), sm_100a.
//
// Per 128-pixel tile of sample b, fused in one persistent CTA:
//   S[q][t]   = sum_d x[d][q] key[d][t]          tcgen05 3xTF32, A = x tile staged in TMEM (thread = pixel),
//                                                B = key^T (K-major, prepared once per CTA in shared memory)
//   p         = softmax_t(mask(S))               in the registers of the thread that owns pixel q (TMEM lane q)
//   attn[t][q] = p[t]                            coalesced 128-byte rows straight from registers
//   out[d][q] = sum_t value[d][t] p[t]           tcgen05 3xTF32, A = p written back to TMEM with tcgen05.st,
//                                                B = value (K-major as stored), accumulator -> registers -> global
// HBM traffic is the compulsory x read and out / attn writes; the CUDA cores only do the softmax.
// The CUDA-core kernels of gag.cu need 4 idf T FLOP of FMA per pixel for this (more than the chip's fp32
// peak at the roofline rate for idf = 128) and stay as the fallback for shapes this kernel does not take.
//
//   warp 0      TMA producer: x boxes [32 d][128 q], un-swizzled, NS-deep ring
//   warp 1      TMEM allocator + MMA issuer (S of tile i+1 is issued before O of tile i: S is double-buffered)
//   warps 2-5   x splitters: shared memory -> hi / lo in TMEM (lane = pixel)
//   warps 6-9   softmax warps: S -> p -> attn, p hi / lo -> TMEM
//   warps 10-13 output warps: O accumulator -> out
// TMEM columns: S0 0-31, S1 32-63, P hi 64-95, P lo 96-127, O 128.., [cross-term accumulators of S, 2 x 32, idf <= 128], x stages (64 columns each).
#include "gemm_tc.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"

namespace eegan {

constexpr int GT_THREADS = 32 * 14;
constexpr int GT_TILE = 128;  // pixels per tile
constexpr int GT_TP = 32;     // words padded to one K-block / 32 accumulator columns
constexpr int GT_XS_MAX = 10; // most shared-memory x stages

struct GagTcArgs {
    const float* key;    // [B][idf][T]
    const float* value;  // [B][idf][T]
    const uint8_t* mask; // [B][T] or null
    float* out;          // [B][idf][Q]
    float* attn;         // [B][T][Q]
    int B, idf, Q, T, mask_mode;
    int nkb;             // ceil(idf / 32)
    int idf_pad;         // idf rounded up to 16 (MMA N of the second contraction)
    int ns;              // TMEM x stages (4, or 2 when idf_pad > 128: the O accumulator takes 256 columns)
    int xs;              // shared-memory x stages (TMA ring; deeper than the TMEM ring to keep enough HBM reads in flight)
    int a_col0;          // first TMEM column of the x stages
    int sx_col0;         // >= 0: the cross terms (lo hi, hi lo) of S accumulate in their own two 32-column buffers from this column
                         // (the tensor core adds with truncation: 48 small additions into one large sum cost ~1e-6 in a
                         // probability; apart, the large sum sees a third of the additions); -1: one accumulator (idf > 128)
};

__device__ __forceinline__ bool gt_elect() {
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

// byte offset of element (row, k) in a K-major SWIZZLE_128B tile of 32-float rows
__device__ __forceinline__ uint32_t sw128_off(int row, int k) { return (uint32_t)(row * 128 + ((((k >> 2) ^ (row & 7))) << 4) + ((k & 3) << 2)); }

template <int TP>
__global__ void __launch_bounds__(GT_THREADS, 1)
gag_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const GagTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_maskbits[1024];  // per mask row: bit t set = word t is padding
    // the warp index through a shuffle is provably warp-uniform: the role branches become uniform branches and the issuer's
    // operands live in uniform registers (otherwise every MMA is wrapped in an elect / broadcast loop)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int tiles_b = (p.Q + GT_TILE - 1) / GT_TILE;
    const int my_tiles = ((int)blockIdx.x < tiles_b) ? (tiles_b - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (my_tiles == 0) return;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    // layout: x stages | key^T hi | key^T lo | value hi | value lo | barriers
    const uint32_t x_bytes = (uint32_t)p.xs * TC_TILE_BYTES;
    const uint32_t kt_bytes = (uint32_t)p.nkb * 4096u;            // nkb tiles of [32 t][32 d]
    const uint32_t vt_bytes = (uint32_t)((p.idf_pad + 7) / 8 * 8) * 128u;  // [idf rows][32 t]
    const uint32_t kt_hi = base + x_bytes, kt_lo = kt_hi + kt_bytes;
    const uint32_t vt_hi = (kt_lo + kt_bytes + 1023u) & ~1023u, vt_lo = (vt_hi + vt_bytes + 1023u) & ~1023u;
    const uint32_t bars = (vt_lo + vt_bytes + 15u) & ~15u;
    auto full = [&](int s) { return bars + 8u * s; };            // [GT_XS_MAX] TMA landed
    auto xfree = [&](int s) { return bars + 8u * (GT_XS_MAX + s); };  // [GT_XS_MAX] splitters copied the stage to registers
    auto conv = [&](int s) { return bars + 8u * (2 * GT_XS_MAX + s); };       // [4] hi/lo in TMEM
    auto empty = [&](int s) { return bars + 8u * (2 * GT_XS_MAX + 4 + s); };  // [4] MMAs done with the TMEM stage
    auto s_full = [&](int a) { return bars + 8u * (2 * GT_XS_MAX + 8 + a); };
    auto s_empty = [&](int a) { return bars + 8u * (2 * GT_XS_MAX + 10 + a); };
    const uint32_t p_full = bars + 8u * (2 * GT_XS_MAX + 12), p_empty = p_full + 8u, o_full = p_full + 16u, o_empty = p_full + 24u;
    const uint32_t tmem_slot = p_full + 32u;

    // ---- one-time: mask bits, key^T / value operand tiles (hi = raw, lo = tf32 residual) ----
    for (int r = threadIdx.x; r < p.B && r < 1024; r += blockDim.x) {
        uint32_t bits = 0;
        if (p.mask)
            for (int t = 0; t < p.T; ++t) bits |= (p.mask[(size_t)r * p.T + t] ? 1u : 0u) << t;
        s_maskbits[r] = bits;
    }
    {
        // zero the operand tiles (word / channel padding), then scatter key / value with coalesced global reads
        float4* z4 = reinterpret_cast<float4*>(gbase + (kt_hi - base));
        const int nz4 = (int)((vt_lo + vt_bytes - kt_hi) >> 4);
        for (int idx = threadIdx.x; idx < nz4; idx += blockDim.x) z4[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        uint8_t* g_kt_hi = gbase + (kt_hi - base);
        uint8_t* g_kt_lo = gbase + (kt_lo - base);
        uint8_t* g_vt_hi = gbase + (vt_hi - base);
        uint8_t* g_vt_lo = gbase + (vt_lo - base);
        const float* kb_ = p.key + (size_t)b * p.idf * p.T;
        const float* vb_ = p.value + (size_t)b * p.idf * p.T;
        // all loads of a thread first (idf T <= 256 x 32 elements over 448 threads), then the scatter: one load latency
        constexpr int PER = (256 * 32 + GT_THREADS - 1) / GT_THREADS;
        float kr[PER], vr[PER];
        const int n = p.idf * p.T;
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int idx = (int)threadIdx.x + r * GT_THREADS;
            kr[r] = idx < n ? __ldg(kb_ + idx) : 0.f;
            vr[r] = idx < n ? __ldg(vb_ + idx) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            const int idx = (int)threadIdx.x + r * GT_THREADS;
            if (idx < n) {
                const int d = idx / p.T, t = idx - d * p.T;
                const float kv = kr[r], vv = vr[r];
                const uint32_t ko = (uint32_t)(d >> 5) * 4096u + sw128_off(t, d & 31);  // key^T: row = word, k = channel
                *reinterpret_cast<float*>(g_kt_hi + ko) = kv;
                *reinterpret_cast<float*>(g_kt_lo + ko) = to_tf32(kv - trunc_tf32(kv));
                const uint32_t vo = sw128_off(d, t);                                    // value: row = channel, k = word
                *reinterpret_cast<float*>(g_vt_hi + vo) = vv;
                *reinterpret_cast<float*>(g_vt_lo + vo) = to_tf32(vv - trunc_tf32(vv));
            }
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_XS_MAX; ++s) {
            mbar_init(full(s), 1);
            mbar_init(xfree(s), 4);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(conv(s), 4);
            mbar_init(empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(s_full(a), 1);
            mbar_init(s_empty(a), 4);
        }
        mbar_init(p_full, 4);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        mbar_init(o_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // operand tiles written by the generic proxy -> visible to UMMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    const int NS = p.ns, XS = p.xs;
    auto tile_q0 = [&](int i) { return ((int)blockIdx.x + i * (int)gridDim.x) * GT_TILE; };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int q0 = tile_q0(i);
                for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                    const int s = it % XS, ph = (it / XS) & 1;
                    mbar_wait(xfree(s), ph ^ 1);
                    mbar_arrive_expect_tx(full(s), TC_TILE_BYTES);
                    tma_load_3d(base + s * TC_TILE_BYTES, &tmx, full(s), q0, kb * 32, b);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop, one elected lane issues =====
        const bool leader = gt_elect();
        {
            const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GT_TP >> 3) << 17) | ((uint32_t)(GT_TILE >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.idf_pad >> 3) << 17) | ((uint32_t)(GT_TILE >> 4) << 24);
            int it = 0;
            auto issue_s = [&](int i) {
                const int sb = i & 1;
                mbar_wait(s_empty(sb), ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_s = tmem_base + (uint32_t)(sb * GT_TP);
                const uint32_t d_x = p.sx_col0 >= 0 ? tmem_base + (uint32_t)(p.sx_col0 + sb * GT_TP) : d_s;
                const bool split = p.sx_col0 >= 0;
                for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                    const int s = it % NS, ph = (it / NS) & 1;
                    mbar_wait(conv(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + (uint32_t)(p.a_col0 + s * 64), a_lo = a_hi + 32;
                    const uint32_t b_hi = kt_hi + (uint32_t)kb * 4096u, b_lo = kt_lo + (uint32_t)kb * 4096u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dbh = umma_desc(b_hi, true, ks, 0, 0), dbl = umma_desc(b_lo, true, ks, 0, 0);
                        if (leader) tc_mma_tf32_ts(d_x, a_lo + ks * 8, dbh, idesc_s, (kb > 0 || ks > 0) ? 1u : 0u);
                        if (leader) tc_mma_tf32_ts(d_x, a_hi + ks * 8, dbl, idesc_s, 1u);
                        if (leader) tc_mma_tf32_ts(d_s, a_hi + ks * 8, dbh, idesc_s, (!split || kb > 0 || ks > 0) ? 1u : 0u);
                    }
                    if (leader) tc_commit(empty(s));
                }
                if (leader) tc_commit(s_full(sb));
            };
            issue_s(0);
            for (int i = 0; i < my_tiles; ++i) {
                if (i + 1 < my_tiles) issue_s(i + 1);
                mbar_wait(p_full, i & 1);
                mbar_wait(o_empty, (i & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_o = tmem_base + 128u, p_hi = tmem_base + 64u, p_lo = tmem_base + 96u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dbh = umma_desc(vt_hi, true, ks, 0, 0), dbl = umma_desc(vt_lo, true, ks, 0, 0);
                    if (leader) tc_mma_tf32_ts(d_o, p_lo + ks * 8, dbh, idesc_o, ks > 0 ? 1u : 0u);
                    if (leader) tc_mma_tf32_ts(d_o, p_hi + ks * 8, dbl, idesc_o, 1u);
                    if (leader) tc_mma_tf32_ts(d_o, p_hi + ks * 8, dbh, idesc_o, 1u);
                }
                if (leader) tc_commit(o_full);
                if (leader) tc_commit(p_empty);
            }
        }
    } else if (warp < 6) {
        // ===== x splitters: shared memory [32 d][128 q] -> TMEM [lane q][32 hi | 32 lo] =====
        const int quarter = warp & 3;
        const uint32_t my_q = (uint32_t)(quarter * 32 + lane) * 4u;
        int it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                const int sx = it % XS, phx = (it / XS) & 1;
                const int s = it % NS, ph = (it / NS) & 1;
                mbar_wait(full(sx), phx);
                const uint32_t sA = base + sx * TC_TILE_BYTES + my_q;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) hi[q] = __float_as_uint(lds_f32(sA + (uint32_t)q * (GT_TILE * 4u)));
                __syncwarp();
                if (lane == 0) mbar_arrive(xfree(sx));  // the stage is in registers: hand the buffer back to the TMA ring
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float x = __uint_as_float(hi[q]);
                    const float h = trunc_tf32(x);
                    hi[q] = __float_as_uint(h);
                    lo[q] = __float_as_uint(to_tf32(x - h));
                }
                mbar_wait(empty(s), ph ^ 1);  // the MMAs that read this TMEM stage are done
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p.a_col0 + s * 64);
                tmem_st32(ta, hi);
                tmem_st32(ta + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(conv(s));
            }
        }
    } else if (warp < 10) {
        // ===== softmax warps =====  (lean on purpose: at idf = 32 a tile is ~1000 cycles of HBM time, and this loop is what
        // every pixel pays — one mask word, exp2 on the shifted score, pointer-increment stores, no per-element range checks)
        const int quarter = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t padbits = p.T >= 32 ? 0u : (0xffffffffu << p.T);  // words >= T count as masked: exp(-inf) = 0
        const int qstep = (int)gridDim.x * GT_TILE;
        int q = tile_q0(0) + quarter * 32 + lane;
        // mask quirk (:114-118): row (b, q) takes mask[(b Q + q) mod B]; mode 1 = mask[b].  The residue is carried
        // from tile to tile instead of a 64-bit modulo per tile.
        int rq = p.mask_mode ? b : (int)(((long long)b * p.Q + q) % p.B);
        const int rstep = p.mask_mode ? 0 : qstep % p.B;
        for (int i = 0; i < my_tiles; ++i) {
            const int sb = i & 1;
            mbar_wait(s_full(sb), (i >> 1) & 1);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(lane_base + (uint32_t)(sb * GT_TP), v);
            if (p.sx_col0 >= 0) {
                uint32_t vx[32];
                tmem_ld32(lane_base + (uint32_t)(p.sx_col0 + sb * GT_TP), vx);
                tmem_ld_wait();
#pragma unroll
                for (int t = 0; t < TP; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) + __uint_as_float(vx[t]));
            } else {
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(sb));
            const uint32_t mbits = padbits | (p.mask ? s_maskbits[rq] : 0u);
            float mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < TP; ++t) {
                const float sv = ((mbits >> t) & 1u) ? -INFINITY : __uint_as_float(v[t]);
                v[t] = __float_as_uint(sv);
                mx = fmaxf(mx, sv);
            }
            float sum = 0.f;
#pragma unroll
            for (int t = 0; t < TP; ++t) {
                float e;  // exp(s - max): all words masked gives -inf - -inf = NaN, as in the reference
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((__uint_as_float(v[t]) - mx) * 1.4426950408889634f));
                v[t] = __float_as_uint(e);
                sum += e;
            }
            const float inv = 1.0f / sum;
            uint32_t lo[32];
            float* ap = p.attn + (size_t)b * p.T * p.Q + q;
            const int nst = q < p.Q ? p.T : 0;  // rows this thread stores
#pragma unroll
            for (int t = 0; t < TP; ++t) {
                const float pv = __uint_as_float(v[t]) * inv;  // masked / padded words: exactly 0 (NaN rows stay NaN, as in the reference)
                if (t < nst) *ap = pv;
                ap += p.Q;
                const float h = trunc_tf32(pv);
                v[t] = __float_as_uint(h);
                lo[t] = __float_as_uint(pv - h);  // exact; the tensor core truncates it to tf32 itself
            }
#pragma unroll
            for (int t = TP; t < 32; ++t) v[t] = lo[t] = 0u;
            mbar_wait(p_empty, (i & 1) ^ 1);  // the previous tile's P has been consumed by its MMAs
            tc_fence_after();
            tmem_st32(lane_base + 64u, v);
            tmem_st32(lane_base + 96u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            q += qstep;
            rq += rstep;
            if (rq >= p.B) rq -= p.B;
        }
    } else {
        // ===== output warps: O[q][d] -> out[b][d][q] =====
        const int quarter = warp & 3;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + 128u;
        for (int i = 0; i < my_tiles; ++i) {
            const int q = tile_q0(i) + quarter * 32 + lane;
            mbar_wait(o_full, i & 1);
            tc_fence_after();
            float* op = p.out + (size_t)b * p.idf * p.Q + q;
            const bool okq = q < p.Q;
            for (int c = 0; c < p.idf_pad; c += 32) {
                uint32_t v[32];
                const bool whole = p.idf - c >= 32;
                if (p.idf_pad - c >= 32) tmem_ld32(lane_base + (uint32_t)c, v);
                else tmem_ld16(lane_base + (uint32_t)c, v);
                tmem_ld_wait();
                if (c + 32 >= p.idf_pad) {  // accumulator fully read
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(o_empty);
                }
                if (whole) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (okq) *op = __uint_as_float(v[j]);
                        op += p.Q;
                    }
                } else if (okq) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c + j < p.idf) op[(size_t)j * p.Q] = __uint_as_float(v[j]);
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

// Does this shape go to the tensor-core kernel?
bool gag_tc_fwd_supported(const float* x, int B, int idf, int Q, int T) {
    return Q % 4 == 0 && idf % 16 == 0 && idf >= 16 && idf <= 256 && T <= 32 && B <= 1024 && B <= 65535 &&
           (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int gag_tc_fwd_launch(const float* x, const float* key, const float* value, const uint8_t* mask, int mask_mode, int B, int idf,
                      int Q, int T, float* out, float* attn, cudaStream_t st) {
    CUtensorMap tmx;
    TcOperand ox{x, nullptr, 0, (long long)Q, (long long)idf * Q, B, Q, idf};  // MN-major: [K = d][rows = q]
    int rc = tc_make_map_plain(&tmx, ox);
    if (rc) return rc;
    GagTcArgs a{};
    a.key = key; a.value = value; a.mask = mask; a.out = out; a.attn = attn;
    a.B = B; a.idf = idf; a.Q = Q; a.T = T; a.mask_mode = mask_mode;
    a.nkb = (idf + 31) / 32;
    a.idf_pad = (idf + 15) / 16 * 16;
    const int o_cols = a.idf_pad <= 128 ? 128 : 256;
    a.sx_col0 = o_cols == 128 ? 128 + o_cols : -1;
    a.a_col0 = 128 + o_cols + (a.sx_col0 >= 0 ? 64 : 0);
    a.ns = (TC_TMEM_COLS - a.a_col0) / 64;  // 4 or 2
    const size_t vt = (size_t)((a.idf_pad + 7) / 8 * 8) * 128;
    const size_t fixed = 2 * (size_t)a.nkb * 4096 + 2 * (vt + 1024) + 1024 + 1024 + 512;
    a.xs = (int)((220 * 1024 - fixed) / TC_TILE_BYTES);
    if (a.xs > GT_XS_MAX) a.xs = GT_XS_MAX;
    const size_t smem = (size_t)a.xs * TC_TILE_BYTES + fixed;
    const int tiles_b = (Q + GT_TILE - 1) / GT_TILE;
    int per_sample = B <= 148 ? 148 / B : 1;
    if (per_sample > tiles_b) per_sample = tiles_b;
    if (per_sample < 1) per_sample = 1;
    auto run = [&](auto kern, SmemGrant& grant) -> int {
        if (int r = grant_dyn_smem(kern, smem, grant, "gag tc fwd")) return r;
        kern<<<dim3(per_sample, B), GT_THREADS, smem, st>>>(tmx, a);
        return check_launch("gag tc fwd");
    };
    static SmemGrant g20, g32;  // words padded to 20 or 32 in the softmax loops
    return T <= 20 ? run(gag_tc_fwd_kernel<20>, g20) : run(gag_tc_fwd_kernel<32>, g32);
}

}  // namespace eegan
