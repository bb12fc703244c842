// tcgen05 "3xFP16" contraction engine for the pair-grid pipeline (contraction engine 3), sm_100a only.
//
//   C[z][m][n] = sum over segments s, reduction batches red, k of  A_s(m,k) * B_s(n,k)
//
// Every operand element x is stored AHEAD OF TIME, by the kernel that produces it, as a pair of
// halves of x' = x * 2^e (one power-of-two scale per tensor, kept on the device):
//     hi = fp16(x'),  lo = fp16(x' - hi)            (hi + lo carries 22 mantissa bits of x')
// i.e. the same 4 bytes per element as the fp32 value, and the same 2^-22 relative error as the
// tf32 hi/lo pair of the 3xTF32 engines (gemm_tc.cu / gemm_ts.cu) — but the three products per
// K-step (lo*hi, hi*lo, hi*hi) run as tcgen05.mma.kind::f16 at twice the tf32 rate, the operand
// tiles arrive from TMA already split and swizzled (no splitter warps, no conv barrier, no TMEM
// staging), and the shared-memory traffic per 128x128x32 block drops from ~128 KB to ~80 KB.
// The epilogue multiplies the accumulator by 2^-(eA+eB) (exact).
//
// Range: fp16 keeps 11 bits down to 2^-14 and 2^-24 absolute below that, so with the tensor's
// maximum scaled into [2^5, 2^15) every element within 2^-8 .. 2^-18 of the maximum still carries the
// full 22 bits and smaller ones an absolute error < 2^-30 of the maximum: fp32-class for a dot product.
// The producers (pair_grid_h.cu) choose the scale from an exact maximum (inputs) or a rigorous
// bound (gradient tensors) so that nothing overflows.
//
// Operands: A is always MN-major ([K][rows], rows contiguous: TMA boxes [32 k][64 rows], SWIZZLE_128B),
// B always K-major ([rows][K]: box [128 rows][32 k], SWIZZLE_64B) — which is what the region-major
// stash of the pair grid gives every contraction (pair_grid_v3.cu).
//
//   warp 0      TMA producer (6 boxes per 32 KB stage: A hi/lo x 2, B hi/lo)
//   warp 1      TMEM allocator + single-thread MMA issuer; tcgen05.commit frees the stage
//   warps 2-9   epilogue: two warps per TMEM lane quarter, one 64-column half of the tile each
// Two TMEM accumulator buffers overlap a tile's epilogue with the next main loop.  A two-segment
// GEMM (GEMM4: dC = DUz^T E + Wp^T dS) gives each segment its own accumulator (DUAL), because the
// two products carry different scales; the epilogue adds them after descaling.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace eegan {

constexpr int H_BM = 128, H_BN = 128, H_BK = 32;
constexpr int H_A_TILE = H_BM * H_BK * 2;                // 8 KB (one of hi / lo)
constexpr int H_B_TILE = H_BN * H_BK * 2;                // 8 KB
constexpr int H_STAGE_BYTES = 2 * H_A_TILE + 2 * H_B_TILE;  // A_hi A_lo B_hi B_lo = 32 KB
constexpr float H_CLAMP = 65504.0f;

struct HOperand {
    const __half* hi;
    const __half* lo;
    long long ld;            // pitch in elements of the non-contiguous index (multiple of 8)
    long long bstride;       // elements between batches (multiple of 8); 0 = not batched
    int nbatch;
    int rows, K;             // logical extents (TMA zero-fills beyond them)
    const float* inv_scale;  // device: 2^-e of the tensor
};

struct HAttnEpi {
    TcAttnEpi base;          // packing metadata, P, Zpart, csz, g1
    __half* out_hi;          // FWD: E' = E * 2^12;  BWD: dS' = dS * *out_scale     [batch][M][ldc]
    __half* out_lo;
    const float* out_scale;  // BWD: device scale of dS;  FWD: unused (constant H_E_SCALE)
    float* emax;             // FWD: device float, running max of E (atomicMax on the bit pattern; E > 0)
};
constexpr float H_E_SCALE = 4096.0f;

struct HGemm {
    HOperand A[2], B[2];
    int nseg;
    float* C;                // PLAIN epilogue output (fp32)
    long long ldc, bC;
    int M, N;
    const int* dynM;
    const int* dynN;
    const int* dynK;
    int epi;                 // TcEpilogue
    HAttnEpi attn;
    int batch, nred, red_total;
};

int h_gemm_launch(const HGemm& g, cudaStream_t st);

// Fused forward of the pair grid (gemm_h.cu, hf_fwd_kernel): S = C^T W -> attention -> U' = E^T C in one persistent launch.
struct HFusedFwd {
    const __half* Ch;        // [Bi][D][Rp] image features (hi, lo)
    const __half* Cl;
    const __half* Wh;        // [NtP][D] packed words
    const __half* Wl;
    int Bi, D, R, Rp, NtP;
    const int* nlive;        // device: live packed columns
    const float* inv_c;      // device: 2^-e of C, W, E
    const float* inv_w;
    const float* inv_e;
    float* U;                // [Bi][NtP][D] OUT: un-normalised word contexts U' = E C^T
    HAttnEpi attn;           // packing metadata + P / E' (hi, lo) / Zpart / emax outputs, g1
};
int h_fused_fwd_launch(const HFusedFwd& f, cudaStream_t st);

// x -> (hi, lo) halves of x * s, clamped to the fp16 range
__device__ __forceinline__ void h_split(float xs, __half& hi, __half& lo) {
    xs = xs > H_CLAMP ? H_CLAMP : (xs < -H_CLAMP ? -H_CLAMP : xs);  // comparisons, not fminf/fmaxf: a NaN stays a NaN, as in the reference
    hi = __float2half_rn(xs);
    lo = __float2half_rn(xs - __half2float(hi));
}

}  // namespace eegan
