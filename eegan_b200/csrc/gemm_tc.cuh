// tcgen05 (5th-gen tensor core) contraction engine for the pair-grid pipeline, sm_100a only.
//
//   C[z][m][n] = sum over segments s, reduction batches red, k of  A_s(m,k) * B_s(n,k)
//
// fp32-accurate "3xTF32": every fp32 operand element x is used as hi = trunc_tf32(x) (the
// tensor core reads only the top 19 bits of an fp32 operand, so the raw tile IS hi) plus
// lo = tf32(x - hi), produced either by the splitter warps in shared memory or ahead of time by
// the kernel that wrote the operand; each K-step issues three tcgen05.mma.kind::tf32
// (lo*hi, hi*lo, hi*hi) into one fp32 accumulator tile in TMEM.  Single-pass TF32/BF16 flips
// argmax word indices; the split keeps ~2^-21 relative error.
//
// Persistent CTAs (one per SM, 448 threads) walk 128 x 128 output tiles, BK = 32 fp32 = one
// 128-byte swizzle row:
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B) of the raw fp32 A/B boxes
//   warp 1      TMEM allocator + single-thread MMA issuer; tcgen05.commit frees the stage
//   warps 2-9   hi/lo splitters (in place, layout-agnostic)
//   warps 10-13 epilogue: tcgen05.ld -> registers -> per-warp smem transpose -> coalesced stores
// mbarrier pipeline per stage: full (TMA landed) -> conv (split done) -> empty (MMAs done);
// two TMEM accumulators (tmem_full / tmem_empty) overlap a tile's epilogue with the next main loop.
// Operands may be K-major ([rows][K], K contiguous) or MN-major ([K][rows], rows contiguous);
// both use the canonical 128B-swizzle UMMA layouts, so no transposed copies are needed.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace eegan {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;            // 16 KB per operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;           // A_hi B_hi A_lo B_lo
constexpr int TC_EPI_PITCH = 65;                            // staging row pitch (floats): odd => conflict-free both ways
constexpr int TC_EPI_STAGE_BYTES = 4 * 32 * TC_EPI_PITCH * 4;  // 4 epilogue warps x [32 rows][64 (+1) columns]
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_EPI_STAGE_BYTES + 1024 /*epilogue scratch*/ + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_SPLIT_WARPS = 8;
constexpr int TC_TMEM_COLS = 512;                           // 2 x 128 accumulator columns + read-window slack
constexpr int TC_THREADS = 32 * (2 + TC_SPLIT_WARPS + 4);

struct TcOperand {
    const float* ptr;    // base (raw fp32; the tensor core uses its top 19 bits = hi)
    const float* lo;     // optional pre-split lo = tf32(x - trunc_tf32(x)), same layout; NULL = split in the kernel
    int kmajor;          // 1: [rows][K] (K contiguous); 0: [K][rows] (rows contiguous)
    long long ld;        // pitch in elements of the non-contiguous index (multiple of 4)
    long long bstride;   // elements between batches (multiple of 4); 0 = not batched
    int nbatch;          // number of batches addressable (>= 1)
    int rows, K;         // logical extents (TMA zero-fills beyond them)
};

// Epilogues.  PLAIN stores the accumulator tile.  The two ATTN epilogues run the word-region
// attention of the DAMSM pair grid (miscc/DAMSM_losses.py:44-54 and its backward) on the
// accumulator while it is still in TMEM: the tile is [128 regions r (TMEM lanes)] x [128 packed word
// columns n' (TMEM columns)], so one thread owns one region and a caption's words sit in its
// registers.  Packed columns are grouped in 64-column bins that hold whole captions.
enum TcEpilogue { TC_EPI_PLAIN = 0, TC_EPI_ATTN_FWD = 1, TC_EPI_ATTN_BWD = 2 };

struct TcAttnEpi {
    const int* nbins;      // device: live 64-column bins
    const int* bin_cap;    // [nbins + 1] first caption of each bin
    const int* bin_used;   // [nbins] columns of the bin that carry words
    const int* col_start;  // [captions] first packed column of caption i
    const int* cap_len;    // [captions] words of caption i (clamped)
    float* P;              // [batch][M][ldc]  FWD: out, P = softmax_words(S);  BWD: in
    float* Zpart;          // FWD: [batch][ceil(M/32)][ldc] per-32-region partial sums of E over regions
    const float* csz;      // BWD: [batch][ldc]  (sum_r A dA) / Z per column
    float g1;
};

struct TcGemm {
    TcOperand A[2], B[2];   // up to two K-concatenated segments (segment 1 unused if nseg == 1)
    int nseg;
    float* C;
    long long ldc, bC;
    int M, N;               // output extents (static upper bounds)
    const int* dynM;        // optional device int: live rows of C / A
    const int* dynN;        // optional device int: live columns of C / rows of B
    const int* dynK;        // optional device int: live K extent (all segments)
    int ts;                 // 1: A operand staged in tensor memory (gemm_ts.cu; MN-major A, K-major B only)
    int epi;                // TcEpilogue; ATTN_*: C = E (FWD) / dS (BWD), both [batch][M][ldc]
    TcAttnEpi attn;
    int batch;              // grid.z
    int nred, red_total;    // reduction batches per z: operand batch index = z*nred + red
};

// Plain (un-swizzled) 2-D fp32 tensor map over a row-major [rows][cols] array with row pitch
// `pitch` elements (multiple of 4) and a [box_rows][box_cols] box; TMA zero-fills out of bounds.
int make_tmap_2d(CUtensorMap* m, const float* ptr, unsigned long long rows, unsigned long long cols, unsigned long long pitch,
                 unsigned box_cols, unsigned box_rows, bool swizzle128 = false);

// Plain (un-swizzled) 3-D fp32 tensor map over [d2][d1][d0] (d0 contiguous; pitches in elements, multiples of 4) with a
// [1][box1][box0] box (gag_tc_bwd.cu: [channels][pixels] units of x / d_out); TMA zero-fills out of bounds.
int make_tmap_3d(CUtensorMap* m, const float* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                 unsigned long long pitch1, unsigned long long pitch2, unsigned box0, unsigned box1);

// Un-swizzled 3-D map of an MN-major operand ([K][rows], rows contiguous): box [32 k][128 rows] (gemm_ts.cu, gag_tc.cu)
int tc_make_map_plain(CUtensorMap* m, const TcOperand& o);

struct TcMaps;
struct TcArgs;
// gemm_ts.cu: launch of the TMEM-staged kernel on prepared maps / arguments
int ts_gemm_dispatch(const TcMaps& maps, const TcArgs& args, unsigned grid, int epi, cudaStream_t st);

// Enqueue the GEMM.  Returns EEGAN_OK or an error code (message via set_error).
int tc_gemm_launch(const TcGemm& g, cudaStream_t st);

}  // namespace eegan
