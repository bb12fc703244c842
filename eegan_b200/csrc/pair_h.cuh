// DAMSM pair grid on the half-pair contraction engine (pair_grid_h.cu): host entry points used by the C ABI in pair_grid.cu.
#pragma once
#include "common.cuh"

namespace eegan {

size_t pair_h_workspace_bytes(int Bi, int Bc, int D, int R, int Tm);
int pair_h_fwd(const float* img, const float* words, const int32_t* cap_lens, int Bi, int Bc, int D, int R, int Tm, float g1,
               float g2, float* m, float* att, int diag_offset, void* workspace, size_t workspace_bytes, cudaStream_t st);
// phases: bit 0 = per-column scalars, dU, GEMM3 (+ attention backward) and GEMM4 -> d_img;  bit 1 = GEMM5 + unpack -> d_words
// (needs the dS stash of bit 0 from an earlier call on the same workspace)
int pair_h_bwd(const float* img, int Bi, int Bc, int D, int R, int Tm, float g1, float g2, const float* dm, float* d_img,
               float* d_words, void* workspace, size_t workspace_bytes, cudaStream_t st, int phases = 3);

}  // namespace eegan
