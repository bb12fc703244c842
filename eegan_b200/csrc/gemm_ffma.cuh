// Generic strided / batched fp32 FFMA GEMM used by the pair-grid pipeline (v1 contraction
// engine: exact fp32 FMA on the CUDA cores).
//
//   C[z][m][n] (+)= sum_{red} sum_k A[z, red](m,k) * B[z, red](k,n)
//
// Operands are addressed by element strides so the same kernel serves the NN, NT and TN
// forms the pipeline needs.  Template flags say which index is contiguous so the
// global->shared copy is coalesced.  The M or K extent can be overridden from device
// memory (the packed column count sum(cap_lens) is only known on the device).
#pragma once
#include "common.cuh"

namespace eegan {

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    int M, N, K;      // static upper bounds
    const int* dynM;  // optional device int: actual M
    const int* dynK;  // optional device int: actual K
    long long sAm, sAk, sBk, sBn, sCm, sCn;  // element strides
    long long bA, bB, bC;                    // per-blockIdx.z strides
    int nred;                                // inner reduction batches per z
    int red_total;                           // if > 0: batch z reduces [z*nred, min((z+1)*nred, red_total))
    long long rA, rB;                        // strides of the reduction batch
    int accumulate;                          // C += instead of C =
};

template <int BM, int BN, int TM, int TN, bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256) gemm_ffma_kernel(GemmArgs g) {
    constexpr int BK = 16, NT = 256;
    static_assert((BM / TM) * (BN / TN) == NT, "thread tiling must give 256 threads");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile must be float4-able");
    constexpr int LA = BM * BK / NT, LB = BN * BK / NT;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int M = g.dynM ? min(*g.dynM, g.M) : g.M;
    const int K = g.dynK ? min(*g.dynK, g.K) : g.K;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= M) return;
    const int z = blockIdx.z;
    const float* A = g.A + (long long)z * g.bA;
    const float* B = g.B + (long long)z * g.bB;
    float* C = g.C + (long long)z * g.bC;

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // micro-tile rows/cols are interleaved in groups of 4 (one float4 per group) so that a
    // quarter-warp's LDS.128 touches 128 contiguous bytes (no bank conflicts).
    auto row_of = [&](int i) { return (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4); };
    auto col_of = [&](int j) { return (j / 4) * (BN / (TN / 4)) + tx * 4 + (j % 4); };

    float ra[LA], rb[LB];
    const int nk = (K + BK - 1) / BK;
    int nred = g.nred;
    if (g.red_total > 0) nred = max(0, min(g.nred, g.red_total - z * g.nred));
    const int total = nred * nk;

    auto gload = [&](int it) {
        const int red = it / nk, k0 = (it - red * nk) * BK;
        const float* Ar = A + (long long)red * g.rA;
        const float* Br = B + (long long)red * g.rB;
#pragma unroll
        for (int p = 0; p < LA; ++p) {
            const int idx = tid + p * NT;
            const int mm = A_KCONTIG ? idx / BK : idx % BM;
            const int kk = A_KCONTIG ? idx % BK : idx / BM;
            const int gm = m0 + mm, gk = k0 + kk;
            ra[p] = (gm < M && gk < K) ? __ldg(Ar + gm * g.sAm + gk * g.sAk) : 0.f;
        }
#pragma unroll
        for (int p = 0; p < LB; ++p) {
            const int idx = tid + p * NT;
            const int nn = B_NCONTIG ? idx % BN : idx / BK;
            const int kk = B_NCONTIG ? idx / BN : idx % BK;
            const int gn = n0 + nn, gk = k0 + kk;
            rb[p] = (gn < g.N && gk < K) ? __ldg(Br + gk * g.sBk + gn * g.sBn) : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int p = 0; p < LA; ++p) {
            const int idx = tid + p * NT;
            const int mm = A_KCONTIG ? idx / BK : idx % BM;
            const int kk = A_KCONTIG ? idx % BK : idx / BM;
            As[buf][kk][mm] = ra[p];
        }
#pragma unroll
        for (int p = 0; p < LB; ++p) {
            const int idx = tid + p * NT;
            const int nn = B_NCONTIG ? idx % BN : idx / BK;
            const int kk = B_NCONTIG ? idx / BN : idx % BK;
            Bs[buf][kk][nn] = rb[p];
        }
    };

    if (total > 0) {
        gload(0);
        sstore(0);
    }
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) gload(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4)
                *reinterpret_cast<float4*>(&a[i]) = *reinterpret_cast<const float4*>(&As[buf][k][row_of(i)]);
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4*>(&b[j]) = *reinterpret_cast<const float4*>(&Bs[buf][k][col_of(j)]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (it + 1 < total) sstore(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + row_of(i);
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + col_of(j);
            if (gn >= g.N) continue;
            float* p = C + gm * g.sCm + gn * g.sCn;
            float v = acc[i][j];
            if (g.accumulate) v += *p;
            *p = v;
        }
    }
}

template <int BM, int BN, int TM, int TN, bool A_KCONTIG, bool B_NCONTIG>
inline void launch_gemm_ffma(const GemmArgs& g, int batch, cudaStream_t st) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batch);
    gemm_ffma_kernel<BM, BN, TM, TN, A_KCONTIG, B_NCONTIG><<<grid, 256, 0, st>>>(g);
}

}  // namespace eegan
