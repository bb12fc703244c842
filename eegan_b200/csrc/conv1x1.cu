// CNN_ENCODER.emb_features (DAMSM.py:162, 229): the 1x1 convolution 768 -> nef (256), no bias, on the 17 x 17
// Inception feature map — the producer of the `img_features` that words_loss consumes.
//
//   fwd   y[b][co][r]  = sum_ci w[co][ci] x[b][ci][r]                      M = co, N = r,  K = ci
//   bwd   dx[b][ci][r] = sum_co w[co][ci] dy[b][co][r]                     M = ci, N = r,  K = co
//         dw[co][ci]   = sum_b sum_r dy[b][co][r] x[b][ci][r]              M = co, N = ci, K = r, reduced over b
//
// All three are batched GEMMs on the fp32-accurate tcgen05 3xTF32 engine (gemm_tc.cu), reading the operands in
// place through K-major / MN-major tensor maps.  TMA needs 16-byte row pitches and R = 289 is odd, so x and dy are
// first re-pitched to Rp = 292 (one coalesced pass; the forward's copy of x is what autograd stashes, so the
// backward never re-reads x).  dw is reduced over the batch in `nsplit` groups (about one wave of tiles) and the
// partials are added in a fixed order: deterministic, no atomics.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace eegan {

static __global__ void __launch_bounds__(256) c1_repitch_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows,
                                                                int R, int Rp) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        const float* sp = src + row * R;
        float* dp = dst + row * Rp;
        for (int r0 = 0; r0 < Rp; r0 += 32 * 10) {
            float v[10];
#pragma unroll
            for (int q = 0; q < 10; ++q) {
                const int r = r0 + 32 * q + lane;
                v[q] = r < R ? __ldg(sp + r) : 0.f;
            }
#pragma unroll
            for (int q = 0; q < 10; ++q) {
                const int r = r0 + 32 * q + lane;
                if (r < Rp) dp[r] = v[q];
            }
        }
    }
}

static __global__ void __launch_bounds__(256) c1_sum_partials_kernel(const float* __restrict__ part, int nsplit, long long n,
                                                                     float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(long long)k * n + i];
    out[i] = s;
}

static int c1_nsplit(int B, int Cin, int Cout) {
    const int tiles = ((Cout + 127) / 128) * ((Cin + 127) / 128);
    int ns = 148 / (tiles > 0 ? tiles : 1);
    if (ns < 1) ns = 1;
    if (ns > B) ns = B;
    const int nred = (B + ns - 1) / ns;
    return (B + nred - 1) / nred;
}

}  // namespace eegan

using namespace eegan;

static int c1_rp(int R) { return (R + 3) / 4 * 4; }

// fwd: B*Cin*Rp floats (the re-pitched x: keep it for the backward).  bwd: B*Cout*Rp + nsplit*Cout*Cin floats.
extern "C" size_t eegan_conv1x1_workspace_bytes(int B, int Cin, int Cout, int R, int backward) {
    if (B <= 0 || Cin <= 0 || Cout <= 0 || R <= 0) return 0;
    const size_t Rp = (size_t)c1_rp(R);
    if (!backward) return align_up((size_t)B * Cin * Rp * sizeof(float), 256);
    return align_up((size_t)B * Cout * Rp * sizeof(float), 256) + align_up((size_t)c1_nsplit(B, Cin, Cout) * Cout * Cin * sizeof(float), 256);
}

extern "C" int eegan_conv1x1_fwd(const float* x, const float* w, int B, int Cin, int Cout, int R, float* y, float* xp,
                                 size_t xp_bytes, void* stream) {
    EEGAN_REQUIRE(x && w && y && xp && B > 0 && Cin > 0 && Cout > 0 && R > 0, "conv1x1 fwd: bad arguments");
    EEGAN_REQUIRE(Cin % 4 == 0, "conv1x1: Cin=%d must be a multiple of 4 (16-byte weight rows for TMA)", Cin);
    const int Rp = c1_rp(R);
    EEGAN_REQUIRE(xp_bytes >= (size_t)B * Cin * Rp * sizeof(float), "conv1x1 fwd: re-pitch buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    c1_repitch_kernel<<<148 * 8, 256, 0, st>>>(x, xp, (long long)B * Cin, R, Rp);
    EEGAN_LAUNCH_CHECK("conv1x1 re-pitch");
    TcGemm g{};
    g.nseg = 1;
    g.A[0] = TcOperand{w, nullptr, 1, Cin, 0, 1, Cout, Cin};                          // [co][ci], K = ci contiguous
    g.B[0] = TcOperand{xp, nullptr, 0, Rp, (long long)Cin * Rp, B, R, Cin};           // [ci][r], N = r contiguous
    g.C = y; g.ldc = R; g.bC = (long long)Cout * R; g.M = Cout; g.N = R; g.batch = B; g.nred = 1;
    return tc_gemm_launch(g, st);
}

extern "C" int eegan_conv1x1_bwd(const float* xp, const float* w, const float* dy, int B, int Cin, int Cout, int R, float* dx,
                                 float* dw, void* workspace, size_t workspace_bytes, void* stream) {
    EEGAN_REQUIRE(xp && w && dy && workspace && B > 0 && Cin > 0 && Cout > 0 && R > 0, "conv1x1 bwd: bad arguments");
    EEGAN_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "conv1x1: channel counts must be multiples of 4");
    EEGAN_REQUIRE(workspace_bytes >= eegan_conv1x1_workspace_bytes(B, Cin, Cout, R, 1), "conv1x1 bwd: workspace too small");
    const int Rp = c1_rp(R);
    cudaStream_t st = (cudaStream_t)stream;
    float* dyp = (float*)workspace;
    float* part = (float*)((char*)workspace + align_up((size_t)B * Cout * Rp * sizeof(float), 256));
    c1_repitch_kernel<<<148 * 8, 256, 0, st>>>(dy, dyp, (long long)B * Cout, R, Rp);
    EEGAN_LAUNCH_CHECK("conv1x1 re-pitch dy");
    if (dx) {
        TcGemm g{};
        g.nseg = 1;
        g.A[0] = TcOperand{w, nullptr, 0, Cin, 0, 1, Cin, Cout};                       // [co][ci] read as [K = co][M = ci]
        g.B[0] = TcOperand{dyp, nullptr, 0, Rp, (long long)Cout * Rp, B, R, Cout};     // [co][r]: [K][N]
        g.C = dx; g.ldc = R; g.bC = (long long)Cin * R; g.M = Cin; g.N = R; g.batch = B; g.nred = 1;
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
    }
    if (dw) {
        const int ns = c1_nsplit(B, Cin, Cout), nred = (B + ns - 1) / ns;
        TcGemm g{};
        g.nseg = 1;
        g.A[0] = TcOperand{dyp, nullptr, 1, Rp, (long long)Cout * Rp, B, Cout, R};    // [co][r], K = r contiguous
        g.B[0] = TcOperand{xp, nullptr, 1, Rp, (long long)Cin * Rp, B, Cin, R};       // [ci][r], K = r contiguous
        g.C = part; g.ldc = Cin; g.bC = (long long)Cout * Cin; g.M = Cout; g.N = Cin; g.batch = ns; g.nred = nred; g.red_total = B;
        int rc = tc_gemm_launch(g, st);
        if (rc) return rc;
        const long long n = (long long)Cout * Cin;
        c1_sum_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, ns, n, dw);
        EEGAN_LAUNCH_CHECK("conv1x1 dw");
    }
    return EEGAN_OK;
}
