// SyncBatchNorm device side (sync_batchnorm/batchnorm.py:48-78, 113-125).  The reference
// makes two extra full passes over x for sum and sum(x**2) (:60-62, the latter materialising
// x**2) and a third to normalise; here the statistics are one fused pass, the cross-replica
// reduction is an NCCL all-reduce of the 2*C floats issued by the host between these calls,
// and normalise / backward are single streaming passes.  x is [N, C, HW] contiguous.
#include <atomic>

#include "common.cuh"

namespace eegan {

constexpr int BN_THREADS = 256;

// grid (C, S): block (c, s) reduces a strided share of the N*HW elements of channel c.
// MODE 0: {sum x, sum x^2};  MODE 1: {sum dy, sum dy*xhat}
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ inv_std, int N, int C, int HW,
                                                               float* __restrict__ out) {
    __shared__ float red[32];
    const int c = blockIdx.x;
    const float mu = MODE ? mean[c] : 0.f, is = MODE ? inv_std[c] : 0.f;
    float a = 0.f, b = 0.f;
    const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     (MODE == 0 || (reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    if (vec) {
        const int hw4 = HW / 4;
        const long long total = (long long)N * hw4;
        for (long long e = (long long)blockIdx.y * BN_THREADS + threadIdx.x; e < total;
             e += (long long)gridDim.y * BN_THREADS) {
            const int n = (int)(e / hw4), k = (int)(e - (long long)n * hw4);
            const size_t off = ((size_t)n * C + c) * HW + (size_t)k * 4;
            const float4 xv = *reinterpret_cast<const float4*>(x + off);
            if (MODE == 0) {
                a += (xv.x + xv.y) + (xv.z + xv.w);
                b = fmaf(xv.x, xv.x, b); b = fmaf(xv.y, xv.y, b); b = fmaf(xv.z, xv.z, b); b = fmaf(xv.w, xv.w, b);
            } else {
                const float4 g = *reinterpret_cast<const float4*>(dy + off);
                a += (g.x + g.y) + (g.z + g.w);
                b = fmaf(g.x, (xv.x - mu) * is, b); b = fmaf(g.y, (xv.y - mu) * is, b);
                b = fmaf(g.z, (xv.z - mu) * is, b); b = fmaf(g.w, (xv.w - mu) * is, b);
            }
        }
    } else {
        const long long total = (long long)N * HW;
        for (long long e = (long long)blockIdx.y * BN_THREADS + threadIdx.x; e < total;
             e += (long long)gridDim.y * BN_THREADS) {
            const int n = (int)(e / HW), k = (int)(e - (long long)n * HW);
            const size_t off = ((size_t)n * C + c) * HW + k;
            const float xv = x[off];
            if (MODE == 0) {
                a += xv;
                b = fmaf(xv, xv, b);
            } else {
                const float g = dy[off];
                a += g;
                b = fmaf(g, (xv - mu) * is, b);
            }
        }
    }
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) {
        atomicAdd(out + c, a);
        atomicAdd(out + C + c, b);
    }
}

__device__ __forceinline__ float dev_count(const float* count_dev, float host_count) {
    // element counts cross the all-reduce as {count / 4096, count % 4096} so the fp32 sum stays exact
    return count_dev ? count_dev[0] * 4096.f + count_dev[1] : host_count;
}

__global__ void bn_finalize_kernel(const float* __restrict__ stats, int C, float count, const float* __restrict__ count_dev,
                                   float eps, float momentum,
                                   int clamp_mode, float* __restrict__ mean, float* __restrict__ inv_std,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float size = dev_count(count_dev, count);
    const float s = stats[c], ss = stats[C + c];
    const float mu = s / size;                 // batchnorm.py:116
    const float sumvar = ss - s * mu;          // :117
    const float bias_var = sumvar / size;      // :119
    mean[c] = mu;
    inv_std[c] = clamp_mode ? 1.0f / sqrtf(fmaxf(bias_var, eps))   // clamp(eps) ** -0.5, :125
                            : 1.0f / sqrtf(bias_var + eps);        // F.batch_norm, :50-53
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;  // :122
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (sumvar / (size - 1.f));  // :118,123
}

// y = (x - mean) * (inv_std * w) + b ;  MODE 1: dx = w*inv_std*(dy - r0/count - xhat*r1/count)
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ inv_std,
                                                              const float* __restrict__ weight,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ red, float host_count,
                                                              const float* __restrict__ count_dev,
                                                              int kill_var_term, float clamp_inv_std, int C, int HW,
                                                              float* __restrict__ out) {
    const int nc = blockIdx.y;  // n*C + c
    const int c = nc % C;
    const float mu = mean[c], is = inv_std[c];
    const float w = weight ? weight[c] : 1.f;
    const float scale = is * w;
    const float shift = (MODE == 0 && bias) ? bias[c] : 0.f;
    float r0 = 0.f, r1 = 0.f;
    if (MODE == 1) {
        const float inv_count = 1.0f / dev_count(count_dev, host_count);
        r0 = red[c] * inv_count;
        // when the variance was clamped (N-replica formula) inv_std is a constant: no var term
        r1 = (kill_var_term && is >= clamp_inv_std) ? 0.f : red[C + c] * inv_count;
    }
    const size_t base = (size_t)nc * HW;
    const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                         (MODE == 1 ? reinterpret_cast<uintptr_t>(dy) : 0)) & 15) == 0);
    if (vec) {
        const int hw4 = HW / 4;
        for (int k = blockIdx.x * BN_THREADS + threadIdx.x; k < hw4; k += gridDim.x * BN_THREADS) {
            const float4 xv = *reinterpret_cast<const float4*>(x + base + (size_t)k * 4);
            float4 o;
            if (MODE == 0) {
                o.x = (xv.x - mu) * scale + shift; o.y = (xv.y - mu) * scale + shift;
                o.z = (xv.z - mu) * scale + shift; o.w = (xv.w - mu) * scale + shift;
            } else {
                const float4 g = *reinterpret_cast<const float4*>(dy + base + (size_t)k * 4);
                o.x = scale * (g.x - r0 - (xv.x - mu) * is * r1); o.y = scale * (g.y - r0 - (xv.y - mu) * is * r1);
                o.z = scale * (g.z - r0 - (xv.z - mu) * is * r1); o.w = scale * (g.w - r0 - (xv.w - mu) * is * r1);
            }
            *reinterpret_cast<float4*>(out + base + (size_t)k * 4) = o;
        }
    } else {
        for (int k = blockIdx.x * BN_THREADS + threadIdx.x; k < HW; k += gridDim.x * BN_THREADS) {
            const float xv = x[base + k];
            out[base + k] = (MODE == 0) ? (xv - mu) * scale + shift
                                        : scale * (dy[base + k] - r0 - (xv - mu) * is * r1);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Single-replica, small-map forms: ONE launch per direction, one CTA per channel.
// 14 of the 24 SyncBN layers of Gen work on maps of 32 x 32 or less: a few MB that a multi-launch
// sequence (memset + reduce + finalize + apply) spends in launch latency.  Here a CTA makes the channel's statistics pass and
// the normalising pass itself; the second pass re-reads the channel from L2 (<= 64K elements = 256 KB per channel).
//   KIND 0: y = (x - mean) * inv_std * w + b              (SynchronizedBatchNorm2d, batchnorm.py:50-53 single replica)
//   KIND 1: y = (g m + 1) xhat + b m                      (affine_ssa, models.py:69-86)
// ---------------------------------------------------------------------------------------
constexpr int BN_SMALL_THREADS = 512;
constexpr long long BN_SMALL_MAX = 65536;  // elements per channel

template <typename F>
__device__ __forceinline__ void bn_small_foreach(int N, int C, int HW, int c, bool vec, F f) {
    if (vec) {
        const int hw4 = HW >> 2;
        const int total = N * hw4;
        for (int e = threadIdx.x; e < total; e += BN_SMALL_THREADS) {
            const int n = e / hw4, k = e - n * hw4;
            f(n, ((size_t)n * C + c) * HW + (size_t)k * 4, k * 4, 4);
        }
    } else {
        const int total = N * HW;
        for (int e = threadIdx.x; e < total; e += BN_SMALL_THREADS) {
            const int n = e / HW, k = e - n * HW;
            f(n, ((size_t)n * C + c) * HW + k, k, 1);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(BN_SMALL_THREADS) bn_fwd_small_kernel(const float* __restrict__ x, const float* __restrict__ weight,
                                                                        const float* __restrict__ bias, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, const float* __restrict__ mask,
                                                                        int N, int C, int HW, float eps, float momentum,
                                                                        float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                        float* __restrict__ y, float* __restrict__ mean,
                                                                        float* __restrict__ inv_std) {
    __shared__ float red[32];
    const int c = blockIdx.x;
    const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                         (KIND == 1 ? reinterpret_cast<uintptr_t>(mask) : 0)) & 15) == 0);
    float a = 0.f, b = 0.f;
    bn_small_foreach(N, C, HW, c, vec, [&](int, size_t off, int, int w) {
        if (w == 4) {
            const float4 v = *reinterpret_cast<const float4*>(x + off);
            a += (v.x + v.y) + (v.z + v.w);
            b = fmaf(v.x, v.x, b); b = fmaf(v.y, v.y, b); b = fmaf(v.z, v.z, b); b = fmaf(v.w, v.w, b);
        } else {
            const float v = x[off];
            a += v;
            b = fmaf(v, v, b);
        }
    });
    a = block_sum(a, red);
    b = block_sum(b, red);
    const float size = (float)N * (float)HW;
    const float mu = a / size, sumvar = b - a * mu;  // batchnorm.py:116-117
    const float is = 1.0f / sqrtf(sumvar / size + eps);
    if (threadIdx.x == 0) {
        mean[c] = mu;
        inv_std[c] = is;
        if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
        if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (sumvar / (size - 1.f));
    }
    const float wv = (KIND == 0 && weight) ? weight[c] : 1.f, bv = (KIND == 0 && bias) ? bias[c] : 0.f;
    const float scale = is * wv;
    bn_small_foreach(N, C, HW, c, vec, [&](int n, size_t off, int k, int w) {
        float g = 0.f, bt = 0.f;
        if (KIND == 1) {
            g = gamma[(size_t)n * C + c];
            bt = beta[(size_t)n * C + c];
        }
        if (w == 4) {
            const float4 v = *reinterpret_cast<const float4*>(x + off);
            float4 o;
            if (KIND == 0) {
                o.x = (v.x - mu) * scale + bv; o.y = (v.y - mu) * scale + bv; o.z = (v.z - mu) * scale + bv; o.w = (v.w - mu) * scale + bv;
            } else {
                const float4 m = *reinterpret_cast<const float4*>(mask + (size_t)n * HW + k);
                o.x = fmaf(fmaf(g, m.x, 1.f), (v.x - mu) * is, bt * m.x); o.y = fmaf(fmaf(g, m.y, 1.f), (v.y - mu) * is, bt * m.y);
                o.z = fmaf(fmaf(g, m.z, 1.f), (v.z - mu) * is, bt * m.z); o.w = fmaf(fmaf(g, m.w, 1.f), (v.w - mu) * is, bt * m.w);
            }
            *reinterpret_cast<float4*>(y + off) = o;
        } else {
            const float v = x[off];
            if (KIND == 0) {
                y[off] = (v - mu) * scale + bv;
            } else {
                const float m = mask[(size_t)n * HW + k];
                y[off] = fmaf(fmaf(g, m, 1.f), (v - mu) * is, bt * m);
            }
        }
    });
}

// dx = w inv_std (dy - S0 / n - xhat S1 / n),  S0 = sum dy, S1 = sum dy xhat;  red[c] = S0, red[C + c] = S1 (d_bias, d_weight)
__global__ void __launch_bounds__(BN_SMALL_THREADS) bn_bwd_small_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                        const float* __restrict__ mean, const float* __restrict__ inv_std,
                                                                        const float* __restrict__ weight, int N, int C, int HW,
                                                                        float* __restrict__ dx, float* __restrict__ red_out) {
    __shared__ float red[32];
    const int c = blockIdx.x;
    const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0);
    const float mu = mean[c], is = inv_std[c];
    float a = 0.f, b = 0.f;
    bn_small_foreach(N, C, HW, c, vec, [&](int, size_t off, int, int w) {
        if (w == 4) {
            const float4 v = *reinterpret_cast<const float4*>(x + off);
            const float4 g = *reinterpret_cast<const float4*>(dy + off);
            a += (g.x + g.y) + (g.z + g.w);
            b = fmaf(g.x, (v.x - mu) * is, b); b = fmaf(g.y, (v.y - mu) * is, b);
            b = fmaf(g.z, (v.z - mu) * is, b); b = fmaf(g.w, (v.w - mu) * is, b);
        } else {
            a += dy[off];
            b = fmaf(dy[off], (x[off] - mu) * is, b);
        }
    });
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) {
        red_out[c] = a;
        red_out[C + c] = b;
    }
    const float inv_count = 1.0f / ((float)N * (float)HW);
    const float r0 = a * inv_count, r1 = b * inv_count;
    const float scale = is * (weight ? weight[c] : 1.f);
    bn_small_foreach(N, C, HW, c, vec, [&](int, size_t off, int, int w) {
        if (w == 4) {
            const float4 v = *reinterpret_cast<const float4*>(x + off);
            const float4 g = *reinterpret_cast<const float4*>(dy + off);
            float4 o;
            o.x = scale * (g.x - r0 - (v.x - mu) * is * r1); o.y = scale * (g.y - r0 - (v.y - mu) * is * r1);
            o.z = scale * (g.z - r0 - (v.z - mu) * is * r1); o.w = scale * (g.w - r0 - (v.w - mu) * is * r1);
            *reinterpret_cast<float4*>(dx + off) = o;
        } else {
            dx[off] = scale * (dy[off] - r0 - (x[off] - mu) * is * r1);
        }
    });
}

static bool bn_small_ok(int N, int C, int HW) { return (long long)N * HW <= BN_SMALL_MAX && C >= 32; }

static int reduce_splits(int N, int C, int HW) {
    const long long per_c = (long long)N * HW;
    long long want = (4LL * 148 + C - 1) / C;             // ~4 CTAs per SM in total
    long long maxs = (per_c + 4 * BN_THREADS - 1) / (4 * BN_THREADS);  // >= 4 elements per thread
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return (int)want;
}

static int validate_bn(int N, int C, int HW) {
    EEGAN_REQUIRE(N > 0 && C > 0 && HW > 0, "syncbn: empty shape N=%d C=%d HW=%d", N, C, HW);
    EEGAN_REQUIRE((long long)N * C <= 0x7fffffffLL, "syncbn: N*C too large");
    return EEGAN_OK;
}

}  // namespace eegan

using namespace eegan;

extern "C" int eegan_syncbn_stats(const float* x, int N, int C, int HW, float* stats, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && stats, "syncbn stats: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(stats, 0, 2 * (size_t)C * sizeof(float), st);
    bn_reduce_kernel<0><<<dim3(C, reduce_splits(N, C, HW)), BN_THREADS, 0, st>>>(x, nullptr, nullptr, nullptr, N, C, HW, stats);
    EEGAN_LAUNCH_CHECK("syncbn stats");
    return EEGAN_OK;
}

__global__ void bn_count_kernel(float* __restrict__ cnt, float hi, float lo) {
    cnt[0] = hi;
    cnt[1] = lo;
}

// stats [2*C + 2]: the statistics of eegan_syncbn_stats followed by the local element count N*HW as the exact pair
// {count / 4096, count % 4096} — the whole buffer is then ONE all-reduce and eegan_syncbn_finalize reads the total from
// its tail (count_dev = stats + 2*C): no host-side scalar writes per layer.
extern "C" int eegan_syncbn_stats_counted(const float* x, int N, int C, int HW, float* stats, void* stream) {
    int rc = eegan_syncbn_stats(x, N, C, HW, stats, stream);
    if (rc) return rc;
    const long long local = (long long)N * HW;
    bn_count_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stats + 2 * (size_t)C, (float)(local / 4096), (float)(local % 4096));
    EEGAN_LAUNCH_CHECK("syncbn stats (count)");
    return EEGAN_OK;
}

extern "C" int eegan_syncbn_finalize(const float* stats, int C, double count, const float* count_dev, float eps, float momentum,
                                     int clamp_mode, float* mean, float* inv_std, float* running_mean,
                                     float* running_var, void* stream) {
    EEGAN_REQUIRE(C > 0 && stats && mean && inv_std, "syncbn finalize: bad arguments");
    EEGAN_REQUIRE(count_dev || count > 1.0, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");  // batchnorm.py:115
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, C, (float)count, count_dev, eps, momentum, clamp_mode,
                                                                          mean, inv_std, running_mean, running_var);
    EEGAN_LAUNCH_CHECK("syncbn finalize");
    return EEGAN_OK;
}

static dim3 apply_grid(int N, int C, int HW) {
    int per = (HW / 4 + BN_THREADS - 1) / BN_THREADS;
    if (per < 1) per = 1;
    if (per > 64) per = 64;
    return dim3(per, N * C);
}

extern "C" int eegan_syncbn_apply(const float* x, const float* mean, const float* inv_std, const float* weight,
                                  const float* bias, int N, int C, int HW, float* y, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && mean && inv_std && y, "syncbn apply: null pointer");
    bn_apply_kernel<0><<<apply_grid(N, C, HW), BN_THREADS, 0, (cudaStream_t)stream>>>(
        x, nullptr, mean, inv_std, weight, bias, nullptr, 1.f, nullptr, 0, 0.f, C, HW, y);
    EEGAN_LAUNCH_CHECK("syncbn apply");
    return EEGAN_OK;
}

extern "C" int eegan_syncbn_bwd_reduce(const float* x, const float* dy, const float* mean, const float* inv_std, int N,
                                       int C, int HW, float* red, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && dy && mean && inv_std && red, "syncbn bwd_reduce: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(red, 0, 2 * (size_t)C * sizeof(float), st);
    bn_reduce_kernel<1><<<dim3(C, reduce_splits(N, C, HW)), BN_THREADS, 0, st>>>(x, dy, mean, inv_std, N, C, HW, red);
    EEGAN_LAUNCH_CHECK("syncbn bwd_reduce");
    return EEGAN_OK;
}

extern "C" int eegan_syncbn_bwd_apply(const float* x, const float* dy, const float* mean, const float* inv_std,
                                      const float* weight, const float* red, double count, const float* count_dev,
                                      float eps, int clamp_mode,
                                      int N, int C, int HW, float* dx, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && dy && mean && inv_std && red && dx && (count_dev || count > 0), "syncbn bwd_apply: bad arguments");
    bn_apply_kernel<1><<<apply_grid(N, C, HW), BN_THREADS, 0, (cudaStream_t)stream>>>(
        x, dy, mean, inv_std, weight, nullptr, red, (float)count, count_dev, clamp_mode, 1.0f / sqrtf(eps), C, HW, dx);
    EEGAN_LAUNCH_CHECK("syncbn bwd_apply");
    return EEGAN_OK;
}

// Single-replica forward / backward in ONE call each (the caller passes work [4*C]: statistics scratch [2C], then the
// mean [C] and inv_std [C] the backward needs).  Small maps take the one-launch kernels above, the rest the same
// stats -> finalize -> apply sequence as the separate entry points.
extern "C" int eegan_syncbn_fwd_fused(const float* x, const float* weight, const float* bias, int N, int C, int HW, float eps,
                                      float momentum, float* running_mean, float* running_var, float* y, float* work, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && y && work, "syncbn fwd_fused: null pointer");
    EEGAN_REQUIRE((long long)N * HW > 1, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");
    float *mean = work + 2 * (size_t)C, *inv_std = work + 3 * (size_t)C;
    cudaStream_t st = (cudaStream_t)stream;
    if (bn_small_ok(N, C, HW)) {
        bn_fwd_small_kernel<0><<<C, BN_SMALL_THREADS, 0, st>>>(x, weight, bias, nullptr, nullptr, nullptr, N, C, HW, eps, momentum,
                                                              running_mean, running_var, y, mean, inv_std);
        return check_launch("syncbn fwd (one launch)");
    }
    if ((rc = eegan_syncbn_stats(x, N, C, HW, work, stream))) return rc;
    if ((rc = eegan_syncbn_finalize(work, C, (double)N * HW, nullptr, eps, momentum, 0, mean, inv_std, running_mean, running_var, stream))) return rc;
    return eegan_syncbn_apply(x, mean, inv_std, weight, bias, N, C, HW, y, stream);
}

extern "C" int eegan_syncbn_bwd_fused(const float* x, const float* dy, const float* work, const float* weight, int N, int C, int HW,
                                      float eps, float* dx, float* red, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && dy && work && dx && red, "syncbn bwd_fused: null pointer");
    const float *mean = work + 2 * (size_t)C, *inv_std = work + 3 * (size_t)C;
    if (bn_small_ok(N, C, HW)) {
        bn_bwd_small_kernel<<<C, BN_SMALL_THREADS, 0, (cudaStream_t)stream>>>(x, dy, mean, inv_std, weight, N, C, HW, dx, red);
        return check_launch("syncbn bwd (one launch)");
    }
    if ((rc = eegan_syncbn_bwd_reduce(x, dy, mean, inv_std, N, C, HW, red, stream))) return rc;
    return eegan_syncbn_bwd_apply(x, dy, mean, inv_std, weight, red, (double)N * HW, nullptr, eps, 0, N, C, HW, dx, stream);
}

// =======================================================================================
// affine_ssa (models.py:43-86): SyncBN(affine=False) followed by the mask-gated modulation
//   out = (gamma[n,c] * mask[n,hw] + 1) * xhat + beta[n,c] * mask[n,hw]
// The reference normalises (three passes), expands gamma / beta to the full tensor and runs four
// more elementwise passes; here the modulation rides on the normalise pass, and the backward is
// the two passes batch norm needs anyway (one reduction, one apply) with d_gamma, d_beta and
// d_mask produced by the reduction pass.  SURVEY.md §8f rank 3.
// =======================================================================================
namespace eegan {

// MODE 0: y = (g m + 1) xhat + b m ;  MODE 1: dx = inv_std * (dxh - r0 - xhat r1), dxh = dy (g m + 1)
template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) ssa_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ inv_std,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               const float* __restrict__ mask,
                                                               const float* __restrict__ red, float host_count,
                                                               const float* __restrict__ count_dev, int kill_var_term,
                                                               float clamp_inv_std, int C, int HW,
                                                               float* __restrict__ out) {
    const int nc = blockIdx.y;  // n*C + c
    const int n = nc / C, c = nc - n * C;
    const float mu = mean[c], is = inv_std[c];
    const float g = gamma[nc], b = MODE == 0 ? beta[nc] : 0.f;
    float r0 = 0.f, r1 = 0.f;
    if (MODE == 1 && red) {
        const float inv_count = 1.0f / dev_count(count_dev, host_count);
        r0 = red[c] * inv_count;
        r1 = (kill_var_term && is >= clamp_inv_std) ? 0.f : red[C + c] * inv_count;
    }
    const size_t base = (size_t)nc * HW;
    const float* mrow = mask + (size_t)n * HW;
    const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                         reinterpret_cast<uintptr_t>(mask) |
                                         (MODE == 1 ? reinterpret_cast<uintptr_t>(dy) : 0)) & 15) == 0);
    if (vec) {
        const int hw4 = HW / 4;
        for (int k = blockIdx.x * BN_THREADS + threadIdx.x; k < hw4; k += gridDim.x * BN_THREADS) {
            const float4 xv = *reinterpret_cast<const float4*>(x + base + (size_t)k * 4);
            const float4 mv = *reinterpret_cast<const float4*>(mrow + (size_t)k * 4);
            float4 o;
            if (MODE == 0) {
                o.x = fmaf(fmaf(g, mv.x, 1.f), (xv.x - mu) * is, b * mv.x); o.y = fmaf(fmaf(g, mv.y, 1.f), (xv.y - mu) * is, b * mv.y);
                o.z = fmaf(fmaf(g, mv.z, 1.f), (xv.z - mu) * is, b * mv.z); o.w = fmaf(fmaf(g, mv.w, 1.f), (xv.w - mu) * is, b * mv.w);
            } else {
                const float4 gv = *reinterpret_cast<const float4*>(dy + base + (size_t)k * 4);
                o.x = is * (gv.x * fmaf(g, mv.x, 1.f) - r0 - (xv.x - mu) * is * r1);
                o.y = is * (gv.y * fmaf(g, mv.y, 1.f) - r0 - (xv.y - mu) * is * r1);
                o.z = is * (gv.z * fmaf(g, mv.z, 1.f) - r0 - (xv.z - mu) * is * r1);
                o.w = is * (gv.w * fmaf(g, mv.w, 1.f) - r0 - (xv.w - mu) * is * r1);
            }
            *reinterpret_cast<float4*>(out + base + (size_t)k * 4) = o;
        }
    } else {
        for (int k = blockIdx.x * BN_THREADS + threadIdx.x; k < HW; k += gridDim.x * BN_THREADS) {
            const float xh = (x[base + k] - mu) * is, m = mrow[k];
            out[base + k] = (MODE == 0) ? fmaf(fmaf(g, m, 1.f), xh, b * m) : is * (dy[base + k] * fmaf(g, m, 1.f) - r0 - xh * r1);
        }
    }
}

// Backward reduction.  CTA = (sample n, tile of 1024 pixels), thread = 4 pixels, loop over channels:
//   dxh = dy (g m + 1);  per (n, c): S0 = sum dxh, S1 = sum dxh xhat, S2 = sum dy xhat m (d_gamma), S3 = sum dy m (d_beta);
//   per pixel: d_mask = sum_c dy (g xhat + b)      (each pixel has one owner thread: no atomics)
// The per-warp parts of S0..S3 for every channel wait in shared memory; after the channel loop the CTA adds
// them up and issues 4 C atomics (red[c], red[C+c] across samples and tiles; d_gamma / d_beta across tiles).
__global__ void __launch_bounds__(BN_THREADS) ssa_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ inv_std,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta,
                                                                    const float* __restrict__ mask, int C, int HW,
                                                                    float* __restrict__ red, float* __restrict__ dgamma,
                                                                    float* __restrict__ dbeta, float* __restrict__ dmask) {
    extern __shared__ float s_part[];  // [C][8 warps][4]
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p0 = blockIdx.x * (BN_THREADS * 4) + threadIdx.x * 4;
    const bool vec = (HW % 4 == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) |
                                         reinterpret_cast<uintptr_t>(mask)) & 15) == 0);
    float m[4], dm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = (p0 + k < HW) ? mask[(size_t)n * HW + p0 + k] : 0.f;
    auto load4 = [&](const float* src, float (&v)[4]) {
        if (vec && p0 + 3 < HW) {
            const float4 t = *reinterpret_cast<const float4*>(src + p0);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (p0 + k < HW) ? src[p0 + k] : 0.f;
        }
    };
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
        const size_t base = ((size_t)n * C + c) * HW;
        float xv[4], gv[4];
        load4(x + base, xv);
        load4(dy + base, gv);
        const float mu = __ldg(mean + c), is = __ldg(inv_std + c), g = __ldg(gamma + (size_t)n * C + c), b = __ldg(beta + (size_t)n * C + c);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xh = (p0 + k < HW) ? (xv[k] - mu) * is : 0.f;
            const float dxh = gv[k] * fmaf(g, m[k], 1.f);
            s0 += dxh;
            s1 = fmaf(dxh, xh, s1);
            s2 = fmaf(gv[k] * xh, m[k], s2);
            s3 = fmaf(gv[k], m[k], s3);
            dm[k] = fmaf(gv[k], fmaf(g, xh, b), dm[k]);
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        if (lane == 0) *reinterpret_cast<float4*>(s_part + ((size_t)c * 8 + warp) * 4) = make_float4(s0, s1, s2, s3);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (p0 + k < HW) dmask[(size_t)n * HW + p0 + k] = dm[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float4 o = *reinterpret_cast<const float4*>(s_part + ((size_t)c * 8 + w) * 4);
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        atomicAdd(red + c, t.x);
        atomicAdd(red + C + c, t.y);
        atomicAdd(dgamma + (size_t)n * C + c, t.z);
        atomicAdd(dbeta + (size_t)n * C + c, t.w);
    }
}

}  // namespace eegan

extern "C" int eegan_ssa_apply(const float* x, const float* mean, const float* inv_std, const float* gamma, const float* beta,
                               const float* mask, int N, int C, int HW, float* y, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && mean && inv_std && gamma && beta && mask && y, "ssa apply: null pointer");
    ssa_apply_kernel<0><<<apply_grid(N, C, HW), BN_THREADS, 0, (cudaStream_t)stream>>>(x, nullptr, mean, inv_std, gamma, beta, mask, nullptr,
                                                                                       1.f, nullptr, 0, 0.f, C, HW, y);
    EEGAN_LAUNCH_CHECK("ssa apply");
    return EEGAN_OK;
}

extern "C" int eegan_ssa_bwd_reduce(const float* x, const float* dy, const float* mean, const float* inv_std, const float* gamma,
                                    const float* beta, const float* mask, int N, int C, int HW, float* red, float* dgamma,
                                    float* dbeta, float* dmask, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && dy && mean && inv_std && gamma && beta && mask && red && dgamma && dbeta && dmask, "ssa bwd_reduce: null pointer");
    EEGAN_REQUIRE(C <= 1536 && N <= 65535, "ssa bwd_reduce: C=%d (<=1536) N=%d (<=65535) unsupported", C, N);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(red, 0, 2 * (size_t)C * sizeof(float), st);
    cudaMemsetAsync(dgamma, 0, (size_t)N * C * sizeof(float), st);
    cudaMemsetAsync(dbeta, 0, (size_t)N * C * sizeof(float), st);
    const size_t smem = (size_t)C * 8 * 4 * sizeof(float);
    static SmemGrant grant;
    if (int rc = grant_dyn_smem(ssa_bwd_reduce_kernel, smem, grant, "ssa bwd_reduce")) return rc;
    ssa_bwd_reduce_kernel<<<dim3((HW + BN_THREADS * 4 - 1) / (BN_THREADS * 4), N), BN_THREADS, smem, st>>>(
        x, dy, mean, inv_std, gamma, beta, mask, C, HW, red, dgamma, dbeta, dmask);
    EEGAN_LAUNCH_CHECK("ssa bwd_reduce");
    return EEGAN_OK;
}

extern "C" int eegan_ssa_bwd_apply(const float* x, const float* dy, const float* mean, const float* inv_std, const float* gamma,
                                   const float* mask, const float* red, double count, const float* count_dev, float eps,
                                   int clamp_mode, int N, int C, int HW, float* dx, void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && dy && mean && inv_std && gamma && mask && dx, "ssa bwd_apply: null pointer");
    EEGAN_REQUIRE(!red || count_dev || count > 0, "ssa bwd_apply: element count missing");
    ssa_apply_kernel<1><<<apply_grid(N, C, HW), BN_THREADS, 0, (cudaStream_t)stream>>>(
        x, dy, mean, inv_std, gamma, nullptr, mask, red, (float)count, count_dev, clamp_mode, 1.0f / sqrtf(eps), C, HW, dx);
    EEGAN_LAUNCH_CHECK("ssa bwd_apply");
    return EEGAN_OK;
}

// Single-replica forward of affine_ssa in one call (work [4*C] as for eegan_syncbn_fwd_fused).
extern "C" int eegan_ssa_fwd_fused(const float* x, const float* gamma, const float* beta, const float* mask, int N, int C, int HW,
                                   float eps, float momentum, float* running_mean, float* running_var, float* y, float* work,
                                   void* stream) {
    int rc = validate_bn(N, C, HW);
    if (rc) return rc;
    EEGAN_REQUIRE(x && gamma && beta && mask && y && work, "ssa fwd_fused: null pointer");
    EEGAN_REQUIRE((long long)N * HW > 1, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");
    float *mean = work + 2 * (size_t)C, *inv_std = work + 3 * (size_t)C;
    if (bn_small_ok(N, C, HW)) {
        bn_fwd_small_kernel<1><<<C, BN_SMALL_THREADS, 0, (cudaStream_t)stream>>>(x, nullptr, nullptr, gamma, beta, mask, N, C, HW, eps,
                                                                                momentum, running_mean, running_var, y, mean, inv_std);
        return check_launch("ssa fwd (one launch)");
    }
    if ((rc = eegan_syncbn_stats(x, N, C, HW, work, stream))) return rc;
    if ((rc = eegan_syncbn_finalize(work, C, (double)N * HW, nullptr, eps, momentum, 0, mean, inv_std, running_mean, running_var, stream))) return rc;
    return eegan_ssa_apply(x, mean, inv_std, gamma, beta, mask, N, C, HW, y, stream);
}
