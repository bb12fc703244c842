// Shared helpers for libeegan_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/eegan_b200.h"

namespace eegan {

// thread-local error text returned by eegan_last_error()
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return EEGAN_ERR_CUDA;
    }
    return EEGAN_OK;
}

#define EEGAN_REQUIRE(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ::eegan::set_error(__VA_ARGS__); \
            return EEGAN_ERR_INVALID;       \
        }                                   \
    } while (0)

#define EEGAN_LAUNCH_CHECK(what)                       \
    do {                                               \
        int _rc = ::eegan::check_launch(what);         \
        if (_rc != EEGAN_OK) return _rc;               \
    } while (0)

// bench-only stage timing (profile.cu); a no-op unless eegan_profile_enable(1) was called
void prof_mark(int stage, cudaStream_t st);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, so the
// "largest size granted so far" is kept PER DEVICE (one zero-initialised SmemGrant per launch site): a second GPU
// driven from the same process (nn.DataParallel's per-device threads) gets its own opt-in.  A race between two host
// threads only repeats the idempotent call.
constexpr int EEGAN_MAX_DEVICES = 64;
struct SmemGrant {
    std::atomic<size_t> per_dev[EEGAN_MAX_DEVICES];
};
template <typename K>
inline int grant_dyn_smem(K kern, size_t bytes, SmemGrant& g, const char* what) {
    if (bytes <= 48 * 1024) return EEGAN_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    const bool tracked = dev < EEGAN_MAX_DEVICES;
    if (tracked && bytes <= g.per_dev[dev].load(std::memory_order_relaxed)) return EEGAN_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("%s: shared-memory opt-in of %zu bytes failed: %s", what, bytes, cudaGetErrorString(e));
        return EEGAN_ERR_CUDA;
    }
    if (tracked) g.per_dev[dev].store(bytes, std::memory_order_relaxed);
    return EEGAN_OK;
}

// SM count of the CURRENT device (cached per device)
int num_sms_current();

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A kernel launched through launch_pdl() may become resident while the previous kernel of the stream is still
// draining: its CTAs run their set-up (barriers, TMEM allocation, tensor-map prefetch) and then block in
// pdl_wait() until the previous grid has completed and its writes are visible.  Every kernel of a PDL chain
// therefore calls pdl_trigger() first (lets the NEXT kernel's CTAs be scheduled once all CTAs of this grid run)
// and pdl_wait() before its first global-memory access.  Both are no-ops in a kernel launched the plain way.
// EEGAN_PDL=0 turns the launch attribute off (A/B timing).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32).  `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : -INFINITY;
    t = warp_max(t);
    return t;
}

}  // namespace eegan
