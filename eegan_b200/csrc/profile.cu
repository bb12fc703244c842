// Optional per-stage device timing for bench.py: when enabled the
// multi-kernel entry points drop a cudaEvent after each stage on the caller's stream;
// eegan_profile_collect() then synchronises those events and returns summed durations per
// stage.  Off by default: the hot path records nothing and never synchronises.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace eegan {

struct Mark {
    cudaEvent_t ev;
    int stage;  // -1 = start of a call
};
// process-wide (autograd runs the backward on its own thread), guarded by a mutex; only ever
// touched when the bench switched profiling on
static std::atomic<bool> g_on{false};
static std::mutex g_mu;
static std::vector<Mark> g_marks;
static std::vector<cudaEvent_t> g_pool;

static const char* kStageNames[EEGAN_PROF_NSTAGES] = {
    "prologue(pack)", "gemm1(S=W.C)+attn_fwd", "attn_softmax(unfused engines)", "gemm2(U=A.C^T)", "cos_lse+att_maps",
    "bwd_scalars+du", "gemm3(dA=dU.C)+attn_bwd", "softmax_bwd(unfused engines)", "gemm4(dC)", "gemm5(dW)", "unpack(d_words)",
};

void prof_mark(int stage, cudaStream_t st) {
    if (!g_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEvent_t ev;
    if (!g_pool.empty()) {
        ev = g_pool.back();
        g_pool.pop_back();
    } else if (cudaEventCreate(&ev) != cudaSuccess) {
        return;
    }
    cudaEventRecord(ev, st);
    g_marks.push_back({ev, stage});
}

}  // namespace eegan

using namespace eegan;

extern "C" int eegan_profile_enable(int on) {
    g_on = on != 0;
    return EEGAN_OK;
}

extern "C" int eegan_profile_nstages(void) { return EEGAN_PROF_NSTAGES; }

extern "C" const char* eegan_profile_stage_name(int stage) {
    return (stage >= 0 && stage < EEGAN_PROF_NSTAGES) ? kStageNames[stage] : "";
}

extern "C" int eegan_profile_collect(double* stage_ms, int* stage_launches) {
    EEGAN_REQUIRE(stage_ms && stage_launches, "profile_collect: null pointer");
    std::lock_guard<std::mutex> lk(g_mu);
    for (int s = 0; s < EEGAN_PROF_NSTAGES; ++s) { stage_ms[s] = 0.0; stage_launches[s] = 0; }
    for (size_t k = 0; k < g_marks.size(); ++k) {
        if (g_marks[k].stage < 0 || k == 0) continue;
        cudaError_t e = cudaEventSynchronize(g_marks[k].ev);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_marks[k - 1].ev, g_marks[k].ev);
        if (e != cudaSuccess) { set_error("profile_collect: %s", cudaGetErrorString(e)); return EEGAN_ERR_CUDA; }
        stage_ms[g_marks[k].stage] += ms;
        stage_launches[g_marks[k].stage] += 1;
    }
    for (auto& m : g_marks) g_pool.push_back(m.ev);
    g_marks.clear();
    return EEGAN_OK;
}
