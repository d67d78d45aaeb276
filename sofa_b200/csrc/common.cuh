// Common definitions of the sofa_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sofa_b200.h"

#define HD __host__ __device__ __forceinline__

namespace sb {

// ---- error handling ---------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#define SB_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return ::sb::fail(SOFAB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define SB_CHECK(cond, msg)                                              \
    do {                                                                 \
        if (!(cond)) return ::sb::fail(SOFAB200_ERR_INVALID, (msg));     \
    } while (0)
#define SB_TRY(call)                      \
    do {                                  \
        int rc_ = (call);                 \
        if (rc_ != SOFAB200_OK) return rc_; \
    } while (0)

// ---- 16-byte aligned quad of Reals: one LDG.128 (float) or two (double) -----------------------
template <class R> struct alignas(16) Quad { R a, b, c, d; };

// ---- device buffer (RAII, context's device) ----------------------------------------------------
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return SOFAB200_OK;
        SB_CUDA(cudaMalloc(&p, count * sizeof(T)));
        return SOFAB200_OK;
    }
    int upload(const std::vector<T>& h, cudaStream_t s) {
        SB_TRY(alloc(h.size()));
        if (!h.empty()) SB_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
        return SOFAB200_OK;
    }
    int zero(cudaStream_t s) {
        if (n) SB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
        return SOFAB200_OK;
    }
};

// (mesh_mass.cu) what a solver node needs to know about a MeshMatrixMass handle
void meshmass_info(const sofab200_meshmass* mm, int* real, size_t* n_nodes, const sofab200_ctx** ctx);

}  // namespace sb

struct sofab200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    cudaStream_t capture_stream = nullptr;   // stream captures run here (the legacy default stream cannot be captured)
    // optional per-class event timing (sofab200_ctx_profile_begin/end)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[SOFAB200_PROFILE_CLASSES];
    void prof_start(int cls) {
        if (!profiling) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, stream);
        prof[cls].push_back({a, b});
    }
    void prof_stop(int cls) {
        if (!profiling) return;
        cudaEventRecord(prof[cls].back().second, stream);
    }
    // optional in-kernel timestamps (sofab200_ctx_trace_begin/end): [CTA][8] globaltimer values of the last launch
    sb::DevBuf<unsigned long long> trace;
    // scratch for reductions (vdot): partial sums + result + counter
    sb::DevBuf<double> red_partials;
    sb::DevBuf<double> red_result;
    sb::DevBuf<unsigned> red_counter;
};
