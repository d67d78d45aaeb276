// Context, error reporting.
#include <cstring>

#include "common.cuh"

namespace sb {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
}  // namespace sb

using namespace sb;

extern "C" {

const char* sofab200_version(void) { return "sofa_b200 0.1 (sm_100a)"; }
const char* sofab200_last_error(void) { return g_last_error.c_str(); }

int sofab200_ctx_create(int device, void* cuda_stream, sofab200_ctx** out) {
    SB_CHECK(out != nullptr, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SOFAB200_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); sofa_b200 has no CPU path");
    SB_CHECK(device >= 0 && device < count, "device ordinal out of range");
    SB_CUDA(cudaSetDevice(device));
    sofab200_ctx* c = new sofab200_ctx();
    c->device = device;
    if (cuda_stream) { c->stream = static_cast<cudaStream_t>(cuda_stream); c->own_stream = false; }
    else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) { delete c; return fail(SOFAB200_ERR_CUDA, cudaGetErrorString(e2)); }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    int rc = c->red_partials.alloc(4096);
    if (rc == SOFAB200_OK) rc = c->red_result.alloc(4);
    if (rc == SOFAB200_OK) rc = c->red_counter.alloc(4);
    if (rc == SOFAB200_OK) rc = c->red_counter.zero(c->stream);
    if (rc != SOFAB200_OK) { delete c; return rc; }
    *out = c;
    return SOFAB200_OK;
}
int sofab200_ctx_destroy(sofab200_ctx* ctx) {
    if (!ctx) return SOFAB200_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
    delete ctx;
    return SOFAB200_OK;
}
int sofab200_ctx_set_stream(sofab200_ctx* ctx, void* cuda_stream) {
    SB_CHECK(ctx != nullptr, "ctx is null");
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (cuda_stream) ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    else { SB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return SOFAB200_OK;
}
int sofab200_ctx_synchronize(sofab200_ctx* ctx) {
    SB_CHECK(ctx != nullptr, "ctx is null");
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SOFAB200_OK;
}
uint64_t sofab200_ctx_launch_count(const sofab200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int sofab200_peer_alloc(sofab200_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char handle[SOFAB200_IPC_HANDLE_BYTES]) {
    SB_CHECK(ctx && dev_ptr && handle && bytes > 0, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SOFAB200_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    SB_CUDA(cudaSetDevice(ctx->device));
    SB_CUDA(cudaMalloc(dev_ptr, bytes));
    SB_CUDA(cudaMemset(*dev_ptr, 0, bytes));
    cudaIpcMemHandle_t h;
    SB_CUDA(cudaIpcGetMemHandle(&h, *dev_ptr));
    std::memcpy(handle, &h, sizeof(h));
    SB_CUDA(cudaDeviceSynchronize());
    return SOFAB200_OK;
}
int sofab200_peer_open(sofab200_ctx* ctx, const unsigned char handle[SOFAB200_IPC_HANDLE_BYTES], void** dev_ptr) {
    SB_CHECK(ctx && dev_ptr && handle, "null argument");
    SB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    SB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SOFAB200_OK;
}
int sofab200_peer_close(sofab200_ctx* ctx, void* dev_ptr) {
    SB_CHECK(ctx != nullptr, "ctx is null");
    if (dev_ptr) SB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return SOFAB200_OK;
}
int sofab200_peer_free(sofab200_ctx* ctx, void* dev_ptr) {
    SB_CHECK(ctx != nullptr, "ctx is null");
    if (dev_ptr) SB_CUDA(cudaFree(dev_ptr));
    return SOFAB200_OK;
}
int sofab200_ctx_trace_begin(sofab200_ctx* ctx) {
    SB_CHECK(ctx != nullptr, "ctx is null");
    SB_TRY(ctx->trace.alloc(size_t(2) * 4096 * 16));
    return ctx->trace.zero(ctx->stream);
}
int sofab200_ctx_trace_end(sofab200_ctx* ctx, uint64_t* out, size_t n) {
    SB_CHECK(ctx && out, "null argument");
    SB_CHECK(ctx->trace.p != nullptr, "trace_begin was not called");
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_CUDA(cudaMemcpy(out, ctx->trace.p, std::min(n, ctx->trace.n) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    ctx->trace.release();
    return SOFAB200_OK;
}
int sofab200_ctx_profile_begin(sofab200_ctx* ctx) {
    SB_CHECK(ctx != nullptr, "ctx is null");
    for (auto& v : ctx->prof) { for (auto& e : v) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } v.clear(); }
    ctx->profiling = true;
    return SOFAB200_OK;
}
int sofab200_ctx_profile_end(sofab200_ctx* ctx, double* total_ms, uint64_t* count) {
    SB_CHECK(ctx && total_ms && count, "null argument");
    ctx->profiling = false;
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < SOFAB200_PROFILE_CLASSES; ++c) {
        double tot = 0.0;
        for (auto& e : ctx->prof[c]) {
            float ms = 0.f;
            SB_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
            tot += ms;
            cudaEventDestroy(e.first); cudaEventDestroy(e.second);
        }
        total_ms[c] = tot; count[c] = ctx->prof[c].size();
        ctx->prof[c].clear();
    }
    return SOFAB200_OK;
}

}  // extern "C"
