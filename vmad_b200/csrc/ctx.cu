// Context lifetime, workspace arena, error reporting.
#include "common.cuh"

namespace vpm {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
}  // namespace vpm

void* vpm_ctx::workspace(size_t bytes) {
    if (bytes <= ws_bytes) return ws;
    // grow-only; the old block may still be in use by enqueued kernels
    VPM_CUDA(cudaStreamSynchronize(stream));
    if (ws) VPM_CUDA(cudaFree(ws));
    ws = nullptr;
    ws_bytes = 0;
    size_t want = bytes + bytes / 4 + (1 << 20);
    VPM_CUDA(cudaMalloc(&ws, want));
    ws_bytes = want;
    return ws;
}

extern "C" {

int vpm_ctx_create(vpm_ctx** out, int device, void* stream) {
    VPM_API_BEGIN
    VPM_REQUIRE(out != nullptr, "ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw vpm::Error("vmadpm: no CUDA device available (this library has no CPU path)");
    VPM_REQUIRE(device >= 0 && device < ndev, "ctx_create: bad device index");
    VPM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VPM_CUDA(cudaGetDeviceProperties(&prop, device));
    vpm_ctx* c = new vpm_ctx();
    c->device = device;
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    c->sm_count = prop.multiProcessorCount;
    VPM_CUDA(cudaMallocHost(&c->host_scalars, 64 * sizeof(double)));
    VPM_CUDA(cudaMalloc(&c->dev_scalars, 64 * sizeof(double)));
    *out = c;
    VPM_API_END
}

int vpm_ctx_destroy(vpm_ctx* ctx) {
    VPM_API_BEGIN
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->plans) cufftDestroy(kv.second);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->host_scalars) cudaFreeHost(ctx->host_scalars);
    if (ctx->dev_scalars) cudaFree(ctx->dev_scalars);
    delete ctx;
    VPM_API_END
}

int vpm_ctx_set_stream(vpm_ctx* ctx, void* stream) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx != nullptr, "ctx is NULL");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (s != ctx->stream) {
        VPM_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->stream = s;
        for (auto& kv : ctx->plans) {
            if (cufftSetStream(kv.second, s) != CUFFT_SUCCESS)
                throw vpm::Error("vmadpm: cufftSetStream failed");
        }
    }
    VPM_API_END
}

int vpm_sync(vpm_ctx* ctx) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx != nullptr, "ctx is NULL");
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    VPM_API_END
}

const char* vpm_last_error(void) { return vpm::g_last_error.c_str(); }

int vpm_version(void) { return 100; }

int64_t vpm_launch_count(void) { return vpm::g_launches.load(); }

int vpm_device_sm_count(vpm_ctx* ctx, int* out_host) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && out_host, "NULL argument");
    *out_host = ctx->sm_count;
    VPM_API_END
}

}  // extern "C"
