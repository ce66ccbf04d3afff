// Context lifetime, workspace arena, error reporting.
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <vector>

namespace vpm {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

// ---- profiling records (one host thread per GPU drives the library) ------------------------
bool g_prof_on = false;
namespace {
struct ProfRec {
    const char* group;
    const char* name;
    double bytes;        // algorithmic bytes stated for this launch (0: none stated)
    double group_bytes;  // stated once per group instance, on its first launch
    cudaEvent_t e0, e1;
};
std::vector<ProfRec> g_prof;
const char* g_prof_group = "";
double g_prof_bytes = 0.0;
double g_prof_group_bytes = 0.0;
}  // namespace

void prof_begin(const char* name, cudaStream_t st) {
    ProfRec r;
    r.group = g_prof_group;
    r.name = name;
    r.bytes = g_prof_bytes;
    r.group_bytes = g_prof_group_bytes;
    g_prof_bytes = 0.0;
    g_prof_group_bytes = 0.0;
    VPM_CUDA(cudaEventCreate(&r.e0));
    VPM_CUDA(cudaEventCreate(&r.e1));
    VPM_CUDA(cudaEventRecord(r.e0, st));
    g_prof.push_back(r);
}

void prof_end(cudaStream_t st) {
    if (!g_prof.empty()) VPM_CUDA(cudaEventRecord(g_prof.back().e1, st));
}

void prof_next_bytes(double bytes) { g_prof_bytes = bytes; }

ProfGroup::ProfGroup(const char* group, double group_bytes) : prev(g_prof_group) {
    g_prof_group = group;
    if (g_prof_on) g_prof_group_bytes = group_bytes;
}
ProfGroup::~ProfGroup() {
    g_prof_group = prev;
    g_prof_group_bytes = 0.0;
}
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

// Launch the next calls on `stream` WITHOUT draining the current one and without re-binding the
// cuFFT plans: for kernels that use neither cuFFT nor the context workspace (the peer puts of a
// pipelined slab transform run beside the 2-D FFTs this way); the caller orders the two streams
// with events.  vpm_ctx_pop_stream restores the stream that was current before the push.
int vpm_ctx_push_stream(vpm_ctx* ctx, void* stream) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx != nullptr, "ctx is NULL");
    VPM_REQUIRE(ctx->pushed_from == nullptr && !ctx->pushed, "a stream is already pushed");
    ctx->pushed_from = ctx->stream;
    ctx->pushed = true;
    ctx->stream = reinterpret_cast<cudaStream_t>(stream);
    VPM_API_END
}

int vpm_ctx_pop_stream(vpm_ctx* ctx) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx != nullptr && ctx->pushed, "no pushed stream");
    ctx->stream = ctx->pushed_from;
    ctx->pushed_from = nullptr;
    ctx->pushed = false;
    VPM_API_END
}

int vpm_ctx_set_stream(vpm_ctx* ctx, void* stream) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx != nullptr, "ctx is NULL");
    VPM_REQUIRE(!ctx->pushed, "vpm_ctx_set_stream while a stream is pushed");
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

int vpm_profile_start(void) {
    VPM_API_BEGIN
    for (auto& r : vpm::g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    vpm::g_prof.clear();
    vpm::g_prof_on = true;
    VPM_API_END
}

int vpm_profile_stop(void) {
    vpm::g_prof_on = false;
    return 0;
}

// One text line per profiled launch: "group\tname\tmilliseconds\tbytes\tgroup_bytes\n".
// Synchronises on the recorded events.  Returns the number of bytes the full text needs
// (call with buf = NULL / cap = 0 first); the records are released when the text fitted.
int64_t vpm_profile_read(char* buf, int64_t cap) {
    try {
        std::string out;
        char line[512];
        for (auto& r : vpm::g_prof) {
            float ms = 0.f;
            VPM_CUDA(cudaEventSynchronize(r.e1));
            VPM_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
            snprintf(line, sizeof(line), "%s\t%s\t%.6f\t%.1f\t%.1f\n", r.group, r.name, (double)ms,
                     r.bytes, r.group_bytes);
            out += line;
        }
        const int64_t need = (int64_t)out.size() + 1;
        if (buf && cap >= need) {
            memcpy(buf, out.c_str(), (size_t)need);
            for (auto& r : vpm::g_prof) {
                cudaEventDestroy(r.e0);
                cudaEventDestroy(r.e1);
            }
            vpm::g_prof.clear();
        }
        return need;
    } catch (const std::exception& e) {
        vpm::g_last_error = e.what();
        return -1;
    }
}

int vpm_device_sm_count(vpm_ctx* ctx, int* out_host) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && out_host, "NULL argument");
    *out_host = ctx->sm_count;
    VPM_API_END
}

}  // extern "C"
