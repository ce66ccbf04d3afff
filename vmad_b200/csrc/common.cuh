// Shared plumbing of libvmadpm: context, error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <atomic>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>

#include "../../include/vmadpm.h"

namespace vpm {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

struct Error : public std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};

inline void check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        throw Error(std::string(what) + ": " + cudaGetErrorString(e) + " at " + file + ":" +
                    std::to_string(line));
    }
}
#define VPM_CUDA(x) ::vpm::check_cuda((x), #x, __FILE__, __LINE__)
#define VPM_REQUIRE(cond, msg)                                                     \
    do {                                                                           \
        if (!(cond)) throw ::vpm::Error(std::string("vmadpm: ") + (msg));          \
    } while (0)

// Every kernel launch goes through this so that vpm_launch_count() is honest.
#define VPM_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                           \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
        ::vpm::g_launches.fetch_add(1, std::memory_order_relaxed);                 \
        VPM_CUDA(cudaGetLastError());                                              \
    } while (0)

// extern "C" bodies: translate exceptions into status codes.
#define VPM_API_BEGIN try {
#define VPM_API_END                                                                \
    return 0;                                                                      \
    }                                                                              \
    catch (const std::exception& e) {                                              \
        ::vpm::g_last_error = e.what();                                            \
        return 1;                                                                  \
    }                                                                              \
    catch (...) {                                                                  \
        ::vpm::g_last_error = "vmadpm: unknown error";                             \
        return 2;                                                                  \
    }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace vpm

struct vpm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    // grow-only scratch arena for sort temporaries / FFT staging / reduction partials
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // pinned host landing zone for reduction results
    double* host_scalars = nullptr;
    double* dev_scalars = nullptr;
    // cuFFT plans keyed by (n0, n1, n2, direction)
    std::map<std::tuple<int64_t, int64_t, int64_t, int>, cufftHandle> plans;

    void* workspace(size_t bytes);
    cufftHandle plan(const int64_t n[3], int dir);
};
