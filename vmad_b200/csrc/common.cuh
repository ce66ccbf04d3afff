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

// ---- in-place profiling (vpm_profile_start / _stop / _read) -------------------------------
// While profiling is on, every launch that goes through VPM_LAUNCH (and every cuFFT exec /
// memset wrapped in VPM_PROF_CALL) is bracketed by two CUDA events recorded on the launching
// stream; an API entry point may name the group the launches belong to and state the
// ALGORITHMIC bytes of its next launch (or of the whole group), which is what bench.py's
// roofline divides by the measured time.  Off by default: one predictable branch per launch.
extern bool g_prof_on;
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
void prof_next_bytes(double bytes);        // algorithmic bytes of the next profiled launch
struct ProfGroup {                         // RAII: launches inside belong to `group`
    const char* prev;
    explicit ProfGroup(const char* group, double group_bytes = 0.0);
    ~ProfGroup();
};
#define VPM_PROF_BYTES(b)                                                          \
    do {                                                                           \
        if (::vpm::g_prof_on) ::vpm::prof_next_bytes((double)(b));                 \
    } while (0)
#define VPM_PROF_GROUP(name, bytes) \
    ::vpm::ProfGroup vpm_prof_group_(name, ::vpm::g_prof_on ? (double)(bytes) : 0.0)
#define VPM_PROF_CALL(name, stream, call)                                          \
    do {                                                                           \
        if (::vpm::g_prof_on) ::vpm::prof_begin(name, stream);                     \
        call;                                                                      \
        if (::vpm::g_prof_on) ::vpm::prof_end(stream);                             \
    } while (0)

// Every kernel launch goes through this so that vpm_launch_count() is honest.
#define VPM_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                           \
        if (::vpm::g_prof_on) ::vpm::prof_begin(#kernel, stream);                  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
        if (::vpm::g_prof_on) ::vpm::prof_end(stream);                             \
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
// local cells of a (slab of a) real mesh: what "8 B/cell" in the algorithmic byte counts refers to
inline double mesh_cells(const vpm_mesh* m) { return m ? (double)m->nloc0 * (double)m->n[1] * (double)m->n[2] : 0.0; }

}  // namespace vpm

struct vpm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pushed_from = nullptr;   // vpm_ctx_push_stream / _pop_stream
    bool pushed = false;
    int sm_count = 148;
    // grow-only scratch arena for sort temporaries / FFT staging / reduction partials
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // pinned host landing zone for reduction results
    double* host_scalars = nullptr;
    double* dev_scalars = nullptr;
    // row written by the take kernels where an index is VPM_ROW_PADDING (vpm_ctx_set_fill_row)
    double fill_row[3] = {0.0, 0.0, 0.0};
    // cuFFT plans keyed by (n0, n1, n2, direction)
    std::map<std::tuple<int64_t, int64_t, int64_t, int>, cufftHandle> plans;

    void* workspace(size_t bytes);
    cufftHandle plan(const int64_t n[3], int dir);
};
