// r2c / c2r: cuFFT (library) for the transform itself, normalisation as a fused scale pass.
//
// Replaces pmesh RealField.r2c / c2r_vjp and ComplexField.c2r / r2c_vjp
// (vmad/lib/fastpm.py:50,53,56,64,67,70).  pmesh normalisation: r2c = FFT / Ntot,
// c2r = un-normalised inverse; the two adjoints are the same transforms with the scales
// swapped (see oracle/pmesh/pm.py::RealField.c2r_vjp for the cotangent convention).
#include "common.cuh"

namespace vpm {

static const char* cufft_str(cufftResult r) {
    switch (r) {
        case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
        case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
        case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
        case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
        case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
        case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
        case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
        case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
        default: return "CUFFT_ERROR";
    }
}
#define VPM_CUFFT(x)                                                                   \
    do {                                                                               \
        cufftResult r__ = (x);                                                         \
        if (r__ != CUFFT_SUCCESS)                                                      \
            throw ::vpm::Error(std::string(#x) + ": " + ::vpm::cufft_str(r__));        \
    } while (0)

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

__global__ void __launch_bounds__(EW_THREADS)
k_scale_inplace(long long n, double a, double* __restrict__ x) {
    long long base = blockIdx.x * (long long)(EW_THREADS * EW_UNROLL) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        long long e = base + (long long)j * EW_THREADS;
        if (e < n) x[e] *= a;
    }
}

}  // namespace vpm

using namespace vpm;

cufftHandle vpm_ctx::plan(const int64_t n[3], int dir) {
    auto key = std::make_tuple(n[0], n[1], n[2], dir);
    auto it = plans.find(key);
    if (it != plans.end()) return it->second;
    for (int d = 0; d < 3; ++d) VPM_REQUIRE(n[d] >= 1 && n[d] < (1LL << 30), "bad FFT size");
    cufftHandle h;
    cufftType type = dir == 0 ? CUFFT_D2Z : CUFFT_Z2D;
    if (n[0] == 1) {
        VPM_CUFFT(cufftPlan2d(&h, (int)n[1], (int)n[2], type));
    } else {
        VPM_CUFFT(cufftPlan3d(&h, (int)n[0], (int)n[1], (int)n[2], type));
    }
    VPM_CUFFT(cufftSetStream(h, stream));
    plans[key] = h;
    return h;
}

extern "C" {

int vpm_r2c(vpm_ctx* ctx, const int64_t n[3], const double* real_in, double* cplx_out,
            double scale) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && n && real_in && cplx_out, "NULL argument");
    cufftHandle h = ctx->plan(n, 0);
    VPM_CUFFT(cufftExecD2Z(h, const_cast<double*>(real_in),
                           reinterpret_cast<cufftDoubleComplex*>(cplx_out)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = 2LL * n[0] * n[1] * (n[2] / 2 + 1);
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, cplx_out);
    }
    VPM_API_END
}

int vpm_c2r(vpm_ctx* ctx, const int64_t n[3], const double* cplx_in, double* real_out,
            double scale) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && n && cplx_in && real_out, "NULL argument");
    cufftHandle h = ctx->plan(n, 1);
    // multi-dimensional Z2D may overwrite its input; operators must not mutate theirs
    size_t bytes = (size_t)16 * n[0] * n[1] * (n[2] / 2 + 1);
    void* stage = ctx->workspace(bytes);
    VPM_CUDA(cudaMemcpyAsync(stage, cplx_in, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    VPM_CUFFT(cufftExecZ2D(h, reinterpret_cast<cufftDoubleComplex*>(stage), real_out));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = (long long)n[0] * n[1] * n[2];
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, real_out);
    }
    VPM_API_END
}

}  // extern "C"
