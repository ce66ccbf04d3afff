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

// dst[b][a][:] = src[a][b][:] for an [A, B, C] array of complex128 (C contiguous):
// pack / unpack around the all-to-all transpose of the slab FFT
__global__ void __launch_bounds__(256)
k_swap_leading_axes(long long A, long long B, long long Cn, const double2* __restrict__ src,
                    double2* __restrict__ dst) {
    const long long total = A * B * Cn;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long c = e % Cn;
        const long long r = e / Cn;   // output row index = b * A + a
        const long long a = r % A;
        const long long b = r / A;
        dst[e] = src[(a * B + b) * Cn + c];
    }
}

}  // namespace vpm

using namespace vpm;

// batched plans for the slab-decomposed transform; kind 10: 2-D D2Z over planes, 11: 2-D Z2D
// over planes, 12: 1-D Z2Z along the leading (strided) axis
static cufftHandle slab_plan(vpm_ctx* ctx, int kind, int64_t a, int64_t b, int64_t c) {
    auto key = std::make_tuple(a, b, c, kind);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) return it->second;
    cufftHandle h;
    if (kind == 10 || kind == 11) {
        // a planes of a (b, c) real mesh
        int n[2] = {(int)b, (int)c};
        VPM_CUFFT(cufftPlanMany(&h, 2, n, nullptr, 1, 0, nullptr, 1, 0,
                                kind == 10 ? CUFFT_D2Z : CUFFT_Z2D, (int)a));
    } else {
        // length-a transforms with stride b (= number of transforms), one element apart
        int n[1] = {(int)a};
        int embed[1] = {(int)a};
        VPM_CUFFT(cufftPlanMany(&h, 1, n, embed, (int)b, 1, embed, (int)b, 1, CUFFT_Z2Z, (int)b));
    }
    VPM_CUFFT(cufftSetStream(h, ctx->stream));
    ctx->plans[key] = h;
    return h;
}

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
    VPM_PROF_GROUP("r2c", 8.0 * ((double)n[0] * n[1] * n[2]) + 16.0 * ((double)n[0] * n[1] * (n[2] / 2 + 1)));
    VPM_REQUIRE(ctx && n && real_in && cplx_out, "NULL argument");
    cufftHandle h = ctx->plan(n, 0);
    VPM_PROF_CALL("cufftExecD2Z", ctx->stream,
                  VPM_CUFFT(cufftExecD2Z(h, const_cast<double*>(real_in),
                                         reinterpret_cast<cufftDoubleComplex*>(cplx_out))));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = 2LL * n[0] * n[1] * (n[2] / 2 + 1);
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, cplx_out);
    }
    VPM_API_END
}

static int c2r_impl(vpm_ctx* ctx, const int64_t n[3], const double* cplx_in, double* real_out,
                    double scale, bool preserve_input) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("c2r", 8.0 * ((double)n[0] * n[1] * n[2]) + 16.0 * ((double)n[0] * n[1] * (n[2] / 2 + 1)));
    VPM_REQUIRE(ctx && n && cplx_in && real_out, "NULL argument");
    cufftHandle h = ctx->plan(n, 1);
    // multi-dimensional Z2D may overwrite its input; operators must not mutate theirs
    void* stage = const_cast<double*>(cplx_in);
    if (preserve_input) {
        size_t bytes = (size_t)16 * n[0] * n[1] * (n[2] / 2 + 1);
        stage = ctx->workspace(bytes);
        VPM_PROF_CALL("memcpy(c2r input)", ctx->stream,
                      VPM_CUDA(cudaMemcpyAsync(stage, cplx_in, bytes, cudaMemcpyDeviceToDevice, ctx->stream)));
    }
    VPM_PROF_CALL("cufftExecZ2D", ctx->stream,
                  VPM_CUFFT(cufftExecZ2D(h, reinterpret_cast<cufftDoubleComplex*>(stage), real_out)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = (long long)n[0] * n[1] * n[2];
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, real_out);
    }
    VPM_API_END
}

int vpm_c2r(vpm_ctx* ctx, const int64_t n[3], const double* cplx_in, double* real_out,
            double scale) {
    return c2r_impl(ctx, n, cplx_in, real_out, scale, true);
}

int vpm_c2r_inplace(vpm_ctx* ctx, const int64_t n[3], double* cplx_scratch, double* real_out,
                    double scale) {
    return c2r_impl(ctx, n, cplx_scratch, real_out, scale, false);
}

int vpm_fft_r2c_planes(vpm_ctx* ctx, int64_t nplanes, int64_t n1, int64_t n2, const double* real_in,
                       double* cplx_out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("r2c_planes", (double)nplanes * n1 * (8.0 * n2 + 16.0 * (n2 / 2 + 1)));
    VPM_REQUIRE(ctx && real_in && cplx_out && nplanes >= 1, "bad argument");
    cufftHandle h = slab_plan(ctx, 10, nplanes, n1, n2);
    VPM_PROF_CALL("cufftExecD2Z", ctx->stream,
                  VPM_CUFFT(cufftExecD2Z(h, const_cast<double*>(real_in),
                                         reinterpret_cast<cufftDoubleComplex*>(cplx_out))));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    VPM_API_END
}

int vpm_fft_c2r_planes(vpm_ctx* ctx, int64_t nplanes, int64_t n1, int64_t n2, double* cplx_in,
                       double* real_out, double scale) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("c2r_planes", (double)nplanes * n1 * (8.0 * n2 + 16.0 * (n2 / 2 + 1)));
    VPM_REQUIRE(ctx && real_out && cplx_in && nplanes >= 1, "bad argument");
    cufftHandle h = slab_plan(ctx, 11, nplanes, n1, n2);
    VPM_PROF_CALL("cufftExecZ2D", ctx->stream,
                  VPM_CUFFT(cufftExecZ2D(h, reinterpret_cast<cufftDoubleComplex*>(cplx_in), real_out)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = (long long)nplanes * n1 * n2;
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, real_out);
    }
    VPM_API_END
}

int vpm_fft_c2c_axis0(vpm_ctx* ctx, int64_t n0, int64_t inner, double* data, int inverse,
                      double scale) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("c2c_axis0", 32.0 * n0 * inner);
    VPM_REQUIRE(ctx && data && n0 >= 1 && inner >= 1, "bad argument");
    VPM_REQUIRE(inner < (1LL << 31), "too many transforms for one plan");
    cufftHandle h = slab_plan(ctx, 12, n0, inner, 0);
    cufftDoubleComplex* d = reinterpret_cast<cufftDoubleComplex*>(data);
    VPM_PROF_CALL("cufftExecZ2Z", ctx->stream,
                  VPM_CUFFT(cufftExecZ2Z(h, d, d, inverse ? CUFFT_INVERSE : CUFFT_FORWARD)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = 2LL * n0 * inner;
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, data);
    }
    VPM_API_END
}

int vpm_fft_c2c_axis0_oop(vpm_ctx* ctx, int64_t n0, int64_t inner, const double* in, double* out,
                          int inverse, double scale) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("c2c_axis0", 32.0 * n0 * inner);
    VPM_REQUIRE(ctx && in && out && n0 >= 1 && inner >= 1, "bad argument");
    VPM_REQUIRE(inner < (1LL << 31), "too many transforms for one plan");
    cufftHandle h = slab_plan(ctx, 12, n0, inner, 0);
    VPM_PROF_CALL("cufftExecZ2Z", ctx->stream,
                  VPM_CUFFT(cufftExecZ2Z(h, reinterpret_cast<cufftDoubleComplex*>(const_cast<double*>(in)),
                                         reinterpret_cast<cufftDoubleComplex*>(out),
                                         inverse ? CUFFT_INVERSE : CUFFT_FORWARD)));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (scale != 1.0) {
        long long nd = 2LL * n0 * inner;
        VPM_LAUNCH(k_scale_inplace, (unsigned)ceil_div(nd, EW_THREADS * EW_UNROLL), EW_THREADS, 0,
                   ctx->stream, nd, scale, out);
    }
    VPM_API_END
}

int vpm_swap_leading_axes(vpm_ctx* ctx, int64_t a, int64_t b, int64_t c, const double* src,
                          double* dst) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("swap_leading_axes", 32.0 * a * b * c);
    VPM_REQUIRE(ctx && src && dst && src != dst, "bad argument");
    long long total = (long long)a * b * c;
    if (total > 0) {
        unsigned grid = (unsigned)std::min<long long>(ceil_div(total, 256), 32LL * ctx->sm_count);
        VPM_LAUNCH(k_swap_leading_axes, grid, 256, 0, ctx->stream, (long long)a, (long long)b,
                   (long long)c, reinterpret_cast<const double2*>(src),
                   reinterpret_cast<double2*>(dst));
    }
    VPM_API_END
}

}  // extern "C"
