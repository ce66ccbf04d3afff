// Streaming elementwise kernels and deterministic reductions on flat fp64 arrays.
//
// These serve the stdlib arithmetic that vmad applies to fields and particle arrays
// (vmad/core/stdlib/operators.py:14-152: add, sub, mul, div, neg), the kick / drift updates
// of vmad/lib/fastpm.py:459-471 (p + f * c fused into one pass), linalg.stack / numpy.take
// (vmad/lib/linalg.py:145-149) and the scalar objectives (linalg.sum, to_scalar).
#include "common.cuh"

namespace vpm {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;
constexpr int RED_THREADS = 256;

#define EW_LOOP(n, ...)                                                                    \
    long long base__ = blockIdx.x * (long long)(EW_THREADS * EW_UNROLL) + threadIdx.x;     \
    _Pragma("unroll") for (int j__ = 0; j__ < EW_UNROLL; ++j__) {                          \
        long long e = base__ + (long long)j__ * EW_THREADS;                                \
        if (e < (n)) { __VA_ARGS__; }                                                          \
    }

template <bool HAS_Y>
__global__ void __launch_bounds__(EW_THREADS)
k_axpby(long long n, double a, const double* __restrict__ x, double b,
        const double* __restrict__ y, double c, double* __restrict__ out) {
    EW_LOOP(n, {
        double v = a * x[e];
        if (HAS_Y) v += b * y[e];
        out[e] = v + c;
    })
}

// drift fused with the position update of the next force evaluation:
//   dx1 = dx + fac * p (one FMA, as k_axpby computes it),  x1 = q + dx1
__global__ void __launch_bounds__(EW_THREADS)
k_drift_pos(long long n, const double* __restrict__ dx, const double* __restrict__ p, double fac,
            const double* __restrict__ q, double* __restrict__ dx1, double* __restrict__ x1) {
    EW_LOOP(n, {
        const double d = fma(fac, p[e], dx[e]);
        dx1[e] = d;
        x1[e] = q[e] + d;
    })
}

template <int OP>  // 0: mul, 1: div
__global__ void __launch_bounds__(EW_THREADS)
k_binary(long long n, const double* __restrict__ x, const double* __restrict__ y,
         double* __restrict__ out) {
    EW_LOOP(n, { out[e] = OP == 0 ? x[e] * y[e] : x[e] / y[e]; })
}

__global__ void __launch_bounds__(EW_THREADS)
k_mul_rows(long long n, int ncomp, const double* __restrict__ x, const double* __restrict__ w,
           double* __restrict__ out) {
    long long total = n * ncomp;
    EW_LOOP(total, { out[e] = x[e] * __ldg(w + e / ncomp); })
}

__global__ void __launch_bounds__(EW_THREADS)
k_cmul(long long n, const double2* __restrict__ x, const double2* __restrict__ y, bool conj,
       double2* __restrict__ out) {
    EW_LOOP(n, {
        double2 a = x[e];
        double2 b = y[e];
        if (conj) b.y = -b.y;
        out[e] = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    })
}

__global__ void __launch_bounds__(EW_THREADS)
k_fill(long long n, double v, double* __restrict__ out) {
    EW_LOOP(n, { out[e] = v; })
}

struct Cols {
    const double* c[8];
};

__global__ void __launch_bounds__(EW_THREADS)
k_stack(long long n, int ncomp, Cols cols, double* __restrict__ out) {
    long long total = n * ncomp;
    EW_LOOP(total, {
        long long i = e / ncomp;
        int c = (int)(e - i * ncomp);
        out[e] = __ldg(cols.c[c] + i);
    })
}

__global__ void __launch_bounds__(EW_THREADS)
k_take_col(long long n, int ncomp, int col, const double* __restrict__ x,
           double* __restrict__ out) {
    EW_LOOP(n, { out[e] = x[e * ncomp + col]; })
}

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum1(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum1(double a) {
    __shared__ double sa[RED_THREADS / 32];
    a = warp_sum1(a);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sa[w] = a;
    __syncthreads();
    if (w == 0) {
        a = lane < RED_THREADS / 32 ? sa[lane] : 0.0;
        a = warp_sum1(a);
    }
    __syncthreads();
    return a;
}

template <bool DOT>
__global__ void __launch_bounds__(RED_THREADS)
k_reduce_partial(long long n, const double* __restrict__ x, const double* __restrict__ y,
                 double* __restrict__ partial) {
    double acc = 0.0;
    for (long long e = blockIdx.x * (long long)RED_THREADS + threadIdx.x; e < n;
         e += (long long)gridDim.x * RED_THREADS)
        acc += DOT ? x[e] * y[e] : x[e];
    acc = block_sum1(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(RED_THREADS)
k_reduce_final(int nparts, const double* __restrict__ partial, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += RED_THREADS) acc += partial[i];
    acc = block_sum1(acc);
    if (threadIdx.x == 0) out[0] = acc;
}

static unsigned ew_grid(long long n) { return (unsigned)ceil_div(n, EW_THREADS * EW_UNROLL); }

static void reduce_to_host(vpm_ctx* ctx, long long n, const double* x, const double* y,
                           double* out_host) {
    if (n <= 0) {
        *out_host = 0.0;
        return;
    }
    int nparts = (int)std::min<long long>(ceil_div(n, RED_THREADS), 4LL * ctx->sm_count);
    double* partial = (double*)ctx->workspace((size_t)nparts * sizeof(double));
    if (y) VPM_LAUNCH(k_reduce_partial<true>, nparts, RED_THREADS, 0, ctx->stream, n, x, y, partial);
    else VPM_LAUNCH(k_reduce_partial<false>, nparts, RED_THREADS, 0, ctx->stream, n, x, y, partial);
    VPM_LAUNCH(k_reduce_final, 1, RED_THREADS, 0, ctx->stream, nparts, partial, ctx->dev_scalars);
    VPM_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_host = ctx->host_scalars[0];
}

}  // namespace vpm

using namespace vpm;

extern "C" {

int vpm_axpby(vpm_ctx* ctx, int64_t n, double a, const double* x, double b, const double* y,
              double c, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("axpby", (16.0 + (y ? 8.0 : 0.0)) * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    if (n > 0) {
        VPM_REQUIRE(x && out, "NULL array");
        if (y) VPM_LAUNCH(k_axpby<true>, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, a, x, b, y, c, out);
        else VPM_LAUNCH(k_axpby<false>, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, a, x, b, y, c, out);
    }
    VPM_API_END
}

int vpm_drift_pos(vpm_ctx* ctx, int64_t n, const double* dx, const double* p, double fac,
                  const double* q, double* dx1, double* x1) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("drift_pos", 40.0 * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    if (n > 0) {
        VPM_REQUIRE(dx && p && q && dx1 && x1, "NULL array");
        VPM_LAUNCH(k_drift_pos, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, dx, p, fac, q, dx1, x1);
    }
    VPM_API_END
}

int vpm_mul(vpm_ctx* ctx, int64_t n, const double* x, const double* y, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("mul", 24.0 * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    if (n > 0) {
        VPM_REQUIRE(x && y && out, "NULL array");
        VPM_LAUNCH(k_binary<0>, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, x, y, out);
    }
    VPM_API_END
}

int vpm_div(vpm_ctx* ctx, int64_t n, const double* x, const double* y, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("div", 24.0 * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    if (n > 0) {
        VPM_REQUIRE(x && y && out, "NULL array");
        VPM_LAUNCH(k_binary<1>, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, x, y, out);
    }
    VPM_API_END
}

int vpm_mul_rows(vpm_ctx* ctx, int64_t n, int ncomp, const double* x, const double* w,
                 double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("mul_rows", (16.0 * ncomp + 8.0) * n);
    VPM_REQUIRE(ctx && ncomp >= 1, "bad argument");
    if (n > 0) {
        VPM_REQUIRE(x && w && out, "NULL array");
        VPM_LAUNCH(k_mul_rows, ew_grid(n * ncomp), EW_THREADS, 0, ctx->stream, (long long)n, ncomp, x, w, out);
    }
    VPM_API_END
}

int vpm_cmul(vpm_ctx* ctx, int64_t n, const double* x, const double* y, int conj_y, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("cmul", 48.0 * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    if (n > 0) {
        VPM_REQUIRE(x && y && out, "NULL array");
        VPM_LAUNCH(k_cmul, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n,
                   reinterpret_cast<const double2*>(x), reinterpret_cast<const double2*>(y),
                   conj_y != 0, reinterpret_cast<double2*>(out));
    }
    VPM_API_END
}

int vpm_fill(vpm_ctx* ctx, int64_t n, double value, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("fill", 8.0 * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    if (n > 0) {
        VPM_REQUIRE(out, "NULL array");
        VPM_LAUNCH(k_fill, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, value, out);
    }
    VPM_API_END
}

int vpm_stack(vpm_ctx* ctx, int64_t n, int ncomp, const double* const* cols_host, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("stack", 16.0 * n * ncomp);
    VPM_REQUIRE(ctx && cols_host, "NULL argument");
    VPM_REQUIRE(ncomp >= 1 && ncomp <= 8, "stack supports 1..8 columns");
    if (n > 0) {
        VPM_REQUIRE(out, "NULL array");
        Cols cols;
        for (int c = 0; c < 8; ++c) cols.c[c] = c < ncomp ? cols_host[c] : nullptr;
        for (int c = 0; c < ncomp; ++c) VPM_REQUIRE(cols.c[c], "NULL column");
        VPM_LAUNCH(k_stack, ew_grid(n * ncomp), EW_THREADS, 0, ctx->stream, (long long)n, ncomp, cols, out);
    }
    VPM_API_END
}

int vpm_take_col(vpm_ctx* ctx, int64_t n, int ncomp, int col, const double* x, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("take_col", 16.0 * n);
    VPM_REQUIRE(ctx && ncomp >= 1 && col >= 0 && col < ncomp, "bad argument");
    if (n > 0) {
        VPM_REQUIRE(x && out, "NULL array");
        VPM_LAUNCH(k_take_col, ew_grid(n), EW_THREADS, 0, ctx->stream, (long long)n, ncomp, col, x, out);
    }
    VPM_API_END
}

int vpm_reduce_sum(vpm_ctx* ctx, int64_t n, const double* x, double* out_host) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("reduce_sum", 8.0 * n);
    VPM_REQUIRE(ctx && out_host && (n <= 0 || x), "NULL argument");
    reduce_to_host(ctx, n, x, nullptr, out_host);
    VPM_API_END
}

int vpm_reduce_dot(vpm_ctx* ctx, int64_t n, const double* x, const double* y, double* out_host) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("reduce_dot", 16.0 * n);
    VPM_REQUIRE(ctx && out_host && (n <= 0 || (x && y)), "NULL argument");
    reduce_to_host(ctx, n, x, y, out_host);
    VPM_API_END
}

}  // extern "C"
