// Real-space form of the library's finite-difference gradient filter.
//
// fourier_space_neg_gradient(d, pm, order=1) (vmad/lib/fastpm.py:316-323) multiplies by
// -i (8 sin w - sin 2w) / (6 h), w = k_d h: that is exactly the Fourier symbol of MINUS the
// 4-point central difference  (D f)(x) = [8 (f(x+h) - f(x-h)) - (f(x+2h) - f(x-2h))] / (12 h).
// So the three force meshes F_d = c2r(neg_gradient_d * p) of `gravity` (:430-432) equal
// -D_d phi with phi = c2r(p): ONE inverse FFT plus a streaming stencil pass instead of THREE
// inverse FFTs, and in the reverse sweep one forward FFT instead of three (D is antisymmetric,
// so the adjoint of F = -D phi is phi_bar = sum_d D_d F_bar_d).  Rounding differs from the
// k-space form by O(eps / (k h)) on the largest scales (< 1e-13 at 512^3), far inside the
// parity bars; tests compare both forms.
// HBM traffic: 8 B read + 24 B written per cell (neighbours come from L1/L2).
#include "common.cuh"

namespace vpm {

struct Grid3 {
    int n0, n1, n2;
    double c0, c1, c2;  // 1 / (12 h_d)
};

__device__ __forceinline__ int wrapi(int i, int n) {
    if (i < 0) i += n;
    if (i >= n) i -= n;
    return i;
}

// F_d = -D_d phi
__global__ void __launch_bounds__(256)
k_fd4_grad3(Grid3 g, const double* __restrict__ phi, double* __restrict__ F0,
            double* __restrict__ F1, double* __restrict__ F2) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int z = (int)(e % g.n2);
    const long long r = e / g.n2;
    const int y = (int)(r % g.n1);
    const int x = (int)(r / g.n1);
    const long long s1 = g.n2, s0 = (long long)g.n1 * g.n2;
    const double* row = phi + (long long)x * s0 + (long long)y * s1;
    {
        const double a = __ldg(row + wrapi(z + 1, g.n2)) - __ldg(row + wrapi(z - 1, g.n2));
        const double b = __ldg(row + wrapi(z + 2, g.n2)) - __ldg(row + wrapi(z - 2, g.n2));
        F2[e] = -(8.0 * a - b) * g.c2;
    }
    {
        const double* col = phi + (long long)x * s0 + z;
        const double a = __ldg(col + wrapi(y + 1, g.n1) * s1) - __ldg(col + wrapi(y - 1, g.n1) * s1);
        const double b = __ldg(col + wrapi(y + 2, g.n1) * s1) - __ldg(col + wrapi(y - 2, g.n1) * s1);
        F1[e] = -(8.0 * a - b) * g.c1;
    }
    {
        const double* pil = phi + (long long)y * s1 + z;
        const double a = __ldg(pil + wrapi(x + 1, g.n0) * s0) - __ldg(pil + wrapi(x - 1, g.n0) * s0);
        const double b = __ldg(pil + wrapi(x + 2, g.n0) * s0) - __ldg(pil + wrapi(x - 2, g.n0) * s0);
        F0[e] = -(8.0 * a - b) * g.c0;
    }
}

// phi_bar = sum_d D_d Fbar_d   (adjoint of k_fd4_grad3)
__global__ void __launch_bounds__(256)
k_fd4_adjoint(Grid3 g, const double* __restrict__ F0, const double* __restrict__ F1,
              const double* __restrict__ F2, double* __restrict__ phibar) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int z = (int)(e % g.n2);
    const long long r = e / g.n2;
    const int y = (int)(r % g.n1);
    const int x = (int)(r / g.n1);
    const long long s1 = g.n2, s0 = (long long)g.n1 * g.n2;
    double acc;
    {
        const double* row = F2 + (long long)x * s0 + (long long)y * s1;
        const double a = __ldg(row + wrapi(z + 1, g.n2)) - __ldg(row + wrapi(z - 1, g.n2));
        const double b = __ldg(row + wrapi(z + 2, g.n2)) - __ldg(row + wrapi(z - 2, g.n2));
        acc = (8.0 * a - b) * g.c2;
    }
    {
        const double* col = F1 + (long long)x * s0 + z;
        const double a = __ldg(col + wrapi(y + 1, g.n1) * s1) - __ldg(col + wrapi(y - 1, g.n1) * s1);
        const double b = __ldg(col + wrapi(y + 2, g.n1) * s1) - __ldg(col + wrapi(y - 2, g.n1) * s1);
        acc += (8.0 * a - b) * g.c1;
    }
    {
        const double* pil = F0 + (long long)y * s1 + z;
        const double a = __ldg(pil + wrapi(x + 1, g.n0) * s0) - __ldg(pil + wrapi(x - 1, g.n0) * s0);
        const double b = __ldg(pil + wrapi(x + 2, g.n0) * s0) - __ldg(pil + wrapi(x - 2, g.n0) * s0);
        acc += (8.0 * a - b) * g.c0;
    }
    phibar[e] = acc;
}

static Grid3 make_grid(const int64_t n[3], const double h[3]) {
    VPM_REQUIRE(n && h, "NULL argument");
    Grid3 g;
    for (int d = 0; d < 3; ++d) VPM_REQUIRE(n[d] >= 1 && n[d] < (1 << 20) && h[d] > 0, "bad grid");
    g.n0 = (int)n[0]; g.n1 = (int)n[1]; g.n2 = (int)n[2];
    g.c0 = 1.0 / (12.0 * h[0]); g.c1 = 1.0 / (12.0 * h[1]); g.c2 = 1.0 / (12.0 * h[2]);
    return g;
}

}  // namespace vpm

using namespace vpm;

extern "C" {

int vpm_fd4_neg_gradient(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* phi,
                         double* f0, double* f1, double* f2) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && phi && f0 && f1 && f2, "NULL argument");
    Grid3 g = make_grid(n, h);
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    VPM_LAUNCH(k_fd4_grad3, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, phi, f0, f1, f2);
    VPM_API_END
}

int vpm_fd4_neg_gradient_adjoint(vpm_ctx* ctx, const int64_t n[3], const double h[3],
                                 const double* f0, const double* f1, const double* f2,
                                 double* phibar) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && phibar && f0 && f1 && f2, "NULL argument");
    Grid3 g = make_grid(n, h);
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    VPM_LAUNCH(k_fd4_adjoint, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, f0, f1, f2, phibar);
    VPM_API_END
}

}  // extern "C"
