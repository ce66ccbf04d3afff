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
#include <algorithm>
#include <initializer_list>

namespace vpm {

struct Grid3 {
    int n0, n1, n2;     // cells written (n0 = local planes)
    double c0, c1, c2;  // 1 / (12 h_d)
    int pad;            // inputs carry `pad` halo planes on each side of axis 0 (0: periodic wrap)
};

__device__ __forceinline__ int wrapi(int i, int n) {
    if (i < 0) i += n;
    if (i >= n) i -= n;
    return i;
}

// F_d = -D_d phi.  phi has n0 + 2 pad planes; with pad > 0 the axis-0 neighbours are halo planes
// delivered by the neighbouring slabs, with pad = 0 the axis wraps periodically.
__global__ void __launch_bounds__(256)
k_fd4_grad3(Grid3 g, const double* __restrict__ phi, double* __restrict__ F0,
            double* __restrict__ F1, double* __restrict__ F2) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int z = (int)(e % g.n2);
    const long long r = e / g.n2;
    const int y = (int)(r % g.n1);
    const int x = (int)(r / g.n1) + g.pad;
    const int nx = g.n0 + 2 * g.pad;
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
        const double a = __ldg(pil + wrapi(x + 1, nx) * s0) - __ldg(pil + wrapi(x - 1, nx) * s0);
        const double b = __ldg(pil + wrapi(x + 2, nx) * s0) - __ldg(pil + wrapi(x - 2, nx) * s0);
        F0[e] = -(8.0 * a - b) * g.c0;
    }
}

// phi_bar = sum_d D_d Fbar_d   (adjoint of k_fd4_grad3).  The three inputs have n0 + 2 pad planes;
// only Fbar_0 needs valid halo planes (the axis-0 difference).
__global__ void __launch_bounds__(256)
k_fd4_adjoint(Grid3 g, const double* __restrict__ F0, const double* __restrict__ F1,
              const double* __restrict__ F2, double* __restrict__ phibar) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int z = (int)(e % g.n2);
    const long long r = e / g.n2;
    const int y = (int)(r % g.n1);
    const int x = (int)(r / g.n1) + g.pad;
    const int nx = g.n0 + 2 * g.pad;
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
        const double a = __ldg(pil + wrapi(x + 1, nx) * s0) - __ldg(pil + wrapi(x - 1, nx) * s0);
        const double b = __ldg(pil + wrapi(x + 2, nx) * s0) - __ldg(pil + wrapi(x - 2, nx) * s0);
        acc += (8.0 * a - b) * g.c0;
    }
    phibar[e] = acc;
}

// ---- the same two stencils, four cells per thread --------------------------------------------
// The kernels above spend ~300 instructions per cell on 64-bit index arithmetic (two divisions to
// find the cell, a wrap and a multiply per neighbour) and issue twelve 8-byte loads: they are bound
// by instruction issue (ncu: 3.0 IPC, 78 % issue slots, 0.45 of the HBM roofline).  Here a CTA
// owns rows (y block, one plane): row pointers are computed once per thread, a thread produces four
// consecutive cells along the contiguous axis from 16-byte loads (8 values of its own row for the
// z difference, 4 x 2 chunks of the y and x neighbour rows) and writes 16-byte vectors.  Same
// operations in the same order per cell, so the results are bit-identical to the scalar kernels.
// Needs n2 % 4 == 0 and 16-byte aligned arrays; otherwise the scalar kernels run.
__device__ __forceinline__ double2 ld2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// d[k] = (8 (v[k+3] - v[k+1]) - (v[k+4] - v[k])) for k = 0..3 given v = values at z-2 .. z+5
#define VPM_FD4_Z(D, L, C0, C1, R)                                     \
    {                                                                  \
        D[0] = 8.0 * (C0.y - L.y) - (C1.x - L.x);                      \
        D[1] = 8.0 * (C1.x - C0.x) - (C1.y - L.y);                     \
        D[2] = 8.0 * (C1.y - C0.y) - (R.x - C0.x);                     \
        D[3] = 8.0 * (R.x - C1.x) - (R.y - C0.y);                      \
    }
// d[k] = 8 (p1[k] - m1[k]) - (p2[k] - m2[k]) for the four cells at z .. z+3 of four neighbour rows
#define VPM_FD4_ROWS(D, M2, M1, P1, P2, Z)                                                     \
    {                                                                                          \
        const double2 m2a = ld2(M2 + Z), m2b = ld2(M2 + Z + 2), m1a = ld2(M1 + Z), m1b = ld2(M1 + Z + 2); \
        const double2 p1a = ld2(P1 + Z), p1b = ld2(P1 + Z + 2), p2a = ld2(P2 + Z), p2b = ld2(P2 + Z + 2); \
        D[0] = 8.0 * (p1a.x - m1a.x) - (p2a.x - m2a.x);                                        \
        D[1] = 8.0 * (p1a.y - m1a.y) - (p2a.y - m2a.y);                                        \
        D[2] = 8.0 * (p1b.x - m1b.x) - (p2b.x - m2b.x);                                        \
        D[3] = 8.0 * (p1b.y - m1b.y) - (p2b.y - m2b.y);                                        \
    }

__global__ void __launch_bounds__(256)
k_fd4_grad3_v4(Grid3 g, const double* __restrict__ phi, double* __restrict__ F0,
               double* __restrict__ F1, double* __restrict__ F2) {
    const int y = blockIdx.x * blockDim.y + threadIdx.y;
    if (y >= g.n1) return;
    const int xo = blockIdx.y, x = xo + g.pad, nx = g.n0 + 2 * g.pad;
    // element offsets as 32-bit integers (the arrays have < 2^31 elements, checked by the caller)
    const int s1 = g.n2, s0 = g.n1 * g.n2;
    const int oc = x * s0 + y * s1;
    const double* rc = phi + oc;
    const int oym2 = x * s0 + wrapi(y - 2, g.n1) * s1, oym1 = x * s0 + wrapi(y - 1, g.n1) * s1;
    const int oyp1 = x * s0 + wrapi(y + 1, g.n1) * s1, oyp2 = x * s0 + wrapi(y + 2, g.n1) * s1;
    const int oxm2 = wrapi(x - 2, nx) * s0 + y * s1, oxm1 = wrapi(x - 1, nx) * s0 + y * s1;
    const int oxp1 = wrapi(x + 1, nx) * s0 + y * s1, oxp2 = wrapi(x + 2, nx) * s0 + y * s1;
    const double *ym2 = phi + oym2, *ym1 = phi + oym1, *yp1 = phi + oyp1, *yp2 = phi + oyp2;
    const double *xm2 = phi + oxm2, *xm1 = phi + oxm1, *xp1 = phi + oxp1, *xp2 = phi + oxp2;
    const int orow = xo * s0 + y * s1;
    const int nq = g.n2 >> 2;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        const int z = q << 2;
        double d[4];
        {
            const double2 c0 = ld2(rc + z), c1 = ld2(rc + z + 2);
            const double2 l = ld2(rc + (q == 0 ? g.n2 - 2 : z - 2)), r = ld2(rc + (q == nq - 1 ? 0 : z + 4));
            VPM_FD4_Z(d, l, c0, c1, r)
            st2(F2 + orow + z, -d[0] * g.c2, -d[1] * g.c2);
            st2(F2 + orow + z + 2, -d[2] * g.c2, -d[3] * g.c2);
        }
        VPM_FD4_ROWS(d, ym2, ym1, yp1, yp2, z)
        st2(F1 + orow + z, -d[0] * g.c1, -d[1] * g.c1);
        st2(F1 + orow + z + 2, -d[2] * g.c1, -d[3] * g.c1);
        VPM_FD4_ROWS(d, xm2, xm1, xp1, xp2, z)
        st2(F0 + orow + z, -d[0] * g.c0, -d[1] * g.c0);
        st2(F0 + orow + z + 2, -d[2] * g.c0, -d[3] * g.c0);
    }
}

__global__ void __launch_bounds__(256, 4)
k_fd4_adjoint_v4(Grid3 g, const double* __restrict__ F0, const double* __restrict__ F1,
                 const double* __restrict__ F2, double* __restrict__ phibar) {
    const int y = blockIdx.x * blockDim.y + threadIdx.y;
    if (y >= g.n1) return;
    const int xo = blockIdx.y, x = xo + g.pad, nx = g.n0 + 2 * g.pad;
    const int s1 = g.n2, s0 = g.n1 * g.n2;
    const double* rc = F2 + (x * s0 + y * s1);
    const int oym2 = x * s0 + wrapi(y - 2, g.n1) * s1, oym1 = x * s0 + wrapi(y - 1, g.n1) * s1;
    const int oyp1 = x * s0 + wrapi(y + 1, g.n1) * s1, oyp2 = x * s0 + wrapi(y + 2, g.n1) * s1;
    const int oxm2 = wrapi(x - 2, nx) * s0 + y * s1, oxm1 = wrapi(x - 1, nx) * s0 + y * s1;
    const int oxp1 = wrapi(x + 1, nx) * s0 + y * s1, oxp2 = wrapi(x + 2, nx) * s0 + y * s1;
    const double *ym2 = F1 + oym2, *ym1 = F1 + oym1, *yp1 = F1 + oyp1, *yp2 = F1 + oyp2;
    const double *xm2 = F0 + oxm2, *xm1 = F0 + oxm1, *xp1 = F0 + oxp1, *xp2 = F0 + oxp2;
    const int orow = xo * s0 + y * s1;
    const int nq = g.n2 >> 2;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        const int z = q << 2;
        double d[4], acc[4];
        {
            const double2 c0 = ld2(rc + z), c1 = ld2(rc + z + 2);
            const double2 l = ld2(rc + (q == 0 ? g.n2 - 2 : z - 2)), r = ld2(rc + (q == nq - 1 ? 0 : z + 4));
            VPM_FD4_Z(d, l, c0, c1, r)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = d[k] * g.c2;
        }
        VPM_FD4_ROWS(d, ym2, ym1, yp1, yp2, z)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += d[k] * g.c1;
        VPM_FD4_ROWS(d, xm2, xm1, xp1, xp2, z)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += d[k] * g.c0;
        st2(phibar + orow + z, acc[0], acc[1]);
        st2(phibar + orow + z + 2, acc[2], acc[3]);
    }
}

static bool fd4_v4_ok(const Grid3& g, std::initializer_list<const void*> arrays) {
    bool ok = g.n2 % 4 == 0 && g.n2 >= 8 && g.n0 <= 65535 &&
              (long long)(g.n0 + 2 * g.pad) * g.n1 * g.n2 < (1LL << 31);
    for (const void* a : arrays) ok = ok && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
    return ok;
}

static void fd4_v4_dims(const Grid3& g, dim3& grid, dim3& block) {
    const int tz = std::min(g.n2 / 4, 256);
    const int ty = std::max(1, 256 / tz);
    block = dim3(tz, ty, 1);
    grid = dim3((unsigned)ceil_div(g.n1, ty), (unsigned)g.n0, 1);
}

// ---- second-order LPT source in real space ------------------------------------------------
// lpt2src (fastpm.py:353-387) needs the six second derivatives P_ij = c2r(neg_grad_i neg_grad_j
// potk) = D_i D_j phi: six inverse FFTs and thirteen transfers in the reference graph.  With
// F_j = -D_j phi already on the mesh (k_fd4_grad3), P_ij = -D_i F_j is one more stencil, and
//   source = 3/7 (P11 P22 + P22 P00 + P00 P11 - P12^2 - P20^2 - P01^2) = 1/2 B(F, F)
// with the symmetric bilinear form B(F, G) = dsource(F)[G]: ONE kernel serves the value
// (G = F, scale 1/2) and the tangent (G = F_, scale 1).
struct Cell3 {
    int x, y, z;
};

// cell e of the n0 x n1 x n2 cells written; x counts planes of the (padded) INPUT arrays
__device__ __forceinline__ Cell3 cell_of(const Grid3& g, long long e) {
    Cell3 c;
    c.z = (int)(e % g.n2);
    const long long r = e / g.n2;
    c.y = (int)(r % g.n1);
    c.x = (int)(r / g.n1) + g.pad;
    return c;
}

// (D_axis f)(cell)
template <int AXIS>
__device__ __forceinline__ double dax(const Grid3& g, const double* __restrict__ f, const Cell3& c) {
    const long long s1 = g.n2, s0 = (long long)g.n1 * g.n2;
    if (AXIS == 2) {
        const double* row = f + (long long)c.x * s0 + (long long)c.y * s1;
        const double a = __ldg(row + wrapi(c.z + 1, g.n2)) - __ldg(row + wrapi(c.z - 1, g.n2));
        const double b = __ldg(row + wrapi(c.z + 2, g.n2)) - __ldg(row + wrapi(c.z - 2, g.n2));
        return (8.0 * a - b) * g.c2;
    } else if (AXIS == 1) {
        const double* col = f + (long long)c.x * s0 + c.z;
        const double a = __ldg(col + wrapi(c.y + 1, g.n1) * s1) - __ldg(col + wrapi(c.y - 1, g.n1) * s1);
        const double b = __ldg(col + wrapi(c.y + 2, g.n1) * s1) - __ldg(col + wrapi(c.y - 2, g.n1) * s1);
        return (8.0 * a - b) * g.c1;
    } else {
        // with pad > 0 the axis-0 neighbours are halo planes (no wrap happens), else periodic
        const int nx = g.n0 + 2 * g.pad;
        const double* pil = f + (long long)c.y * s1 + c.z;
        const double a = __ldg(pil + wrapi(c.x + 1, nx) * s0) - __ldg(pil + wrapi(c.x - 1, nx) * s0);
        const double b = __ldg(pil + wrapi(c.x + 2, nx) * s0) - __ldg(pil + wrapi(c.x - 2, nx) * s0);
        return (8.0 * a - b) * g.c0;
    }
}

struct Hess {
    double p00, p11, p22, p12, p20, p01;
};

// P_ij = -D_i F_j
__device__ __forceinline__ Hess hessian(const Grid3& g, const double* __restrict__ F0,
                                        const double* __restrict__ F1,
                                        const double* __restrict__ F2, const Cell3& c) {
    Hess h;
    h.p00 = -dax<0>(g, F0, c);
    h.p11 = -dax<1>(g, F1, c);
    h.p22 = -dax<2>(g, F2, c);
    h.p12 = -dax<1>(g, F2, c);
    h.p20 = -dax<2>(g, F0, c);
    h.p01 = -dax<0>(g, F1, c);
    return h;
}

// out = scale * B(F, G);  SAME: G is F (the value, scale 1/2)
template <bool SAME>
__global__ void __launch_bounds__(256)
k_lpt2_bilinear(Grid3 g, const double* __restrict__ F0, const double* __restrict__ F1,
                const double* __restrict__ F2, const double* __restrict__ G0,
                const double* __restrict__ G1, const double* __restrict__ G2, double scale,
                double* __restrict__ out) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const Cell3 c = cell_of(g, e);
    const Hess p = hessian(g, F0, F1, F2, c);
    const Hess q = SAME ? p : hessian(g, G0, G1, G2, c);
    const double diag = p.p11 * q.p22 + q.p11 * p.p22 + p.p22 * q.p00 + q.p22 * p.p00 +
                        p.p00 * q.p11 + q.p00 * p.p11;
    const double off = p.p12 * q.p12 + p.p20 * q.p20 + p.p01 * q.p01;
    out[e] = scale * ((3.0 / 7.0) * diag - (6.0 / 7.0) * off);
}

// reverse sweep, first half: W_ij = dsource/dP_ij * sbar  (six meshes, W = [00, 11, 22, 12, 20, 01])
__global__ void __launch_bounds__(256)
k_lpt2_adjoint_w(Grid3 g, const double* __restrict__ F0, const double* __restrict__ F1,
                 const double* __restrict__ F2, const double* __restrict__ sbar,
                 double* __restrict__ W) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const Cell3 c = cell_of(g, e);
    const Hess p = hessian(g, F0, F1, F2, c);
    const double s = sbar[e];
    // the six meshes carry the same halo planes as the inputs (filled by the caller across ranks)
    const long long plane = (long long)g.n1 * g.n2;
    const long long comp = (long long)(g.n0 + 2 * g.pad) * plane;
    const long long o = e + g.pad * plane;
    W[o] = (3.0 / 7.0) * (p.p11 + p.p22) * s;
    W[comp + o] = (3.0 / 7.0) * (p.p22 + p.p00) * s;
    W[2 * comp + o] = (3.0 / 7.0) * (p.p00 + p.p11) * s;
    W[3 * comp + o] = -(6.0 / 7.0) * p.p12 * s;
    W[4 * comp + o] = -(6.0 / 7.0) * p.p20 * s;
    W[5 * comp + o] = -(6.0 / 7.0) * p.p01 * s;
}

// second half: P_ij = -D_i F_j  =>  Fbar_j = sum_i D_i W_ij
//   Fbar_0 = D_0 W00 + D_2 W20,  Fbar_1 = D_1 W11 + D_0 W01,  Fbar_2 = D_2 W22 + D_1 W12
__global__ void __launch_bounds__(256)
k_lpt2_adjoint_f(Grid3 g, const double* __restrict__ W, double* __restrict__ Fb0,
                 double* __restrict__ Fb1, double* __restrict__ Fb2) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= total) return;
    const Cell3 c = cell_of(g, e);
    const long long plane = (long long)g.n1 * g.n2;
    const long long comp = (long long)(g.n0 + 2 * g.pad) * plane;
    const long long o = e + g.pad * plane;      // outputs padded like the inputs
    Fb0[o] = dax<0>(g, W, c) + dax<2>(g, W + 4 * comp, c);
    Fb1[o] = dax<1>(g, W + comp, c) + dax<0>(g, W + 5 * comp, c);
    Fb2[o] = dax<2>(g, W + 2 * comp, c) + dax<1>(g, W + 3 * comp, c);
}

static Grid3 make_grid(const int64_t n[3], const double h[3]) {
    VPM_REQUIRE(n && h, "NULL argument");
    Grid3 g;
    for (int d = 0; d < 3; ++d) VPM_REQUIRE(n[d] >= 1 && n[d] < (1 << 20) && h[d] > 0, "bad grid");
    g.n0 = (int)n[0]; g.n1 = (int)n[1]; g.n2 = (int)n[2];
    g.c0 = 1.0 / (12.0 * h[0]); g.c1 = 1.0 / (12.0 * h[1]); g.c2 = 1.0 / (12.0 * h[2]);
    g.pad = 0;
    return g;
}

}  // namespace vpm

using namespace vpm;

static void launch_fd4_grad3(vpm_ctx* ctx, const Grid3& g, const double* phi, double* f0, double* f1, double* f2) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    if (total <= 0) return;
    if (fd4_v4_ok(g, {phi, f0, f1, f2})) {
        dim3 grid, block;
        fd4_v4_dims(g, grid, block);
        VPM_LAUNCH(k_fd4_grad3_v4, grid, block, 0, ctx->stream, g, phi, f0, f1, f2);
    } else {
        VPM_LAUNCH(k_fd4_grad3, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, phi, f0, f1, f2);
    }
}

static void launch_fd4_adjoint(vpm_ctx* ctx, const Grid3& g, const double* f0, const double* f1, const double* f2,
                               double* phibar) {
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    if (total <= 0) return;
    if (fd4_v4_ok(g, {f0, f1, f2, phibar})) {
        dim3 grid, block;
        fd4_v4_dims(g, grid, block);
        VPM_LAUNCH(k_fd4_adjoint_v4, grid, block, 0, ctx->stream, g, f0, f1, f2, phibar);
    } else {
        VPM_LAUNCH(k_fd4_adjoint, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, f0, f1, f2, phibar);
    }
}

extern "C" {

int vpm_fd4_neg_gradient(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* phi,
                         double* f0, double* f1, double* f2) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("fd4_neg_gradient", 32.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && phi && f0 && f1 && f2, "NULL argument");
    Grid3 g = make_grid(n, h);
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    launch_fd4_grad3(ctx, g, phi, f0, f1, f2);
    VPM_API_END
}

int vpm_fd4_neg_gradient_adjoint(vpm_ctx* ctx, const int64_t n[3], const double h[3],
                                 const double* f0, const double* f1, const double* f2,
                                 double* phibar) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("fd4_neg_gradient_adjoint", 32.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && phibar && f0 && f1 && f2, "NULL argument");
    Grid3 g = make_grid(n, h);
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    launch_fd4_adjoint(ctx, g, f0, f1, f2, phibar);
    VPM_API_END
}

int vpm_fd4_neg_gradient_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad,
                              const double* phi_padded, double* f0, double* f1, double* f2) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("fd4_neg_gradient", 32.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && phi_padded && f0 && f1 && f2, "NULL argument");
    VPM_REQUIRE(pad == 0 || pad >= 2, "the stencil needs two halo planes per side");
    Grid3 g = make_grid(n, h);
    g.pad = pad;
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    if (total > 0) launch_fd4_grad3(ctx, g, phi_padded, f0, f1, f2);
    VPM_API_END
}

int vpm_fd4_neg_gradient_adjoint_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad,
                                      const double* f0_padded, const double* f1_padded,
                                      const double* f2_padded, double* phibar) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("fd4_neg_gradient_adjoint", 32.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && phibar && f0_padded && f1_padded && f2_padded, "NULL argument");
    VPM_REQUIRE(pad == 0 || pad >= 2, "the stencil needs two halo planes per side");
    Grid3 g = make_grid(n, h);
    g.pad = pad;
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    if (total > 0) launch_fd4_adjoint(ctx, g, f0_padded, f1_padded, f2_padded, phibar);
    VPM_API_END
}

int vpm_lpt2_source(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* f0,
                    const double* f1, const double* f2, const double* g0, const double* g1,
                    const double* g2, double scale, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("lpt2_source", 56.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && f0 && f1 && f2 && out, "NULL argument");
    VPM_REQUIRE((g0 && g1 && g2) || (!g0 && !g1 && !g2), "pass all of g0, g1, g2 or none");
    Grid3 g = make_grid(n, h);
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const unsigned grid = (unsigned)ceil_div(total, 256);
    if (g0)
        VPM_LAUNCH(k_lpt2_bilinear<false>, grid, 256, 0, ctx->stream, g, f0, f1, f2, g0, g1, g2, scale, out);
    else
        VPM_LAUNCH(k_lpt2_bilinear<true>, grid, 256, 0, ctx->stream, g, f0, f1, f2, f0, f1, f2, scale, out);
    VPM_API_END
}

int vpm_lpt2_source_adjoint(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* f0,
                            const double* f1, const double* f2, const double* sbar, double* work6,
                            double* fb0, double* fb1, double* fb2) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("lpt2_source_adjoint", 104.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && f0 && f1 && f2 && sbar && work6 && fb0 && fb1 && fb2, "NULL argument");
    Grid3 g = make_grid(n, h);
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const unsigned grid = (unsigned)ceil_div(total, 256);
    VPM_LAUNCH(k_lpt2_adjoint_w, grid, 256, 0, ctx->stream, g, f0, f1, f2, sbar, work6);
    VPM_LAUNCH(k_lpt2_adjoint_f, grid, 256, 0, ctx->stream, g, work6, fb0, fb1, fb2);
    VPM_API_END
}

}  // extern "C"

extern "C" {

// slab forms of the 2LPT source and its adjoint: every mesh argument except `out` / `sbar` has
// `pad` (>= 2) halo planes on each side of axis 0, filled by the caller from the neighbouring
// slabs; n = the cells written (the local slab)
int vpm_lpt2_source_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad, const double* f0,
                         const double* f1, const double* f2, const double* g0, const double* g1,
                         const double* g2, double scale, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("lpt2_source", 56.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && f0 && f1 && f2 && out, "NULL argument");
    VPM_REQUIRE((g0 && g1 && g2) || (!g0 && !g1 && !g2), "pass all of g0, g1, g2 or none");
    VPM_REQUIRE(pad >= 2, "the stencil needs two halo planes per side");
    Grid3 g = make_grid(n, h);
    g.pad = pad;
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    const unsigned grid = (unsigned)ceil_div(total, 256);
    if (total > 0) {
        if (g0)
            VPM_LAUNCH(k_lpt2_bilinear<false>, grid, 256, 0, ctx->stream, g, f0, f1, f2, g0, g1, g2, scale, out);
        else
            VPM_LAUNCH(k_lpt2_bilinear<true>, grid, 256, 0, ctx->stream, g, f0, f1, f2, f0, f1, f2, scale, out);
    }
    VPM_API_END
}

int vpm_lpt2_adjoint_w_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad, const double* f0,
                            const double* f1, const double* f2, const double* sbar, double* w6_padded) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("lpt2_source_adjoint", 80.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && f0 && f1 && f2 && sbar && w6_padded, "NULL argument");
    VPM_REQUIRE(pad >= 2, "the stencil needs two halo planes per side");
    Grid3 g = make_grid(n, h);
    g.pad = pad;
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    if (total > 0)
        VPM_LAUNCH(k_lpt2_adjoint_w, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, f0, f1, f2, sbar, w6_padded);
    VPM_API_END
}

int vpm_lpt2_adjoint_f_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad,
                            const double* w6_padded, double* fb0, double* fb1, double* fb2) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("lpt2_source_adjoint", 72.0 * ((double)n[0] * n[1] * n[2]));
    VPM_REQUIRE(ctx && w6_padded && fb0 && fb1 && fb2, "NULL argument");
    VPM_REQUIRE(pad >= 2, "the stencil needs two halo planes per side");
    Grid3 g = make_grid(n, h);
    g.pad = pad;
    const long long total = (long long)g.n0 * g.n1 * g.n2;
    if (total > 0)
        VPM_LAUNCH(k_lpt2_adjoint_f, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, w6_padded, fb0, fb1, fb2);
    VPM_API_END
}

}  // extern "C"
