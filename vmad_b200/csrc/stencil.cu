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

// ---- the same stencils marching along axis 0 (2.5-D blocking) --------------------------------
// The four-cells-per-thread kernels above still read every value nine times through L2 (own row,
// four y neighbours, four x neighbours): 96 B of L2 traffic per cell, 12.9 GB per 512^3 launch in
// 1.2 ms = the L2 bandwidth of the chip (~11 TB/s), at 0.55 of the HBM roofline.  Here a CTA owns a
// (MARCH_TY rows x MARCH_TZ cells) tile and walks along axis 0: the x neighbours of a thread's two
// cells are the values it loaded two, one, minus one and minus two planes ago (a rolling window in
// registers), the y neighbours come from a double-buffered shared-memory copy of the centre plane
// (tile rows + two halo rows on each side), the z neighbours from the same rows (tile edges: from
// global memory).  L2 reads fall from 9 to ~1.6 per value; the results are bit-identical.
// Two cells per thread (one 16-byte vector), so a 8 x 128-cell tile has 512 threads; the loads of the
// NEXT plane (own cells and halo rows) are issued one iteration before they are used.
constexpr int MARCH_TY = 8, MARCH_TZ2 = 64, MARCH_TZ = 2 * MARCH_TZ2;

struct MarchTile {
    double2 row[2][MARCH_TY + 4][MARCH_TZ2];     // 16-byte elements at unit stride: conflict-free
};

// d[k] for the two cells (z, z + 1) given the chunks at z - 2 (l), z (c), z + 2 (r)
#define VPM_FD4_Z2(D, L, C, R)                       \
    {                                                \
        D[0] = 8.0 * (C.y - L.y) - (R.x - L.x);      \
        D[1] = 8.0 * (R.x - C.x) - (R.y - L.y);      \
    }
#define VPM_FD4_5(D, M2, M1, P1, P2)                 \
    {                                                \
        D[0] = 8.0 * (P1.x - M1.x) - (P2.x - M2.x);  \
        D[1] = 8.0 * (P1.y - M1.y) - (P2.y - M2.y);  \
    }

struct MarchGeo {
    int z, y0, y, yc, hy, hr, nx, xo0, xo1, zl, zr;
    bool zin, live, ledge, redge;
    int s0, s1, rowoff, haloff;      // element offsets fit 32 bits (checked by the caller)
};

__device__ __forceinline__ bool march_geo(const Grid3& g, int planes_per_seg, MarchGeo& m) {
    m.z = (blockIdx.x * MARCH_TZ2 + threadIdx.x) << 1;
    m.y0 = blockIdx.y * MARCH_TY;
    m.y = m.y0 + threadIdx.y;
    m.zin = m.z < g.n2;
    m.live = m.zin && m.y < g.n1;
    m.nx = g.n0 + 2 * g.pad;
    m.xo0 = blockIdx.z * planes_per_seg;
    m.xo1 = min(g.n0, m.xo0 + planes_per_seg);
    // rows past the edge of a partial tile hold the wrapped rows (the y neighbours of the last live
    // rows) and write nothing
    m.yc = wrapi(m.y, g.n1);
    m.s1 = g.n2;
    m.s0 = g.n1 * g.n2;
    const int zc = m.zin ? m.z : 0;
    m.rowoff = m.yc * m.s1 + zc;
    // threads of the first / last two tile rows also carry one halo row each
    const int ty = threadIdx.y;
    m.hy = -1;
    m.hr = 0;
    if (ty < 2) { m.hy = wrapi(m.y0 - 2 + ty, g.n1); m.hr = ty; }
    else if (ty >= MARCH_TY - 2) { m.hy = wrapi(m.y0 + ty + 2, g.n1); m.hr = ty + 4; }
    m.haloff = max(m.hy, 0) * m.s1 + zc;
    // tile-edge threads take their z neighbours from global memory (periodic wrap included)
    m.ledge = threadIdx.x == 0;
    m.redge = threadIdx.x == MARCH_TZ2 - 1 || m.z + 2 >= g.n2;
    m.zl = m.z == 0 ? g.n2 - 2 : m.z - 2;
    m.zr = m.z + 2 >= g.n2 ? 0 : m.z + 2;
    return m.xo0 < m.xo1;
}

__global__ void __launch_bounds__(MARCH_TY * MARCH_TZ2, 2)
k_fd4_grad3_march(Grid3 g, int planes_per_seg, const double* __restrict__ phi, double* __restrict__ F0,
                  double* __restrict__ F1, double* __restrict__ F2) {
    __shared__ MarchTile sm;
    MarchGeo m;
    if (!march_geo(g, planes_per_seg, m)) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    double2 a[5];              // planes x-2 .. x+2 of the thread's two cells
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = ld2(phi + wrapi(m.xo0 + g.pad - 2 + k, m.nx) * m.s0 + m.rowoff);
    double2 nxt = ld2(phi + wrapi(m.xo0 + g.pad + 2, m.nx) * m.s0 + m.rowoff);
    double2 halo = make_double2(0.0, 0.0);
    if (m.hy >= 0) halo = ld2(phi + (m.xo0 + g.pad) * m.s0 + m.haloff);
    int b = 0;
    for (int xo = m.xo0; xo < m.xo1; ++xo, b ^= 1) {
        const int x = xo + g.pad;
        a[4] = nxt;
        if (m.zin) {
            sm.row[b][ty + 2][tx] = a[2];
            if (m.hy >= 0) sm.row[b][m.hr][tx] = halo;
        }
        if (xo + 1 < m.xo1) {       // requests for the next plane: a whole iteration to arrive
            nxt = ld2(phi + wrapi(x + 3, m.nx) * m.s0 + m.rowoff);
            if (m.hy >= 0) halo = ld2(phi + (x + 1) * m.s0 + m.haloff);
        }
        __syncthreads();
        if (m.live) {
            const int o = xo * m.s0 + m.y * m.s1 + m.z;
            double d[2];
            {
                const double* rc = phi + (x * m.s0 + m.y * m.s1);
                const double2 l = m.ledge ? ld2(rc + m.zl) : sm.row[b][ty + 2][tx - 1];
                const double2 r = m.redge ? ld2(rc + m.zr) : sm.row[b][ty + 2][tx + 1];
                VPM_FD4_Z2(d, l, a[2], r)
                st2(F2 + o, -d[0] * g.c2, -d[1] * g.c2);
            }
            {
                const double2 m2 = sm.row[b][ty][tx], m1 = sm.row[b][ty + 1][tx];
                const double2 p1 = sm.row[b][ty + 3][tx], p2 = sm.row[b][ty + 4][tx];
                VPM_FD4_5(d, m2, m1, p1, p2)
                st2(F1 + o, -d[0] * g.c1, -d[1] * g.c1);
            }
            VPM_FD4_5(d, a[0], a[1], a[3], a[4])
            st2(F0 + o, -d[0] * g.c0, -d[1] * g.c0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = a[k + 1];
    }
}

__global__ void __launch_bounds__(MARCH_TY * MARCH_TZ2, 2)
k_fd4_adjoint_march(Grid3 g, int planes_per_seg, const double* __restrict__ F0, const double* __restrict__ F1,
                    const double* __restrict__ F2, double* __restrict__ phibar) {
    __shared__ MarchTile sm;
    MarchGeo m;
    if (!march_geo(g, planes_per_seg, m)) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    double2 a[5];              // F0 at planes x-2 .. x+2
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = ld2(F0 + wrapi(m.xo0 + g.pad - 2 + k, m.nx) * m.s0 + m.rowoff);
    double2 nxt = ld2(F0 + wrapi(m.xo0 + g.pad + 2, m.nx) * m.s0 + m.rowoff);
    // F1 (y stencil: shared-memory tile) and F2 (z stencil: the thread's own row) of the centre plane
    double2 f1 = ld2(F1 + (m.xo0 + g.pad) * m.s0 + m.rowoff);
    double2 halo = make_double2(0.0, 0.0);
    if (m.hy >= 0) halo = ld2(F1 + (m.xo0 + g.pad) * m.s0 + m.haloff);
    const int zoff = m.yc * m.s1;
    const int zc = m.zin ? m.z : 0;
    double2 f2l = ld2(F2 + (m.xo0 + g.pad) * m.s0 + zoff + (m.zin ? m.zl : 0));
    double2 f2c = ld2(F2 + (m.xo0 + g.pad) * m.s0 + zoff + zc);
    double2 f2r = ld2(F2 + (m.xo0 + g.pad) * m.s0 + zoff + (m.zin ? m.zr : 0));
    int b = 0;
    for (int xo = m.xo0; xo < m.xo1; ++xo, b ^= 1) {
        const int x = xo + g.pad;
        a[4] = nxt;
        if (m.zin) {
            sm.row[b][ty + 2][tx] = f1;
            if (m.hy >= 0) sm.row[b][m.hr][tx] = halo;
        }
        double d[2], acc[2];
        VPM_FD4_Z2(d, f2l, f2c, f2r)
        acc[0] = d[0] * g.c2;
        acc[1] = d[1] * g.c2;
        if (xo + 1 < m.xo1) {       // requests for the next plane: a whole iteration to arrive
            nxt = ld2(F0 + wrapi(x + 3, m.nx) * m.s0 + m.rowoff);
            const int pl = (x + 1) * m.s0;
            f1 = ld2(F1 + pl + m.rowoff);
            if (m.hy >= 0) halo = ld2(F1 + pl + m.haloff);
            f2l = ld2(F2 + pl + zoff + (m.zin ? m.zl : 0));
            f2c = ld2(F2 + pl + zoff + zc);
            f2r = ld2(F2 + pl + zoff + (m.zin ? m.zr : 0));
        }
        __syncthreads();
        if (m.live) {
            {
                const double2 m2 = sm.row[b][ty][tx], m1 = sm.row[b][ty + 1][tx];
                const double2 p1 = sm.row[b][ty + 3][tx], p2 = sm.row[b][ty + 4][tx];
                VPM_FD4_5(d, m2, m1, p1, p2)
                acc[0] += d[0] * g.c1;
                acc[1] += d[1] * g.c1;
            }
            VPM_FD4_5(d, a[0], a[1], a[3], a[4])
            acc[0] += d[0] * g.c0;
            acc[1] += d[1] * g.c0;
            st2(phibar + (xo * m.s0 + m.y * m.s1 + m.z), acc[0], acc[1]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = a[k + 1];
    }
}

static int g_fd4_algo = 2;      // 0 scalar, 1 four cells per thread, 2 marching gradient + four-cell adjoint (default: the
                                // measured best of each, 0.85 vs 1.20 ms and 0.86 vs 1.02 ms per 512^3 pass), 3 marching for both

// single-wrap index arithmetic: tile rows reach MARCH_TY + 1 past the last row, planes 2 past the ends
static bool fd4_march_ok(const Grid3& g) { return g.n1 >= MARCH_TY + 4 && g.n0 + 2 * g.pad >= 4; }

static void fd4_march_dims(const vpm_ctx* ctx, const Grid3& g, dim3& grid, dim3& block, int& pps) {
    block = dim3(MARCH_TZ2, MARCH_TY, 1);
    const int gx = (int)ceil_div(g.n2, MARCH_TZ), gy = (int)ceil_div(g.n1, MARCH_TY);
    // enough CTAs for ~6 per SM (3 waves of 2), but segments of at least 16 planes (4 planes of run-in each)
    int nseg = (int)std::min<long long>(std::max<long long>(1, ceil_div(6LL * ctx->sm_count, (long long)gx * gy)),
                                        std::max(1, g.n0 / 16));
    pps = (int)ceil_div(g.n0, nseg);
    nseg = (int)ceil_div(g.n0, pps);
    grid = dim3(gx, gy, nseg);
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
    if (g_fd4_algo >= 2 && fd4_v4_ok(g, {phi, f0, f1, f2}) && fd4_march_ok(g)) {
        dim3 grid, block;
        int pps;
        fd4_march_dims(ctx, g, grid, block, pps);
        VPM_LAUNCH(k_fd4_grad3_march, grid, block, 0, ctx->stream, g, pps, phi, f0, f1, f2);
    } else if (g_fd4_algo >= 1 && fd4_v4_ok(g, {phi, f0, f1, f2})) {
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
    if (g_fd4_algo >= 3 && fd4_v4_ok(g, {f0, f1, f2, phibar}) && fd4_march_ok(g)) {
        dim3 grid, block;
        int pps;
        fd4_march_dims(ctx, g, grid, block, pps);
        VPM_LAUNCH(k_fd4_adjoint_march, grid, block, 0, ctx->stream, g, pps, f0, f1, f2, phibar);
    } else if (g_fd4_algo >= 1 && fd4_v4_ok(g, {f0, f1, f2, phibar})) {
        dim3 grid, block;
        fd4_v4_dims(g, grid, block);
        VPM_LAUNCH(k_fd4_adjoint_v4, grid, block, 0, ctx->stream, g, f0, f1, f2, phibar);
    } else {
        VPM_LAUNCH(k_fd4_adjoint, (unsigned)ceil_div(total, 256), 256, 0, ctx->stream, g, f0, f1, f2, phibar);
    }
}

extern "C" {

int vpm_fd4_set_algorithm(int algo) {
    if (algo < 0 || algo > 3) return 1;
    g_fd4_algo = algo;
    return 0;
}

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
