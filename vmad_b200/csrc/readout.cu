// CIC readout (gather), its gradient and its tangent.
//
// Replaces pmesh RealField.readout / readout_vjp / readout_jvp and the particle half of
// ParticleMesh.paint_vjp (vmad/lib/fastpm.py:229,252,256,263).
//
// One thread per particle.  The eight mesh values of a particle are four pairs of cells that
// are adjacent along the contiguous axis, so each pair is fetched as one 16-byte access when
// it is aligned and does not wrap (ld.global.nc.v2.f64), else as two 8-byte loads.  Particles
// arrive cell-sorted from `exchange` (or in lattice order), so neighbouring threads touch
// neighbouring cells and the mesh is read from HBM once; the re-use is served by L1/L2.
// Algorithmic HBM bytes: 24 B/particle read + 8 B/cell read + 8 B (value) or 24 B (gradient,
// or three fused fields) per particle written.
#include "cic.cuh"

namespace vpm {

struct Stencil {
    long long off[4];  // element offsets of rows (a, b) = (0,0) (0,1) (1,0) (1,1); -1: not local
    int i2, j2;        // the two cells along the contiguous axis
    double f0, f1, f2;
};

__device__ __forceinline__ Stencil make_stencil(const Geo& g, const double* __restrict__ x,
                                                long long p) {
    Stencil s;
    int i0, i1;
    cell_axis(__ldg(x + 3 * p + 0), g.s0, g.n0, i0, s.f0);
    cell_axis(__ldg(x + 3 * p + 1), g.s1, g.n1, i1, s.f1);
    cell_axis(__ldg(x + 3 * p + 2), g.s2, g.n2, s.i2, s.f2);
    s.j2 = wrap_inc(s.i2, g.n2);
    const int j1 = wrap_inc(i1, g.n1);
    const int la = local_plane(g, i0), lb = local_plane(g, wrap_inc(i0, g.n0));
    s.off[0] = la < 0 ? -1 : (long long)la * g.stride0 + (long long)i1 * g.stride1;
    s.off[1] = la < 0 ? -1 : (long long)la * g.stride0 + (long long)j1 * g.stride1;
    s.off[2] = lb < 0 ? -1 : (long long)lb * g.stride0 + (long long)i1 * g.stride1;
    s.off[3] = lb < 0 ? -1 : (long long)lb * g.stride0 + (long long)j1 * g.stride1;
    return s;
}

// the pair (field[row + i2], field[row + j2])
__device__ __forceinline__ void load_pair(const double* __restrict__ field, long long row, int i2,
                                          int j2, double& v0, double& v1) {
    const double* p = field + row + i2;
    if (j2 == i2 + 1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        v0 = t.x;
        v1 = t.y;
    } else {
        v0 = __ldg(p);
        v1 = __ldg(field + row + j2);
    }
}

// value = sum over corners in (a, b, c) order of field * ((w0 * w1) * w2)
__device__ __forceinline__ double gather_value(const Stencil& s, const double* __restrict__ field) {
    double acc = 0.0;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        if (s.off[ab] < 0) continue;
        const double w0 = (ab >> 1) ? s.f0 : 1.0 - s.f0;
        const double w1 = (ab & 1) ? s.f1 : 1.0 - s.f1;
        const double wxy = w0 * w1;
        double v0, v1;
        load_pair(field, s.off[ab], s.i2, s.j2, v0, v1);
        acc += v0 * (wxy * (1.0 - s.f2));
        acc += v1 * (wxy * s.f2);
    }
    return acc;
}

// value and the three derivatives d/dx_d (chain factor Nmesh_d / BoxSize_d included)
__device__ __forceinline__ void gather_value_grad(const Geo& g, const Stencil& s,
                                                  const double* __restrict__ field, double& val,
                                                  double& g0, double& g1, double& g2) {
    val = g0 = g1 = g2 = 0.0;
    const double w2a = 1.0 - s.f2, w2b = s.f2;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        if (s.off[ab] < 0) continue;
        const int a = ab >> 1, b = ab & 1;
        const double w0 = a ? s.f0 : 1.0 - s.f0;
        const double w1 = b ? s.f1 : 1.0 - s.f1;
        const double d0 = a ? g.s0 : -g.s0;
        const double d1 = b ? g.s1 : -g.s1;
        double v0, v1;
        load_pair(field, s.off[ab], s.i2, s.j2, v0, v1);
        const double wxy = w0 * w1;
        val += v0 * (wxy * w2a);
        val += v1 * (wxy * w2b);
        g0 += v0 * ((d0 * w1) * w2a);
        g0 += v1 * ((d0 * w1) * w2b);
        g1 += v0 * ((w0 * d1) * w2a);
        g1 += v1 * ((w0 * d1) * w2b);
        g2 += v0 * (wxy * -g.s2);
        g2 += v1 * (wxy * g.s2);
    }
}

__global__ void __launch_bounds__(256)
k_readout(Geo g, const double* __restrict__ field, const double* __restrict__ x, long long np,
          double* __restrict__ value) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    value[p] = gather_value(s, field);
}

__global__ void __launch_bounds__(256)
k_readout3(Geo g, const double* __restrict__ fa, const double* __restrict__ fb,
           const double* __restrict__ fc, const double* __restrict__ x, long long np,
           const int* __restrict__ out_row, double* __restrict__ value3,
           const double* __restrict__ y_in, double fac, double* __restrict__ y_out) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    const double a = gather_value(s, fa);
    const double b = gather_value(s, fb);
    const double c = gather_value(s, fc);
    // out_row: the layout's slot -> home row map (Layout.gather fused into the readout)
    const long long o = out_row ? (long long)__ldg(out_row + p) : p;
    if (value3) {
        value3[3 * o + 0] = a;
        value3[3 * o + 1] = b;
        value3[3 * o + 2] = c;
    }
    if (y_out) {   // the kick that consumes the force: y_out = y_in + fac * value (one FMA each)
        y_out[3 * o + 0] = fma(fac, a, __ldg(y_in + 3 * o + 0));
        y_out[3 * o + 1] = fma(fac, b, __ldg(y_in + 3 * o + 1));
        y_out[3 * o + 2] = fma(fac, c, __ldg(y_in + 3 * o + 2));
    }
}

template <bool HAS_W, bool WANT_VALUE>
__global__ void __launch_bounds__(256)
k_readout_grad(Geo g, const double* __restrict__ field, const double* __restrict__ x,
               const double* __restrict__ weight, double weight0, long long np,
               double* __restrict__ value, double* __restrict__ grad) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    double val, g0, g1, g2;
    gather_value_grad(g, s, field, val, g0, g1, g2);
    const double w = HAS_W ? __ldg(weight + p) : weight0;
    if (WANT_VALUE) value[p] = val;
    grad[3 * p + 0] = g0 * w;
    grad[3 * p + 1] = g1 * w;
    grad[3 * p + 2] = g2 * w;
}

// grad[p, :] = sum_d w[p, d] * grad_readout(F_d, x_p) + grad_readout(R, x_p): the particle half
// of the reverse sweep of a fused force evaluation, one pass over the particles for the four
// meshes (F_0, F_1, F_2: the force components; R: the density cotangent)
__global__ void __launch_bounds__(256)
k_readout_grad4(Geo g, const double* __restrict__ F0, const double* __restrict__ F1,
                const double* __restrict__ F2, const double* __restrict__ R,
                const double* __restrict__ x, const double* __restrict__ w3, long long np,
                const int* __restrict__ out_row, const double* __restrict__ base,
                double* __restrict__ grad, const double* __restrict__ y_in, double fac,
                double* __restrict__ y_out) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    const double wa = __ldg(w3 + 3 * p + 0), wb = __ldg(w3 + 3 * p + 1), wc = __ldg(w3 + 3 * p + 2);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const double w2a = 1.0 - s.f2, w2b = s.f2;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        if (s.off[ab] < 0) continue;
        const int a = ab >> 1, b = ab & 1;
        const double w0 = a ? s.f0 : 1.0 - s.f0;
        const double w1 = b ? s.f1 : 1.0 - s.f1;
        const double d0 = a ? g.s0 : -g.s0;
        const double d1 = b ? g.s1 : -g.s1;
        double u0, u1, t0, t1;
        // combined mesh value seen by this particle: m = wa F0 + wb F1 + wc F2 + R
        load_pair(F0, s.off[ab], s.i2, s.j2, t0, t1);
        u0 = wa * t0;
        u1 = wa * t1;
        load_pair(F1, s.off[ab], s.i2, s.j2, t0, t1);
        u0 += wb * t0;
        u1 += wb * t1;
        load_pair(F2, s.off[ab], s.i2, s.j2, t0, t1);
        u0 += wc * t0;
        u1 += wc * t1;
        load_pair(R, s.off[ab], s.i2, s.j2, t0, t1);
        u0 += t0;
        u1 += t1;
        const double wxy = w0 * w1;
        a0 += u0 * ((d0 * w1) * w2a) + u1 * ((d0 * w1) * w2b);
        a1 += u0 * ((w0 * d1) * w2a) + u1 * ((w0 * d1) * w2b);
        a2 += (u1 - u0) * (wxy * g.s2);
    }
    const long long o = out_row ? (long long)__ldg(out_row + p) : p;
    if (base) {   // cotangent arriving at the same rows from the other consumers of dx
        a0 += __ldg(base + 3 * o + 0);
        a1 += __ldg(base + 3 * o + 1);
        a2 += __ldg(base + 3 * o + 2);
    }
    grad[3 * o + 0] = a0;
    grad[3 * o + 1] = a1;
    grad[3 * o + 2] = a2;
    if (y_out) {   // the drift's momentum cotangent: y_out = y_in + fac * grad
        y_out[3 * o + 0] = fma(fac, a0, __ldg(y_in + 3 * o + 0));
        y_out[3 * o + 1] = fma(fac, a1, __ldg(y_in + 3 * o + 1));
        y_out[3 * o + 2] = fma(fac, a2, __ldg(y_in + 3 * o + 2));
    }
}

template <bool HAS_FDOT, bool HAS_XDOT>
__global__ void __launch_bounds__(256)
k_readout_jvp(Geo g, const double* __restrict__ field, const double* __restrict__ field_dot,
              const double* __restrict__ x, const double* __restrict__ xdot, long long np,
              double* __restrict__ value) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    double r = 0.0;
    if (HAS_FDOT) r += gather_value(s, field_dot);
    if (HAS_XDOT) {
        double val, g0, g1, g2;
        gather_value_grad(g, s, field, val, g0, g1, g2);
        r += (__ldg(xdot + 3 * p + 0) * g0 + __ldg(xdot + 3 * p + 1) * g1) +
             __ldg(xdot + 3 * p + 2) * g2;
    }
    value[p] = r;
}

}  // namespace vpm

using namespace vpm;

extern "C" {

int vpm_readout(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field, const double* x,
                int64_t np, double* value) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout", 32.0 * np + 8.0 * mesh_cells(mesh));
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(field && x && value, "NULL array");
        VPM_LAUNCH(k_readout, (unsigned)ceil_div(np, 256), 256, 0, ctx->stream, g, field, x,
                   (long long)np, value);
    }
    VPM_API_END
}

int vpm_readout3(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                 const double* f2, const double* x, int64_t np, const int32_t* out_row,
                 double* value3, const double* y_in, double fac, double* y_out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout3", 24.0 * np + 24.0 * mesh_cells(mesh) + (out_row ? 4.0 : 0.0) * np + (value3 ? 24.0 : 0.0) * np + (y_out ? 48.0 : 0.0) * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(f0 && f1 && f2 && x && (value3 || y_out), "NULL array");
        VPM_REQUIRE(!y_out || y_in, "y_out needs y_in");
        VPM_LAUNCH(k_readout3, (unsigned)ceil_div(np, 256), 256, 0, ctx->stream, g, f0, f1, f2, x,
                   (long long)np, out_row, value3, y_in, fac, y_out);
    }
    VPM_API_END
}

int vpm_readout_grad(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field, const double* x,
                     const double* weight, double weight0, int64_t np, double* value,
                     double* grad) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout_grad", 48.0 * np + 8.0 * mesh_cells(mesh) + (weight ? 8.0 : 0.0) * np + (value ? 8.0 : 0.0) * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(field && x && grad, "NULL array");
        unsigned grid = (unsigned)ceil_div(np, 256);
        cudaStream_t st = ctx->stream;
        if (weight && value) VPM_LAUNCH((k_readout_grad<true, true>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad);
        else if (weight) VPM_LAUNCH((k_readout_grad<true, false>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad);
        else if (value) VPM_LAUNCH((k_readout_grad<false, true>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad);
        else VPM_LAUNCH((k_readout_grad<false, false>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad);
    }
    VPM_API_END
}

int vpm_readout_grad4(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                      const double* f2, const double* r, const double* x, const double* w3,
                      int64_t np, const int32_t* out_row, const double* base, double* grad,
                      const double* y_in, double fac, double* y_out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout_grad4", 72.0 * np + 32.0 * mesh_cells(mesh) + (out_row ? 4.0 : 0.0) * np + (base ? 24.0 : 0.0) * np + (y_out ? 48.0 : 0.0) * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(f0 && f1 && f2 && r && x && w3 && grad, "NULL array");
        VPM_REQUIRE(!y_out || y_in, "y_out needs y_in");
        VPM_LAUNCH(k_readout_grad4, (unsigned)ceil_div(np, 256), 256, 0, ctx->stream, g, f0, f1, f2, r,
                   x, w3, (long long)np, out_row, base, grad, y_in, fac, y_out);
    }
    VPM_API_END
}

int vpm_readout_jvp(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field,
                    const double* field_dot, const double* x, const double* xdot, int64_t np,
                    double* value) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout_jvp", 32.0 * np + (field_dot ? 8.0 * mesh_cells(mesh) : 0.0) + (xdot ? 24.0 * np + 8.0 * mesh_cells(mesh) : 0.0));
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(x && value, "NULL array");
        VPM_REQUIRE(!xdot || field, "field is required with xdot");
        unsigned grid = (unsigned)ceil_div(np, 256);
        cudaStream_t st = ctx->stream;
        if (field_dot && xdot) VPM_LAUNCH((k_readout_jvp<true, true>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value);
        else if (field_dot) VPM_LAUNCH((k_readout_jvp<true, false>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value);
        else if (xdot) VPM_LAUNCH((k_readout_jvp<false, true>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value);
        else VPM_LAUNCH((k_readout_jvp<false, false>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value);
    }
    VPM_API_END
}

}  // extern "C"
