// CIC paint (scatter) and its tangent, for cell-sorted particles.
//
// Replaces pmesh ParticleMesh.paint / paint_jvp and the mesh half of RealField.readout_vjp
// (vmad/lib/fastpm.py:224,238,256).
//
// Formulation ("row gather"): with particles sorted by the key ((i0*n1)+i1)*n2+i2 the particles
// of one row of cells (fixed i0, i1; all i2) are contiguous, and cell_start[] delimits every
// cell.  One warp produces one OUTPUT row (o0, o1, :) 32 cells at a time.  An output cell
// receives mass from the four source rows (o0-a, o1-b), a, b in {0,1}; within a source row,
// cell z feeds output cells z (weight 1-f2) and z+1 (weight f2).  Lane l owns source cell z0+l,
// sums its particles into S0 (-> out z0+l) and S1 (-> out z0+l+1); one shuffle moves S1 to the
// right neighbour and a register carries it across 32-cell chunks.  Consequences:
//   * no atomics and no shared memory; every output cell is written exactly once, coalesced
//     (256 B per warp store) -> the zero fill of the mesh is fused away;
//   * the summation order is fixed -> run-to-run bit-identical meshes;
//   * a particle is read by the (up to) four warps whose output rows it touches; those warps
//     sit in the same CTA (consecutive o1), so the re-reads are L1/L2 hits, not HBM traffic.
// Algorithmic HBM bytes: 24 B/particle (+8 B with per-particle mass) + 8 B/cell written
// (+4 B/cell of cell_start read).
#include "cic.cuh"

namespace vpm {

constexpr int PAINT_WARPS = 8;

// MODE 0: value (mass * W); MODE 1: tangent (mass * sum_d v_d dW/dx_d [+ vmass * W])
template <int MODE, bool HAS_MASS, bool HAS_VMASS>
__global__ void __launch_bounds__(PAINT_WARPS * 32)
k_paint_rows(Geo g, const double* __restrict__ xs, const double* __restrict__ mass, double mass0,
             const double* __restrict__ vpos, const double* __restrict__ vmass,
             const int* __restrict__ cell_start, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // block -> (o0, group of PAINT_WARPS consecutive o1)
    const int groups1 = (g.n1 + PAINT_WARPS - 1) / PAINT_WARPS;
    const int o0 = blockIdx.x / groups1;
    const int o1 = (blockIdx.x - o0 * groups1) * PAINT_WARPS + warp;
    if (o1 >= g.n1) return;

    // the four source rows
    int row_base[4];
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        const int a = ab >> 1, b = ab & 1;
        int sp;
        if (g.ghost) {
            sp = o0 + 1 - a;  // source plane index counted from the lower ghost plane
        } else {
            sp = o0 - a;
            if (sp < 0) sp += g.n0;
        }
        int s1 = o1 - b;
        if (s1 < 0) s1 += g.n1;
        row_base[ab] = (sp * g.n1 + s1) * g.n2;
    }
    double* orow = out + (long long)o0 * g.stride0 + (long long)o1 * g.stride1;

    // carry into chunk 0 comes from the last cell of the row (periodic in z)
    double first_value = 0.0;  // lane 0 keeps out[0] until the wrap-around carry is known
    double carry = 0.0;
    const int nchunk = (g.n2 + 31) >> 5;
    for (int ch = 0; ch < nchunk; ++ch) {
        const int z = (ch << 5) + lane;
        double t0 = 0.0, t1 = 0.0;
        if (z < g.n2) {
#pragma unroll
            for (int ab = 0; ab < 4; ++ab) {
                const int a = ab >> 1, b = ab & 1;
                const int c = row_base[ab] + z;
                const int beg = __ldg(cell_start + c);
                const int end = __ldg(cell_start + c + 1);
                double S0 = 0.0, S1 = 0.0;
                for (int p = beg; p < end; ++p) {
                    const double x0 = __ldg(xs + 3LL * p + 0);
                    const double x1 = __ldg(xs + 3LL * p + 1);
                    const double x2 = __ldg(xs + 3LL * p + 2);
                    double u, fl, f0, f1, f2;
                    u = __dmul_rn(x0, g.s0); fl = floor(u); f0 = u - fl;
                    u = __dmul_rn(x1, g.s1); fl = floor(u); f1 = u - fl;
                    u = __dmul_rn(x2, g.s2); fl = floor(u); f2 = u - fl;
                    const double w0 = a ? f0 : 1.0 - f0;
                    const double w1 = b ? f1 : 1.0 - f1;
                    const double m = HAS_MASS ? __ldg(mass + p) : mass0;
                    if (MODE == 0) {
                        const double wxy = w0 * w1;
                        S0 += m * (wxy * (1.0 - f2));
                        S1 += m * (wxy * f2);
                    } else {
                        const double d0 = a ? g.s0 : -g.s0;
                        const double d1 = b ? g.s1 : -g.s1;
                        const double v0 = __ldg(vpos + 3LL * p + 0);
                        const double v1 = __ldg(vpos + 3LL * p + 1);
                        const double v2 = __ldg(vpos + 3LL * p + 2);
                        const double w2a = 1.0 - f2, w2b = f2;
                        // c = v0*(dw0*w1*w2) + v1*(w0*dw1*w2) + v2*(w0*w1*dw2)
                        const double gxy = v0 * (d0 * w1) + v1 * (w0 * d1);
                        const double wxy = w0 * w1;
                        double c0 = gxy * w2a + v2 * (wxy * -g.s2);
                        double c1 = gxy * w2b + v2 * (wxy * g.s2);
                        c0 *= m;
                        c1 *= m;
                        if (HAS_VMASS) {
                            const double vm = __ldg(vmass + p);
                            c0 += vm * (wxy * w2a);
                            c1 += vm * (wxy * w2b);
                        }
                        S0 += c0;
                        S1 += c1;
                    }
                }
                t0 += S0;
                t1 += S1;
            }
        }
        double left = __shfl_up_sync(0xffffffffu, t1, 1);
        if (lane == 0) left = carry;
        // the last valid lane of this chunk holds the carry for the next one
        const int last = min(31, g.n2 - 1 - (ch << 5));
        carry = __shfl_sync(0xffffffffu, t1, last);
        const double v = t0 + left;
        if (z < g.n2) {
            if (z == 0) first_value = v;  // completed after the loop
            else orow[z] = v;
        }
    }
    if (lane == 0) orow[0] = first_value + carry;  // periodic wrap: cell n2-1 feeds cell 0
}

// Baseline: thread per particle, 8 fp64 reductions into global memory (RED.E.ADD.F64).
template <bool HAS_MASS>
__global__ void k_paint_atomic(Geo g, const double* __restrict__ x, const double* __restrict__ mass,
                               double mass0, long long np, double* __restrict__ out) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    int i0, i1, i2;
    double f0, f1, f2;
    cell_axis(x[3 * p + 0], g.s0, g.n0, i0, f0);
    cell_axis(x[3 * p + 1], g.s1, g.n1, i1, f1);
    cell_axis(x[3 * p + 2], g.s2, g.n2, i2, f2);
    const double m = HAS_MASS ? mass[p] : mass0;
    const int j1 = wrap_inc(i1, g.n1), j2 = wrap_inc(i2, g.n2);
    const int la = local_plane(g, i0), lb = local_plane(g, wrap_inc(i0, g.n0));
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int l0 = a ? lb : la;
        if (l0 < 0) continue;
        const double w0 = a ? f0 : 1.0 - f0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double wxy = w0 * (b ? f1 : 1.0 - f1);
            double* row = out + (long long)l0 * g.stride0 + (long long)(b ? j1 : i1) * g.stride1;
            atomicAdd(row + i2, m * (wxy * (1.0 - f2)));
            atomicAdd(row + j2, m * (wxy * f2));
        }
    }
}

__global__ void k_zero_mesh(Geo g, double* __restrict__ out) {
    long long total = (long long)g.nloc0 * g.n1 * g.n2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long r = e / g.n2;
        int z = (int)(e - r * g.n2);
        int o0 = (int)(r / g.n1);
        int o1 = (int)(r - (long long)o0 * g.n1);
        out[(long long)o0 * g.stride0 + (long long)o1 * g.stride1 + z] = 0.0;
    }
}

}  // namespace vpm

using namespace vpm;

template <int MODE>
static void launch_paint_rows(vpm_ctx* ctx, const Geo& g, const double* xs, const double* mass,
                              double mass0, const double* vpos, const double* vmass,
                              const int32_t* cell_start, double* out) {
    const int groups1 = (g.n1 + PAINT_WARPS - 1) / PAINT_WARPS;
    const long long blocks = (long long)g.nloc0 * groups1;
    if (blocks == 0) return;
    VPM_REQUIRE(blocks < (1LL << 31), "mesh too large for one launch");
    const unsigned grid = (unsigned)blocks;
    const int threads = PAINT_WARPS * 32;
    cudaStream_t st = ctx->stream;
    if (MODE == 0) {
        if (mass) VPM_LAUNCH((k_paint_rows<0, true, false>), grid, threads, 0, st, g, xs, mass, mass0, vpos, vmass, cell_start, out);
        else VPM_LAUNCH((k_paint_rows<0, false, false>), grid, threads, 0, st, g, xs, mass, mass0, vpos, vmass, cell_start, out);
    } else {
        if (mass && vmass) VPM_LAUNCH((k_paint_rows<1, true, true>), grid, threads, 0, st, g, xs, mass, mass0, vpos, vmass, cell_start, out);
        else if (mass) VPM_LAUNCH((k_paint_rows<1, true, false>), grid, threads, 0, st, g, xs, mass, mass0, vpos, vmass, cell_start, out);
        else if (vmass) VPM_LAUNCH((k_paint_rows<1, false, true>), grid, threads, 0, st, g, xs, mass, mass0, vpos, vmass, cell_start, out);
        else VPM_LAUNCH((k_paint_rows<1, false, false>), grid, threads, 0, st, g, xs, mass, mass0, vpos, vmass, cell_start, out);
    }
}

extern "C" {

int vpm_paint_sorted(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs, const double* mass,
                     double mass0, int64_t np, const int32_t* cell_start, double* out) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && out && cell_start, "NULL argument");
    VPM_REQUIRE(np == 0 || xs, "xs is NULL");
    Geo g = make_geo(mesh);
    launch_paint_rows<0>(ctx, g, xs, mass, mass0, nullptr, nullptr, cell_start, out);
    VPM_API_END
}

int vpm_paint_jvp_sorted(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs, const double* mass,
                         double mass0, const double* vpos, const double* vmass, int64_t np,
                         const int32_t* cell_start, double* out) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && out && cell_start, "NULL argument");
    VPM_REQUIRE(np == 0 || (xs && vpos), "xs / vpos is NULL");
    Geo g = make_geo(mesh);
    launch_paint_rows<1>(ctx, g, xs, mass, mass0, vpos, vmass, cell_start, out);
    VPM_API_END
}

int vpm_paint_atomic(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x, const double* mass,
                     double mass0, int64_t np, double* out) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && out, "NULL argument");
    Geo g = make_geo(mesh);
    long long total = (long long)g.nloc0 * g.n1 * g.n2;
    if (total > 0) {
        unsigned zgrid = (unsigned)std::min<long long>(ceil_div(total, 256), 8LL * ctx->sm_count);
        VPM_LAUNCH(k_zero_mesh, zgrid, 256, 0, ctx->stream, g, out);
    }
    if (np > 0) {
        VPM_REQUIRE(x, "x is NULL");
        unsigned grid = (unsigned)ceil_div(np, 256);
        if (mass) VPM_LAUNCH(k_paint_atomic<true>, grid, 256, 0, ctx->stream, g, x, mass, mass0, (long long)np, out);
        else VPM_LAUNCH(k_paint_atomic<false>, grid, 256, 0, ctx->stream, g, x, mass, mass0, (long long)np, out);
    }
    VPM_API_END
}

}  // extern "C"
