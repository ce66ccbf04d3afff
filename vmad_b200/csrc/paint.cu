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
#include <cstdlib>

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

// -------------------------------------------------------------------------------------------
// Tile kernel (the production path).
//
// A CTA owns an output tile of TX x TY x TZ cells and accumulates it in shared memory; nothing
// but this CTA writes those cells, so the mesh needs no zero fill and no atomics, and the
// result is bit-reproducible.  The particles that touch the tile live in the (TX+1) x (TY+1)
// source rows (one halo row/plane on the low side) and, because of the cell sort, each row's
// particles for the tile's z-range are ONE contiguous run of the particle array.  Warps take
// source rows from a shared counter and walk the run 32 particles at a time with
// LANES OVER PARTICLES: loads are coalesced and independent of cell_start (no dependent-load
// chain), every lane does the same amount of fp64 work (no divergence from uneven cell
// occupancy), and particles of the same cell -- adjacent lanes, thanks to the sort -- are
// combined with a segmented warp scan whose depth adapts to the longest run in the warp.  The
// last lane of each run adds the run's eight corner sums into shared memory with plain
// read-modify-writes: for a fixed corner (a, b) a target row is fed by exactly one source row,
// i.e. one warp, and within the warp the targets z (phase 1) and z+1 (phase 2) are distinct.
// Four partial tiles W_ab are kept and summed in a fixed order when the tile is stored.
// -------------------------------------------------------------------------------------------
constexpr int TILE_WARPS = 8;

struct TileDims {
    int TX, TY, TZ;     // tile extents in cells
    int ntx, nty, ntz;  // number of tiles along each axis
};

// wrapped cell index along the contiguous axis; the common case (particle inside the box or
// one box length outside) costs a compare, the general case is the exact 64-bit modulo that
// the key kernel uses, so the two always agree.
// (wrap_cell / wrap_fast live in cic.cuh)

// MASSKIND 0: mass == 1 (no multiply); 1: scalar mass0; 2: per-particle array
template <int MODE, int MASSKIND, bool HAS_VMASS>
__global__ void __launch_bounds__(TILE_WARPS * 32, 4)
k_paint_tile(Geo g, TileDims td, const double* __restrict__ xs, const double* __restrict__ mass,
             double mass0, const double* __restrict__ vpos, const double* __restrict__ vmass,
             const int* __restrict__ cell_start, double* __restrict__ out) {
    extern __shared__ __align__(16) double W[];  // [4][TX*TY*TZ]
    __shared__ int next_item;
    const int lane = threadIdx.x & 31;
    const int TY = td.TY, TZ = td.TZ;
    const int tsize = td.TX * TY * TZ;

    // tile coordinates: y fastest, so that CTAs that share halo rows run close together
    int t = blockIdx.x;
    const int ty = t % td.nty; t /= td.nty;
    const int tz = t % td.ntz; t /= td.ntz;
    const int ox0 = t * td.TX, oy0 = ty * TY, oz0 = tz * TZ;
    const int ex = min(td.TX, g.nloc0 - ox0), ey = min(TY, g.n1 - oy0), ez = min(TZ, g.n2 - oz0);

    {
        double2* W2 = reinterpret_cast<double2*>(W);
        const double2 z2 = make_double2(0.0, 0.0);
        for (int i = threadIdx.x; i < 2 * tsize; i += blockDim.x) W2[i] = z2;
    }
    if (threadIdx.x == 0) next_item = 0;
    __syncthreads();

    const int nrows = (ex + 1) * (ey + 1);
    const int n2 = g.n2;
    const bool full_row = (ez == n2);
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&next_item, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nrows) break;
        const int rsx = item / (ey + 1) - 1;  // source row relative to the tile: [-1, ex) x [-1, ey)
        const int rsy = item - (rsx + 1) * (ey + 1) - 1;
        int sp = ox0 + rsx;  // local output plane of the source row (may be -1)
        if (g.ghost) {
            sp += 1;         // key planes are counted from the lower ghost plane
        } else if (sp < 0) {
            sp += g.n0;
        }
        int sy = oy0 + rsy;
        if (sy < 0) sy += g.n1;
        const int rowbase = (sp * g.n1 + sy) * n2;
        // which corners (a, b) of this source row land inside the tile (warp-uniform)
        const bool va0 = rsx >= 0, va1 = rsx + 1 < ex;
        const bool vb0 = rsy >= 0, vb1 = rsy + 1 < ey;
        const bool v00 = va0 && vb0, v01 = va0 && vb1, v10 = va1 && vb0, v11 = va1 && vb1;
        // word offsets of the target rows in the four partial tiles, z origin folded in
        const int o00 = 0 * tsize + ((rsx + 0) * TY + (rsy + 0)) * TZ - oz0;
        const int o01 = 1 * tsize + ((rsx + 0) * TY + (rsy + 1)) * TZ - oz0;
        const int o10 = 2 * tsize + ((rsx + 1) * TY + (rsy + 0)) * TZ - oz0;
        const int o11 = 3 * tsize + ((rsx + 1) * TY + (rsy + 1)) * TZ - oz0;
        // particle ranges: source cells [oz0 - 1, oz0 + ez - 1] (periodic), contiguous per range
        int rb0, re0, rb1 = 0, re1 = 0;
        if (full_row) {
            rb0 = __ldg(cell_start + rowbase);
            re0 = __ldg(cell_start + rowbase + n2);
        } else {
            rb0 = __ldg(cell_start + rowbase + max(oz0 - 1, 0));
            re0 = __ldg(cell_start + rowbase + oz0 + ez);
            if (oz0 == 0) {
                rb1 = __ldg(cell_start + rowbase + n2 - 1);
                re1 = __ldg(cell_start + rowbase + n2);
            }
        }
        for (int r = 0; r < 2; ++r) {
            const int rb = r ? rb1 : rb0, re = r ? re1 : re0;
            for (int p0 = rb; p0 < re; p0 += 32) {
                const int p = p0 + lane;
                const bool active = p < re;
                int iz = -1 - lane;  // inactive lanes: unique keys, never merged
                // S<ab><c>: this particle's contribution to corner (a, b, c)
                double S000 = 0.0, S001 = 0.0, S010 = 0.0, S011 = 0.0;
                double S100 = 0.0, S101 = 0.0, S110 = 0.0, S111 = 0.0;
                if (active) {
                    const double* xp = xs + 3LL * p;
                    const double x0 = __ldg(xp + 0), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
                    double u, fl;
                    u = __dmul_rn(x0, g.s0); fl = floor(u); const double f0 = u - fl;
                    u = __dmul_rn(x1, g.s1); fl = floor(u); const double f1 = u - fl;
                    u = __dmul_rn(x2, g.s2); fl = floor(u); const double f2 = u - fl;
                    iz = wrap_cell(fl, n2);
                    const double e0 = 1.0 - f0, e1 = 1.0 - f1, e2 = 1.0 - f2;
                    if (MODE == 0) {
                        // per-corner weight (w0 * w1) * w2, times the mass
                        double a0 = e0, a1 = f0;
                        if (MASSKIND != 0) {
                            const double m = MASSKIND == 2 ? __ldg(mass + p) : mass0;
                            a0 *= m;
                            a1 *= m;
                        }
                        if (v00) { const double w = a0 * e1; S000 = w * e2; S001 = w * f2; }
                        if (v01) { const double w = a0 * f1; S010 = w * e2; S011 = w * f2; }
                        if (v10) { const double w = a1 * e1; S100 = w * e2; S101 = w * f2; }
                        if (v11) { const double w = a1 * f1; S110 = w * e2; S111 = w * f2; }
                    } else {
                        const double* vp = vpos + 3LL * p;
                        const double m = MASSKIND == 2 ? __ldg(mass + p) : (MASSKIND == 1 ? mass0 : 1.0);
                        const double v0 = m * __ldg(vp + 0), v1 = m * __ldg(vp + 1), v2 = m * __ldg(vp + 2);
                        const double vm = HAS_VMASS ? __ldg(vmass + p) : 0.0;
                        // corner value: v0 dW/dx0 + v1 dW/dx1 + v2 dW/dx2 (+ vm W)
                        const double q0 = v0 * g.s0, q1 = v1 * g.s1, q2 = v2 * g.s2;
#define VPM_TANGENT(W0, D0, W1, D1, SA, SB)                                          \
    {                                                                                \
        const double wxy = (W0) * (W1);                                              \
        const double gxy = (D0) * (W1) + (W0) * (D1);                                \
        const double base = HAS_VMASS ? vm * wxy : 0.0;                              \
        SA = gxy * e2 - q2 * wxy + base * e2;                                        \
        SB = gxy * f2 + q2 * wxy + base * f2;                                        \
    }
                        if (v00) VPM_TANGENT(e0, -q0, e1, -q1, S000, S001)
                        if (v01) VPM_TANGENT(e0, -q0, f1, q1, S010, S011)
                        if (v10) VPM_TANGENT(f0, q0, e1, -q1, S100, S101)
                        if (v11) VPM_TANGENT(f0, q0, f1, q1, S110, S111)
#undef VPM_TANGENT
                    }
                }
                // segmented inclusive scan over runs of equal iz (sorted => runs are contiguous);
                // the number of rounds adapts to the longest run in this warp
                for (int d = 1; d < 32; d <<= 1) {
                    const int izd = __shfl_up_sync(0xffffffffu, iz, d);
                    const bool same = (lane >= d) && (izd == iz);
                    if (!__any_sync(0xffffffffu, same)) break;
#define VPM_SCAN(V)                                                                  \
    {                                                                                \
        const double tv = __shfl_up_sync(0xffffffffu, V, d);                         \
        if (same) V += tv;                                                           \
    }
                    if (v00) { VPM_SCAN(S000) VPM_SCAN(S001) }
                    if (v01) { VPM_SCAN(S010) VPM_SCAN(S011) }
                    if (v10) { VPM_SCAN(S100) VPM_SCAN(S101) }
                    if (v11) { VPM_SCAN(S110) VPM_SCAN(S111) }
#undef VPM_SCAN
                }
                const int iznext = __shfl_down_sync(0xffffffffu, iz, 1);
                const bool tail = active && (lane == 31 || iznext != iz);
                // phase 1: corner c = 0 -> cell iz; phase 2: c = 1 -> cell iz + 1 (periodic)
                {
                    const int zl = iz - oz0;
                    if (tail && zl >= 0 && zl < ez) {
                        if (v00) W[o00 + iz] += S000;
                        if (v01) W[o01 + iz] += S010;
                        if (v10) W[o10 + iz] += S100;
                        if (v11) W[o11 + iz] += S110;
                    }
                }
                __syncwarp();
                {
                    int zt = iz + 1;
                    if (zt == n2) zt = 0;
                    const int zl = zt - oz0;
                    if (tail && zl >= 0 && zl < ez) {
                        if (v00) W[o00 + zt] += S001;
                        if (v01) W[o01 + zt] += S011;
                        if (v10) W[o10 + zt] += S101;
                        if (v11) W[o11 + zt] += S111;
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    // store the tile: fixed summation order over the four partial tiles; one warp per row
    const int warp = threadIdx.x >> 5;
    for (int r = warp; r < ex * ey; r += TILE_WARPS) {
        const int x = r / ey;
        const int y = r - x * ey;
        const double* w = W + (x * TY + y) * TZ;
        double* o = out + (long long)(ox0 + x) * g.stride0 + (long long)(oy0 + y) * g.stride1 + oz0;
        for (int z = lane; z < ez; z += 32)
            o[z] = ((w[z] + w[tsize + z]) + w[2 * tsize + z]) + w[3 * tsize + z];
    }
}

// -------------------------------------------------------------------------------------------
// Warp-aggregated scatter (the fast path).
//
// Lanes over cell-sorted particles, no tiles: a warp takes 32 consecutive particles, sums the
// particles of each cell with the same segmented scan as the tile kernel, folds the c = 1
// (cell z + 1) sums of a run into the c = 0 sums of the next run when that run is the
// neighbouring cell of the same row, and the last lane of each run issues at most eight fp64
// reductions (RED.E.ADD.F64) into the mesh.  Every particle is read exactly once (no halo
// re-reads) and the instruction count per particle is about a third of the tile kernel's; the
// price is a zero fill of the mesh beforehand and a summation order that depends on the order
// in which L2 receives the reductions (results agree to rounding, not bit for bit).
// -------------------------------------------------------------------------------------------

// V += tv on the lanes where pred != 0 (one predicated DADD, no select)
#define VPM_ADD_IF(V, TV, PRED)                                                      \
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p add.f64 %0, %0, %1;\n\t}"   \
        : "+d"(V) : "d"(TV), "r"(PRED))

template <int MODE, int MASSKIND, bool HAS_VMASS>
__global__ void __launch_bounds__(256)
k_paint_agg(Geo g, const double* __restrict__ xs, const double* __restrict__ mass, double mass0,
            int mstride, const double* __restrict__ vpos, const double* __restrict__ vmass,
            int np, double* __restrict__ out, int cpw) {
    const int lane = threadIdx.x & 31;
    const int nchunk = (np + 31) >> 5;
    const int st0 = (int)g.stride0, st1 = (int)g.stride1;  // local slab < 2^31 elements (checked)
    // A warp takes `cpw` CONSECUTIVE chunks of 32 particles and CTAs start in index order, so the
    // kernel sweeps the cell-sorted particles as one front: a mesh plane receives its a = 1
    // contributions (particles of plane i0 - 1) and its a = 0 contributions (plane i0) within a few
    // megabytes of traffic and stays in L2 in between.  (A grid-stride loop over the whole array had
    // ~55 fronts in flight, 660 MB of live mesh planes: every plane was fetched and written back
    // twice -- 21.1 GB of DRAM traffic per 512^3 three-column launch against 12.9 GB.)
    // software pipeline: the positions of the warp's next chunk are requested before the
    // current chunk is processed (the loads are the longest-latency step of the loop)
    int ch = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * cpw;
    const int chend = min(nchunk, ch + cpw);
    double nx0 = 0.0, nx1 = 0.0, nx2 = 0.0;
    if (ch < chend && (ch << 5) + lane < np) {
        const double* xp = xs + 3LL * ((ch << 5) + lane);
        nx0 = __ldg(xp + 0); nx1 = __ldg(xp + 1); nx2 = __ldg(xp + 2);
    }
    for (; ch < chend; ++ch) {
        const int p = (ch << 5) + lane;
        const bool active = p < np;
        const double x0 = nx0, x1 = nx1, x2 = nx2;
        {
            if (ch + 1 < chend) {
                const int pn = ((ch + 1) << 5) + lane;
                if (pn < np) {
                    const double* xp = xs + 3LL * pn;
                    nx0 = __ldg(xp + 0); nx1 = __ldg(xp + 1); nx2 = __ldg(xp + 2);
                }
            }
        }
        int key = -1 - lane;  // inactive lanes: unique keys, never merged
        int i0 = 0, i1 = 0, iz = 0;
        double S000 = 0.0, S001 = 0.0, S010 = 0.0, S011 = 0.0;
        double S100 = 0.0, S101 = 0.0, S110 = 0.0, S111 = 0.0;
        if (active) {
            double u, fl;
            u = __dmul_rn(x0, g.s0); fl = floor(u); const double f0 = u - fl; i0 = wrap_fast(fl, g.n0);
            u = __dmul_rn(x1, g.s1); fl = floor(u); const double f1 = u - fl; i1 = wrap_fast(fl, g.n1);
            u = __dmul_rn(x2, g.s2); fl = floor(u); const double f2 = u - fl; iz = wrap_fast(fl, g.n2);
            key = (i0 * g.n1 + i1) * g.n2 + iz;
            const double e0 = 1.0 - f0, e1 = 1.0 - f1, e2 = 1.0 - f2;
            if (MODE == 0) {
                double a0 = e0, a1 = f0;
                if (MASSKIND != 0) {
                    const double m = MASSKIND == 2 ? __ldg(mass + (long long)p * mstride) : mass0;
                    a0 *= m;
                    a1 *= m;
                }
                double w;
                w = a0 * e1; S000 = w * e2; S001 = w * f2;
                w = a0 * f1; S010 = w * e2; S011 = w * f2;
                w = a1 * e1; S100 = w * e2; S101 = w * f2;
                w = a1 * f1; S110 = w * e2; S111 = w * f2;
            } else {
                const double* vp = vpos + 3LL * p;
                const double m = MASSKIND == 2 ? __ldg(mass + (long long)p * mstride)
                                               : (MASSKIND == 1 ? mass0 : 1.0);
                const double q0 = m * __ldg(vp + 0) * g.s0, q1 = m * __ldg(vp + 1) * g.s1,
                             q2 = m * __ldg(vp + 2) * g.s2;
                const double vm = HAS_VMASS ? __ldg(vmass + p) : 0.0;
#define VPM_TANGENT(W0, D0, W1, D1, SA, SB)                                          \
    {                                                                                \
        const double wxy = (W0) * (W1);                                              \
        const double gxy = (D0) * (W1) + (W0) * (D1);                                \
        const double base = HAS_VMASS ? vm * wxy : 0.0;                              \
        SA = gxy * e2 - q2 * wxy + base * e2;                                        \
        SB = gxy * f2 + q2 * wxy + base * f2;                                        \
    }
                VPM_TANGENT(e0, -q0, e1, -q1, S000, S001)
                VPM_TANGENT(e0, -q0, f1, q1, S010, S011)
                VPM_TANGENT(f0, q0, e1, -q1, S100, S101)
                VPM_TANGENT(f0, q0, f1, q1, S110, S111)
#undef VPM_TANGENT
            }
        }
        // segmented inclusive scan over runs of equal cell key (adjacent lanes, thanks to the
        // sort); skipped when every lane starts its own run, rounds adapt to the longest run
        const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, (lane == 0) || (kprev != key));
        if (heads != 0xffffffffu) {
            for (int d = 1; d < 32; d <<= 1) {
                const int kd = __shfl_up_sync(0xffffffffu, key, d);
                const int same = (lane >= d) && (kd == key);
                if (!__any_sync(0xffffffffu, same)) break;
#define VPM_SCAN(V)                                                                  \
    {                                                                                \
        const double tv = __shfl_up_sync(0xffffffffu, V, d);                         \
        VPM_ADD_IF(V, tv, same);                                                     \
    }
                VPM_SCAN(S000) VPM_SCAN(S001) VPM_SCAN(S010) VPM_SCAN(S011)
                VPM_SCAN(S100) VPM_SCAN(S101) VPM_SCAN(S110) VPM_SCAN(S111)
#undef VPM_SCAN
            }
        }
        const int knext = __shfl_down_sync(0xffffffffu, key, 1);
        const bool tail = active && (lane == 31 || knext != key);
        // z-merge: a run whose successor in this warp is cell z + 1 of the same row hands its
        // c = 1 sums to that run instead of writing them.  The run before mine ends at lane
        // (head of my run) - 1.
        const bool hand_over = tail && lane < 31 && knext == key + 1 && iz + 1 < g.n2;
        const int myhead = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const int sl = max(myhead - 1, 0);
        const int ksrc = __shfl_sync(0xffffffffu, key, sl);
        const int take = tail && myhead > 0 && ksrc == key - 1 && iz > 0;
        {
            const double t0 = __shfl_sync(0xffffffffu, S001, sl);
            const double t1 = __shfl_sync(0xffffffffu, S011, sl);
            const double t2 = __shfl_sync(0xffffffffu, S101, sl);
            const double t3 = __shfl_sync(0xffffffffu, S111, sl);
            VPM_ADD_IF(S000, t0, take);
            VPM_ADD_IF(S010, t1, take);
            VPM_ADD_IF(S100, t2, take);
            VPM_ADD_IF(S110, t3, take);
        }
        if (tail) {
            const int j1 = wrap_inc(i1, g.n1), jz = wrap_inc(iz, g.n2);
            const int la = local_plane(g, i0), lb = local_plane(g, wrap_inc(i0, g.n0));
            const int ra = la * st0, rb = lb * st0;
            const int c0 = i1 * st1, c1 = j1 * st1;
            if (la >= 0) {
                atomicAdd(out + (ra + c0 + iz), S000);
                atomicAdd(out + (ra + c1 + iz), S010);
                if (!hand_over) {
                    atomicAdd(out + (ra + c0 + jz), S001);
                    atomicAdd(out + (ra + c1 + jz), S011);
                }
            }
            if (lb >= 0) {
                atomicAdd(out + (rb + c0 + iz), S100);
                atomicAdd(out + (rb + c1 + iz), S110);
                if (!hand_over) {
                    atomicAdd(out + (rb + c0 + jz), S101);
                    atomicAdd(out + (rb + c1 + jz), S111);
                }
            }
        }
    }
}

// Three columns of an interleaved [np, mstride] mass array into three meshes in ONE pass
// (reverse sweep of the force operator: the three paints of the force cotangent share the
// positions).  Same algorithm as k_paint_agg; position loads, cell arithmetic, run detection,
// scan control and target addresses are shared by the three columns, only the weighted sums,
// their shuffles and the reductions are per column (~45 % fewer instructions than 3 launches).
__global__ void __launch_bounds__(256, 3)
k_paint_agg3(Geo g, const double* __restrict__ xs, const double* __restrict__ mass, int mstride,
             int np, double* __restrict__ out, long long colstride, int cpw) {
    const int lane = threadIdx.x & 31;
    const int nchunk = (np + 31) >> 5;
    const int st0 = (int)g.stride0, st1 = (int)g.stride1;
    // consecutive chunks per warp, CTAs in index order: one front through the sorted particles
    // (see k_paint_agg)
    const int ch0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * cpw;
    const int chend = min(nchunk, ch0 + cpw);
    for (int ch = ch0; ch < chend; ++ch) {
        const int p = (ch << 5) + lane;
        const bool active = p < np;
        int key = -1 - lane;
        int i0 = 0, i1 = 0, iz = 0;
        double S[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < 8; ++q) S[c][q] = 0.0;
        if (active) {
            const double* xp = xs + 3LL * p;
            const double x0 = __ldg(xp + 0), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
            const double* mp = mass + (long long)p * mstride;
            const double m0 = __ldg(mp + 0), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
            double u, fl;
            u = __dmul_rn(x0, g.s0); fl = floor(u); const double f0 = u - fl; i0 = wrap_fast(fl, g.n0);
            u = __dmul_rn(x1, g.s1); fl = floor(u); const double f1 = u - fl; i1 = wrap_fast(fl, g.n1);
            u = __dmul_rn(x2, g.s2); fl = floor(u); const double f2 = u - fl; iz = wrap_fast(fl, g.n2);
            key = (i0 * g.n1 + i1) * g.n2 + iz;
            const double e0 = 1.0 - f0, e1 = 1.0 - f1, e2 = 1.0 - f2;
            double W[8];   // index a * 4 + b * 2 + cz
            double w;
            w = e0 * e1; W[0] = w * e2; W[1] = w * f2;
            w = e0 * f1; W[2] = w * e2; W[3] = w * f2;
            w = f0 * e1; W[4] = w * e2; W[5] = w * f2;
            w = f0 * f1; W[6] = w * e2; W[7] = w * f2;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                S[0][q] = m0 * W[q];
                S[1][q] = m1 * W[q];
                S[2][q] = m2 * W[q];
            }
        }
        const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, (lane == 0) || (kprev != key));
        if (heads != 0xffffffffu) {
            for (int d = 1; d < 32; d <<= 1) {
                const int kd = __shfl_up_sync(0xffffffffu, key, d);
                const int same = (lane >= d) && (kd == key);
                if (!__any_sync(0xffffffffu, same)) break;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double tv = __shfl_up_sync(0xffffffffu, S[c][q], d);
                        VPM_ADD_IF(S[c][q], tv, same);
                    }
            }
        }
        const int knext = __shfl_down_sync(0xffffffffu, key, 1);
        const bool tail = active && (lane == 31 || knext != key);
        const bool hand_over = tail && lane < 31 && knext == key + 1 && iz + 1 < g.n2;
        const int myhead = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const int sl = max(myhead - 1, 0);
        const int ksrc = __shfl_sync(0xffffffffu, key, sl);
        const int take = tail && myhead > 0 && ksrc == key - 1 && iz > 0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ab = 0; ab < 4; ++ab) {
                const double t = __shfl_sync(0xffffffffu, S[c][2 * ab + 1], sl);
                VPM_ADD_IF(S[c][2 * ab], t, take);
            }
        if (tail) {
            const int j1 = wrap_inc(i1, g.n1), jz = wrap_inc(iz, g.n2);
            const int la = local_plane(g, i0), lb = local_plane(g, wrap_inc(i0, g.n0));
            const int ra = la * st0, rb = lb * st0;
            const int c0 = i1 * st1, c1 = j1 * st1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double* o = out + c * colstride;
                if (la >= 0) {
                    atomicAdd(o + (ra + c0 + iz), S[c][0]);
                    atomicAdd(o + (ra + c1 + iz), S[c][2]);
                    if (!hand_over) {
                        atomicAdd(o + (ra + c0 + jz), S[c][1]);
                        atomicAdd(o + (ra + c1 + jz), S[c][3]);
                    }
                }
                if (lb >= 0) {
                    atomicAdd(o + (rb + c0 + iz), S[c][4]);
                    atomicAdd(o + (rb + c1 + iz), S[c][6]);
                    if (!hand_over) {
                        atomicAdd(o + (rb + c0 + jz), S[c][5]);
                        atomicAdd(o + (rb + c1 + jz), S[c][7]);
                    }
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------
// Block gather (no atomics, no zero fill, fixed summation order).
//
// A CTA produces a 2 x 2 x n2 block of OUTPUT cells: planes (l0, l0+1), rows (j0, j0+1), every z;
// thread t owns the column z = t.  Those cells receive mass from the 3 x 3 source rows
// (l0-1..l0+1) x (j0-1..j0+1): a particle of source row (ra, rb) feeds output (oa, ob) when
// ra - oa and rb - ob are in {0, 1} taken from the low side, with weight f (upper neighbour) or
// 1 - f (own cell) per axis -- 16 (row, output) pairs for 9 row visits, so a particle is
// visited 2.25 times instead of the 4 of the row formulation, its position is decoded once per
// visit for up to four outputs, and the loads of a warp are the consecutive particles of 32
// consecutive cells (coalesced).  Cell z feeds output z (1 - f2) and z + 1 (f2): the second sum
// moves one lane up by shuffle and across warps / the periodic wrap through shared memory.
// The cell ranges of all nine rows and the first particle of each are requested before any
// arithmetic (one load latency per phase instead of nine).  NCOL = 3 paints the three columns
// of an interleaved [np, mstride] mass array into three meshes sharing positions and weights
// (reverse sweep of the force operator).
// -------------------------------------------------------------------------------------------
template <int NCOL, int MASSKIND>
__global__ void __launch_bounds__(512)
k_paint_gather(Geo g, const double* __restrict__ xs, const double* __restrict__ mass, double mass0,
               int mstride, const int* __restrict__ cs, double* __restrict__ out,
               long long colstride) {
    extern __shared__ double sF[];  // [warps][NCOL * 4]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int z = threadIdx.x;
    const bool zv = z < g.n2;
    const int l0 = 2 * blockIdx.y, j0 = 2 * blockIdx.x;
    int sp[3], jy[3];
    bool vx[3] = {true, true, false}, vy[3] = {true, true, false};
    if (g.ghost) {
        sp[0] = l0; sp[1] = l0 + 1; sp[2] = l0 + 2;     // keys count from the lower ghost plane
    } else {
        sp[0] = l0 == 0 ? g.n0 - 1 : l0 - 1; sp[1] = l0; sp[2] = l0 + 1;
    }
    vx[2] = l0 + 1 < g.nloc0;
    jy[0] = j0 == 0 ? g.n1 - 1 : j0 - 1; jy[1] = j0; jy[2] = j0 + 1;
    vy[2] = j0 + 1 < g.n1;

    // phase A: the nine cell ranges; phase B: the first particle of each
    int lo[9], hi[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int ra = r / 3, rb = r % 3;
        lo[r] = hi[r] = 0;
        if (zv && vx[ra] && vy[rb]) {
            const int key = (sp[ra] * g.n1 + jy[rb]) * g.n2 + z;
            lo[r] = __ldg(cs + key);
            hi[r] = __ldg(cs + key + 1);
        }
    }
    double X0[9], X1[9], X2[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        X0[r] = X1[r] = X2[r] = 0.0;
        if (lo[r] < hi[r]) {
            const double* xp = xs + 3LL * lo[r];
            X0[r] = __ldg(xp + 0); X1[r] = __ldg(xp + 1); X2[r] = __ldg(xp + 2);
        }
    }
    double E[NCOL][2][2], F[NCOL][2][2];
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) { E[c][q >> 1][q & 1] = 0.0; F[c][q >> 1][q & 1] = 0.0; }

#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const int ra = r / 3, rb = r % 3;
        double x0 = X0[r], x1 = X1[r], x2 = X2[r];
        for (int k = lo[r]; k < hi[r];) {
            double u;
            u = __dmul_rn(x0, g.s0); const double f0 = u - floor(u);
            u = __dmul_rn(x1, g.s1); const double f1 = u - floor(u);
            u = __dmul_rn(x2, g.s2); const double f2 = u - floor(u);
            const double e0 = 1.0 - f0, e1 = 1.0 - f1, e2 = 1.0 - f2;
            double m[NCOL];
#pragma unroll
            for (int c = 0; c < NCOL; ++c)
                m[c] = MASSKIND == 2 ? __ldg(mass + (long long)k * mstride + c)
                                     : (MASSKIND == 1 ? mass0 : 1.0);
            ++k;
            if (k < hi[r]) {   // next particle of the cell: requested before the arithmetic
                const double* xp = xs + 3LL * k;
                x0 = __ldg(xp + 0); x1 = __ldg(xp + 1); x2 = __ldg(xp + 2);
            }
            // source row ra feeds output plane oa = ra (as the lower cell: weight f0) and
            // oa = ra - 1 (as its own cell: weight 1 - f0); same along y
#pragma unroll
            for (int oa = 0; oa < 2; ++oa) {
                if (ra != oa && ra != oa + 1) continue;
                const double wx = (ra == oa) ? f0 : e0;
#pragma unroll
                for (int ob = 0; ob < 2; ++ob) {
                    if (rb != ob && rb != ob + 1) continue;
                    const double wxy = wx * ((rb == ob) ? f1 : e1);
#pragma unroll
                    for (int c = 0; c < NCOL; ++c) {
                        const double wm = MASSKIND == 0 ? wxy : wxy * m[c];
                        E[c][oa][ob] += wm * e2;
                        F[c][oa][ob] += wm * f2;
                    }
                }
            }
        }
    }
    // the f2 sums belong to the next cell along z: one lane up, across warps through shared
    // memory, periodic wrap from the last cell of the row
    const int lastlane = (warp == nwarp - 1) ? ((g.n2 - 1) & 31) : 31;
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double f = F[c][q >> 1][q & 1];
            if (lane == lastlane) sF[warp * (NCOL * 4) + c * 4 + q] = f;
            F[c][q >> 1][q & 1] = __shfl_up_sync(0xffffffffu, f, 1);
        }
    __syncthreads();
    if (!zv) return;
    const int pw = warp == 0 ? nwarp - 1 : warp - 1;
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
#pragma unroll
        for (int oa = 0; oa < 2; ++oa) {
            if (oa == 1 && !vx[2]) continue;
#pragma unroll
            for (int ob = 0; ob < 2; ++ob) {
                if (ob == 1 && !vy[2]) continue;
                const double left = lane == 0 ? sF[pw * (NCOL * 4) + c * 4 + oa * 2 + ob] : F[c][oa][ob];
                out[c * colstride + (long long)(l0 + oa) * g.stride0 + (long long)(j0 + ob) * g.stride1 + z] =
                    E[c][oa][ob] + left;
            }
        }
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

template <int MODE, int MK, bool HV>
static void launch_tile_variant(vpm_ctx* ctx, const Geo& g, const TileDims& td, size_t smem,
                                const double* xs, const double* mass, double mass0,
                                const double* vpos, const double* vmass,
                                const int32_t* cell_start, double* out) {
    auto kern = k_paint_tile<MODE, MK, HV>;
    static bool configured = false;  // one attribute call per instantiation
    if (!configured) {
        VPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
    }
    const long long blocks = (long long)td.ntx * td.nty * td.ntz;
    VPM_REQUIRE(blocks < (1LL << 31), "mesh too large for one launch");
    VPM_LAUNCH(kern, (unsigned)blocks, TILE_WARPS * 32, smem, ctx->stream, g, td, xs, mass, mass0,
               vpos, vmass, cell_start, out);
}

static int g_tile_override[3] = {0, 0, 0};
static int g_paint_algo = 0;  // 0: warp-aggregated reductions (fast); 1: tile kernel (bit-reproducible)

static void zero_mesh(vpm_ctx* ctx, const Geo& g, double* out) {
    const long long total = (long long)g.nloc0 * g.n1 * g.n2;
    if (total == 0) return;
    if (g.stride1 == g.n2 && g.stride0 == (long long)g.n1 * g.n2) {
        VPM_PROF_CALL("memset(mesh)", ctx->stream,
                      VPM_CUDA(cudaMemsetAsync(out, 0, (size_t)total * sizeof(double), ctx->stream)));
    } else {
        unsigned zgrid = (unsigned)std::min<long long>(ceil_div(total, 256), 8LL * ctx->sm_count);
        VPM_LAUNCH(k_zero_mesh, zgrid, 256, 0, ctx->stream, g, out);
    }
}

// consecutive chunks per warp of the aggregated scatter: up to 16 (measured at 512^3: 1 / 2 / 4 / 8 / 16 / 32
// chunks -> 4.92 / 4.92 / 4.77 / 4.70 / 4.63 / 5.65 ms per three-column launch), fewer when the problem is small
// (at least ~16 CTAs per SM in the grid)
static int paint_chunks_per_warp(const vpm_ctx* ctx, long long nchunk) {
    static const int cap = [] { const char* e = getenv("VMAD_B200_PAINT_CPW"); return e ? std::max(1, atoi(e)) : 16; }();
    return (int)std::max<long long>(1, std::min<long long>(cap, nchunk / (8LL * 16 * ctx->sm_count)));
}

template <int MODE>
static void launch_paint_agg(vpm_ctx* ctx, const Geo& g, const double* xs, const double* mass,
                             double mass0, const double* vpos, const double* vmass, long long np,
                             double* out, int mstride = 1) {
    zero_mesh(ctx, g, out);
    if (np == 0) return;
    VPM_REQUIRE(np < (1LL << 31) - 64, "particle count must fit int32");
    VPM_REQUIRE((long long)g.nloc0 * g.stride0 < (1LL << 31), "local mesh slab must have < 2^31 elements");
    const long long nchunk = (np + 31) / 32;
    const int cpw = paint_chunks_per_warp(ctx, nchunk);
    const unsigned grid = (unsigned)ceil_div(nchunk, 8LL * cpw);
    const int mk = mass ? 2 : (mass0 == 1.0 ? 0 : 1);
    cudaStream_t st = ctx->stream;
#define VPM_AGG(MODE_, MK_, HV_) VPM_LAUNCH((k_paint_agg<MODE_, MK_, HV_>), grid, 256, 0, st, g, xs, mass, mass0, mstride, vpos, vmass, (int)np, out, cpw)
    if (MODE == 0) {
        if (mk == 2) VPM_AGG(0, 2, false);
        else if (mk == 1) VPM_AGG(0, 1, false);
        else VPM_AGG(0, 0, false);
    } else {
        if (vmass) {
            if (mk == 2) VPM_AGG(1, 2, true);
            else if (mk == 1) VPM_AGG(1, 1, true);
            else VPM_AGG(1, 0, true);
        } else {
            if (mk == 2) VPM_AGG(1, 2, false);
            else if (mk == 1) VPM_AGG(1, 1, false);
            else VPM_AGG(1, 0, false);
        }
    }
#undef VPM_AGG
}

template <int MODE>
static void launch_paint_tile(vpm_ctx* ctx, const Geo& g, const double* xs, const double* mass,
                              double mass0, const double* vpos, const double* vmass,
                              const int32_t* cell_start, double* out) {
    if (g.nloc0 == 0) return;
    TileDims td;
    td.TX = std::min(g_tile_override[0] > 0 ? g_tile_override[0] : 4, g.nloc0);
    td.TY = std::min(g_tile_override[1] > 0 ? g_tile_override[1] : 8, g.n1);
    td.TZ = std::min(g_tile_override[2] > 0 ? g_tile_override[2] : 32, g.n2);
    if ((td.TX * td.TY * td.TZ) & 1) td.TZ += 1;  // the zero fill writes double2
    td.ntx = (g.nloc0 + td.TX - 1) / td.TX;
    td.nty = (g.n1 + td.TY - 1) / td.TY;
    td.ntz = (g.n2 + td.TZ - 1) / td.TZ;
    const size_t smem = (size_t)4 * td.TX * td.TY * td.TZ * sizeof(double);
    VPM_REQUIRE(smem <= 160 * 1024, "paint tile does not fit shared memory");
    const int mk = mass ? 2 : (mass0 == 1.0 ? 0 : 1);
#define VPM_TILE(MODE_, MK_, HV_) launch_tile_variant<MODE_, MK_, HV_>(ctx, g, td, smem, xs, mass, mass0, vpos, vmass, cell_start, out)
    if (MODE == 0) {
        if (mk == 2) VPM_TILE(0, 2, false);
        else if (mk == 1) VPM_TILE(0, 1, false);
        else VPM_TILE(0, 0, false);
    } else {
        if (vmass) {
            if (mk == 2) VPM_TILE(1, 2, true);
            else if (mk == 1) VPM_TILE(1, 1, true);
            else VPM_TILE(1, 0, true);
        } else {
            if (mk == 2) VPM_TILE(1, 2, false);
            else if (mk == 1) VPM_TILE(1, 1, false);
            else VPM_TILE(1, 0, false);
        }
    }
#undef VPM_TILE
}

// the block-gather kernel needs the cell offsets, one thread per z cell in a CTA, and either a
// ghost-keyed slab or the whole periodic mesh
static bool gather_ok(const Geo& g, const int32_t* cell_start) {
    return cell_start != nullptr && g.n2 <= 512 && (g.ghost || (g.start0 == 0 && g.nloc0 == g.n0));
}

template <int NCOL>
static void launch_paint_gather(vpm_ctx* ctx, const Geo& g, const double* xs, const double* mass,
                                double mass0, int mstride, const int32_t* cell_start, double* out,
                                long long colstride) {
    if (g.nloc0 == 0) return;
    const int warps = (g.n2 + 31) / 32;
    dim3 grid((unsigned)((g.n1 + 1) / 2), (unsigned)((g.nloc0 + 1) / 2));
    VPM_REQUIRE(grid.y <= 65535, "too many planes for one launch");
    const size_t smem = (size_t)warps * NCOL * 4 * sizeof(double);
    const int mk = mass ? 2 : (mass0 == 1.0 ? 0 : 1);
    cudaStream_t st = ctx->stream;
#define VPM_GATHER(MK_) VPM_LAUNCH((k_paint_gather<NCOL, MK_>), grid, warps * 32, smem, st, g, xs, mass, mass0, mstride, cell_start, out, colstride)
    if (mk == 2) VPM_GATHER(2);
    else if (mk == 1) VPM_GATHER(1);
    else VPM_GATHER(0);
#undef VPM_GATHER
}

extern "C" {

int vpm_paint_sorted(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs, const double* mass,
                     double mass0, int64_t np, const int32_t* cell_start, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("paint", 24.0 * np + (mass ? 8.0 : 0.0) * np + 8.0 * mesh_cells(mesh));
    VPM_REQUIRE(ctx && out && (cell_start || g_paint_algo != 1), "NULL argument");
    VPM_REQUIRE(np == 0 || xs, "xs is NULL");
    Geo g = make_geo(mesh);
    if (g_paint_algo == 2 && gather_ok(g, cell_start))
        launch_paint_gather<1>(ctx, g, xs, mass, mass0, 1, cell_start, out, 0);
    else if (g_paint_algo == 1) launch_paint_tile<0>(ctx, g, xs, mass, mass0, nullptr, nullptr, cell_start, out);
    else launch_paint_agg<0>(ctx, g, xs, mass, mass0, nullptr, nullptr, np, out);
    VPM_API_END
}

int vpm_paint_sorted_rows(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs, const double* mass,
                          double mass0, int64_t np, const int32_t* cell_start, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("paint_rows", 24.0 * np + (mass ? 8.0 : 0.0) * np + 8.0 * mesh_cells(mesh));
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
    VPM_PROF_GROUP("paint_jvp", 48.0 * np + (mass ? 8.0 : 0.0) * np + (vmass ? 8.0 : 0.0) * np + 8.0 * mesh_cells(mesh));
    VPM_REQUIRE(ctx && out && cell_start, "NULL argument");
    VPM_REQUIRE(np == 0 || (xs && vpos), "xs / vpos is NULL");
    Geo g = make_geo(mesh);
    if (g_paint_algo == 0) launch_paint_agg<1>(ctx, g, xs, mass, mass0, vpos, vmass, np, out);
    else launch_paint_tile<1>(ctx, g, xs, mass, mass0, vpos, vmass, cell_start, out);
    VPM_API_END
}

int vpm_paint_sorted_col(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs, const double* mass2d,
                         int ncomp, int col, int64_t np, const int32_t* cell_start, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("paint", 32.0 * np + 8.0 * mesh_cells(mesh));
    VPM_REQUIRE(ctx && out && mass2d, "NULL argument");
    VPM_REQUIRE(ncomp >= 1 && col >= 0 && col < ncomp, "bad column");
    VPM_REQUIRE(np == 0 || xs, "xs is NULL");
    Geo g = make_geo(mesh);
    if (g_paint_algo == 2 && gather_ok(g, cell_start))
        launch_paint_gather<1>(ctx, g, xs, mass2d + col, 1.0, ncomp, cell_start, out, 0);
    else
        launch_paint_agg<0>(ctx, g, xs, mass2d + col, 1.0, nullptr, nullptr, np, out, ncomp);
    VPM_API_END
}

int vpm_paint_sorted_cols3(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs, const double* mass2d,
                           int ncomp, int64_t np, const int32_t* cell_start, double* out,
                           int64_t out_colstride) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("paint_cols3", 48.0 * np + 24.0 * mesh_cells(mesh));
    VPM_REQUIRE(ctx && out && mass2d, "NULL argument");
    VPM_REQUIRE(ncomp >= 3, "need at least three columns");
    VPM_REQUIRE(np == 0 || xs, "xs is NULL");
    Geo g = make_geo(mesh);
    if (g_paint_algo == 2 && gather_ok(g, cell_start)) {
        launch_paint_gather<3>(ctx, g, xs, mass2d, 1.0, ncomp, cell_start, out, out_colstride);
    } else {
        // one pass for the three columns (shared positions, cell arithmetic and run detection)
        for (int c = 0; c < 3; ++c) zero_mesh(ctx, g, out + c * out_colstride);
        if (np > 0) {
            VPM_REQUIRE(np < (1LL << 31) - 64, "particle count must fit int32");
            VPM_REQUIRE((long long)g.nloc0 * g.stride0 < (1LL << 31), "local mesh slab must have < 2^31 elements");
            const long long nchunk = (np + 31) / 32;
            const int cpw = paint_chunks_per_warp(ctx, nchunk);
            const unsigned grid = (unsigned)ceil_div(nchunk, 8LL * cpw);
            VPM_LAUNCH(k_paint_agg3, grid, 256, 0, ctx->stream, g, xs, mass2d, ncomp, (int)np, out,
                       (long long)out_colstride, cpw);
        }
    }
    VPM_API_END
}

int vpm_paint_set_algorithm(int algo) {
    VPM_API_BEGIN
    VPM_REQUIRE(algo >= 0 && algo <= 2, "paint algorithm must be 0 (aggregated), 1 (tile) or 2 (block gather)");
    g_paint_algo = algo;
    VPM_API_END
}

int vpm_paint_set_tile(int tx, int ty, int tz) {
    g_tile_override[0] = tx;
    g_tile_override[1] = ty;
    g_tile_override[2] = tz;
    return 0;
}

int vpm_paint_atomic(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x, const double* mass,
                     double mass0, int64_t np, double* out) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("paint_atomic", 24.0 * np + (mass ? 8.0 : 0.0) * np + 8.0 * mesh_cells(mesh));
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
