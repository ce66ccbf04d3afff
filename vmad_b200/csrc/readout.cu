// CIC readout (gather), its gradient and its tangent.
//
// Replaces pmesh RealField.readout / readout_vjp / readout_jvp and the particle half of
// ParticleMesh.paint_vjp (vmad/lib/fastpm.py:229,252,256,263).
//
// One thread per particle.  The eight mesh values of a particle are four pairs of cells that
// are adjacent along the contiguous axis; the pair loads are branch-free (what was loaded from
// a plane that is not held is discarded by a select), so the loads of a particle are issued
// back to back.  How a pair is fetched (two 8-byte loads, one 16-byte load where aligned, an
// aligned chunk + an odd-lane load) and how the particle rows move (plain loads, a bulk-copied
// shared-memory tile, 16 + 8-byte accesses, streaming hints) are VARIANTS of the two hot
// kernels, all bit-identical; the default is the measured best (benchmarks/readout_variants.py:
// the plainest one -- two 8-byte loads, plain rows, 48 registers, 5 CTAs per SM).  Particles
// arrive cell-sorted from `exchange` (or in lattice order), so neighbouring threads touch
// neighbouring cells and the mesh is read from HBM once; the re-use is served by L1/L2.
// Algorithmic HBM bytes: 24 B/particle read + 8 B/cell read + 8 B (value) or 24 B (gradient,
// or three fused fields) per particle written.
#include "cic.cuh"
#include "tile.cuh"
#include <initializer_list>

namespace vpm {

struct Stencil {
    int off[4];        // element offsets of rows (a, b) = (0,0) (0,1) (1,0) (1,1); the local slab has < 2^31 elements (checked)
    bool ok[2];        // plane a = 0 / 1 is held by this slab (always, on one rank)
    int i2, j2;        // the two cells along the contiguous axis
    double f0, f1, f2;
};

__device__ __forceinline__ Stencil make_stencil(const Geo& g, double x0, double x1, double x2) {
    Stencil s;
    int i0, i1;
    double u, fl;
    u = __dmul_rn(x0, g.s0); fl = floor(u); s.f0 = __dsub_rn(u, fl); i0 = wrap_fast(fl, g.n0);
    u = __dmul_rn(x1, g.s1); fl = floor(u); s.f1 = __dsub_rn(u, fl); i1 = wrap_fast(fl, g.n1);
    u = __dmul_rn(x2, g.s2); fl = floor(u); s.f2 = __dsub_rn(u, fl); s.i2 = wrap_fast(fl, g.n2);
    s.j2 = wrap_inc(s.i2, g.n2);
    const int j1 = wrap_inc(i1, g.n1);
    const int la = local_plane(g, i0), lb = local_plane(g, wrap_inc(i0, g.n0));
    s.ok[0] = la >= 0;
    s.ok[1] = lb >= 0;
    const int st0 = (int)g.stride0, st1 = (int)g.stride1;
    const int ra = max(la, 0) * st0, rb = max(lb, 0) * st0;
    s.off[0] = ra + i1 * st1;
    s.off[1] = ra + j1 * st1;
    s.off[2] = rb + i1 * st1;
    s.off[3] = rb + j1 * st1;
    return s;
}

__device__ __forceinline__ Stencil make_stencil(const Geo& g, const double* __restrict__ x,
                                                long long p) {
    return make_stencil(g, __ldg(x + 3 * p + 0), __ldg(x + 3 * p + 1), __ldg(x + 3 * p + 2));
}

// the pair (field[row + i2], field[row + j2]); zeros when the row is not held here.
// AL (rows 16-byte aligned, n2 even): branch-free -- the aligned 16-byte chunk that holds cell i2,
// plus an 8-byte load of cell j2 on the lanes where i2 is odd (the pair then straddles two chunks
// or wraps).  No control flow, so the loads of all pairs of a particle are issued back to back and
// their latencies overlap.
template <int PM>   // 0: two 8-byte loads; 1: branch-free aligned chunk + odd-lane load; 2: 16-byte load where the pair is aligned, else two 8-byte loads
__device__ __forceinline__ void load_pair(const double* __restrict__ field, int row, bool ok, int i2,
                                          int j2, double& v0, double& v1) {
    // the address is always valid (rows of planes that are not held are clamped to plane 0 by
    // make_stencil); what was loaded is discarded by a select
    const double* q = field + row;
    if (PM == 1) {
        const bool odd = (i2 & 1) != 0;
        const double2 A = __ldg(reinterpret_cast<const double2*>(q + (i2 & ~1)));
        const double B = __ldg(q + (odd ? j2 : i2));      // even lanes: same sector as A (L1 hit)
        v0 = odd ? A.y : A.x;
        v1 = odd ? B : A.y;
    } else if (PM == 2) {
        const double* p = q + i2;
        if (j2 == i2 + 1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const double2 t = __ldg(reinterpret_cast<const double2*>(p));
            v0 = t.x;
            v1 = t.y;
        } else {
            v0 = __ldg(p);
            v1 = __ldg(q + j2);
        }
    } else {
        v0 = __ldg(q + i2);
        v1 = __ldg(q + j2);
    }
    v0 = ok ? v0 : 0.0;
    v1 = ok ? v1 : 0.0;
}

// fields whose rows can be read with aligned 16-byte accesses
inline bool rows_aligned(const Geo& g, std::initializer_list<const void*> fields) {
    bool al = (g.n2 % 2 == 0) && (g.stride0 % 2 == 0) && (g.stride1 % 2 == 0);
    for (const void* f : fields) al = al && (f == nullptr || (reinterpret_cast<uintptr_t>(f) & 15) == 0);
    return al;
}

// value = sum over corners in (a, b, c) order of field * ((w0 * w1) * w2)
template <int AL>
__device__ __forceinline__ double gather_value(const Stencil& s, const double* __restrict__ field) {
    double acc = 0.0;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        const double w0 = (ab >> 1) ? s.f0 : 1.0 - s.f0;
        const double w1 = (ab & 1) ? s.f1 : 1.0 - s.f1;
        const double wxy = w0 * w1;
        double v0, v1;
        load_pair<AL>(field, s.off[ab], s.ok[ab >> 1], s.i2, s.j2, v0, v1);
        acc += v0 * (wxy * (1.0 - s.f2));
        acc += v1 * (wxy * s.f2);
    }
    return acc;
}

// value and the three derivatives d/dx_d (chain factor Nmesh_d / BoxSize_d included)
template <int AL>
__device__ __forceinline__ void gather_value_grad(const Geo& g, const Stencil& s,
                                                  const double* __restrict__ field, double& val,
                                                  double& g0, double& g1, double& g2) {
    val = g0 = g1 = g2 = 0.0;
    const double w2a = 1.0 - s.f2, w2b = s.f2;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
        const int a = ab >> 1, b = ab & 1;
        const double w0 = a ? s.f0 : 1.0 - s.f0;
        const double w1 = b ? s.f1 : 1.0 - s.f1;
        const double d0 = a ? g.s0 : -g.s0;
        const double d1 = b ? g.s1 : -g.s1;
        double v0, v1;
        load_pair<AL>(field, s.off[ab], s.ok[a], s.i2, s.j2, v0, v1);
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

template <int AL>
__global__ void __launch_bounds__(256)
k_readout(Geo g, const double* __restrict__ field, const double* __restrict__ x, long long np,
          double* __restrict__ value) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    value[p] = gather_value<AL>(s, field);
}

// Variant bits of the two hot kernels (ablation, vpm_readout_set_variant):
//   1  TILE   the positions (and weights) of the CTA's ROW_TILE particles arrive in shared memory by
//             one bulk asynchronous copy (tile.cuh) instead of 24-byte-strided loads per thread
//   2  EARLY  the scattered rows a particle needs (y_in, base at its home row) are requested
//             before the mesh is touched, so the latencies overlap instead of adding up
//   4  SPLIT  scattered 24-byte rows move as one 16-byte + one 8-byte access
//   32 STREAM scattered rows with streaming (evict-first) loads / stores
//   8/16      pair loads: 0 = 16-byte load where the pair is aligned, else two 8-byte loads
//             (branches); 8 = branch-free aligned chunk + odd-lane load; 16 = two 8-byte loads
constexpr int ROW_TILE = 256;
constexpr int RV_TILE = 1, RV_EARLY = 2, RV_SPLIT = 4, RV_PAIR_AL = 8, RV_PAIR_8 = 16, RV_STREAM = 32;
__host__ __device__ constexpr int rv_pair_mode(int v) { return (v & RV_PAIR_AL) ? 1 : ((v & RV_PAIR_8) ? 0 : 2); }

template <int V>
__device__ __forceinline__ void fetch_row3(const double* __restrict__ base, long long o, bool scattered,
                                           double& a, double& b, double& c) {
    if ((V & RV_STREAM) && scattered) load_row3_cs(base, o, a, b, c);
    else if ((V & RV_SPLIT) && scattered) load_row3(base, o, a, b, c);
    else { a = __ldg(base + 3 * o); b = __ldg(base + 3 * o + 1); c = __ldg(base + 3 * o + 2); }
}
template <int V>
__device__ __forceinline__ void write_row3(double* __restrict__ base, long long o, double a, double b, double c) {
    if (V & RV_STREAM) store_row3_cs(base, o, a, b, c);
    else if (V & RV_SPLIT) store_row3(base, o, a, b, c);
    else { base[3 * o] = a; base[3 * o + 1] = b; base[3 * o + 2] = c; }
}

template <int V, int MB>
__global__ void __launch_bounds__(ROW_TILE, MB)
k_readout3(Geo g, const double* __restrict__ fa, const double* __restrict__ fb,
           const double* __restrict__ fc, const double* __restrict__ x, long long np,
           const int* __restrict__ out_row, double* __restrict__ value3,
           const double* __restrict__ y_in, double fac, double* __restrict__ y_out, int dist) {
    // dist != 0 (layout over several ranks): value3 is stored in SLOT order for the rows whose home is
    // on another rank (they travel from there; the other slots are not written), y_out[o] = y_in[o] (0 when NULL) + fac * value only where
    // the home row o = out_row[p] is on this rank (o >= 0)
    constexpr int PM = rv_pair_mode(V);
    __shared__ RowTile<(V & RV_TILE) ? ROW_TILE : 1, 1> tile;
    const long long p0 = blockIdx.x * (long long)ROW_TILE;
    const long long p = p0 + threadIdx.x;
    const bool live = p < np;
    // out_row: the layout's slot -> home row map (Layout.gather fused into the readout)
    long long o = p;
    if ((V & RV_EARLY) && out_row && live) o = (long long)__ldg(out_row + p);
    if (V & RV_TILE) {
        const double* const arrs[1] = {x};
        tile.load(arrs, p0, np);
    }
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    if ((V & RV_EARLY) && y_out && y_in && live && o >= 0) fetch_row3<V>(y_in, o, out_row != nullptr, y0, y1, y2);
    double a = 0.0, b = 0.0, c = 0.0;
    if (live) {
        double x0, x1, x2;
        if (V & RV_TILE) tile.get(0, x, p0, p, x0, x1, x2);
        else { x0 = __ldg(x + 3 * p); x1 = __ldg(x + 3 * p + 1); x2 = __ldg(x + 3 * p + 2); }
        const Stencil s = make_stencil(g, x0, x1, x2);
        a = gather_value<PM>(s, fa);
        b = gather_value<PM>(s, fb);
        c = gather_value<PM>(s, fc);
    }
    if (!(V & RV_EARLY) && live) {
        if (out_row) o = (long long)__ldg(out_row + p);
        if (y_out && y_in && o >= 0) fetch_row3<V>(y_in, o, out_row != nullptr, y0, y1, y2);
    }
    if (dist) {
        if (!live) return;
        if (value3 && o < 0) { value3[3 * p + 0] = a; value3[3 * p + 1] = b; value3[3 * p + 2] = c; }
        if (y_out && o >= 0) write_row3<V>(y_out, o, fma(fac, a, y0), fma(fac, b, y1), fma(fac, c, y2));
        return;
    }
    if (out_row || !(V & RV_TILE)) {      // scattered rows (or no tile to stage through)
        if (!live) return;
        if (value3) write_row3<V>(value3, o, a, b, c);
        // the kick that consumes the force: y_out = y_in + fac * value (one FMA each)
        if (y_out) write_row3<V>(y_out, o, fma(fac, a, y0), fma(fac, b, y1), fma(fac, c, y2));
        return;
    }
    // rows in slot order: through the tile, one bulk store per array
    if (value3) {
        if (tile.can_flush(value3, p0, np)) {
            tile.put(0, threadIdx.x, a, b, c);
            tile.flush(0, value3, p0);
        } else if (live) {
            value3[3 * p + 0] = a; value3[3 * p + 1] = b; value3[3 * p + 2] = c;
        }
    }
    if (y_out) {
        const double r0 = fma(fac, a, y0), r1 = fma(fac, b, y1), r2 = fma(fac, c, y2);
        if (tile.can_flush(y_out, p0, np)) {
            __syncthreads();        // the bulk store of value3 has read the tile (wait_group.read by thread 0)
            tile.put(0, threadIdx.x, r0, r1, r2);
            tile.flush(0, y_out, p0);
        } else if (live) {
            y_out[3 * p + 0] = r0; y_out[3 * p + 1] = r1; y_out[3 * p + 2] = r2;
        }
    }
}

template <int AL, bool HAS_W, bool WANT_VALUE>
__global__ void __launch_bounds__(256)
k_readout_grad(Geo g, const double* __restrict__ field, const double* __restrict__ x,
               const double* __restrict__ weight, double weight0, long long np,
               double* __restrict__ value, double* __restrict__ grad) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    double val, g0, g1, g2;
    gather_value_grad<AL>(g, s, field, val, g0, g1, g2);
    const double w = HAS_W ? __ldg(weight + p) : weight0;
    if (WANT_VALUE) value[p] = val;
    grad[3 * p + 0] = g0 * w;
    grad[3 * p + 1] = g1 * w;
    grad[3 * p + 2] = g2 * w;
}

// grad[p, :] = sum_d w[p, d] * grad_readout(F_d, x_p) + grad_readout(R, x_p): the particle half
// of the reverse sweep of a fused force evaluation, one pass over the particles for the four
// meshes (F_0, F_1, F_2: the force components; R: the density cotangent)
template <int V, int MB>
__global__ void __launch_bounds__(ROW_TILE, MB)
k_readout_grad4(Geo g, const double* __restrict__ F0, const double* __restrict__ F1,
                const double* __restrict__ F2, const double* __restrict__ R,
                const double* __restrict__ x, const double* __restrict__ w3, long long np,
                const int* __restrict__ out_row, const double* __restrict__ base,
                double* __restrict__ grad, const double* __restrict__ y_in, double fac,
                double* __restrict__ y_out, int dist) {
    // dist != 0: grad is stored in SLOT order WITHOUT base for the rows whose home is elsewhere, y_out[o] = base[o] (0 when NULL) + grad
    // only where the home row o = out_row[p] is on this rank (o >= 0); y_in / fac unused
    constexpr int PM = rv_pair_mode(V);
    __shared__ RowTile<(V & RV_TILE) ? ROW_TILE : 1, 2> tile;
    const long long p0 = blockIdx.x * (long long)ROW_TILE;
    const long long p = p0 + threadIdx.x;
    const bool live = p < np;
    long long o = p;
    if ((V & RV_EARLY) && out_row && live) o = (long long)__ldg(out_row + p);
    if (V & RV_TILE) {
        const double* const arrs[2] = {x, w3};
        tile.load(arrs, p0, np);
    }
    // rows reached through the permutation (EARLY: requested now, consumed after the mesh)
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, y0 = 0.0, y1 = 0.0, y2 = 0.0;
    if ((V & RV_EARLY) && live && o >= 0) {
        if (base) fetch_row3<V>(base, o, out_row != nullptr, b0, b1, b2);
        if (y_out && !dist) fetch_row3<V>(y_in, o, out_row != nullptr, y0, y1, y2);
    }
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (live) {
        double x0, x1, x2, wa, wb, wc;
        if (V & RV_TILE) {
            tile.get(0, x, p0, p, x0, x1, x2);
            tile.get(1, w3, p0, p, wa, wb, wc);
        } else {
            x0 = __ldg(x + 3 * p); x1 = __ldg(x + 3 * p + 1); x2 = __ldg(x + 3 * p + 2);
            wa = __ldg(w3 + 3 * p); wb = __ldg(w3 + 3 * p + 1); wc = __ldg(w3 + 3 * p + 2);
        }
        const Stencil s = make_stencil(g, x0, x1, x2);
        const double w2a = 1.0 - s.f2, w2b = s.f2;
        // combined mesh value seen by this particle at its eight cells: m = wa F0 + wb F1 + wc F2 + R
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) {
            const int a = ab >> 1, b = ab & 1;
            const bool ok = s.ok[a];
            double u0, u1, t0, t1, r0, r1, q0, q1, z0, z1;
            load_pair<PM>(F0, s.off[ab], ok, s.i2, s.j2, t0, t1);
            load_pair<PM>(F1, s.off[ab], ok, s.i2, s.j2, r0, r1);
            load_pair<PM>(F2, s.off[ab], ok, s.i2, s.j2, q0, q1);
            load_pair<PM>(R, s.off[ab], ok, s.i2, s.j2, z0, z1);
            u0 = wa * t0;
            u1 = wa * t1;
            u0 += wb * r0;
            u1 += wb * r1;
            u0 += wc * q0;
            u1 += wc * q1;
            u0 += z0;
            u1 += z1;
            const double w0 = a ? s.f0 : 1.0 - s.f0;
            const double w1 = b ? s.f1 : 1.0 - s.f1;
            const double d0 = a ? g.s0 : -g.s0;
            const double d1 = b ? g.s1 : -g.s1;
            const double wxy = w0 * w1;
            a0 += u0 * ((d0 * w1) * w2a) + u1 * ((d0 * w1) * w2b);
            a1 += u0 * ((w0 * d1) * w2a) + u1 * ((w0 * d1) * w2b);
            a2 += (u1 - u0) * (wxy * g.s2);
        }
        if (!(V & RV_EARLY)) {
            if (out_row) o = (long long)__ldg(out_row + p);
            if (o >= 0) {
                if (base) fetch_row3<V>(base, o, out_row != nullptr, b0, b1, b2);
                if (y_out && !dist) fetch_row3<V>(y_in, o, out_row != nullptr, y0, y1, y2);
            }
        }
        // (gradient) + base: the sum the separate accumulation pass would form
        if (base && !dist) { a0 += b0; a1 += b1; a2 += b2; }
    }
    if (dist) {
        if (!live) return;
        if (grad && o < 0) { grad[3 * p + 0] = a0; grad[3 * p + 1] = a1; grad[3 * p + 2] = a2; }
        if (y_out && o >= 0) write_row3<V>(y_out, o, b0 + a0, b1 + a1, b2 + a2);
        return;
    }
    if (out_row || !(V & RV_TILE)) {
        if (!live) return;
        write_row3<V>(grad, o, a0, a1, a2);
        // the drift's momentum cotangent: y_out = y_in + fac * grad
        if (y_out) write_row3<V>(y_out, o, fma(fac, a0, y0), fma(fac, a1, y1), fma(fac, a2, y2));
        return;
    }
    if (tile.can_flush(grad, p0, np)) {
        tile.put(0, threadIdx.x, a0, a1, a2);
        tile.flush(0, grad, p0);
    } else if (live) {
        grad[3 * p + 0] = a0; grad[3 * p + 1] = a1; grad[3 * p + 2] = a2;
    }
    if (y_out) {
        const double r0 = fma(fac, a0, y0), r1 = fma(fac, a1, y1), r2 = fma(fac, a2, y2);
        if (tile.can_flush(y_out, p0, np)) {
            tile.put(1, threadIdx.x, r0, r1, r2);
            tile.flush(1, y_out, p0);
        } else if (live) {
            y_out[3 * p + 0] = r0; y_out[3 * p + 1] = r1; y_out[3 * p + 2] = r2;
        }
    }
}

template <int AL, bool HAS_FDOT, bool HAS_XDOT>
__global__ void __launch_bounds__(256)
k_readout_jvp(Geo g, const double* __restrict__ field, const double* __restrict__ field_dot,
              const double* __restrict__ x, const double* __restrict__ xdot, long long np,
              double* __restrict__ value) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= np) return;
    Stencil s = make_stencil(g, x, p);
    double r = 0.0;
    if (HAS_FDOT) r += gather_value<AL>(s, field_dot);
    if (HAS_XDOT) {
        double val, g0, g1, g2;
        gather_value_grad<AL>(g, s, field, val, g0, g1, g2);
        r += (__ldg(xdot + 3 * p + 0) * g0 + __ldg(xdot + 3 * p + 1) * g1) +
             __ldg(xdot + 3 * p + 2) * g2;
    }
    value[p] = r;
}

}  // namespace vpm

using namespace vpm;

// every combination of the variant bits (TILE, EARLY, SPLIT) x pair mode {0, 8, 16}
// variant id = bits | (minimum CTAs per SM asked of the compiler << 6)
#define VPM_RV_CASES(X) X(0, 1) X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(5, 1) X(6, 1) X(7, 1) X(8, 1) X(9, 1) X(10, 1) X(11, 1) \
    X(12, 1) X(13, 1) X(14, 1) X(15, 1) X(16, 1) X(17, 1) X(18, 1) X(19, 1) X(20, 1) X(21, 1) X(22, 1) X(23, 1) \
    X(16, 5) X(20, 5) X(16, 6) X(20, 6) X(16, 7) X(20, 7) X(16, 8) X(20, 8) X(48, 5) X(48, 4) X(16, 4)
static int g_readout_variant = RV_PAIR_8 | (5 << 6);   // measured best (benchmarks/readout_variants.py, profiles/r2_readout_variants_512.txt)

extern "C" {

int vpm_readout_set_variant(int v) {
    if (v < 0 || (v >> 6) > 8) return 1;
    g_readout_variant = (v >> 6) ? v : (v | (1 << 6));      // unknown combinations fail at the launch
    return 0;
}

int vpm_readout(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field, const double* x,
                int64_t np, double* value) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout", 32.0 * np + 8.0 * mesh_cells(mesh));
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(field && x && value, "NULL array");
        if (rows_aligned(g, {field})) VPM_LAUNCH(k_readout<2>, (unsigned)ceil_div(np, 256), 256, 0, ctx->stream, g, field, x, (long long)np, value);
        else VPM_LAUNCH(k_readout<0>, (unsigned)ceil_div(np, 256), 256, 0, ctx->stream, g, field, x, (long long)np, value);
    }
    VPM_API_END
}

static int readout3_impl(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                         const double* f2, const double* x, int64_t np, const int32_t* out_row,
                         double* value3, const double* y_in, double fac, double* y_out, int dist) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout3", 24.0 * np + 24.0 * mesh_cells(mesh) + (out_row ? 4.0 : 0.0) * np + (value3 ? 24.0 : 0.0) * np + (y_out ? 48.0 : 0.0) * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(f0 && f1 && f2 && x && (value3 || y_out), "NULL array");
        VPM_REQUIRE(!y_out || y_in || dist, "y_out needs y_in");
        VPM_REQUIRE(!dist || out_row, "home rows are required");
        int v = g_readout_variant;
        if (!rows_aligned(g, {f0, f1, f2}) && (v & RV_PAIR_AL)) v = (v & ~RV_PAIR_AL) | RV_PAIR_8;
        const unsigned grid = (unsigned)ceil_div(np, ROW_TILE);
#define VPM_R3(VV, MB) case VV | (MB << 6): VPM_LAUNCH((k_readout3<VV, MB>), grid, ROW_TILE, 0, ctx->stream, g, f0, f1, f2, x, (long long)np, out_row, value3, y_in, fac, y_out, dist); break;
        switch (v) { VPM_RV_CASES(VPM_R3) default: VPM_REQUIRE(false, "bad readout variant"); }
#undef VPM_R3
    }
    VPM_API_END
}

int vpm_readout3(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                 const double* f2, const double* x, int64_t np, const int32_t* out_row,
                 double* value3, const double* y_in, double fac, double* y_out) {
    return readout3_impl(ctx, mesh, f0, f1, f2, x, np, out_row, value3, y_in, fac, y_out, 0);
}

int vpm_readout3_home(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                      const double* f2, const double* x, int64_t np, const int32_t* home_row,
                      double* value3_slots, const double* acc_in, double scale, double* acc_out) {
    return readout3_impl(ctx, mesh, f0, f1, f2, x, np, home_row, value3_slots, acc_in, scale, acc_out, 1);
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
        const bool al = rows_aligned(g, {field});
        if (weight && value) { if (al) VPM_LAUNCH((k_readout_grad<2, true, true>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); else VPM_LAUNCH((k_readout_grad<0, true, true>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); }
        else if (weight) { if (al) VPM_LAUNCH((k_readout_grad<2, true, false>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); else VPM_LAUNCH((k_readout_grad<0, true, false>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); }
        else if (value) { if (al) VPM_LAUNCH((k_readout_grad<2, false, true>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); else VPM_LAUNCH((k_readout_grad<0, false, true>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); }
        else { if (al) VPM_LAUNCH((k_readout_grad<2, false, false>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); else VPM_LAUNCH((k_readout_grad<0, false, false>), grid, 256, 0, st, g, field, x, weight, weight0, (long long)np, value, grad); }
    }
    VPM_API_END
}

static int readout_grad4_impl(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                              const double* f2, const double* r, const double* x, const double* w3,
                              int64_t np, const int32_t* out_row, const double* base, double* grad,
                              const double* y_in, double fac, double* y_out, int dist) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("readout_grad4", 72.0 * np + 32.0 * mesh_cells(mesh) + (out_row ? 4.0 : 0.0) * np + (base ? 24.0 : 0.0) * np + (y_out ? 48.0 : 0.0) * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(f0 && f1 && f2 && r && x && w3 && (grad || (dist && y_out)), "NULL array");
        VPM_REQUIRE(!y_out || y_in || dist, "y_out needs y_in");
        VPM_REQUIRE(!dist || out_row, "home rows are required");
        int v = g_readout_variant;
        if (!rows_aligned(g, {f0, f1, f2, r}) && (v & RV_PAIR_AL)) v = (v & ~RV_PAIR_AL) | RV_PAIR_8;
        const unsigned grid = (unsigned)ceil_div(np, ROW_TILE);
#define VPM_G4(VV, MB) case VV | (MB << 6): VPM_LAUNCH((k_readout_grad4<VV, MB>), grid, ROW_TILE, 0, ctx->stream, g, f0, f1, f2, r, x, w3, (long long)np, out_row, base, grad, y_in, fac, y_out, dist); break;
        switch (v) { VPM_RV_CASES(VPM_G4) default: VPM_REQUIRE(false, "bad readout variant"); }
#undef VPM_G4
    }
    VPM_API_END
}

int vpm_readout_grad4(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                      const double* f2, const double* r, const double* x, const double* w3,
                      int64_t np, const int32_t* out_row, const double* base, double* grad,
                      const double* y_in, double fac, double* y_out) {
    return readout_grad4_impl(ctx, mesh, f0, f1, f2, r, x, w3, np, out_row, base, grad, y_in, fac, y_out, 0);
}

int vpm_readout_grad4_home(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                           const double* f2, const double* r, const double* x, const double* w3,
                           int64_t np, const int32_t* home_row, double* grad_slots,
                           const double* acc_in, double* acc_out) {
    return readout_grad4_impl(ctx, mesh, f0, f1, f2, r, x, w3, np, home_row, acc_in, grad_slots, nullptr, 0.0,
                              acc_out, 1);
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
        const bool al = rows_aligned(g, {field, field_dot});
        if (field_dot && xdot) { if (al) VPM_LAUNCH((k_readout_jvp<2, true, true>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); else VPM_LAUNCH((k_readout_jvp<0, true, true>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); }
        else if (field_dot) { if (al) VPM_LAUNCH((k_readout_jvp<2, true, false>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); else VPM_LAUNCH((k_readout_jvp<0, true, false>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); }
        else if (xdot) { if (al) VPM_LAUNCH((k_readout_jvp<2, false, true>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); else VPM_LAUNCH((k_readout_jvp<0, false, true>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); }
        else { if (al) VPM_LAUNCH((k_readout_jvp<2, false, false>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); else VPM_LAUNCH((k_readout_jvp<0, false, false>), grid, 256, 0, st, g, field, field_dot, x, xdot, (long long)np, value); }
    }
    VPM_API_END
}

}  // extern "C"
