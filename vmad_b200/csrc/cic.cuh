// Position -> (cell, fractional offset) arithmetic shared by every particle kernel.
// Mirrors oracle/pmesh/pm.py::ParticleMesh.cell_index operation for operation:
//   u = x * scale (one rounded fp64 multiply, never contracted into an FMA),
//   i = floor(u), f = u - i (exact), and the *integer* is wrapped periodically.
#pragma once
#include "common.cuh"

namespace vpm {

struct Geo {
    int n0, n1, n2;          // global mesh
    double s0, s1, s2;       // Nmesh / BoxSize
    int start0, nloc0;       // local slab of planes
    long long stride0, stride1;
    int ghost;               // 1: keys count nloc0+1 source planes (lower ghost first)
    int nsp;                 // number of source planes keyed: ghost ? nloc0 + 1 : n0
};

inline Geo make_geo(const vpm_mesh* m) {
    VPM_REQUIRE(m != nullptr, "mesh is NULL");
    for (int d = 0; d < 3; ++d)
        VPM_REQUIRE(m->n[d] >= 1 && m->n[d] <= (1 << 20), "mesh size out of range");
    VPM_REQUIRE(m->nloc0 >= 0 && m->start0 >= 0 && m->start0 + m->nloc0 <= m->n[0],
                "bad slab");
    Geo g;
    g.n0 = (int)m->n[0]; g.n1 = (int)m->n[1]; g.n2 = (int)m->n[2];
    g.s0 = m->scale[0]; g.s1 = m->scale[1]; g.s2 = m->scale[2];
    g.start0 = (int)m->start0; g.nloc0 = (int)m->nloc0;
    g.stride0 = m->stride0; g.stride1 = m->stride1;
    g.ghost = m->ghost;
    g.nsp = m->ghost ? g.nloc0 + 1 : g.n0;
    VPM_REQUIRE(g.stride1 >= g.n2 && g.stride0 >= g.stride1 * (long long)g.n1, "bad strides");
    VPM_REQUIRE((long long)g.nsp * g.n1 * g.n2 + 1 < (1LL << 31), "too many cells for int32 keys");
    VPM_REQUIRE((long long)g.nloc0 * g.stride0 < (1LL << 31), "local mesh slab must have < 2^31 elements");
    return g;
}

// wrapped cell index and fractional offset along one axis
__device__ __forceinline__ void cell_axis(double x, double scale, int n, int& i, double& f) {
    double u = __dmul_rn(x, scale);
    double fl = floor(u);
    f = __dsub_rn(u, fl);
    double dn = (double)n;
    if (fl >= 0.0 && fl < dn) {
        i = (int)fl;
    } else if (fl >= -dn && fl < 0.0) {
        i = (int)fl + n;
    } else if (fl >= dn && fl < 2.0 * dn) {
        i = (int)fl - n;
    } else {
        // far outside the box (or not finite): exact 64-bit modulo, NaN/Inf -> cell 0
        long long ll = (fabs(fl) < 9.0e18) ? (long long)fl : 0LL;
        ll %= (long long)n;
        if (ll < 0) ll += n;
        i = (int)ll;
    }
}

// Periodic wrap of floor(u) given as a double.  The common case (inside the box or within one
// box length outside) costs a compare, the general case is the exact 64-bit modulo of cell_axis,
// so the two always agree.
__device__ __forceinline__ int wrap_cell(double fl, int n) {
    int i = __double2int_rn(fl);  // exact for |fl| < 2^31, saturates beyond
    if ((unsigned)i >= (unsigned)n) {
        if (i < 0 && i >= -n) {
            i += n;
        } else if (i >= n && i < 2 * n) {
            i -= n;
        } else {
            long long ll = (fabs(fl) < 9.0e18) ? (long long)fl : 0LL;
            ll %= (long long)n;
            if (ll < 0) ll += n;
            i = (int)ll;
        }
    }
    return i;
}

// wrapped cell index, branch-free in the common case (inside the box or one box length out)
__device__ __forceinline__ int wrap_fast(double fl, int n) {
    int i = __double2int_rn(fl);  // exact for |fl| < 2^31, saturates beyond
    i += (i < 0) ? n : 0;
    i -= (i >= n) ? n : 0;
    if (__builtin_expect((unsigned)i >= (unsigned)n, 0)) i = wrap_cell(fl, n);
    return i;
}

__device__ __forceinline__ int wrap_inc(int i, int n) { return (i + 1 == n) ? 0 : i + 1; }

// Source-plane index used in cell keys: plane i0 itself on a periodic single slab, or the
// offset from the lower ghost plane on a slab; returns -1 when the particle's floor plane
// is not keyed on this slab.
__device__ __forceinline__ int source_plane(const Geo& g, int i0) {
    if (!g.ghost) return i0;
    int sp = i0 - g.start0 + 1;
    if (sp < 0) sp += g.n0;
    if (sp >= g.n0) sp -= g.n0;
    return (sp <= g.nloc0) ? sp : -1;
}

// Local plane index of global plane i0, or -1 when it is not held here.
__device__ __forceinline__ int local_plane(const Geo& g, int i0) {
    int l = i0 - g.start0;
    return (l >= 0 && l < g.nloc0) ? l : -1;
}

}  // namespace vpm
