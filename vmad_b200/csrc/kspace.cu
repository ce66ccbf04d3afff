// k-space passes over the Hermitian-compressed complex mesh [n0, n1, nh].
//
// Replaces pmesh ComplexField.apply as driven by vmad/lib/fastpm.py:77-87 (apply_transfer and
// its vjp/jvp), the library filters fourier_space_laplace (:328-334) and
// fourier_space_neg_gradient (:310-326), and ComplexField.cnorm / cdot (:15,519).
//
// The wavenumber vectors and the per-axis gradient tables are evaluated ONCE on the host with
// numpy (the same numbers the oracle uses) and handed in as small device tables, so the device
// does no trigonometry and the Green's function -1/k^2 is bit-identical to numpy's:
// k2 = (kx*kx + ky*ky) + kz*kz with separately rounded products and sums.
// All kernels are single streaming passes: 16 B read + 16 B written per mode per output.
#include "common.cuh"

namespace vpm {

constexpr int KS_THREADS = 256;

struct CShape {
    long long n0, n1, nh, total;
};

__device__ __forceinline__ void decompose(const CShape& s, long long e, int& i, int& j, int& k) {
    long long row = e / s.nh;
    k = (int)(e - row * s.nh);
    i = (int)(row / s.n1);
    j = (int)(row - (long long)i * s.n1);
}

__device__ __forceinline__ double green(const double* __restrict__ kk0, const double* __restrict__ kk1,
                                        const double* __restrict__ kk2, int i, int j, int k) {
    const double a = __ldg(kk0 + i), b = __ldg(kk1 + j), c = __ldg(kk2 + k);
    const double k2 = __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c));
    return (k2 == 0.0) ? 0.0 : -(1.0 / k2);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ double2 load_tab(const double* __restrict__ t, int idx, bool conj) {
    double2 v = __ldg(reinterpret_cast<const double2*>(t) + idx);
    if (conj) v.y = -v.y;
    return v;
}

template <bool LAPLACE>
__global__ void __launch_bounds__(KS_THREADS)
k_transfer(CShape s, const double2* __restrict__ in, double2* __restrict__ out, double scale,
           const double* __restrict__ kk0, const double* __restrict__ kk1,
           const double* __restrict__ kk2, const double* __restrict__ a0,
           const double* __restrict__ a1, const double* __restrict__ a2, bool conj) {
    long long e = blockIdx.x * (long long)KS_THREADS + threadIdx.x;
    if (e >= s.total) return;
    int i, j, k;
    decompose(s, e, i, j, k);
    double2 v = in[e];
    if (LAPLACE) {
        const double L = green(kk0, kk1, kk2, i, j, k);
        v.x *= L;
        v.y *= L;
    }
    if (scale != 1.0) {
        v.x *= scale;
        v.y *= scale;
    }
    if (a0) v = cmul(v, load_tab(a0, i, conj));
    if (a1) v = cmul(v, load_tab(a1, j, conj));
    if (a2) v = cmul(v, load_tab(a2, k, conj));
    out[e] = v;
}

template <bool TCOMPLEX>
__global__ void __launch_bounds__(KS_THREADS)
k_transfer_table(CShape s, const double2* __restrict__ in, double2* __restrict__ out,
                 const double* __restrict__ table, long long t0, long long t1, long long t2,
                 bool conj) {
    long long e = blockIdx.x * (long long)KS_THREADS + threadIdx.x;
    if (e >= s.total) return;
    int i, j, k;
    decompose(s, e, i, j, k);
    const long long ti = i * t0 + j * t1 + k * t2;
    double2 v = in[e];
    if (TCOMPLEX) {
        double2 t = __ldg(reinterpret_cast<const double2*>(table) + ti);
        if (conj) t.y = -t.y;
        v = cmul(v, t);
    } else {
        const double t = __ldg(table + ti);
        v.x *= t;
        v.y *= t;
    }
    out[e] = v;
}

__global__ void __launch_bounds__(KS_THREADS)
k_force_kspace(CShape s, const double2* __restrict__ rhok, double scale,
               const double* __restrict__ kk0, const double* __restrict__ kk1,
               const double* __restrict__ kk2, const double* __restrict__ a0,
               const double* __restrict__ a1, const double* __restrict__ a2,
               double2* __restrict__ p_out, double2* __restrict__ g0, double2* __restrict__ g1,
               double2* __restrict__ g2) {
    long long e = blockIdx.x * (long long)KS_THREADS + threadIdx.x;
    if (e >= s.total) return;
    int i, j, k;
    decompose(s, e, i, j, k);
    double2 v = rhok[e];
    const double L = green(kk0, kk1, kk2, i, j, k) * scale;
    v.x *= L;
    v.y *= L;
    if (p_out) p_out[e] = v;
    g0[e] = cmul(v, load_tab(a0, i, false));
    g1[e] = cmul(v, load_tab(a1, j, false));
    g2[e] = cmul(v, load_tab(a2, k, false));
}

// adjoint of k_force_kspace: rhok_bar = scale * L * (sum_d conj(a_d) g_d_bar + p_bar)
__global__ void __launch_bounds__(KS_THREADS)
k_force_kspace_vjp(CShape s, const double2* __restrict__ g0, const double2* __restrict__ g1,
                   const double2* __restrict__ g2, const double2* __restrict__ p_bar, double scale,
                   const double* __restrict__ kk0, const double* __restrict__ kk1,
                   const double* __restrict__ kk2, const double* __restrict__ a0,
                   const double* __restrict__ a1, const double* __restrict__ a2,
                   double2* __restrict__ rhok_bar) {
    long long e = blockIdx.x * (long long)KS_THREADS + threadIdx.x;
    if (e >= s.total) return;
    int i, j, k;
    decompose(s, e, i, j, k);
    double2 acc = cmul(g0[e], load_tab(a0, i, true));
    double2 t = cmul(g1[e], load_tab(a1, j, true));
    acc.x += t.x;
    acc.y += t.y;
    t = cmul(g2[e], load_tab(a2, k, true));
    acc.x += t.x;
    acc.y += t.y;
    if (p_bar) {
        const double2 pb = p_bar[e];
        acc.x += pb.x;
        acc.y += pb.y;
    }
    const double L = green(kk0, kk1, kk2, i, j, k) * scale;
    rhok_bar[e] = make_double2(acc.x * L, acc.y * L);
}

// ---- Hermitian-weighted reduction sum_k w_k x conj(y) -----------------------------------
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level sum of two values; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
    __shared__ double sa[RED_THREADS / 32], sb[RED_THREADS / 32];
    a = warp_sum(a);
    b = warp_sum(b);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        sa[w] = a;
        sb[w] = b;
    }
    __syncthreads();
    if (w == 0) {
        a = lane < RED_THREADS / 32 ? sa[lane] : 0.0;
        b = lane < RED_THREADS / 32 ? sb[lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RED_THREADS)
k_cdot_partial(CShape s, long long n_last, const double2* __restrict__ x,
               const double2* __restrict__ y, double* __restrict__ partial) {
    double re = 0.0, im = 0.0;
    const bool even = (n_last % 2) == 0;
    for (long long e = blockIdx.x * (long long)RED_THREADS + threadIdx.x; e < s.total;
         e += (long long)gridDim.x * RED_THREADS) {
        const int k = (int)(e % s.nh);
        const double w = (k == 0 || (even && k == n_last / 2)) ? 1.0 : 2.0;
        const double2 a = x[e], b = y[e];
        re += w * (a.x * b.x + a.y * b.y);
        im += w * (a.y * b.x - a.x * b.y);
    }
    block_sum2(re, im);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x + 0] = re;
        partial[2 * blockIdx.x + 1] = im;
    }
}

__global__ void __launch_bounds__(RED_THREADS)
k_final2(int nparts, const double* __restrict__ partial, double* __restrict__ out) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nparts; i += RED_THREADS) {
        a += partial[2 * i + 0];
        b += partial[2 * i + 1];
    }
    block_sum2(a, b);
    if (threadIdx.x == 0) {
        out[0] = a;
        out[1] = b;
    }
}

// ---- binned transfer functions (apply_digitized, vmad/lib/fastpm.py:97-213) -----------------
// mode m belongs to bin ind[m] (evaluated once on the host by the user's digitizer, like any
// tf(k)); the per-bin table t has nb complex entries.
//   y[m] = x[m] * T(m),  T(m) = (cj[m] ? conj(t[ind[m]]) : t[ind[m]]) inside [0, nb), else dflt
__global__ void __launch_bounds__(KS_THREADS)
k_digitized_mul(long long n, const double2* __restrict__ x, const int* __restrict__ ind,
                const unsigned char* __restrict__ cj, int nb, const double2* __restrict__ t,
                double dflt, bool conj_all, double2* __restrict__ y) {
    long long e = blockIdx.x * (long long)KS_THREADS + threadIdx.x;
    if (e >= n) return;
    const int b = ind[e];
    double2 T = make_double2(dflt, 0.0);
    if (b >= 0 && b < nb) {
        T = t[b];
        if ((cj && cj[e]) != conj_all) T.y = -T.y;
    }
    y[e] = cmul(x[e], T);
}

// the vjp with respect to the per-bin table: out[b] += mult(m) Re(ybar[m] conj(pre x[m] E(m))) over
// the modes of bin b, mult = 2 for modes that stand for two modes of the full spectrum
// (fastpm.py:174,188); E = 1 (amplitude mode) or the bin's phase factor and pre = i (phase mode).
// Block-level bins in shared memory, one global reduction per bin and block.
constexpr int DIG_MAXBINS = 4096;
__global__ void __launch_bounds__(KS_THREADS)
k_digitized_bin(long long n, long long nh, long long n_last, const double2* __restrict__ x,
                const double2* __restrict__ ybar, const int* __restrict__ ind,
                const unsigned char* __restrict__ cj, int nb, const double2* __restrict__ eit,
                double* __restrict__ out) {
    extern __shared__ double bins[];
    for (int b = threadIdx.x; b < nb; b += KS_THREADS) bins[b] = 0.0;
    __syncthreads();
    for (long long e = blockIdx.x * (long long)KS_THREADS + threadIdx.x; e < n;
         e += (long long)gridDim.x * KS_THREADS) {
        const int b = ind[e];
        if (b < 0 || b >= nb) continue;
        const long long k = e % nh;
        const double mult = (k == 0 || k == n_last / 2) ? 1.0 : 2.0;
        double2 v = x[e];
        if (eit) {                       // y = x E, then i y
            double2 E = eit[b];
            if (cj && cj[e]) E.y = -E.y;
            v = cmul(v, E);
            v = make_double2(-v.y, v.x);
        }
        const double2 g = ybar[e];
        atomicAdd(bins + b, mult * (g.x * v.x + g.y * v.y));     // Re(g conj(v))
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += KS_THREADS)
        if (bins[b] != 0.0) atomicAdd(out + b, bins[b]);
}

static CShape make_cshape(const int64_t nc[3]) {
    VPM_REQUIRE(nc && nc[0] >= 1 && nc[1] >= 1 && nc[2] >= 1, "bad complex shape");
    CShape s;
    s.n0 = nc[0]; s.n1 = nc[1]; s.nh = nc[2];
    s.total = s.n0 * s.n1 * s.nh;
    return s;
}

}  // namespace vpm

using namespace vpm;

extern "C" {

int vpm_transfer(vpm_ctx* ctx, const int64_t nc[3], const double* in, double* out, double scale,
                 int laplace, const double* kk0, const double* kk1, const double* kk2,
                 const double* a0, const double* a1, const double* a2, int conj_tables) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("transfer", 32.0 * ((double)nc[0] * nc[1] * nc[2]));
    VPM_REQUIRE(ctx && in && out, "NULL argument");
    CShape s = make_cshape(nc);
    VPM_REQUIRE(!laplace || (kk0 && kk1 && kk2), "laplace needs the three wavenumber tables");
    unsigned grid = (unsigned)ceil_div(s.total, KS_THREADS);
    const double2* i2 = reinterpret_cast<const double2*>(in);
    double2* o2 = reinterpret_cast<double2*>(out);
    if (laplace) VPM_LAUNCH(k_transfer<true>, grid, KS_THREADS, 0, ctx->stream, s, i2, o2, scale, kk0, kk1, kk2, a0, a1, a2, conj_tables != 0);
    else VPM_LAUNCH(k_transfer<false>, grid, KS_THREADS, 0, ctx->stream, s, i2, o2, scale, kk0, kk1, kk2, a0, a1, a2, conj_tables != 0);
    VPM_API_END
}

int vpm_transfer_table(vpm_ctx* ctx, const int64_t nc[3], const double* in, double* out,
                       const double* table, const int64_t tstride[3], int table_is_complex,
                       int conj_table) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("transfer", 32.0 * ((double)nc[0] * nc[1] * nc[2]));
    VPM_REQUIRE(ctx && in && out && table && tstride, "NULL argument");
    CShape s = make_cshape(nc);
    unsigned grid = (unsigned)ceil_div(s.total, KS_THREADS);
    const double2* i2 = reinterpret_cast<const double2*>(in);
    double2* o2 = reinterpret_cast<double2*>(out);
    if (table_is_complex) VPM_LAUNCH(k_transfer_table<true>, grid, KS_THREADS, 0, ctx->stream, s, i2, o2, table, (long long)tstride[0], (long long)tstride[1], (long long)tstride[2], conj_table != 0);
    else VPM_LAUNCH(k_transfer_table<false>, grid, KS_THREADS, 0, ctx->stream, s, i2, o2, table, (long long)tstride[0], (long long)tstride[1], (long long)tstride[2], conj_table != 0);
    VPM_API_END
}

int vpm_force_kspace(vpm_ctx* ctx, const int64_t nc[3], const double* rhok, double scale,
                     const double* kk0, const double* kk1, const double* kk2, const double* a0,
                     const double* a1, const double* a2, double* p_out, double* g0, double* g1,
                     double* g2) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("force_kspace", (64.0 + (p_out ? 16.0 : 0.0)) * ((double)nc[0] * nc[1] * nc[2]));
    VPM_REQUIRE(ctx && rhok && kk0 && kk1 && kk2 && a0 && a1 && a2 && g0 && g1 && g2, "NULL argument");
    CShape s = make_cshape(nc);
    unsigned grid = (unsigned)ceil_div(s.total, KS_THREADS);
    VPM_LAUNCH(k_force_kspace, grid, KS_THREADS, 0, ctx->stream, s,
               reinterpret_cast<const double2*>(rhok), scale, kk0, kk1, kk2, a0, a1, a2,
               reinterpret_cast<double2*>(p_out), reinterpret_cast<double2*>(g0),
               reinterpret_cast<double2*>(g1), reinterpret_cast<double2*>(g2));
    VPM_API_END
}

int vpm_force_kspace_vjp(vpm_ctx* ctx, const int64_t nc[3], const double* g0, const double* g1,
                         const double* g2, const double* p_bar, double scale, const double* kk0,
                         const double* kk1, const double* kk2, const double* a0, const double* a1,
                         const double* a2, double* rhok_bar) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("force_kspace_vjp", (64.0 + (p_bar ? 16.0 : 0.0)) * ((double)nc[0] * nc[1] * nc[2]));
    VPM_REQUIRE(ctx && g0 && g1 && g2 && kk0 && kk1 && kk2 && a0 && a1 && a2 && rhok_bar, "NULL argument");
    CShape s = make_cshape(nc);
    unsigned grid = (unsigned)ceil_div(s.total, KS_THREADS);
    VPM_LAUNCH(k_force_kspace_vjp, grid, KS_THREADS, 0, ctx->stream, s,
               reinterpret_cast<const double2*>(g0), reinterpret_cast<const double2*>(g1),
               reinterpret_cast<const double2*>(g2), reinterpret_cast<const double2*>(p_bar), scale,
               kk0, kk1, kk2, a0, a1, a2, reinterpret_cast<double2*>(rhok_bar));
    VPM_API_END
}

int vpm_digitized_mul(vpm_ctx* ctx, int64_t n, const double* x, const int32_t* ind, const uint8_t* conj_mask,
                      int nbins, const double* table, double dflt, int conj_table, double* y) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("digitized_mul", 37.0 * n);
    VPM_REQUIRE(ctx && nbins >= 0, "bad argument");
    if (n > 0) {
        VPM_REQUIRE(x && ind && y && (nbins == 0 || table), "NULL argument");
        VPM_LAUNCH(k_digitized_mul, (unsigned)ceil_div(n, KS_THREADS), KS_THREADS, 0, ctx->stream, (long long)n,
                   reinterpret_cast<const double2*>(x), ind, conj_mask, nbins,
                   reinterpret_cast<const double2*>(table), dflt, conj_table != 0, reinterpret_cast<double2*>(y));
    }
    VPM_API_END
}

int vpm_digitized_bin(vpm_ctx* ctx, const int64_t nc[3], int64_t n_last, const double* x, const double* ybar,
                      const int32_t* ind, const uint8_t* conj_mask, int nbins, const double* eit, double* out) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && out && nbins >= 1 && nbins <= DIG_MAXBINS, "bad argument (up to 4096 bins)");
    CShape s = make_cshape(nc);
    VPM_PROF_GROUP("digitized_bin", 37.0 * s.total);
    VPM_CUDA(cudaMemsetAsync(out, 0, (size_t)nbins * sizeof(double), ctx->stream));
    VPM_REQUIRE(x && ybar && ind, "NULL argument");
    const int grid = (int)std::min<long long>(ceil_div(s.total, KS_THREADS), 8LL * ctx->sm_count);
    VPM_LAUNCH(k_digitized_bin, grid, KS_THREADS, (size_t)nbins * sizeof(double), ctx->stream, s.total, s.nh,
               (long long)n_last, reinterpret_cast<const double2*>(x), reinterpret_cast<const double2*>(ybar), ind,
               conj_mask, nbins, reinterpret_cast<const double2*>(eit), out);
    VPM_API_END
}

int vpm_reduce_cdot(vpm_ctx* ctx, const int64_t nc[3], int64_t n_last, const double* x,
                    const double* y, double* out_host) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("reduce_cdot", 32.0 * ((double)nc[0] * nc[1] * nc[2]));
    VPM_REQUIRE(ctx && x && y && out_host, "NULL argument");
    CShape s = make_cshape(nc);
    int nparts = (int)std::min<long long>(ceil_div(s.total, RED_THREADS), 4LL * ctx->sm_count);
    double* partial = (double*)ctx->workspace((size_t)nparts * 2 * sizeof(double));
    VPM_LAUNCH(k_cdot_partial, nparts, RED_THREADS, 0, ctx->stream, s, (long long)n_last,
               reinterpret_cast<const double2*>(x), reinterpret_cast<const double2*>(y), partial);
    VPM_LAUNCH(k_final2, 1, RED_THREADS, 0, ctx->stream, nparts, partial, ctx->dev_scalars);
    VPM_CUDA(cudaMemcpyAsync(ctx->host_scalars, ctx->dev_scalars, 2 * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    out_host[0] = ctx->host_scalars[0];
    out_host[1] = ctx->host_scalars[1];
    VPM_API_END
}

}  // extern "C"
