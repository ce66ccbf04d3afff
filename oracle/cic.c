/*
 * oracle/cic.c -- plain C restatement of the CIC window loops of the oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Same formulas, in the same order, as the
 * numpy versions in oracle/pmesh/pm.py (_cic_paint, _cic_paint_gradient, _cic_readout,
 * _cic_readout_gradient), which restate pmesh's ResampleWindow 'cic' paint / readout as called
 * from /root/reference/vmad/lib/fastpm.py:224,229,238,252,256,263:
 *
 *     u_d = x_d * scale_d;  i_d = floor(u_d);  f_d = u_d - i_d;  i_d wrapped as an integer;
 *     weight(a,b,c) = (w0 * w1) * w2,  w_d = 1 - f_d (corner 0) or f_d (corner 1);
 *     dW/dx_d      = product with w_d replaced by -scale_d (corner 0) or +scale_d (corner 1).
 *
 * It exists so that the CPU baseline reported next to the GPU numbers is a compiled loop and
 * not numpy fancy indexing.  Meshes are 3-D, C order ([1, N1, N2] for 2-D).  nthreads > 1 uses
 * OpenMP (paint then accumulates with atomic updates: results are summation-order dependent at
 * the 1e-16 level); nthreads == 1 is strictly sequential in particle order.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -o oracle/_build/libcic_oracle.so oracle/cic.c
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t i[3][2]; /* wrapped indices of the two cells per axis */
    double w[3][2];  /* weights */
    double d[3][2];  /* derivative of the weights */
} stencil_t;

static inline int64_t wrap(int64_t i, int64_t n) {
    int64_t r = i % n;
    return r < 0 ? r + n : r;
}

static inline void make_stencil(const double* x, const int64_t* n, const double* scale,
                                stencil_t* s) {
    for (int d = 0; d < 3; ++d) {
        double u = x[d] * scale[d];
        double fl = floor(u);
        double f = u - fl;
        int64_t i = (int64_t)fl;
        s->i[d][0] = wrap(i, n[d]);
        s->i[d][1] = wrap(i + 1, n[d]);
        s->w[d][0] = 1.0 - f;
        s->w[d][1] = f;
        s->d[d][0] = -scale[d];
        s->d[d][1] = scale[d];
    }
}

void cic_paint(const double* pos, const double* mass, double mass0, int64_t np,
               const int64_t* n, const double* scale, double* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t p = 0; p < np; ++p) {
        stencil_t s;
        make_stencil(pos + 3 * p, n, scale, &s);
        const double m = mass ? mass[p] : mass0;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const double w = (s.w[0][a] * s.w[1][b]) * s.w[2][c];
                    const int64_t flat = (s.i[0][a] * n[1] + s.i[1][b]) * n[2] + s.i[2][c];
                    const double v = m * w;
                    if (nthreads > 1) {
#pragma omp atomic
                        out[flat] += v;
                    } else {
                        out[flat] += v;
                    }
                }
    }
}

void cic_paint_gradient(const double* pos, const double* mass, double mass0, const double* vpos,
                        int64_t np, const int64_t* n, const double* scale, double* out,
                        int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t p = 0; p < np; ++p) {
        stencil_t s;
        make_stencil(pos + 3 * p, n, scale, &s);
        const double m = mass ? mass[p] : mass0;
        const double* v = vpos + 3 * p;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const double g0 = (s.d[0][a] * s.w[1][b]) * s.w[2][c];
                    const double g1 = (s.w[0][a] * s.d[1][b]) * s.w[2][c];
                    const double g2 = (s.w[0][a] * s.w[1][b]) * s.d[2][c];
                    const double cc = (v[0] * g0 + v[1] * g1) + v[2] * g2;
                    const int64_t flat = (s.i[0][a] * n[1] + s.i[1][b]) * n[2] + s.i[2][c];
                    const double val = m * cc;
                    if (nthreads > 1) {
#pragma omp atomic
                        out[flat] += val;
                    } else {
                        out[flat] += val;
                    }
                }
    }
}

void cic_readout(const double* mesh, const double* pos, int64_t np, const int64_t* n,
                 const double* scale, double* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t p = 0; p < np; ++p) {
        stencil_t s;
        make_stencil(pos + 3 * p, n, scale, &s);
        double acc = 0.0;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const double w = (s.w[0][a] * s.w[1][b]) * s.w[2][c];
                    const int64_t flat = (s.i[0][a] * n[1] + s.i[1][b]) * n[2] + s.i[2][c];
                    acc += mesh[flat] * w;
                }
        out[p] = acc;
    }
}

void cic_readout_gradient(const double* mesh, const double* pos, int64_t np, const int64_t* n,
                          const double* scale, double* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static) if (nthreads > 1)
    for (int64_t p = 0; p < np; ++p) {
        stencil_t s;
        make_stencil(pos + 3 * p, n, scale, &s);
        double g[3] = {0.0, 0.0, 0.0};
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const int64_t flat = (s.i[0][a] * n[1] + s.i[1][b]) * n[2] + s.i[2][c];
                    const double mv = mesh[flat];
                    g[0] += mv * ((s.d[0][a] * s.w[1][b]) * s.w[2][c]);
                    g[1] += mv * ((s.w[0][a] * s.d[1][b]) * s.w[2][c]);
                    g[2] += mv * ((s.w[0][a] * s.w[1][b]) * s.d[2][c]);
                }
        out[3 * p + 0] = g[0];
        out[3 * p + 1] = g[1];
        out[3 * p + 2] = g[2];
    }
}

/* out[idx[i], :] += data[i, :]  (Layout.gather, mode 'sum'); sequential: idx may repeat */
void scatter_add_rows(const double* data, const int64_t* idx, int64_t n, int64_t ncomp,
                      double* out) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t c = 0; c < ncomp; ++c) out[idx[i] * ncomp + c] += data[i * ncomp + c];
}

/* wrapped cell key ((i0 * n1) + i1) * n2 + i2 of every particle */
void cic_cell_keys(const double* pos, int64_t np, const int64_t* n, const double* scale,
                   int64_t* keys) {
    for (int64_t p = 0; p < np; ++p) {
        int64_t key = 0;
        for (int d = 0; d < 3; ++d) {
            double u = pos[3 * p + d] * scale[d];
            key = key * n[d] + wrap((int64_t)floor(u), n[d]);
        }
        keys[p] = key;
    }
}
