/*
 * vmadpm.h -- C ABI of libvmadpm.so, the B200 (sm_100a) particle-mesh backend that
 * replaces the pmesh calls made by vmad's FastPM operator library.
 *
 * Every entry point names the reference interface it stands in for, as
 *   vmad/lib/fastpm.py:<line>  (the call site in /root/reference)  and the pmesh
 * method that call site delegates to (pmesh is an un-vendored dependency of the
 * reference; SURVEY.md Appendix A lists the protocol).
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on failure; vpm_last_error() then
 *     returns a message (thread-local, valid until the next failing call);
 *   - all pointers except `host_*` / `out_host` are DEVICE pointers owned by the
 *     caller; the library never frees or retains them beyond the call's kernels;
 *   - all work is enqueued on the context's stream and is asynchronous; the only
 *     calls that wait for the device are vpm_sync, vpm_layout_count (returns a
 *     host integer) and the vpm_reduce_* family (return host scalars);
 *   - fp64 throughout ("f8"), complex128 stored interleaved (re, im);
 *   - meshes are 3-D, C order, last axis contiguous.  A 2-D mesh [N1, N2] is passed
 *     as [1, N1, N2] with scale[0] = 1 and positions' first coordinate = 0;
 *   - slab decomposition: a rank holds planes [start0, start0 + nloc0) of axis 0 of
 *     the real mesh.  Single-rank: start0 = 0, nloc0 = n[0], ghost = 0.
 */
#ifndef VMADPM_H
#define VMADPM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vpm_ctx vpm_ctx;

/* Geometry of the (local slab of the) real mesh. */
typedef struct vpm_mesh {
    int64_t n[3];      /* global mesh size Nmesh                                  */
    double scale[3];   /* Nmesh / BoxSize: u = x * scale (one rounded multiply)   */
    int64_t start0;    /* first plane of axis 0 held locally                      */
    int64_t nloc0;     /* number of planes held locally                           */
    int64_t stride0;   /* element stride between planes of the real array         */
    int64_t stride1;   /* element stride between rows   (>= n[2]; n[2]+2 when the
                          array is the padded in-place c2r output)                */
    int32_t ghost;     /* 0: periodic single slab; 1: slab with a lower ghost
                          plane of *particles* (cell keys count nloc0 + 1 planes) */
    int32_t reserved;
} vpm_mesh;

/* ---- context -------------------------------------------------------------- */
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream), 0 = legacy
 * default stream. */
int vpm_ctx_create(vpm_ctx** out, int device, void* stream);
int vpm_ctx_destroy(vpm_ctx* ctx);
int vpm_ctx_set_stream(vpm_ctx* ctx, void* stream);
/* Launch the following calls on `stream` without draining the current stream or re-binding the FFT
 * plans -- only for entry points that use neither cuFFT nor the context workspace (the peer puts
 * of the pipelined slab transform); the caller orders the streams with events.  pop restores. */
int vpm_ctx_push_stream(vpm_ctx* ctx, void* stream);
int vpm_ctx_pop_stream(vpm_ctx* ctx);
int vpm_sync(vpm_ctx* ctx);
const char* vpm_last_error(void);
int vpm_version(void);
/* number of kernels (own kernels + cuFFT executions) launched by this library */
int64_t vpm_launch_count(void);
int vpm_device_sm_count(vpm_ctx* ctx, int* out_host);

/* In-place profiling for the roofline report (no reference counterpart; bench.py's measurement
 * contract): between _start and _stop every kernel launch / cuFFT execution / memset of this
 * library is bracketed by two CUDA events on the launching stream.  vpm_profile_read waits for
 * them and writes one text line per launch, "group\tkernel\tmilliseconds\talgorithmic bytes of
 * the launch\talgorithmic bytes of the group instance\n" (0 where an entry point states none);
 * it returns the buffer size the text needs (-1 on error) and releases the records once the text
 * fitted into buf. */
int vpm_profile_start(void);
int vpm_profile_stop(void);
int64_t vpm_profile_read(char* buf, int64_t cap);

/* ---- decompose / exchange / gather  (fastpm.py:272,286,289,301,304; pmesh
 *      ParticleMesh.decompose, Layout.exchange, Layout.gather) ---------------- */
/* cell key of every particle: ((i0' * n1) + i1) * n2 + i2 of the wrapped floor cell,
 * i0' = i0 (ghost = 0) or (i0 - start0 + 1) mod n0 (ghost = 1). int32. */
int vpm_cell_keys(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x, int64_t np,
                  int32_t* keys);
/* number of cells keyed by vpm_cell_keys for this mesh (host value, no sync) */
int64_t vpm_layout_ncell(const vpm_mesh* mesh);
/* stable counting sort by cell key.  Outputs: keys[np], perm[np] (slot -> input
 * index; equal keys keep input order), cell_start[ncell + 1] (first slot of each
 * cell; cell_start[ncell] = np). */
int vpm_layout_build(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x, int64_t np,
                     int32_t* keys, int32_t* perm, int32_t* cell_start);
/* An index equal to VPM_ROW_PADDING marks a slot that holds no particle (the unused tail of a
 * fixed-capacity segment, vpm_route_build_static): the take functions write the context's fill
 * row there when the array has three columns (0 otherwise; vpm_ctx_set_fill_row, default 0), the
 * scatter / put functions skip it like any negative index. */
#define VPM_ROW_PADDING (-2147483647 - 1)
int vpm_ctx_set_fill_row(vpm_ctx* ctx, const double* row3_host);
/* sorted_to_recv[t] = VPM_ROW_PADDING for every slot t >= sum(recv_counts[0 .. nranks)) (device
 * counts): the padding rows of a fixed-capacity receive order sort last */
int vpm_route_mark_padding(vpm_ctx* ctx, int32_t* sorted_to_recv, int64_t n, const int32_t* recv_counts,
                           int nranks);
/* dst[i, :] = src[perm[i], :]            (Layout.exchange on one rank) */
int vpm_take_rows(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n,
                  int ncomp, double* dst);
/* dst[i, :] = scale * src[perm[i], :]  (exchange of a cotangent fused with the constant factor of
 * the kick it comes from: reverse sweep of the fused drift-force-kick step) */
int vpm_take_rows_scaled(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n,
                         int ncomp, double scale, double* dst);
/* dst[perm[i], :] = src[i, :]  for a permutation (Layout.gather on one rank) */
int vpm_scatter_rows(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n,
                     int ncomp, double* dst);
/* dst[idx[i], :] += src[i, :]  (gather with duplicated copies; dst pre-zeroed).  Rows with a
 * negative index are skipped (vpm_scatter_rows too). */
int vpm_scatter_add_rows(vpm_ctx* ctx, const double* src, const int32_t* idx, int64_t n,
                         int ncomp, double* dst);
/* dst[idx[i], :] += scale * src[i, :]  (gather across ranks fused with a kick / drift update: dst
 * starts as a copy of the array being updated) */
int vpm_scatter_add_rows_scaled(vpm_ctx* ctx, const double* src, const int32_t* idx, int64_t n,
                                int ncomp, double scale, double* dst);
/* dst[i, :] = idx[i] >= 0 ? a[idx[i], :] : b[-1 - idx[i], :]   (Layout.exchange across ranks: rows
 * that stay on this rank are read from the home array a, the rest from the receive area b) */
int vpm_take_rows2(vpm_ctx* ctx, const double* a, const double* b, const int32_t* idx, int64_t n,
                   int ncomp, double* dst);
/* ... times a constant (the kick factor riding on the exchange of a momentum cotangent) */
int vpm_take_rows2_scaled(vpm_ctx* ctx, const double* a, const double* b, const int32_t* idx, int64_t n,
                          int ncomp, double scale, double* dst);
/* Index composition for a cross-rank layout.  sorted_to_recv = the local sort permutation over the
 * receive order; the self segment is rows [recv_lo, recv_hi) of the receive order and rows
 * [send_lo, send_lo + recv_hi - recv_lo) of the send order.
 *   take[k]   = send_idx[send_lo + r - recv_lo] if r = sorted_to_recv[k] is a self row, else -1 - r
 *   remote[j] = send_idx[j] outside the self segment, -1 inside
 *   inv_remote[r'] (optional, nrecv - (recv_hi - recv_lo) ints) = the sorted slot k of receive
 *               row r outside the self segment (r' = r below the segment, r - its length above),
 *               -1 for rows no slot refers to */
int vpm_route_compose(vpm_ctx* ctx, const int32_t* sorted_to_recv, int64_t nrecv,
                      const int32_t* send_idx, int64_t nsend, int64_t recv_lo, int64_t recv_hi,
                      int64_t send_lo, int32_t* take, int32_t* remote, int32_t* inv_remote);

/* Routing for the slab decomposition over `nranks` ranks (rank r owns planes
 * [r n0/nranks, (r+1) n0/nranks)): a particle is sent to the owner of its floor plane and, when
 * different, to the owner of the next plane.  send_idx[send_start[d] .. send_start[d+1]) are the
 * particle indices bound for rank d, in original order (stable).  send_start_host[nranks + 1]
 * is delivered to the host (this call synchronises: the counts size the all-to-all, as the
 * MPI_Alltoall of counts does inside pmesh's decompose).  send_capacity >= 2 np is always safe. */
int vpm_route_build(vpm_ctx* ctx, const vpm_mesh* mesh, int nranks, const double* x, int64_t np,
                    int32_t* send_idx, int64_t send_capacity, int32_t* send_start_host);
/* The same routing with FIXED-CAPACITY segments and no host synchronisation (what makes a
 * multi-rank step capturable in a CUDA graph): destination d owns positions
 * [seg_start_host[d], seg_start_host[d+1]) of send_idx; positions beyond the rows actually bound
 * for d hold -1.  counts (device, nranks + 1 ints; counts[nranks] must be zero-initialised by the
 * caller once) receives the number of rows per destination and, in counts[nranks], a sticky flag
 * that is set when a segment was too small (the surplus rows are dropped: the caller checks the
 * flag when it next synchronises and re-plans the capacities). */
int vpm_route_build_static(vpm_ctx* ctx, const vpm_mesh* mesh, int nranks, const double* x, int64_t np,
                           const int32_t* seg_start_host, int32_t* send_idx, int32_t* counts);

/* ---- paint  (fastpm.py:224,238,256; pmesh ParticleMesh.paint / paint_jvp and the
 *      mesh half of RealField.readout_vjp) ---------------------------------- */
/* Particles sorted by cell key (xs = x[perm]); fills every local cell (no prior zero fill
 * needed).  mass: per-particle array (sorted like xs) or NULL -> mass0.
 * Algorithms (vpm_paint_set_algorithm): 0 warp-aggregated fp64 reductions into L2 -- needs no
 * cell_start (may be NULL), summation order not fixed (results agree to rounding); 1 shared-memory
 * tiles (ablation); 2 block gather -- no atomics, no zero fill, every cell written exactly once,
 * run-to-run bit-identical; needs cell_start and n[2] <= 512, otherwise falls back to 0. */
int vpm_paint_set_algorithm(int algo);
int vpm_paint_sorted(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs,
                     const double* mass, double mass0, int64_t np,
                     const int32_t* cell_start, double* out);
/* mass = column `col` of an interleaved [np, ncomp] array (no gather of the column needed);
 * used by the reverse sweep of the fused force operator (three paints of the columns of the
 * force cotangent).  Always the aggregated-reduction algorithm. */
int vpm_paint_sorted_col(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs,
                         const double* mass2d, int ncomp, int col, int64_t np,
                         const int32_t* cell_start, double* out);
/* the three columns of an interleaved [np, ncomp] array into three meshes, out + c *
 * out_colstride (c = 0, 1, 2), in ONE pass over the particles: positions, cell arithmetic and run
 * detection are shared by the columns (aggregated reductions, or the block gather when that
 * algorithm is selected). */
int vpm_paint_sorted_cols3(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs,
                           const double* mass2d, int ncomp, int64_t np,
                           const int32_t* cell_start, double* out, int64_t out_colstride);
/* Tuning knob (benchmarks only): output tile extents of the sorted paint kernel in cells;
 * 0 keeps the default. */
int vpm_paint_set_tile(int tx, int ty, int tz);
/* Same contract, earlier "row gather" formulation (one warp per output row, lanes over cells);
 * kept for ablation against the tile kernel. */
int vpm_paint_sorted_rows(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs,
                          const double* mass, double mass0, int64_t np,
                          const int32_t* cell_start, double* out);
/* sum_d mass * vpos[:, d] * dW/dx_d  (+ vmass * W when vmass != NULL) */
int vpm_paint_jvp_sorted(vpm_ctx* ctx, const vpm_mesh* mesh, const double* xs,
                         const double* mass, double mass0, const double* vpos,
                         const double* vmass, int64_t np, const int32_t* cell_start,
                         double* out);
/* Baseline scatter for unsorted particles: zero-fill + 8 RED.F64 per particle. */
int vpm_paint_atomic(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x,
                     const double* mass, double mass0, int64_t np, double* out);

/* ---- readout  (fastpm.py:229,252,256,263; pmesh RealField.readout / readout_vjp /
 *      readout_jvp and the particle half of ParticleMesh.paint_vjp) ------------ */
/* value[p] = sum_c field[c] W(x_p - c) */
int vpm_readout(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field, const double* x,
                int64_t np, double* value);
/* three fields at once, written interleaved [np, 3] (fuses linalg.stack,
 * fastpm.py:349,437).  out_row (may be NULL): row p is stored at row out_row[p] -- Layout.gather
 * of a one-rank layout (its slot -> home row permutation) fused into the store.  y_out (may be
 * NULL): y_out[row] = y_in[row] + fac * value[row] at the rows written -- the kick that consumes
 * the force (fastpm.py:470-471); value3 may then be NULL when the force itself is not needed. */
int vpm_readout3(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                 const double* f2, const double* x, int64_t np, const int32_t* out_row,
                 double* value3, const double* y_in, double fac, double* y_out);
/* The same readout under a layout that spans several ranks (Layout.gather, fastpm.py:286-304, split
 * into a local store and a cross-rank sum): value3_slots[p] = value where home_row[p] < 0 (slot
 * order: the rows whose home is on another rank travel from there; other slots are not written)
 * and, where home_row[p] >= 0 (the particle's home row
 * is on this rank; negative = elsewhere or an empty slot), acc_out[home_row[p]] =
 * acc_in[home_row[p]] (0 when acc_in is NULL) + scale * value. */
int vpm_readout3_home(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                      const double* f2, const double* x, int64_t np, const int32_t* home_row,
                      double* value3_slots, const double* acc_in, double scale, double* acc_out);
/* value[p] (optional, may be NULL) and grad[p, d] = w_p * sum_c field[c] dW/dx_d,
 * w_p = weight[p] or weight0 when weight == NULL. */
int vpm_readout_grad(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field,
                     const double* x, const double* weight, double weight0, int64_t np,
                     double* value, double* grad);
/* grad[p, :] = sum_d w3[p, d] * gradreadout(f_d, x_p) + gradreadout(r, x_p): the particle half
 * of the reverse sweep of the fused force operator (readout.vjp x3 + paint.vjp, fastpm.py:229,
 * 256) in one pass over the particles; out_row as in vpm_readout3; base (may be NULL): an
 * [np, 3] array added row by row at the rows written (the cotangent reaching dx from its other
 * consumers -- saves the separate accumulation pass); y_out (may be NULL): y_out[row] =
 * y_in[row] + fac * grad[row] (the momentum cotangent of the drift, fastpm.py:463-464). */
int vpm_readout_grad4(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                      const double* f2, const double* r, const double* x, const double* w3,
                      int64_t np, const int32_t* out_row, const double* base, double* grad,
                      const double* y_in, double fac, double* y_out);
/* vpm_readout_grad4 under a layout that spans several ranks: grad_slots[p] = the gradient (slot
 * order, no base; only where home_row[p] < 0) and acc_out[home_row[p]] = acc_in[home_row[p]] (0 when NULL) + gradient where
 * home_row[p] >= 0; see vpm_readout3_home. */
int vpm_readout_grad4_home(vpm_ctx* ctx, const vpm_mesh* mesh, const double* f0, const double* f1,
                           const double* f2, const double* r, const double* x, const double* w3,
                           int64_t np, const int32_t* home_row, double* grad_slots,
                           const double* acc_in, double* acc_out);
/* Implementation variant of vpm_readout3 / vpm_readout_grad4 (same results bit for bit): bits
 * 1 positions staged per CTA by a bulk asynchronous copy, 2 scattered rows requested before the
 * mesh, 4 scattered rows as 16 + 8 byte accesses, 8 / 16 pair-load mode, 32 streaming hints on the
 * scattered rows; | (minimum CTAs per SM << 6) (csrc/readout.cu).  Used by
 * the ablation in benchmarks/readout_variants.py; the default is the measured best. */
int vpm_readout_set_variant(int variant);
/* value_[p] = readout(field_dot, x) + sum_d xdot[p, d] * gradreadout_d(field, x);
 * either term may be absent (NULL). */
int vpm_readout_jvp(vpm_ctx* ctx, const vpm_mesh* mesh, const double* field,
                    const double* field_dot, const double* x, const double* xdot,
                    int64_t np, double* value);

/* ---- FFT  (fastpm.py:50,53,56,64,67,70; pmesh RealField.r2c / c2r_vjp,
 *      ComplexField.c2r / r2c_vjp).  cuFFT does the transform; `scale` is applied
 *      by a fused pass (1.0 skips it).  n = real mesh shape (n[0] = 1 for 2-D). --- */
int vpm_r2c(vpm_ctx* ctx, const int64_t n[3], const double* real_in, double* cplx_out,
            double scale);
/* preserves cplx_in (copies into the context workspace first) */
int vpm_c2r(vpm_ctx* ctx, const int64_t n[3], const double* cplx_in, double* real_out,
            double scale);

/* same transform, but the complex input is a scratch buffer the caller gives up (cuFFT may
 * overwrite it): saves the staging copy for temporaries such as the force gradient fields */
int vpm_c2r_inplace(vpm_ctx* ctx, const int64_t n[3], double* cplx_scratch, double* real_out,
                    double scale);

/* Pieces of the slab-decomposed transform (the local transforms are cuFFT; the transpose
 * between them is an all-to-all done by the caller):
 *   r2c: vpm_fft_r2c_planes -> vpm_swap_leading_axes (pack) -> all-to-all -> vpm_fft_c2c_axis0
 *   c2r: vpm_fft_c2c_axis0 (inverse) -> all-to-all -> vpm_swap_leading_axes (unpack)
 *        -> vpm_fft_c2r_planes
 * Local real slab [nplanes, n1, n2]; local complex slab [n0, n1 / P, n2/2 + 1] ("transposed
 * out": the complex mesh is sharded along axis 1). */
int vpm_fft_r2c_planes(vpm_ctx* ctx, int64_t nplanes, int64_t n1, int64_t n2, const double* real_in,
                       double* cplx_out);
/* destroys cplx_in */
int vpm_fft_c2r_planes(vpm_ctx* ctx, int64_t nplanes, int64_t n1, int64_t n2, double* cplx_in,
                       double* real_out, double scale);
/* in place, `inner` transforms of length n0 with stride `inner` */
int vpm_fft_c2c_axis0(vpm_ctx* ctx, int64_t n0, int64_t inner, double* data, int inverse,
                      double scale);
/* out-of-place variant (the input may be a peer-exchange receive buffer) */
int vpm_fft_c2c_axis0_oop(vpm_ctx* ctx, int64_t n0, int64_t inner, const double* in, double* out,
                          int inverse, double scale);
/* Peer-memory transpose (replaces the MPI_Alltoall inside pfft): the pack kernel stores block d
 * of the local array straight into rank d's receive buffer over NVLink,
 *   peer_bufs[d][dst_off + i * inner + c] = src[i * s_row + d * s_dst + c]   (complex128),
 * peer_bufs / peer_flags are device pointers valid in THIS process for every rank's receive
 * buffer / flag block (CUDA IPC mappings; entry [me] is the local one). */
/* receive areas live outside the caller's allocator so that they can be exported: alloc returns
 * a zeroed device buffer and its 64-byte CUDA IPC handle; open maps another process's buffer. */
int vpm_peer_alloc(vpm_ctx* ctx, int64_t nbytes, void** ptr, unsigned char* handle64);
int vpm_peer_open(vpm_ctx* ctx, const unsigned char* handle64, void** ptr);
int vpm_peer_close(vpm_ctx* ctx, void* ptr);
int vpm_peer_free(vpm_ctx* ctx, void* ptr);
/* Flag protocol carried by a put launch (NULL or mode 0: none).  Every rank owns a flag block
 * of VPM_PEER_FLAG_WORDS int64 (vpm_peer_alloc, zero-initialised, mapped by every peer):
 * data[b][s] ("rank s delivered its transfer number e on receive buffer b"), ack[b][s] ("rank s
 * finished reading what transfer e put into its buffer b"), a four-word mailbox per (b, s) and the
 * rank's own protocol state (transfers completed per buffer, buffer of the last transfer).  The
 * transfer numbering lives in that state ON THE DEVICE -- the host only alternates `buf` -- so a
 * CUDA graph holding put launches can be replayed.  peer_flags[d] = rank d's block as mapped
 * here, my_flags = my own.  mode bits for a transfer on receive buffer `buf`:
 *   RAISE_ACK  at kernel start acknowledge my previous transfer to every rank (its consumer
 *              precedes this launch on the stream);
 *   WAIT_ACK   before storing wait until every rank acknowledged the previous transfer that used
 *              buffer `buf`;
 *   SIGNAL     after the stores the last thread block raises data[buf][rank] on every rank;
 *   WAIT_DATA  ... and then waits until all of my data flags arrived, so that the kernel's
 *              completion means "my receive area is complete".
 * A wait that exceeds the spin limit (default 2^31 naps of 64 ns; vpm_peer_set_spin_limit)
 * prints a message and traps: a dead peer produces a launch failure, not a hung device. */
#define VPM_PEER_RAISE_ACK 1
#define VPM_PEER_WAIT_ACK 2
#define VPM_PEER_SIGNAL 4
#define VPM_PEER_WAIT_DATA 8
#define VPM_PEER_FLAG_WORDS 128
typedef struct vpm_peer_sync {
    const uint64_t* peer_flags;
    int64_t* my_flags;
    int32_t rank;
    int32_t mode;
    int32_t buf;
    int32_t reserved;
} vpm_peer_sync;
int vpm_peer_put(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                 int64_t s_dst, const double* src, const uint64_t* peer_bufs, int64_t dst_off,
                 const vpm_peer_sync* sync);
/* the same with a destination row stride: peer_bufs[d][dst_off + i * d_row + c] = src[i * s_row +
 * d * s_dst + c] -- the transpose of the inverse slab FFT lands in [planes, N1, N2/2+1] order in the
 * receive area, so the batched 2-D Z2D reads it in place (no unpack pass). */
int vpm_peer_put_strided(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                         int64_t s_dst, int64_t d_row, const double* src, const uint64_t* peer_bufs,
                         int64_t dst_off, const vpm_peer_sync* sync);
/* `count` contiguous complex128 elements (= 2 count doubles) to every rank whose bit is set in
 * dest_mask, stored at element dst_off of its receive area (halo planes for the real-space
 * stencils: the two boundary planes of a slab go to the neighbouring rank only); sync as above. */
int vpm_peer_put_to(vpm_ctx* ctx, int nranks, uint32_t dest_mask, int64_t count, const double* src,
                    const uint64_t* peer_bufs, int64_t dst_off, const vpm_peer_sync* sync);
/* nblocks (<= 8) contiguous blocks of `count` complex128 elements in ONE launch: block k goes from
 * src + src_off[k] (elements) to element dst_off[k] of rank dest[k]'s receive area (host arrays).
 * The two boundary-plane pairs of a slab travel to the previous and the next rank this way. */
int vpm_peer_put_blocks(vpm_ctx* ctx, int nranks, int nblocks, int64_t count, const double* src,
                        const int32_t* dest, const int64_t* src_off, const int64_t* dst_off,
                        const uint64_t* peer_bufs, const vpm_peer_sync* sync);
/* out_a[0 .. count) = a, out_b[0 .. count) = b (complex128 elements) in one launch */
int vpm_copy2(vpm_ctx* ctx, int64_t count, const double* a, double* out_a, const double* b, double* out_b);
/* Particle rows (Layout.exchange / Layout.gather across ranks; replaces pmesh's MPI_Alltoallv):
 * pack and transfer in one kernel.  Position p of the routing order lies in segment d when
 * seg_start[d] <= p < seg_start[d+1] (host array, nranks+1) and is stored to row
 * dst_base[d] + p - seg_start[d] of peer d's receive area (ncomp doubles per row).
 * idx_is_src=1: position p, source row idx[p] (exchange); idx_is_src=0: p = idx[j], source row j
 * (gather, idx = the local sort permutation); idx_is_src=2: position p, source row idx[t] with
 * idx = the inverse permutation over the positions outside skip_seg (gather; vpm_route_compose's
 * inv_remote).  Rows of segment skip_seg (-1: none) are not transferred -- rows staying on this
 * rank bypass the receive area (vpm_route_compose) -- and in modes 1 and 2 not even visited.
 * Fixed-capacity segments (vpm_route_build_static; the counts stay on the device):
 *   seg_count  (device, nranks ints or NULL) gather direction: only the first seg_count[d] rows
 *              of segment d hold particles;
 *   mail       (device, nranks ints or NULL) posted to the peers' mailboxes with this transfer
 *              (the per-destination row counts of the routing; vpm_peer_read_mail);
 *   sentinel_rows (host, nranks x 3 doubles or NULL) exchange direction: a position without a
 *              particle (idx[j] < 0) in the segment of rank d is stored as row d of this array
 *              -- a position that lies on no plane rank d holds -- instead of being skipped. */
int vpm_peer_put_rows(vpm_ctx* ctx, int nranks, int64_t n, int ncomp, const double* src,
                      const int32_t* idx, int idx_is_src, int skip_seg, const int32_t* seg_start,
                      const int64_t* dst_base, const int32_t* seg_count, const int32_t* mail,
                      const double* sentinel_rows, const uint64_t* peer_bufs, const vpm_peer_sync* sync);
/* counts[s] = what rank s posted with the last transfer on receive buffer `buf` */
int vpm_peer_read_mail(vpm_ctx* ctx, int nranks, const int64_t* my_flags, int buf, int32_t* counts);
int vpm_peer_set_spin_limit(vpm_ctx* ctx, uint64_t spins);
/* dst[b][a][:] = src[a][b][:] for [a, b, c] complex128 */
int vpm_swap_leading_axes(vpm_ctx* ctx, int64_t a, int64_t b, int64_t c, const double* src,
                          double* dst);

/* ---- k-space  (fastpm.py:79,83,87 apply_transfer; :310-334 filters; pmesh
 *      ComplexField.apply) ---------------------------------------------------- */
/* out = in * scale * L * a0[i] * a1[j] * a2[k]
 *   L      = -1 / (kk0[i]^2 + kk1[j]^2 + kk2[k]^2)  (0 where the sum is 0) when
 *            laplace != 0, else 1                      (fourier_space_laplace)
 *   a_d    = complex table of length n_d (nh for d = 2) or NULL (= 1); conj_tables
 *            conjugates them (apply_transfer.vjp, fastpm.py:81-83)
 * nc = complex array shape [n0, n1, nh]. */
int vpm_transfer(vpm_ctx* ctx, const int64_t nc[3], const double* in, double* out,
                 double scale, int laplace, const double* kk0, const double* kk1,
                 const double* kk2, const double* a0, const double* a1, const double* a2,
                 int conj_tables);
/* out = in * table (general, broadcast strides in complex elements; table real or
 * complex; conj optional) -- arbitrary tf(k) evaluated once on the host. */
int vpm_transfer_table(vpm_ctx* ctx, const int64_t nc[3], const double* in, double* out,
                       const double* table, const int64_t tstride[3], int table_is_complex,
                       int conj_table);
/* Fused force pass of `gravity` (fastpm.py:424-431): from rho_k (un-normalised cuFFT
 * output allowed via `scale`) produce p = scale * L * rho_k (optional, may be NULL) and
 * the three gradient fields g_d = p * a_d[.] in one read of rho_k. */
int vpm_force_kspace(vpm_ctx* ctx, const int64_t nc[3], const double* rhok, double scale,
                     const double* kk0, const double* kk1, const double* kk2,
                     const double* a0, const double* a1, const double* a2,
                     double* p_out, double* g0, double* g1, double* g2);

/* adjoint of vpm_force_kspace: rhok_bar = scale * L * (sum_d conj(a_d) g_d_bar + p_bar);
 * p_bar may be NULL. */
int vpm_force_kspace_vjp(vpm_ctx* ctx, const int64_t nc[3], const double* g0, const double* g1,
                         const double* g2, const double* p_bar, double scale, const double* kk0,
                         const double* kk1, const double* kk2, const double* a0, const double* a1,
                         const double* a2, double* rhok_bar);

/* implementation of the two stencils: 0 scalar (any shape), 1 four cells per thread, 2 (default)
 * gradient marching along axis 0 with a register window and a shared-memory tile + four-cell
 * adjoint, 3 marching for both; falls back by shape / alignment; all produce identical results */
int vpm_fd4_set_algorithm(int algo);
/* Real-space form of fourier_space_neg_gradient(order=1) (fastpm.py:316-323): its multiplier is
 * the Fourier symbol of minus the 4-point central difference, so the three force meshes of
 * `gravity` are f_d = -D_d phi, phi = c2r(p): one inverse FFT + this stencil instead of three
 * inverse FFTs.  n: real mesh shape (single slab), h: cell sizes BoxSize / Nmesh. */
int vpm_fd4_neg_gradient(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* phi,
                         double* f0, double* f1, double* f2);
/* adjoint: phibar = sum_d D_d fbar_d */
int vpm_fd4_neg_gradient_adjoint(vpm_ctx* ctx, const int64_t n[3], const double h[3],
                                 const double* f0, const double* f1, const double* f2,
                                 double* phibar);
/* the same two stencils on a slab: the inputs carry `pad` (>= 2) halo planes on each side of axis 0
 * (n[0] + 2 pad planes; delivered by the neighbouring slabs), n[0] = local planes written; only f0
 * needs valid halo planes in the adjoint. */
int vpm_fd4_neg_gradient_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad,
                              const double* phi_padded, double* f0, double* f1, double* f2);
int vpm_fd4_neg_gradient_adjoint_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad,
                                      const double* f0_padded, const double* f1_padded,
                                      const double* f2_padded, double* phibar);
/* Second-order LPT source (fastpm.py:353-387) in real space.  With f_j = -D_j phi (the output of
 * vpm_fd4_neg_gradient) the second derivatives are P_ij = -D_i f_j and
 *   source = 3/7 (P11 P22 + P22 P00 + P00 P11 - P12^2 - P20^2 - P01^2) = 1/2 B(f, f);
 * out = scale * B(f, g), B the symmetric bilinear form dsource(f)[g]; g0 = g1 = g2 = NULL means
 * g = f.  Value: g = NULL, scale = 0.5; tangent: g = f_, scale = 1.  Replaces six inverse FFTs
 * and thirteen transfers of the reference graph. */
int vpm_lpt2_source(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* f0,
                    const double* f1, const double* f2, const double* g0, const double* g1,
                    const double* g2, double scale, double* out);
/* its adjoint: cotangent sbar of the source -> cotangents fb_j of f_j (work6: 6 meshes of scratch) */
int vpm_lpt2_source_adjoint(vpm_ctx* ctx, const int64_t n[3], const double h[3], const double* f0,
                            const double* f1, const double* f2, const double* sbar, double* work6,
                            double* fb0, double* fb1, double* fb2);
/* Slab forms (several ranks): every mesh argument except out / sbar has `pad` (>= 2) halo planes on
 * each side of axis 0 which the caller fills from the neighbouring slabs; n = the local cells
 * written; the adjoint comes in two halves because W_00 and W_01 need their halo planes filled in
 * between (the axis-0 difference of the second half).  w6_padded: [6][n0 + 2 pad][n1][n2]; fb0..2
 * padded alike (interiors written). */
int vpm_lpt2_source_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad, const double* f0,
                         const double* f1, const double* f2, const double* g0, const double* g1,
                         const double* g2, double scale, double* out);
int vpm_lpt2_adjoint_w_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad, const double* f0,
                            const double* f1, const double* f2, const double* sbar, double* w6_padded);
int vpm_lpt2_adjoint_f_slab(vpm_ctx* ctx, const int64_t n[3], const double h[3], int pad,
                            const double* w6_padded, double* fb0, double* fb1, double* fb2);

/* Binned transfer functions (fastpm.py:97-213 apply_digitized; pmesh ComplexField.apply with a
 * digitizer).  ind[m] = bin of stored mode m (int32, evaluated once on the host by the user's
 * digitizer), conj_mask[m] != 0: use the conjugate of the bin's entry (NULL: never).
 *   y[m] = x[m] * T(m),  T(m) = table[ind[m]] (conjugated per mask, and once more when conj_table)
 *                               for 0 <= ind[m] < nbins, else dflt  (1: leave the mode alone) */
int vpm_digitized_mul(vpm_ctx* ctx, int64_t n, const double* x, const int32_t* ind, const uint8_t* conj_mask,
                      int nbins, const double* table, double dflt, int conj_table, double* y);
/* vjp with respect to the per-bin table (fastpm.py:170-196): out[b] = sum over the modes of bin b
 * of mult(m) Re(ybar[m] conj(v[m])), v = x (eit NULL: amplitude mode) or i x E(m), E the bin's
 * phase factor eit[b] conjugated per mask (phase mode); mult = 2 for the stored modes that stand
 * for two modes of the full spectrum (last-axis index not 0 and not n_last / 2).  out (nbins
 * doubles, device) is overwritten; local modes only -- the caller all-reduces across ranks. */
int vpm_digitized_bin(vpm_ctx* ctx, const int64_t nc[3], int64_t n_last, const double* x, const double* ybar,
                      const int32_t* ind, const uint8_t* conj_mask, int nbins, const double* eit, double* out);
/* ---- elementwise and reductions on flat fp64 arrays (the stdlib arithmetic that
 *      vmad applies to fields / particle arrays: vmad/core/stdlib/operators.py:14-152;
 *      kick / drift of fastpm.py:459-471) ------------------------------------- */
/* out = a * x + b * y + c   (y may be NULL) */
int vpm_axpby(vpm_ctx* ctx, int64_t n, double a, const double* x, double b,
              const double* y, double c, double* out);
/* dx1 = dx + fac * p  and  x1 = q + dx1 in one pass (drift + the position update of the next force
 * evaluation, fastpm.py:411,463-464) */
int vpm_drift_pos(vpm_ctx* ctx, int64_t n, const double* dx, const double* p, double fac,
                  const double* q, double* dx1, double* x1);
/* out = x * y */
int vpm_mul(vpm_ctx* ctx, int64_t n, const double* x, const double* y, double* out);
/* out = x / y */
int vpm_div(vpm_ctx* ctx, int64_t n, const double* x, const double* y, double* out);
/* out[i, :] = x[i, :] * w[i]   (row scaling of an [n, ncomp] array) */
int vpm_mul_rows(vpm_ctx* ctx, int64_t n, int ncomp, const double* x, const double* w,
                 double* out);
/* complex: out = x * y or x * conj(y) */
int vpm_cmul(vpm_ctx* ctx, int64_t n, const double* x, const double* y, int conj_y,
             double* out);
int vpm_fill(vpm_ctx* ctx, int64_t n, double value, double* out);
/* [n] x ncomp separate arrays -> interleaved [n, ncomp]  (linalg.stack axis=-1) */
int vpm_stack(vpm_ctx* ctx, int64_t n, int ncomp, const double* const* cols_host,
              double* out);
/* out[i] = x[i, col] of an [n, ncomp] array  (numpy.take(..., axis=-1)) */
int vpm_take_col(vpm_ctx* ctx, int64_t n, int ncomp, int col, const double* x, double* out);

/* deterministic two-stage reductions; result delivered to the host (synchronises) */
int vpm_reduce_sum(vpm_ctx* ctx, int64_t n, const double* x, double* out_host);
int vpm_reduce_dot(vpm_ctx* ctx, int64_t n, const double* x, const double* y,
                   double* out_host);
/* sum over stored modes of w_k * x conj(y): out_host[0] = real, [1] = imag.
 * w_k = 2 except on the planes k = 0 and (n_last even) k = n_last/2 of the half axis
 * (pmesh ComplexField.cnorm / cdot, fastpm.py:15,519). nc = [n0, n1, nh]. */
int vpm_reduce_cdot(vpm_ctx* ctx, const int64_t nc[3], int64_t n_last, const double* x,
                    const double* y, double* out_host);

#ifdef __cplusplus
}
#endif
#endif /* VMADPM_H */
