// Peer-memory exchange for the slab-FFT transpose: the pack kernel stores each destination
// block straight into the peer GPU's receive buffer over NVLink (P2P stores through CUDA IPC
// mappings), so packing and transfer are one kernel and no NCCL send/recv kernel or host-side
// collective call sits between the local FFT stages.  Ranks are synchronised with flags in
// peer memory: after its puts a rank raises data[me] = epoch on every peer; a consumer spins
// (one tiny kernel on its own GPU -- the producer runs on a different GPU) until all of its
// data flags reach the epoch; consumed buffers are acknowledged the same way so that the
// double-buffered receive areas can be reused.  Replaces pfft's MPI_Alltoall transposes.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace vpm {

struct PeerPtrs {
    unsigned long long p[8];
};

// The whole flag protocol rides on the put launches (one launch per transfer instead of four)
// and its state lives ON THE DEVICE, so that a captured CUDA graph can be replayed: a rank's
// flag block (VPM_PEER_FLAG_WORDS int64, allocated by vpm_peer_alloc, mapped by every peer) holds
//   data[b][s]  "rank s delivered its transfer number e on receive buffer b"
//   ack[b][s]   "rank s finished reading what transfer e put into ITS buffer b"
//   mail[b][s]  four words rank s may post together with a transfer (per-destination row counts)
//   eb[b], last_b   my own state: transfers completed on buffer b, buffer of my last transfer
// For a transfer on buffer b (the host alternates b; the numbering e = eb[b] + 1 is read from the
// device, in lock step on all ranks because all ranks issue the same sequence of transfers):
//   start  - CTA 0 acknowledges my previous transfer (buffer last_b, number eb[last_b]) to every
//            peer: stream order guarantees that its consumer has finished reading my receive area;
//          - every CTA makes sure all peers acknowledged transfer e - 1 of buffer b, the previous
//            use of the receive area about to be overwritten (double buffering);
//   end    - the last CTA to finish (device counter) raises data[b][me] = e on every peer, waits
//            until all of MY data flags reached e, then advances eb[b] and last_b: when the kernel
//            completes, the receive area is complete, and the next kernel of the stream (cuFFT,
//            the local gather) can read it.
// Every put raises its acknowledgement BEFORE it waits for anything, and waits only for
// acknowledgements that the peers' next put raises unconditionally, so the scheme cannot deadlock
// (also when a replayed graph starts on the buffer the previous replay ended on).
constexpr int W_DATA = 0, W_ACK = 16, W_MAIL = 32, W_EB = 96, W_LASTB = 98, W_COUNTER = 99, W_ERROR = 100;
static_assert(W_ERROR < VPM_PEER_FLAG_WORDS, "flag block too small");

struct PeerSync {
    PeerPtrs flags;               // every rank's flag block as mapped here
    long long* mine;              // my own flag block
    int rank;
    int mode;                     // VPM_PEER_* bits
    int buf;                      // receive buffer of this transfer (0 / 1)
};

__device__ unsigned long long g_peer_spin_limit = 1ull << 31;   // ~2 minutes of 64 ns naps

__device__ __forceinline__ void peer_spin(const volatile long long* f, long long want, const char* what) {
    unsigned long long spins = 0;
    while (*f < want) {
        if (++spins > g_peer_spin_limit) {
            printf("vmadpm: peer %s wait timed out (flag = %lld, want %lld)\n", what, *f, want);
            __trap();
        }
        __nanosleep(64);
    }
}

// number of the transfer this launch belongs to (my previous launches on the stream wrote eb[])
__device__ __forceinline__ long long peer_epoch(const PeerSync& sy) { return sy.mine[W_EB + sy.buf] + 1; }

__device__ __forceinline__ void peer_acquire(int P, const PeerSync& sy) {
    if (!sy.mode) return;
    const long long e = peer_epoch(sy);
    if ((sy.mode & VPM_PEER_RAISE_ACK) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 &&
        (int)threadIdx.x < P) {
        const int pb = (int)sy.mine[W_LASTB];
        const long long pe = sy.mine[W_EB + pb];
        if (pe > 0) {
            volatile long long* f = reinterpret_cast<volatile long long*>(sy.flags.p[threadIdx.x]);
            f[W_ACK + 8 * pb + sy.rank] = pe;
        }
    }
    if (sy.mode & VPM_PEER_WAIT_ACK) {
        if ((int)threadIdx.x < P) peer_spin(sy.mine + W_ACK + 8 * sy.buf + threadIdx.x, e - 1, "ack");
        __syncthreads();
    }
}

__device__ __forceinline__ void peer_publish(int P, const PeerSync& sy) {
    __threadfence_system();
    if (!(sy.mode & VPM_PEER_SIGNAL)) return;
    const long long e = peer_epoch(sy);     // read before the last CTA advances it
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* counter = reinterpret_cast<unsigned*>(sy.mine + W_COUNTER);
        const unsigned done = atomicAdd(counter, 1u);
        if (done == gridDim.x * gridDim.y * gridDim.z - 1) {
            *counter = 0;
            __threadfence_system();
            for (int d = 0; d < P; ++d) {
                volatile long long* f = reinterpret_cast<volatile long long*>(sy.flags.p[d]);
                f[W_DATA + 8 * sy.buf + sy.rank] = e;
            }
            __threadfence_system();
            if (sy.mode & VPM_PEER_WAIT_DATA) {
                for (int d = 0; d < P; ++d) peer_spin(sy.mine + W_DATA + 8 * sy.buf + d, e, "data");
                __threadfence_system();
            }
            sy.mine[W_EB + sy.buf] = e;
            sy.mine[W_LASTB] = sy.buf;
        }
    }
}

// dst_d[dst_off + i * d_row + c] = src[i * s_row + d * s_dst + c]   for every peer d
// grid = (chunks of a row, rows, peers): no index divisions, 16-byte coalesced loads / stores
__global__ void __launch_bounds__(256)
k_peer_put(int P, long long rows, long long inner, long long s_row, long long s_dst, long long d_row,
           const double2* __restrict__ src, PeerPtrs dst, long long dst_off, unsigned mask,
           PeerSync sy, int rot) {
    peer_acquire(P, sy);
    // destinations in rotated order (rank + 1 first): when the grid does not fit in one wave, the
    // ranks do not all start on the same receiver
    const int d = (int)((blockIdx.z + rot) % P);
    double2* __restrict__ out = reinterpret_cast<double2*>(dst.p[d]) + dst_off;
    // destinations outside the mask get no data; their CTAs still take part in the flag protocol
    const long long nrows = ((mask >> d) & 1u) ? rows : 0;
    for (long long i = blockIdx.y; i < nrows; i += gridDim.y) {
        const double2* __restrict__ in = src + i * s_row + d * s_dst;
        double2* __restrict__ o = out + i * d_row;
        // four independent 16-B loads in flight per thread before the (posted) stores
        const long long stride = (long long)gridDim.x * blockDim.x;
        long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        for (; c + 3 * stride < inner; c += 4 * stride) {
            const double2 v0 = in[c], v1 = in[c + stride], v2 = in[c + 2 * stride], v3 = in[c + 3 * stride];
            o[c] = v0; o[c + stride] = v1; o[c + 2 * stride] = v2; o[c + 3 * stride] = v3;
        }
        for (; c < inner; c += stride) o[c] = in[c];
    }
    peer_publish(P, sy);
}

// up to 8 contiguous blocks of `count` complex elements, block k going from src + src_off[k] to
// element dst_off[k] of rank dest[k]'s receive area: the halo planes of a slab travel to both
// neighbours in ONE launch (and one pass of the flag protocol)
struct PeerBlocks {
    int dest[8];
    long long src_off[8], dst_off[8];
};

__global__ void __launch_bounds__(256)
k_peer_put_blocks(int P, int nblocks, long long count, const double2* __restrict__ src, PeerBlocks blk,
                  PeerPtrs dst, PeerSync sy) {
    peer_acquire(P, sy);
    const int k = blockIdx.y;
    if (k < nblocks) {
        const double2* __restrict__ in = src + blk.src_off[k];
        double2* __restrict__ out = reinterpret_cast<double2*>(dst.p[blk.dest[k]]) + blk.dst_off[k];
        for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < count;
             c += (long long)gridDim.x * blockDim.x)
            out[c] = in[c];
    }
    peer_publish(P, sy);
}

// two copies of `count` complex elements in one launch (the received halo planes into the padded array)
__global__ void __launch_bounds__(256)
k_copy2(long long count, const double2* __restrict__ a, double2* __restrict__ oa,
        const double2* __restrict__ b, double2* __restrict__ ob) {
    const double2* __restrict__ in = blockIdx.y ? b : a;
    double2* __restrict__ out = blockIdx.y ? ob : oa;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < count;
         c += (long long)gridDim.x * blockDim.x)
        out[c] = in[c];
}

struct PeerSegs {
    int start[9];            // segment boundaries in the routing order (P + 1 used)
    long long base[8];       // first destination row of my segment in peer d's receive area
    const int* cnt;          // fixed-capacity segments, gather direction: only rows [start[d], start[d] +
                             // cnt[d]) of segment d hold particles (NULL: every row does)
    const int* mail;         // my per-destination row counts, posted to peer d's mailbox (NULL: none)
    int fill;                // exchange direction: positions without a particle (idx < 0) are stored as
    double sentinel[8][3];   // the destination's row (a position off ITS slab); else they are skipped
};

// Particle rows to their destination ranks, pack and transfer fused.  Routing order position p
// (a row of the send order / of the receive order) belongs to segment d when
// start[d] <= p < start[d + 1] and lands in row base[d] + p - start[d] of peer d.
//   mode 1 (exchange):            position p,            source row = idx[p]
//   mode 0 (gather):              position p = idx[j],   source row = j        (j over all rows)
//   mode 2 (gather, by position): position p,            source row = idx[t]   (idx = the inverse
//                                 of the sort permutation over the positions that are not skipped)
// In modes 1 and 2 the loop runs over the positions OUTSIDE the skipped segment only (t -> p jumps
// over it): rows that stay on their rank cost nothing here.
template <int NC>
__global__ void __launch_bounds__(256)
k_peer_put_rows(int P, int n, int ncomp, const double* __restrict__ src,
                const int* __restrict__ idx, int mode, int skip_seg, PeerSegs seg,
                PeerPtrs dst, PeerSync sy) {
    peer_acquire(P, sy);
    if (seg.mail && blockIdx.x == 0 && (int)threadIdx.x < P) {
        // (made visible by the fences of peer_publish before the data flag is raised)
        volatile long long* f = reinterpret_cast<volatile long long*>(sy.flags.p[threadIdx.x]);
        f[W_MAIL + 4 * (8 * sy.buf + sy.rank)] = seg.mail[threadIdx.x];
    }
    const int nc = NC > 0 ? NC : ncomp;
    const bool jump = mode != 0 && skip_seg >= 0;
    const int skip_lo = jump ? seg.start[skip_seg] : 0;
    const int skip_len = jump ? seg.start[skip_seg + 1] - seg.start[skip_seg] : 0;
    const long long total = (long long)(n - skip_len) * nc;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(e / nc);
        const int c = (int)(e - (long long)t * nc);
        int p, row;
        if (mode == 0) {
            p = idx[t];
            row = t;
        } else {
            p = t < skip_lo ? t : t + skip_len;
            row = mode == 1 ? idx[p] : idx[t];
        }
        if (p < 0 || p >= seg.start[P]) continue;
        int d = 0;
#pragma unroll
        for (int u = 1; u < 8; ++u) d += (u < P && p >= seg.start[u]) ? 1 : 0;
        if (d == skip_seg) continue;
        const int off = p - seg.start[d];
        if (seg.cnt && off >= seg.cnt[d]) continue;
        double v;
        if (row >= 0) v = src[(long long)row * nc + c];
        else if (seg.fill) v = seg.sentinel[d][c < 3 ? c : 2];
        else continue;
        double* out = reinterpret_cast<double*>(dst.p[d]);
        out[(seg.base[d] + off) * nc + c] = v;
    }
    peer_publish(P, sy);
}

// the counts the peers posted with transfer (buffer b) -> out[0 .. P): rows received from rank s
__global__ void k_peer_read_mail(int P, const long long* mine, int buf, int* __restrict__ out) {
    const int s = threadIdx.x;
    if (s < P) out[s] = (int)reinterpret_cast<const volatile long long*>(mine)[W_MAIL + 4 * (8 * buf + s)];
}

}  // namespace vpm

using namespace vpm;

static PeerPtrs make_ptrs(int P, const uint64_t* p) {
    PeerPtrs r;
    for (int i = 0; i < 8; ++i) r.p[i] = i < P ? p[i] : 0;
    return r;
}

static bool put_rotate() {
    static const bool on = [] { const char* e = getenv("VMAD_B200_PUT_ROTATE"); return !(e && e[0] == '0'); }();
    return on;
}

static PeerSync make_sync(int P, const vpm_peer_sync* h) {
    PeerSync sy;
    memset(&sy, 0, sizeof(sy));
    if (h && h->mode) {
        VPM_REQUIRE(h->peer_flags && h->my_flags, "peer sync: flag blocks missing");
        VPM_REQUIRE(!(h->mode & VPM_PEER_WAIT_DATA) || (h->mode & VPM_PEER_SIGNAL),
                    "peer sync: waiting for data implies signalling");
        VPM_REQUIRE(h->buf == 0 || h->buf == 1, "peer sync: buffer must be 0 or 1");
        sy.flags = make_ptrs(P, h->peer_flags);
        sy.mine = reinterpret_cast<long long*>(h->my_flags);
        sy.rank = h->rank;
        sy.mode = h->mode;
        sy.buf = h->buf;
    }
    return sy;
}

extern "C" {

int vpm_peer_alloc(vpm_ctx* ctx, int64_t nbytes, void** ptr, unsigned char* handle64) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && ptr && handle64 && nbytes > 0, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    VPM_CUDA(cudaMalloc(&p, (size_t)nbytes));
    VPM_CUDA(cudaMemset(p, 0, (size_t)nbytes));
    cudaIpcMemHandle_t h;
    VPM_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *ptr = p;
    VPM_API_END
}

int vpm_peer_open(vpm_ctx* ctx, const unsigned char* handle64, void** ptr) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && ptr && handle64, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    VPM_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    VPM_API_END
}

int vpm_peer_close(vpm_ctx* ctx, void* ptr) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx, "bad argument");
    if (ptr) VPM_CUDA(cudaIpcCloseMemHandle(ptr));
    VPM_API_END
}

int vpm_peer_free(vpm_ctx* ctx, void* ptr) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx, "bad argument");
    if (ptr) VPM_CUDA(cudaFree(ptr));
    VPM_API_END
}

static int peer_put_impl(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                         int64_t s_dst, int64_t d_row, const double* src, const uint64_t* peer_bufs,
                         int64_t dst_off, unsigned mask, const vpm_peer_sync* sync);

int vpm_peer_put(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                 int64_t s_dst, const double* src, const uint64_t* peer_bufs, int64_t dst_off,
                 const vpm_peer_sync* sync) {
    return peer_put_impl(ctx, nranks, rows, inner, s_row, s_dst, inner, src, peer_bufs, dst_off, 0xffffffffu, sync);
}

int vpm_peer_put_strided(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                         int64_t s_dst, int64_t d_row, const double* src, const uint64_t* peer_bufs,
                         int64_t dst_off, const vpm_peer_sync* sync) {
    return peer_put_impl(ctx, nranks, rows, inner, s_row, s_dst, d_row, src, peer_bufs, dst_off, 0xffffffffu, sync);
}

int vpm_peer_put_to(vpm_ctx* ctx, int nranks, uint32_t dest_mask, int64_t count, const double* src,
                    const uint64_t* peer_bufs, int64_t dst_off, const vpm_peer_sync* sync) {
    // `count` contiguous complex128 elements to every rank in dest_mask, at dst_off of its area
    return peer_put_impl(ctx, nranks, 1, count, 0, 0, count, src, peer_bufs, dst_off, dest_mask, sync);
}

static int peer_put_impl(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                         int64_t s_dst, int64_t d_row, const double* src, const uint64_t* peer_bufs,
                         int64_t dst_off, unsigned mask, const vpm_peer_sync* sync) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("peer_put", 32.0 * rows * inner * nranks);
    VPM_REQUIRE(ctx && src && peer_bufs && nranks >= 1 && nranks <= 8, "bad argument");
    const long long total = (long long)rows * inner * nranks;
    if (total > 0 || sync) {   // an empty put still carries the flag traffic
        // about 16 CTAs per SM in flight: rows along y (capped), the rest along x
        const long long gy = std::max<long long>(1, std::min<long long>(rows, 1024));
        const long long want = std::max<long long>(1, 8LL * ctx->sm_count / (gy * nranks));
        const long long gx = std::max<long long>(1, std::min<long long>(ceil_div(inner, 256), want));
        dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nranks);
        VPM_LAUNCH(k_peer_put, grid, 256, 0, ctx->stream, nranks, (long long)rows, (long long)inner,
                   (long long)s_row, (long long)s_dst, (long long)d_row, reinterpret_cast<const double2*>(src),
                   make_ptrs(nranks, peer_bufs), (long long)dst_off, mask, make_sync(nranks, sync),
                   (put_rotate() && sync) ? (sync->rank + 1) % nranks : 0);
    }
    VPM_API_END
}

int vpm_peer_put_blocks(vpm_ctx* ctx, int nranks, int nblocks, int64_t count, const double* src,
                        const int32_t* dest, const int64_t* src_off, const int64_t* dst_off,
                        const uint64_t* peer_bufs, const vpm_peer_sync* sync) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("peer_put", 32.0 * count * nblocks);
    VPM_REQUIRE(ctx && src && dest && src_off && dst_off && peer_bufs, "NULL argument");
    VPM_REQUIRE(nranks >= 1 && nranks <= 8 && nblocks >= 1 && nblocks <= 8 && count >= 0, "bad argument");
    VPM_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0, "source must be 16-byte aligned");
    PeerBlocks blk;
    for (int k = 0; k < 8; ++k) {
        blk.dest[k] = k < nblocks ? dest[k] : 0;
        blk.src_off[k] = k < nblocks ? src_off[k] : 0;
        blk.dst_off[k] = k < nblocks ? dst_off[k] : 0;
        VPM_REQUIRE(blk.dest[k] >= 0 && blk.dest[k] < nranks, "bad destination rank");
    }
    const long long gx = std::max<long long>(1, std::min<long long>(ceil_div(count, 256), 4LL * ctx->sm_count / nblocks));
    dim3 grid((unsigned)gx, (unsigned)nblocks);
    VPM_LAUNCH(k_peer_put_blocks, grid, 256, 0, ctx->stream, nranks, nblocks, (long long)count,
               reinterpret_cast<const double2*>(src), blk, make_ptrs(nranks, peer_bufs), make_sync(nranks, sync));
    VPM_API_END
}

int vpm_copy2(vpm_ctx* ctx, int64_t count, const double* a, double* out_a, const double* b, double* out_b) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("copy2", 64.0 * count);
    VPM_REQUIRE(ctx && count >= 0, "bad argument");
    if (count > 0) {
        VPM_REQUIRE(a && out_a && b && out_b, "NULL array");
        const long long gx = std::max<long long>(1, std::min<long long>(ceil_div(count, 256), 2LL * ctx->sm_count));
        dim3 grid((unsigned)gx, 2);
        VPM_LAUNCH(k_copy2, grid, 256, 0, ctx->stream, (long long)count, reinterpret_cast<const double2*>(a),
                   reinterpret_cast<double2*>(out_a), reinterpret_cast<const double2*>(b),
                   reinterpret_cast<double2*>(out_b));
    }
    VPM_API_END
}

int vpm_peer_put_rows(vpm_ctx* ctx, int nranks, int64_t n, int ncomp, const double* src,
                      const int32_t* idx, int idx_is_src, int skip_seg, const int32_t* seg_start,
                      const int64_t* dst_base, const int32_t* seg_count, const int32_t* mail,
                      const double* sentinel_rows, const uint64_t* peer_bufs, const vpm_peer_sync* sync) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("peer_put_rows", (16.0 * ncomp + 4.0) * n);
    VPM_REQUIRE(ctx && seg_start && dst_base && peer_bufs && nranks >= 1 && nranks <= 8,
                "bad argument");
    VPM_REQUIRE(n >= 0 && n < (1LL << 31) && ncomp >= 1, "row count / columns out of range");
    VPM_REQUIRE(!mail || sync, "posting counts needs the flag protocol");
    if (n > 0 || sync) {
        VPM_REQUIRE(n == 0 || (src && idx), "bad argument");
        const PeerSync sy = make_sync(nranks, sync);
        PeerSegs seg;
        for (int i = 0; i < 9; ++i) seg.start[i] = i <= nranks ? seg_start[i] : 0x7fffffff;
        for (int i = 0; i < 8; ++i) seg.base[i] = i < nranks ? dst_base[i] : 0;
        seg.cnt = seg_count;
        seg.mail = mail;
        seg.fill = sentinel_rows ? 1 : 0;
        for (int d = 0; d < 8; ++d)
            for (int i = 0; i < 3; ++i) seg.sentinel[d][i] = (sentinel_rows && d < nranks) ? sentinel_rows[3 * d + i] : 0.0;
        const PeerPtrs dst = make_ptrs(nranks, peer_bufs);
        VPM_REQUIRE(idx_is_src >= 0 && idx_is_src <= 2, "mode must be 0, 1 or 2");
        VPM_REQUIRE(skip_seg < nranks, "bad skip segment");
        long long rows = n;
        if (idx_is_src != 0 && skip_seg >= 0) rows -= seg.start[skip_seg + 1] - seg.start[skip_seg];
        const long long total = std::max<long long>(rows, 0) * ncomp;
        unsigned grid = (unsigned)std::max<long long>(
            1, std::min<long long>(ceil_div(total, 256), 16LL * ctx->sm_count));
        if (ncomp == 1)
            VPM_LAUNCH(k_peer_put_rows<1>, grid, 256, 0, ctx->stream, nranks, (int)n, ncomp, src, idx,
                       idx_is_src, skip_seg, seg, dst, sy);
        else if (ncomp == 3)
            VPM_LAUNCH(k_peer_put_rows<3>, grid, 256, 0, ctx->stream, nranks, (int)n, ncomp, src, idx,
                       idx_is_src, skip_seg, seg, dst, sy);
        else
            VPM_LAUNCH(k_peer_put_rows<0>, grid, 256, 0, ctx->stream, nranks, (int)n, ncomp, src, idx,
                       idx_is_src, skip_seg, seg, dst, sy);
    }
    VPM_API_END
}

int vpm_peer_read_mail(vpm_ctx* ctx, int nranks, const int64_t* my_flags, int buf, int32_t* counts) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && my_flags && counts && nranks >= 1 && nranks <= 8 && (buf == 0 || buf == 1), "bad argument");
    VPM_LAUNCH(k_peer_read_mail, 1, 32, 0, ctx->stream, nranks, reinterpret_cast<const long long*>(my_flags), buf, counts);
    VPM_API_END
}

int vpm_peer_set_spin_limit(vpm_ctx* ctx, uint64_t spins) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && spins > 0, "bad argument");
    unsigned long long v = spins;
    VPM_CUDA(cudaMemcpyToSymbol(g_peer_spin_limit, &v, sizeof(v)));
    VPM_API_END
}

}  // extern "C"
