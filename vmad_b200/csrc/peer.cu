// Peer-memory exchange for the slab-FFT transpose: the pack kernel stores each destination
// block straight into the peer GPU's receive buffer over NVLink (P2P stores through CUDA IPC
// mappings), so packing and transfer are one kernel and no NCCL send/recv kernel or host-side
// collective call sits between the local FFT stages.  Ranks are synchronised with flags in
// peer memory: after its puts a rank raises data[me] = epoch on every peer; a consumer spins
// (one tiny kernel on its own GPU -- the producer runs on a different GPU) until all of its
// data flags reach the epoch; consumed buffers are acknowledged the same way so that the
// double-buffered receive areas can be reused.  Replaces pfft's MPI_Alltoall transposes.
#include "common.cuh"
#include <cstring>

namespace vpm {

struct PeerPtrs {
    unsigned long long p[8];
};

// The whole flag protocol rides on the put launches (one launch per transfer instead of four):
//   start  - CTA 0 raises ack[me] = epoch - 1 on every peer (stream order guarantees that the
//            consumer of the previous transfer has finished reading my receive area);
//          - every CTA makes sure all peers acknowledged transfer epoch - 2, the previous use of
//            the receive area about to be overwritten (double buffering);
//   end    - the last CTA to finish (device counter) raises data[me] = epoch on every peer and
//            then waits until all of MY data flags reached epoch: when the kernel completes,
//            the receive area is complete, and the next kernel of the stream (cuFFT, the local
//            gather) can read it.
// A rank's put never waits for something that only a later kernel of a peer provides (acks of
// e - 2 are raised by the peer's put e - 1 at the latest), so the scheme cannot deadlock.
struct PeerSync {
    PeerPtrs flags;               // every rank's flag array: [0, P) data flags, [P, 2P) ack flags
    const long long* my_flags;    // my own flag array
    int rank;
    int mode;                     // VPM_PEER_* bits
    long long epoch;
    unsigned* counter;            // zero-initialised device word, reset by the last CTA
};

__device__ __forceinline__ void peer_spin(const volatile long long* f, long long want, const char* what) {
    unsigned long long spins = 0;
    while (*f < want) {
        if (++spins > (1ull << 31)) {
            printf("vmadpm: peer %s wait timed out (flag = %lld, want %lld)\n", what, *f, want);
            __trap();
        }
        __nanosleep(64);
    }
}

__device__ __forceinline__ void peer_acquire(int P, const PeerSync& sy) {
    if ((sy.mode & VPM_PEER_RAISE_ACK) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 &&
        (int)threadIdx.x < P) {
        volatile long long* f = reinterpret_cast<volatile long long*>(sy.flags.p[threadIdx.x]);
        f[P + sy.rank] = sy.epoch - 1;
    }
    if (sy.mode & VPM_PEER_WAIT_ACK) {
        if ((int)threadIdx.x < P) peer_spin(sy.my_flags + P + threadIdx.x, sy.epoch - 2, "ack");
        __syncthreads();
    }
}

__device__ __forceinline__ void peer_publish(int P, const PeerSync& sy) {
    __threadfence_system();
    if (!(sy.mode & VPM_PEER_SIGNAL)) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(sy.counter, 1u);
        if (done == gridDim.x * gridDim.y * gridDim.z - 1) {
            *sy.counter = 0;
            __threadfence_system();
            for (int d = 0; d < P; ++d) {
                volatile long long* f = reinterpret_cast<volatile long long*>(sy.flags.p[d]);
                f[sy.rank] = sy.epoch;
            }
            __threadfence_system();
            if (sy.mode & VPM_PEER_WAIT_DATA) {
                for (int d = 0; d < P; ++d) peer_spin(sy.my_flags + d, sy.epoch, "data");
                __threadfence_system();
            }
        }
    }
}

// dst_d[dst_off + i * inner + c] = src[i * s_row + d * s_dst + c]   for every peer d
// grid = (chunks of a row, rows, peers): no index divisions, 16-byte coalesced loads / stores
__global__ void __launch_bounds__(256)
k_peer_put(int P, long long rows, long long inner, long long s_row, long long s_dst,
           const double2* __restrict__ src, PeerPtrs dst, long long dst_off, unsigned mask,
           PeerSync sy) {
    peer_acquire(P, sy);
    const int d = blockIdx.z;
    double2* __restrict__ out = reinterpret_cast<double2*>(dst.p[d]) + dst_off;
    // destinations outside the mask get no data; their CTAs still take part in the flag protocol
    const long long nrows = ((mask >> d) & 1u) ? rows : 0;
    for (long long i = blockIdx.y; i < nrows; i += gridDim.y) {
        const double2* __restrict__ in = src + i * s_row + d * s_dst;
        double2* __restrict__ o = out + i * inner;
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

struct PeerSegs {
    int start[9];            // segment boundaries in the routing order (P + 1 used)
    long long base[8];       // first destination row of my segment in peer d's receive area
};

// Particle rows to their destination ranks, pack and transfer fused.  Routing order position p
// (a row of the send order / of the receive order) belongs to segment d when
// start[d] <= p < start[d + 1] and lands in row base[d] + p - start[d] of peer d.
//   idx_is_src = 1 (exchange): position p = j,       source row = idx[j]
//   idx_is_src = 0 (gather):   position p = idx[j],  source row = j
template <int NC>
__global__ void __launch_bounds__(256)
k_peer_put_rows(int P, int n, int ncomp, const double* __restrict__ src,
                const int* __restrict__ idx, int idx_is_src, int skip_seg, PeerSegs seg,
                PeerPtrs dst, PeerSync sy) {
    peer_acquire(P, sy);
    const int nc = NC > 0 ? NC : ncomp;
    const long long total = (long long)n * nc;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e / nc);
        const int c = (int)(e - (long long)j * nc);
        const int k = idx[j];
        const int p = idx_is_src ? j : k;
        const int row = idx_is_src ? k : j;
        int d = 0;
#pragma unroll
        for (int t = 1; t < 8; ++t) d += (t < P && p >= seg.start[t]) ? 1 : 0;
        if (d == skip_seg) continue;
        double* out = reinterpret_cast<double*>(dst.p[d]);
        out[(seg.base[d] + (p - seg.start[d])) * nc + c] = src[(long long)row * nc + c];
    }
    peer_publish(P, sy);
}

// raise flag[slot] = epoch on every peer (after the stores of earlier kernels on this stream)
__global__ void k_peer_signal(int P, int slot, PeerPtrs flags, long long epoch) {
    __threadfence_system();
    const int d = threadIdx.x;
    if (d < P) {
        volatile long long* f = reinterpret_cast<volatile long long*>(flags.p[d]);
        f[slot] = epoch;
    }
    __threadfence_system();
}

// spin until every one of my P flags has reached `epoch`; traps instead of hanging forever
__global__ void k_peer_wait(int P, const long long* my_flags, long long epoch) {
    const int d = threadIdx.x;
    if (d < P) {
        const volatile long long* f = my_flags;
        unsigned long long spins = 0;
        while (f[d] < epoch) {
            if (++spins > (1ull << 31)) {
                printf("vmadpm: peer wait timed out (flag %d = %lld, want %lld)\n", d, f[d], epoch);
                __trap();
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

}  // namespace vpm

using namespace vpm;

static PeerPtrs make_ptrs(int P, const uint64_t* p) {
    PeerPtrs r;
    for (int i = 0; i < 8; ++i) r.p[i] = i < P ? p[i] : 0;
    return r;
}

static PeerSync make_sync(int P, const vpm_peer_sync* h) {
    PeerSync sy;
    memset(&sy, 0, sizeof(sy));
    if (h && h->mode) {
        VPM_REQUIRE(h->peer_flags && h->my_flags, "peer sync: flag arrays missing");
        VPM_REQUIRE(!(h->mode & VPM_PEER_SIGNAL) || h->counter, "peer sync: signalling needs a counter word");
        VPM_REQUIRE(!(h->mode & VPM_PEER_WAIT_DATA) || (h->mode & VPM_PEER_SIGNAL),
                    "peer sync: waiting for data implies signalling");
        sy.flags = make_ptrs(P, h->peer_flags);
        sy.my_flags = reinterpret_cast<const long long*>(h->my_flags);
        sy.rank = h->rank;
        sy.mode = h->mode;
        sy.epoch = h->epoch;
        sy.counter = h->counter;
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
                         int64_t s_dst, const double* src, const uint64_t* peer_bufs, int64_t dst_off,
                         unsigned mask, const vpm_peer_sync* sync);

int vpm_peer_put(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                 int64_t s_dst, const double* src, const uint64_t* peer_bufs, int64_t dst_off,
                 const vpm_peer_sync* sync) {
    return peer_put_impl(ctx, nranks, rows, inner, s_row, s_dst, src, peer_bufs, dst_off, 0xffffffffu, sync);
}

int vpm_peer_put_to(vpm_ctx* ctx, int nranks, uint32_t dest_mask, int64_t count, const double* src,
                    const uint64_t* peer_bufs, int64_t dst_off, const vpm_peer_sync* sync) {
    // `count` contiguous complex128 elements to every rank in dest_mask, at dst_off of its area
    return peer_put_impl(ctx, nranks, 1, count, 0, 0, src, peer_bufs, dst_off, dest_mask, sync);
}

static int peer_put_impl(vpm_ctx* ctx, int nranks, int64_t rows, int64_t inner, int64_t s_row,
                         int64_t s_dst, const double* src, const uint64_t* peer_bufs, int64_t dst_off,
                         unsigned mask, const vpm_peer_sync* sync) {
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
                   (long long)s_row, (long long)s_dst, reinterpret_cast<const double2*>(src),
                   make_ptrs(nranks, peer_bufs), (long long)dst_off, mask, make_sync(nranks, sync));
    }
    VPM_API_END
}

int vpm_peer_put_rows(vpm_ctx* ctx, int nranks, int64_t n, int ncomp, const double* src,
                      const int32_t* idx, int idx_is_src, int skip_seg, const int32_t* seg_start,
                      const int64_t* dst_base, const uint64_t* peer_bufs, const vpm_peer_sync* sync) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("peer_put_rows", (16.0 * ncomp + 4.0) * n);
    VPM_REQUIRE(ctx && seg_start && dst_base && peer_bufs && nranks >= 1 && nranks <= 8,
                "bad argument");
    VPM_REQUIRE(n >= 0 && n < (1LL << 31) && ncomp >= 1, "row count / columns out of range");
    if (n > 0 || sync) {
        VPM_REQUIRE(n == 0 || (src && idx), "bad argument");
        const PeerSync sy = make_sync(nranks, sync);
        PeerSegs seg;
        for (int i = 0; i < 9; ++i) seg.start[i] = i <= nranks ? seg_start[i] : 0x7fffffff;
        for (int i = 0; i < 8; ++i) seg.base[i] = i < nranks ? dst_base[i] : 0;
        const PeerPtrs dst = make_ptrs(nranks, peer_bufs);
        const long long total = (long long)n * ncomp;
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

int vpm_peer_signal(vpm_ctx* ctx, int nranks, int slot, const uint64_t* peer_flags, int64_t epoch) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && peer_flags && nranks >= 1 && nranks <= 8, "bad argument");
    VPM_LAUNCH(k_peer_signal, 1, 32, 0, ctx->stream, nranks, slot, make_ptrs(nranks, peer_flags),
               (long long)epoch);
    VPM_API_END
}

int vpm_peer_wait(vpm_ctx* ctx, int nranks, const int64_t* my_flags, int64_t epoch) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && my_flags && nranks >= 1 && nranks <= 8, "bad argument");
    VPM_LAUNCH(k_peer_wait, 1, 32, 0, ctx->stream, nranks,
               reinterpret_cast<const long long*>(my_flags), (long long)epoch);
    VPM_API_END
}

}  // extern "C"
