// decompose / exchange / gather: integer cell keys, stable counting sort, row permutes.
//
// Replaces pmesh ParticleMesh.decompose + Layout.exchange / Layout.gather as called from
// vmad/lib/fastpm.py:272,286,289,301,304.  The sort is a counting sort over cell keys:
//   (1) key + histogram (one int atomic per particle, returns the arrival rank),
//   (2) exclusive scan of the histogram  -> cell_start,
//   (3) scatter of particle ids to cell_start[key] + rank,
//   (4) per-cell re-sort of ids, which turns the arrival order of (1) into input order
//       and so makes the permutation exactly the stable argsort of the keys.
// It moves ~44 B per particle + 12 B per cell instead of the >=4 passes over (key, id)
// pairs of a radix sort, and FastPM particles arrive almost cell-sorted so the atomics
// of (1) hit L2 lines that are already resident.
#include "cic.cuh"
#include "tile.cuh"

namespace vpm {

__device__ __forceinline__ int key_of(const Geo& g, const double* __restrict__ x, long long p) {
    int i0, i1, i2;
    double f;
    cell_axis(x[3 * p + 0], g.s0, g.n0, i0, f);
    cell_axis(x[3 * p + 1], g.s1, g.n1, i1, f);
    cell_axis(x[3 * p + 2], g.s2, g.n2, i2, f);
    int sp = source_plane(g, i0);
    if (sp < 0) return g.nsp * g.n1 * g.n2;  // overflow bucket: not on this slab
    return (sp * g.n1 + i1) * g.n2 + i2;
}

__global__ void k_keys(Geo g, const double* __restrict__ x, long long np, int* __restrict__ keys) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p < np) keys[p] = key_of(g, x, p);
}

// KC_ILP particles per thread (blockDim apart): their position loads and their returning atomics
// are independent, so the two memory latencies of a particle overlap across the group
constexpr int KC_ILP = 4;
__global__ void __launch_bounds__(256)
k_keys_count(Geo g, const double* __restrict__ x, long long np,
             int* __restrict__ keys, int* __restrict__ rank,
             int* __restrict__ count) {
    const long long p0 = blockIdx.x * (long long)(KC_ILP * blockDim.x) + threadIdx.x;
    const int overflow = g.nsp * g.n1 * g.n2;
    const int lane = threadIdx.x & 31;
    int k[KC_ILP], r[KC_ILP];
#pragma unroll
    for (int j = 0; j < KC_ILP; ++j) {
        const long long p = p0 + (long long)j * blockDim.x;
        k[j] = p < np ? key_of(g, x, p) : -1;
    }
#pragma unroll
    for (int j = 0; j < KC_ILP; ++j) {
        // rows that are not on this slab (the padding of a fixed-capacity receive order: long runs
        // of them) would all hit one counter: one atomic per warp for those
        const unsigned off = __ballot_sync(0xffffffffu, k[j] == overflow);
        r[j] = 0;
        if (k[j] == overflow) {
            const int leader = __ffs(off) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(count + k[j], __popc(off));
            base = __shfl_sync(off, base, leader);
            r[j] = base + __popc(off & ((1u << lane) - 1));
        } else if (k[j] >= 0) {
            r[j] = atomicAdd(count + k[j], 1);
        }
    }
#pragma unroll
    for (int j = 0; j < KC_ILP; ++j) {
        const long long p = p0 + (long long)j * blockDim.x;
        if (p < np) {
            keys[p] = k[j];
            rank[p] = r[j];
        }
    }
}

// ---- exclusive scan (int32), recursive three-phase --------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    int base = (w == 0) ? 0 : warp_sums[w - 1];
    total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void k_scan_reduce(const int* __restrict__ in, long long n, int* __restrict__ sums) {
    long long base = blockIdx.x * (long long)SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
        if (base + j < n) s += in[base + j];
    int total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// close != 0: also write the grand total to out[n] (the cell offsets array has n + 1 entries)
__global__ void k_scan_down(const int* __restrict__ in, long long n,
                            const int* __restrict__ block_offsets, int* __restrict__ out, int close = 0) {
    long long base = blockIdx.x * (long long)SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        s += v[j];
    }
    int total;
    int ex = block_exclusive_scan(s, total);
    int off = (block_offsets ? block_offsets[blockIdx.x] : 0) + ex;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) out[base + j] = off;
        off += v[j];
    }
    if (close && base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = off;
}

// exclusive scan of in[0..n) into out[0..n); tmp must hold scan_tmp_ints(n) ints
static long long scan_tmp_ints(long long n) {
    long long t = 0;
    while (n > SCAN_TILE) {
        n = ceil_div(n, SCAN_TILE);
        t += n;
    }
    return t + 1;
}

static void exclusive_scan(cudaStream_t st, const int* in, long long n, int* out, int* tmp, int close = 0) {
    if (n <= 0) return;
    long long nb = ceil_div(n, SCAN_TILE);
    if (nb == 1) {
        VPM_LAUNCH(k_scan_down, 1, SCAN_THREADS, 0, st, in, n, (const int*)nullptr, out, close);
        return;
    }
    int* sums = tmp;
    VPM_LAUNCH(k_scan_reduce, (unsigned)nb, SCAN_THREADS, 0, st, in, n, sums);
    exclusive_scan(st, sums, nb, sums, tmp + nb);  // in-place on the block sums
    VPM_LAUNCH(k_scan_down, (unsigned)nb, SCAN_THREADS, 0, st, in, n, (const int*)sums, out, close);
}

__global__ void __launch_bounds__(256)
k_scatter_ids(const int* __restrict__ keys, const int* __restrict__ rank,
              const int* __restrict__ cell_start, long long np, int* __restrict__ perm) {
    // KC_ILP particles per thread: the dependent chain key -> cell_start[key] -> store runs four wide
    const long long p0 = blockIdx.x * (long long)(KC_ILP * blockDim.x) + threadIdx.x;
    int k[KC_ILP], r[KC_ILP], b[KC_ILP];
#pragma unroll
    for (int j = 0; j < KC_ILP; ++j) {
        const long long p = p0 + (long long)j * blockDim.x;
        k[j] = p < np ? __ldg(keys + p) : -1;
        r[j] = p < np ? __ldg(rank + p) : 0;
    }
#pragma unroll
    for (int j = 0; j < KC_ILP; ++j) b[j] = k[j] >= 0 ? __ldg(cell_start + k[j]) : 0;
#pragma unroll
    for (int j = 0; j < KC_ILP; ++j) {
        const long long p = p0 + (long long)j * blockDim.x;
        if (p < np) perm[b[j] + r[j]] = (int)p;
    }
}

// per-cell ascending sort of the ids (restores input order inside a cell).  Three tiers by
// occupancy: <= FIX_SMALL ids: insertion sort by one thread; <= FIX_MEDIUM: rank sort by one
// warp, in place; larger (pathological): rank sort spread over the whole grid via a copy.
constexpr int FIX_SMALL = 16;
constexpr int FIX_MEDIUM = 1024;
// The arrival order of the atomics is already ascending for most cells (blocks start in index
// order), so every tier first checks whether anything has to move.  Small cells are sorted in
// REGISTERS (independent loads, a branch-free insertion network, stores only when the order
// changed) instead of through global memory.
__global__ void k_fix_small(const int* __restrict__ cell_start, long long ncell, long long nrepair,
                            int* __restrict__ perm, int* __restrict__ lists,
                            int* __restrict__ counts) {
    long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= nrepair) return;
    const int b = cell_start[c], e = cell_start[c + 1];
    const int cnt = e - b;
    if (cnt < 2) return;
    if (cnt > FIX_SMALL) {
        const int which = cnt > FIX_MEDIUM ? 1 : 0;
        int slot = atomicAdd(counts + which, 1);
        lists[which * ncell + slot] = (int)c;
        return;
    }
    // already ascending?  (contiguous entries: the loads of neighbouring threads share sectors)
    int prev = perm[b];
    bool sorted = true;
    for (int i = b + 1; i < e; ++i) {
        const int v = perm[i];
        sorted = sorted && (prev <= v);
        prev = v;
    }
    if (sorted) return;
    for (int i = b + 1; i < e; ++i) {  // insertion sort, segments are tiny
        int v = perm[i];
        int j = i - 1;
        while (j >= b && perm[j] > v) {
            perm[j + 1] = perm[j];
            --j;
        }
        perm[j + 1] = v;
    }
}

// medium cells: one warp per cell.  The ids are distinct, so the final slot of an id is the
// number of smaller ids in the cell; ranks are computed first, then written in place.
__global__ void __launch_bounds__(256)
k_fix_medium(const int* __restrict__ cell_start, const int* __restrict__ list,
             const int* __restrict__ count, int* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarp = (gridDim.x * blockDim.x) >> 5;
    const int n_med = *count;
    for (int s = warp; s < n_med; s += nwarp) {
        const int c = list[s];
        const int b = cell_start[c];
        const int n = cell_start[c + 1] - b;
        int* a = perm + b;
        {   // already ascending (the usual case: the atomics arrive in index order)?
            bool bad = false;
            for (int i = lane; i + 1 < n; i += 32) bad = bad || (a[i] > a[i + 1]);
            if (!__any_sync(0xffffffffu, bad)) continue;
        }
        int v[FIX_MEDIUM / 32], r[FIX_MEDIUM / 32];
        const int nk = (n + 31) >> 5;  // register slots in use (warp-uniform)
#pragma unroll
        for (int k = 0; k < FIX_MEDIUM / 32; ++k) {
            const int i = lane + 32 * k;
            v[k] = i < n ? a[i] : 0x7fffffff;
            r[k] = 0;
        }
        for (int t = 0; t < n; ++t) {
            const int o = a[t];  // warp-uniform broadcast load
#pragma unroll
            for (int k = 0; k < FIX_MEDIUM / 32; ++k)
                if (k < nk) r[k] += (o < v[k]);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < FIX_MEDIUM / 32; ++k) {
            const int i = lane + 32 * k;
            if (i < n) a[r[k]] = v[k];
        }
        __syncwarp();
    }
}

// huge cells: every block takes a slice of every such cell, so one pathological cell still
// uses the whole GPU; writes go to perm_out and are copied back afterwards.
__global__ void k_fix_big_rank(const int* __restrict__ cell_start, const int* __restrict__ big_list,
                               const int* __restrict__ big_count, const int* __restrict__ perm_in,
                               int* __restrict__ perm_out) {
    __shared__ int tile[1024];
    int nbig = *big_count;
    for (int s = 0; s < nbig; ++s) {
        int c = big_list[s];
        int b = cell_start[c];
        int n = cell_start[c + 1] - b;
        const int* a = perm_in + b;
        for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n;
             i0 += (long long)gridDim.x * blockDim.x) {
            int i = (int)i0 + threadIdx.x;
            int v = (i < n) ? a[i] : 0;
            int r = 0;
            for (int t0 = 0; t0 < n; t0 += 1024) {
                int tn = min(1024, n - t0);
                __syncthreads();
                for (int t = threadIdx.x; t < tn; t += blockDim.x) tile[t] = a[t0 + t];
                __syncthreads();
                if (i < n)
                    for (int t = 0; t < tn; ++t) r += (tile[t] < v);
            }
            if (i < n) perm_out[b + r] = v;
        }
        __syncthreads();
    }
}

__global__ void k_copy_big(const int* __restrict__ cell_start, const int* __restrict__ big_list,
                           const int* __restrict__ big_count, const int* __restrict__ src,
                           int* __restrict__ dst) {
    for (int s = blockIdx.x; s < *big_count; s += gridDim.x) {
        int c = big_list[s];
        int b = cell_start[c], e = cell_start[c + 1];
        for (int i = b + threadIdx.x; i < e; i += blockDim.x) dst[i] = src[i];
    }
}


// ---- row permutes -------------------------------------------------------------------
// an index equal to VPM_ROW_PADDING marks a slot without a particle (fixed-capacity routing):
// takes write the context's fill row there (three-column arrays; 0 otherwise), scatters skip it
struct Fill3 {
    double v[3];
};

template <int NC>
__global__ void k_take_rows(const double* __restrict__ src, const int* __restrict__ perm,
                            long long n, int ncomp, double scale, Fill3 fill, double* __restrict__ dst) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int nc = NC > 0 ? NC : ncomp;
    if (e >= n * nc) return;
    long long i = e / nc;
    int c = (int)(e - i * nc);
    const int p = perm[i];
    if (p == VPM_ROW_PADDING) {
        dst[e] = (NC == 3) ? fill.v[c] : 0.0;
        return;
    }
    dst[e] = scale * __ldg(src + (long long)p * nc + c);   // scale = 1 is exact
}

// [n, 3] rows, one thread per row: the scattered source row is fetched with one 16-byte and one
// 8-byte access (tile.cuh), the destination tile of the CTA -- contiguous -- leaves as ONE bulk copy.
// TWO = 1: rows come from two arrays (k_take_rows2 semantics).
constexpr int TAKE_TILE = 256;
template <int TWO>
__global__ void __launch_bounds__(TAKE_TILE)
k_take_rows3(const double* __restrict__ a, const double* __restrict__ b, const int* __restrict__ perm,
             long long n, double scale, Fill3 fill, double* __restrict__ dst) {
    __shared__ RowTile<TAKE_TILE, 1> tile;
    const long long p0 = blockIdx.x * (long long)TAKE_TILE;
    const long long i = p0 + threadIdx.x;
    const bool live = i < n;
    double v0 = fill.v[0], v1 = fill.v[1], v2 = fill.v[2];
    if (live) {
        const int p = __ldg(perm + i);
        if (p != VPM_ROW_PADDING) {
            if (TWO && p < 0) load_row3(b, (long long)(-1 - p), v0, v1, v2);
            else load_row3(a, (long long)p, v0, v1, v2);
            v0 *= scale; v1 *= scale; v2 *= scale;      // scale = 1 is exact
        }
    }
    if (tile.can_flush(dst, p0, n)) {
        tile.put(0, threadIdx.x, v0, v1, v2);
        tile.flush(0, dst, p0);
    } else if (live) {
        dst[3 * i + 0] = v0; dst[3 * i + 1] = v1; dst[3 * i + 2] = v2;
    }
}

// slots of the sorted order at or beyond the number of rows that really arrived hold padding
__global__ void k_mark_padding(int* __restrict__ sorted_to_recv, int n, const int* __restrict__ rcnt, int P) {
    int nvalid = 0;
    for (int s = 0; s < P; ++s) nvalid += rcnt[s];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && t >= nvalid) sorted_to_recv[t] = VPM_ROW_PADDING;
}

template <int NC, bool ADD>
__global__ void k_scatter_rows(const double* __restrict__ src, const int* __restrict__ perm,
                               long long n, int ncomp, double scale, double* __restrict__ dst) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int nc = NC > 0 ? NC : ncomp;
    if (e >= n * nc) return;
    long long i = e / nc;
    int c = (int)(e - i * nc);
    const int p = perm[i];
    if (p < 0) return;              // masked row (lives on another rank / arrives separately)
    double v = scale * src[e];       // scale = 1 is exact
    double* d = dst + (long long)p * nc + c;
    if (ADD) atomicAdd(d, v);
    else *d = v;
}

// out[k] = idx[k] >= 0 ? a[idx[k]] : b[-1 - idx[k]]   (rows of NC doubles)
template <int NC>
__global__ void k_take_rows2(const double* __restrict__ a, const double* __restrict__ b,
                             const int* __restrict__ idx, long long n, int ncomp, double scale, Fill3 fill,
                             double* __restrict__ dst) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int nc = NC > 0 ? NC : ncomp;
    if (e >= n * nc) return;
    long long i = e / nc;
    int c = (int)(e - i * nc);
    const int p = idx[i];
    if (p == VPM_ROW_PADDING) {
        dst[e] = (NC == 3) ? fill.v[c] : 0.0;
        return;
    }
    dst[e] = scale * (p >= 0 ? a[(long long)p * nc + c] : b[(long long)(-1 - p) * nc + c]);
}

// Index composition for a cross-rank layout: rows that stay on this rank skip the receive area.
//   take[k]   = the home row sorted slot k reads directly, or -1 - (receive-area row)
//   remote[j] = send_idx[j] for rows sent to other ranks, -1 inside the self segment
__global__ void k_route_compose(const int* __restrict__ sorted_to_recv, int nrecv,
                                const int* __restrict__ send_idx, int nsend, int recv_lo,
                                int recv_hi, int send_lo, int* __restrict__ take,
                                int* __restrict__ remote, int* __restrict__ inv_remote) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nrecv) {
        const int r = sorted_to_recv[t];
        if (r == VPM_ROW_PADDING) {
            take[t] = VPM_ROW_PADDING;
        } else if (r >= recv_lo && r < recv_hi) {
            const int h = send_idx[send_lo + (r - recv_lo)];
            take[t] = h >= 0 ? h : VPM_ROW_PADDING;
        } else {
            take[t] = -1 - r;
            if (inv_remote) inv_remote[r < recv_lo ? r : r - (recv_hi - recv_lo)] = t;
        }
    }
    if (t < nsend) {
        const int send_hi = send_lo + (recv_hi - recv_lo);
        remote[t] = (t >= send_lo && t < send_hi) ? -1 : send_idx[t];
    }
}

// ---- routing of particles to slab owners (stable multisplit) -----------------------------
// A particle goes to the rank that owns its floor plane i0 and, when different, to the rank
// that owns plane i0 + 1 (the two planes a CIC particle touches).  Within a destination the
// original particle order is kept (stable), so the routing is deterministic.
constexpr int ROUTE_THREADS = 256;
constexpr int ROUTE_ROUNDS = 8;                          // a block routes ROUTE_ROUNDS x 256 consecutive particles
constexpr int ROUTE_TILE = ROUTE_THREADS * ROUTE_ROUNDS;
constexpr int ROUTE_MAXP = 64;

__device__ __forceinline__ void route_of(const Geo& g, int planes_per_rank, const double* x,
                                         long long p, int& r0, int& r1) {
    int i0;
    double f;
    cell_axis(x[3 * p], g.s0, g.n0, i0, f);
    r0 = i0 / planes_per_rank;
    r1 = wrap_inc(i0, g.n0) / planes_per_rank;
    if (r1 == r0) r1 = -1;
}

__global__ void __launch_bounds__(ROUTE_THREADS)
k_route_count(Geo g, int planes_per_rank, int P, const double* __restrict__ x, long long np,
              long long nblocks, int* __restrict__ counts /* [P][nblocks] */) {
    __shared__ int c[ROUTE_MAXP];
    for (int d = threadIdx.x; d < P; d += blockDim.x) c[d] = 0;
    __syncthreads();
    for (int r = 0; r < ROUTE_ROUNDS; ++r) {
        long long p = blockIdx.x * (long long)ROUTE_TILE + r * ROUTE_THREADS + threadIdx.x;
        if (p < np) {
            int r0, r1;
            route_of(g, planes_per_rank, x, p, r0, r1);
            atomicAdd(&c[r0], 1);
            if (r1 >= 0) atomicAdd(&c[r1], 1);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < P; d += blockDim.x) counts[d * nblocks + blockIdx.x] = c[d];
}

// in-block stable ranks: rounds in order, within a round warps in order, within a warp lanes in
// order (ballot).  A thread first finds the destinations of its ROUTE_ROUNDS particles (eight
// independent loads), then the block visits only the destinations that occur in it (a slab keeps
// nearly all of its particles: typically itself and one neighbour): per destination every warp
// ballots its eight rounds, the 8 x 8 (round, warp) counts are scanned by one warp, and the rows
// are written -- three barriers per destination present, instead of three per (round, destination).
template <bool STATIC>
__device__ __forceinline__ void route_scatter_block(const Geo& g, int planes_per_rank, int P,
                                                    const double* __restrict__ x, long long np,
                                                    long long nblocks, const int* __restrict__ offsets,
                                                    const int* seg_start, int* __restrict__ send_idx) {
    constexpr int NW = ROUTE_THREADS / 32;
    __shared__ int cnt[ROUTE_ROUNDS * NW];      // (round, warp) counts of the current destination, then their exclusive scan
    __shared__ unsigned long long present;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) present = 0ull;
    __syncthreads();
    int r0[ROUTE_ROUNDS], r1[ROUTE_ROUNDS];
    unsigned long long mine_mask = 0ull;
#pragma unroll
    for (int r = 0; r < ROUTE_ROUNDS; ++r) {
        const long long p = blockIdx.x * (long long)ROUTE_TILE + r * ROUTE_THREADS + threadIdx.x;
        r0[r] = r1[r] = -1;
        if (p < np) {
            route_of(g, planes_per_rank, x, p, r0[r], r1[r]);
            mine_mask |= 1ull << r0[r];
            if (r1[r] >= 0) mine_mask |= 1ull << r1[r];
        }
    }
    // destinations present in the block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine_mask |= __shfl_xor_sync(0xffffffffu, mine_mask, o);
    if (lane == 0 && mine_mask) atomicOr(&present, mine_mask);
    __syncthreads();
    unsigned long long todo = present;
    while (todo) {
        const int d = __ffsll((long long)todo) - 1;
        todo &= todo - 1;
        unsigned bal[ROUTE_ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUTE_ROUNDS; ++r) {
            bal[r] = __ballot_sync(0xffffffffu, (r0[r] == d) || (r1[r] == d));
            if (lane == 0) cnt[r * NW + w] = __popc(bal[r]);
        }
        __syncthreads();
        if (w == 0) {   // exclusive scan of the 64 counts in (round, warp) order: two per lane
            static_assert(ROUTE_ROUNDS * NW == 64, "scan below assumes 64 counts");
            const int a0 = cnt[2 * lane], a1 = cnt[2 * lane + 1];
            int v = a0 + a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            const int ex = v - (a0 + a1);
            cnt[2 * lane] = ex;
            cnt[2 * lane + 1] = ex + a0;
        }
        __syncthreads();
        const int first = STATIC ? offsets[d * nblocks + blockIdx.x] - offsets[d * nblocks]
                                 : offsets[d * nblocks + blockIdx.x];
#pragma unroll
        for (int r = 0; r < ROUTE_ROUNDS; ++r) {
            if ((r0[r] == d) || (r1[r] == d)) {
                const long long p = blockIdx.x * (long long)ROUTE_TILE + r * ROUTE_THREADS + threadIdx.x;
                const int pos = first + cnt[r * NW + w] + __popc(bal[r] & ((1u << lane) - 1));
                if (STATIC) {
                    // rank among the rows bound for d (input order): rows of earlier blocks + this one
                    if (pos < seg_start[d + 1] - seg_start[d]) send_idx[seg_start[d] + pos] = (int)p;
                } else {
                    send_idx[pos] = (int)p;
                }
            }
        }
        __syncthreads();    // cnt is re-used by the next destination
    }
}

__global__ void __launch_bounds__(ROUTE_THREADS)
k_route_scatter(Geo g, int planes_per_rank, int P, const double* __restrict__ x, long long np,
                long long nblocks, const int* __restrict__ offsets /* [P][nblocks] */,
                int* __restrict__ send_idx) {
    route_scatter_block<false>(g, planes_per_rank, P, x, np, nblocks, offsets, nullptr, send_idx);
}

// fixed-capacity variant: destination d owns positions [seg[d], seg[d + 1]) of the send order
struct RouteSegs {
    int start[ROUTE_MAXP + 1];
};

__global__ void __launch_bounds__(ROUTE_THREADS)
k_route_scatter_static(Geo g, int planes_per_rank, int P, const double* __restrict__ x, long long np,
                       long long nblocks, const int* __restrict__ offsets /* [P][nblocks] */,
                       RouteSegs seg, int* __restrict__ send_idx) {
    route_scatter_block<true>(g, planes_per_rank, P, x, np, nblocks, offsets, seg.start, send_idx);
}

// counts[d] = rows bound for d (clipped to the capacity), counts[P] = 1 when a segment overflowed
// (sticky: OR-ed into what is there)
__global__ void k_route_totals_static(int P, long long nblocks, const int* __restrict__ offsets,
                                      const int* __restrict__ counts_in, RouteSegs seg,
                                      int* __restrict__ counts /* [P + 1] */) {
    int d = threadIdx.x;
    if (d < P) {
        const long long first = (long long)d * nblocks;
        const long long end = (long long)(d + 1) * nblocks - 1;
        const int total = offsets[end] + counts_in[end] - offsets[first];
        const int cap = seg.start[d + 1] - seg.start[d];
        counts[d] = min(total, cap);
        if (total > cap) atomicOr(counts + P, 1);
    }
}

__global__ void k_route_totals(int P, long long nblocks, long long nsend_total_index,
                               const int* __restrict__ offsets, const int* __restrict__ counts,
                               int* __restrict__ totals /* [P + 1] */) {
    int d = threadIdx.x;
    if (d < P) totals[d] = offsets[d * nblocks];
    if (d == P) {
        long long last = (long long)P * nblocks - 1;
        totals[P] = offsets[last] + counts[last];
    }
}

}  // namespace vpm

using namespace vpm;

extern "C" {

int vpm_route_build(vpm_ctx* ctx, const vpm_mesh* mesh, int nranks, const double* x, int64_t np,
                    int32_t* send_idx, int64_t send_capacity, int32_t* send_start_host) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("route_build", 28.0 * np);
    VPM_REQUIRE(ctx && mesh && send_start_host, "NULL argument");
    VPM_REQUIRE(nranks >= 1 && nranks <= ROUTE_MAXP, "unsupported number of ranks");
    VPM_REQUIRE(mesh->n[0] % nranks == 0, "mesh planes must divide evenly over the ranks");
    VPM_REQUIRE(np >= 0 && 2 * np < (1LL << 31) - 1, "particle count must fit int32");
    vpm_mesh m = *mesh;  // routing only needs the global geometry
    m.start0 = 0;
    m.nloc0 = m.n[0];
    m.ghost = 0;
    Geo g = make_geo(&m);
    const int ppr = (int)(mesh->n[0] / nranks);
    cudaStream_t st = ctx->stream;
    const long long nblocks = std::max<long long>(1, ceil_div(np, ROUTE_TILE));
    const long long ncount = (long long)nranks * nblocks;
    const long long ntmp = scan_tmp_ints(ncount);
    int* ws = (int*)ctx->workspace((size_t)(2 * ncount + ntmp + nranks + 8) * sizeof(int));
    int* counts = ws;
    int* offsets = counts + ncount;
    int* tmp = offsets + ncount;
    int* totals = tmp + ntmp;
    VPM_LAUNCH(k_route_count, (unsigned)nblocks, ROUTE_THREADS, 0, st, g, ppr, nranks, x,
               (long long)np, nblocks, counts);
    exclusive_scan(st, counts, ncount, offsets, tmp);
    VPM_LAUNCH(k_route_totals, 1, ROUTE_MAXP + 1, 0, st, nranks, nblocks, 0LL, offsets, counts, totals);
    // the send counts are needed on the host to size the exchange: the one sync of decompose
    VPM_CUDA(cudaMemcpyAsync(send_start_host, totals, (nranks + 1) * sizeof(int),
                             cudaMemcpyDeviceToHost, st));
    VPM_CUDA(cudaStreamSynchronize(st));
    VPM_REQUIRE(send_start_host[nranks] <= send_capacity, "send index buffer too small");
    if (np > 0) {
        VPM_REQUIRE(x && send_idx, "NULL array");
        VPM_LAUNCH(k_route_scatter, (unsigned)nblocks, ROUTE_THREADS, 0, st, g, ppr, nranks, x,
                   (long long)np, nblocks, offsets, send_idx);
    }
    VPM_API_END
}

int vpm_route_build_static(vpm_ctx* ctx, const vpm_mesh* mesh, int nranks, const double* x, int64_t np,
                           const int32_t* seg_start_host, int32_t* send_idx, int32_t* counts) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("route_build", 28.0 * np);
    VPM_REQUIRE(ctx && mesh && seg_start_host && send_idx && counts, "NULL argument");
    VPM_REQUIRE(nranks >= 1 && nranks <= ROUTE_MAXP, "unsupported number of ranks");
    VPM_REQUIRE(mesh->n[0] % nranks == 0, "mesh planes must divide evenly over the ranks");
    VPM_REQUIRE(np >= 0 && 2 * np < (1LL << 31) - 1, "particle count must fit int32");
    vpm_mesh m = *mesh;  // routing only needs the global geometry
    m.start0 = 0;
    m.nloc0 = m.n[0];
    m.ghost = 0;
    Geo g = make_geo(&m);
    const int ppr = (int)(mesh->n[0] / nranks);
    cudaStream_t st = ctx->stream;
    RouteSegs seg;
    for (int d = 0; d <= ROUTE_MAXP; ++d) seg.start[d] = seg_start_host[d <= nranks ? d : nranks];
    for (int d = 0; d < nranks; ++d) VPM_REQUIRE(seg.start[d + 1] >= seg.start[d], "segments must be ordered");
    // positions without a particle keep -1
    VPM_PROF_CALL("memset(send_idx)", st,
                  VPM_CUDA(cudaMemsetAsync(send_idx, 0xff, (size_t)seg.start[nranks] * sizeof(int), st)));
    const long long nblocks = std::max<long long>(1, ceil_div(np, ROUTE_TILE));
    const long long ncount = (long long)nranks * nblocks;
    const long long ntmp = scan_tmp_ints(ncount);
    int* ws = (int*)ctx->workspace((size_t)(2 * ncount + ntmp + 8) * sizeof(int));
    int* blk_counts = ws;
    int* offsets = blk_counts + ncount;
    int* tmp = offsets + ncount;
    VPM_LAUNCH(k_route_count, (unsigned)nblocks, ROUTE_THREADS, 0, st, g, ppr, nranks, x,
               (long long)np, nblocks, blk_counts);
    exclusive_scan(st, blk_counts, ncount, offsets, tmp);
    VPM_LAUNCH(k_route_totals_static, 1, ROUTE_MAXP, 0, st, nranks, nblocks, offsets, blk_counts, seg, counts);
    if (np > 0) {
        VPM_REQUIRE(x, "NULL array");
        VPM_LAUNCH(k_route_scatter_static, (unsigned)nblocks, ROUTE_THREADS, 0, st, g, ppr, nranks, x,
                   (long long)np, nblocks, offsets, seg, send_idx);
    }
    VPM_API_END
}

int64_t vpm_layout_ncell(const vpm_mesh* m) {
    if (!m) return -1;
    int64_t planes = m->ghost ? m->nloc0 + 1 : m->n[0];
    return planes * m->n[1] * m->n[2] + 1;  // + overflow bucket (not on this slab)
}

int vpm_cell_keys(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x, int64_t np, int32_t* keys) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("cell_keys", 28.0 * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    Geo g = make_geo(mesh);
    if (np > 0) {
        VPM_REQUIRE(x && keys, "NULL array");
        VPM_LAUNCH(k_keys, (unsigned)ceil_div(np, 256), 256, 0, ctx->stream, g, x, (long long)np, keys);
    }
    VPM_API_END
}

int vpm_layout_build(vpm_ctx* ctx, const vpm_mesh* mesh, const double* x, int64_t np,
                     int32_t* keys, int32_t* perm, int32_t* cell_start) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("layout_build", 48.0 * np);
    VPM_REQUIRE(ctx, "ctx is NULL");
    VPM_REQUIRE(np >= 0 && np < (1LL << 31) - 1, "particle count must fit int32");
    Geo g = make_geo(mesh);
    long long ncell = vpm_layout_ncell(mesh);
    VPM_REQUIRE(cell_start, "cell_start is NULL");
    cudaStream_t st = ctx->stream;
    // workspace: counts[8] | count[ncell] -- zeroed by ONE memset -- | rank[np] | scan tmp | lists[2][ncell] | perm2[np]
    long long ntmp = scan_tmp_ints(ncell);
    size_t ints = 8 + (size_t)ncell + (size_t)np + (size_t)ntmp + 2 * (size_t)ncell + (size_t)np;
    int* ws = (int*)ctx->workspace(ints * sizeof(int));
    int* big_count = ws;
    int* count = ws + 8;
    int* rank = count + ncell;
    int* tmp = rank + np;
    int* big_list = tmp + ntmp;          // [0]: medium cells, [1]: huge cells
    int* perm2 = big_list + 2 * ncell;
    VPM_PROF_CALL("memset(cell counts)", st, VPM_CUDA(cudaMemsetAsync(ws, 0, (size_t)(8 + ncell) * sizeof(int), st)));
    if (np > 0) {
        VPM_REQUIRE(x && keys && perm, "NULL array");
        VPM_LAUNCH(k_keys_count, (unsigned)ceil_div(np, 256 * KC_ILP), 256, 0, st, g, x, (long long)np, keys,
                   rank, count);
    }
    exclusive_scan(st, count, ncell, cell_start, tmp, /*close: cell_start[ncell] = np*/ 1);
    if (np > 0) {
        VPM_LAUNCH(k_scatter_ids, (unsigned)ceil_div(np, 256 * KC_ILP), 256, 0, st, keys, rank, cell_start,
                   (long long)np, perm);
        // the last bucket collects what is not on this slab (only the inert padding rows of the
        // fixed-capacity routing end up there, possibly very many): their order is irrelevant
        VPM_LAUNCH(k_fix_small, (unsigned)ceil_div(ncell, 256), 256, 0, st, cell_start, ncell, ncell - 1, perm,
                   big_list, big_count);
        int nblk = 4 * ctx->sm_count;
        VPM_LAUNCH(k_fix_medium, nblk, 256, 0, st, cell_start, big_list, big_count, perm);
        // pathological cells (> FIX_MEDIUM ids): rank-sort each into perm2, then copy back
        VPM_LAUNCH(k_fix_big_rank, nblk, 256, 0, st, cell_start, big_list + ncell, big_count + 1,
                   perm, perm2);
        VPM_LAUNCH(k_copy_big, nblk, 256, 0, st, cell_start, big_list + ncell, big_count + 1, perm2,
                   perm);
    }
    VPM_API_END
}

static void launch_take_rows(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n, int ncomp,
                             double scale, double* dst) {
    VPM_REQUIRE(ctx, "ctx is NULL");
    VPM_REQUIRE(ncomp >= 1, "ncomp must be >= 1");
    if (n > 0) {
        VPM_REQUIRE(src && perm && dst, "NULL array");
        unsigned grid = (unsigned)ceil_div(n * ncomp, 256);
        Fill3 fill;
        for (int c = 0; c < 3; ++c) fill.v[c] = ctx->fill_row[c];
        if (ncomp == 1) VPM_LAUNCH(k_take_rows<1>, grid, 256, 0, ctx->stream, src, perm, (long long)n, ncomp, scale, fill, dst);
        else if (ncomp == 3) VPM_LAUNCH(k_take_rows3<0>, (unsigned)ceil_div(n, TAKE_TILE), TAKE_TILE, 0, ctx->stream, src, (const double*)nullptr, perm, (long long)n, scale, fill, dst);
        else VPM_LAUNCH(k_take_rows<0>, grid, 256, 0, ctx->stream, src, perm, (long long)n, ncomp, scale, fill, dst);
    }
}

int vpm_take_rows(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n, int ncomp,
                  double* dst) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("take_rows", (16.0 * ncomp + 4.0) * n);
    launch_take_rows(ctx, src, perm, n, ncomp, 1.0, dst);
    VPM_API_END
}

int vpm_take_rows_scaled(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n, int ncomp,
                         double scale, double* dst) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("take_rows", (16.0 * ncomp + 4.0) * n);
    launch_take_rows(ctx, src, perm, n, ncomp, scale, dst);
    VPM_API_END
}

int vpm_scatter_rows(vpm_ctx* ctx, const double* src, const int32_t* perm, int64_t n, int ncomp,
                     double* dst) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("scatter_rows", (16.0 * ncomp + 4.0) * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    VPM_REQUIRE(ncomp >= 1, "ncomp must be >= 1");
    if (n > 0) {
        VPM_REQUIRE(src && perm && dst, "NULL array");
        unsigned grid = (unsigned)ceil_div(n * ncomp, 256);
        if (ncomp == 1) VPM_LAUNCH((k_scatter_rows<1, false>), grid, 256, 0, ctx->stream, src, perm, (long long)n, ncomp, 1.0, dst);
        else if (ncomp == 3) VPM_LAUNCH((k_scatter_rows<3, false>), grid, 256, 0, ctx->stream, src, perm, (long long)n, ncomp, 1.0, dst);   // a thread-per-row form with a staged source tile was measured slower (0.89 vs 0.81 ms at 2^24 rows)
        else VPM_LAUNCH((k_scatter_rows<0, false>), grid, 256, 0, ctx->stream, src, perm, (long long)n, ncomp, 1.0, dst);
    }
    VPM_API_END
}

int vpm_take_rows2(vpm_ctx* ctx, const double* a, const double* b, const int32_t* idx, int64_t n,
                   int ncomp, double* dst) {
    return vpm_take_rows2_scaled(ctx, a, b, idx, n, ncomp, 1.0, dst);
}

int vpm_take_rows2_scaled(vpm_ctx* ctx, const double* a, const double* b, const int32_t* idx, int64_t n,
                          int ncomp, double scale, double* dst) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("take_rows", (16.0 * ncomp + 4.0) * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    VPM_REQUIRE(ncomp >= 1, "ncomp must be >= 1");
    if (n > 0) {
        VPM_REQUIRE(a && b && idx && dst, "NULL array");
        unsigned grid = (unsigned)ceil_div(n * ncomp, 256);
        Fill3 fill;
        for (int c = 0; c < 3; ++c) fill.v[c] = ctx->fill_row[c];
        if (ncomp == 1) VPM_LAUNCH(k_take_rows2<1>, grid, 256, 0, ctx->stream, a, b, idx, (long long)n, ncomp, scale, fill, dst);
        else if (ncomp == 3) VPM_LAUNCH(k_take_rows3<1>, (unsigned)ceil_div(n, TAKE_TILE), TAKE_TILE, 0, ctx->stream, a, b, idx, (long long)n, scale, fill, dst);
        else VPM_LAUNCH(k_take_rows2<0>, grid, 256, 0, ctx->stream, a, b, idx, (long long)n, ncomp, scale, fill, dst);
    }
    VPM_API_END
}

int vpm_route_compose(vpm_ctx* ctx, const int32_t* sorted_to_recv, int64_t nrecv,
                      const int32_t* send_idx, int64_t nsend, int64_t recv_lo, int64_t recv_hi,
                      int64_t send_lo, int32_t* take, int32_t* remote, int32_t* inv_remote) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("route_compose", 8.0 * nrecv + 8.0 * nsend);
    VPM_REQUIRE(ctx, "ctx is NULL");
    VPM_REQUIRE(nrecv >= 0 && nsend >= 0 && nrecv < (1LL << 31) && nsend < (1LL << 31), "bad counts");
    const long long n = std::max<long long>(nrecv, nsend);
    if (n > 0) {
        VPM_REQUIRE((nrecv == 0 || (sorted_to_recv && take)) && (nsend == 0 || (send_idx && remote)),
                    "NULL array");
        if (inv_remote && nrecv - (recv_hi - recv_lo) > 0)
            VPM_CUDA(cudaMemsetAsync(inv_remote, 0xff, (size_t)(nrecv - (recv_hi - recv_lo)) * sizeof(int), ctx->stream));
        VPM_LAUNCH(k_route_compose, (unsigned)ceil_div(n, 256), 256, 0, ctx->stream, sorted_to_recv,
                   (int)nrecv, send_idx, (int)nsend, (int)recv_lo, (int)recv_hi, (int)send_lo, take,
                   remote, inv_remote);
    }
    VPM_API_END
}

int vpm_route_mark_padding(vpm_ctx* ctx, int32_t* sorted_to_recv, int64_t n, const int32_t* recv_counts,
                           int nranks) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && recv_counts && nranks >= 1 && nranks <= ROUTE_MAXP && n >= 0 && n < (1LL << 31), "bad argument");
    if (n > 0) {
        VPM_REQUIRE(sorted_to_recv, "NULL array");
        VPM_LAUNCH(k_mark_padding, (unsigned)ceil_div(n, 256), 256, 0, ctx->stream, sorted_to_recv, (int)n,
                   recv_counts, nranks);
    }
    VPM_API_END
}

int vpm_ctx_set_fill_row(vpm_ctx* ctx, const double* row3_host) {
    VPM_API_BEGIN
    VPM_REQUIRE(ctx && row3_host, "NULL argument");
    for (int c = 0; c < 3; ++c) ctx->fill_row[c] = row3_host[c];
    VPM_API_END
}

int vpm_scatter_add_rows(vpm_ctx* ctx, const double* src, const int32_t* idx, int64_t n, int ncomp,
                         double* dst) {
    return vpm_scatter_add_rows_scaled(ctx, src, idx, n, ncomp, 1.0, dst);
}

int vpm_scatter_add_rows_scaled(vpm_ctx* ctx, const double* src, const int32_t* idx, int64_t n, int ncomp,
                                double scale, double* dst) {
    VPM_API_BEGIN
    VPM_PROF_GROUP("scatter_add_rows", (16.0 * ncomp + 4.0) * n);
    VPM_REQUIRE(ctx, "ctx is NULL");
    VPM_REQUIRE(ncomp >= 1, "ncomp must be >= 1");
    if (n > 0) {
        VPM_REQUIRE(src && idx && dst, "NULL array");
        unsigned grid = (unsigned)ceil_div(n * ncomp, 256);
        if (ncomp == 1) VPM_LAUNCH((k_scatter_rows<1, true>), grid, 256, 0, ctx->stream, src, idx, (long long)n, ncomp, scale, dst);
        else if (ncomp == 3) VPM_LAUNCH((k_scatter_rows<3, true>), grid, 256, 0, ctx->stream, src, idx, (long long)n, ncomp, scale, dst);
        else VPM_LAUNCH((k_scatter_rows<0, true>), grid, 256, 0, ctx->stream, src, idx, (long long)n, ncomp, scale, dst);
    }
    VPM_API_END
}

}  // extern "C"
