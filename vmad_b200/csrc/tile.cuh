// Particle tiles staged through shared memory by the TMA engine (sm_100a).
//
// Every particle kernel reads rows of [np, 3] fp64 arrays (positions, cotangents): a thread
// that loads its own row issues three 8-byte loads with a 24-byte stride, i.e. a warp touches
// 768 contiguous bytes with three instructions of six L1 wavefronts each.  Here the CTA's tile
// of rows -- one contiguous block of global memory -- is copied into shared memory by ONE bulk
// asynchronous copy (cp.async.bulk.shared::cluster.global, 1-D TMA: 16-byte granular, fully
// coalesced, no register staging, no L1 traffic), completion is signalled on an mbarrier, and
// the threads then read their rows from shared memory (8-byte reads at a 24-byte stride are
// bank-conflict free: the 16 lanes of a half-warp hit 16 distinct even banks).
// Tiles that are not whole, or not 16-byte aligned, fall back to plain loads.
//
// Where this pays (measured): contiguous tiles that LEAVE an SM as a whole -- the
// output tile of the row gather (put / flush: one cp.async.bulk.global.shared::cta per CTA, 1.9x).
// In front of a gather whose addresses depend on the staged rows (the readouts) the CTA-wide wait
// costs more than the strided loads it replaces; there the staged load is an ablation variant only.
#pragma once
#include <stdint.h>

namespace vpm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    // make the initialised barrier visible to the asynchronous (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    while (!mbar_try_wait(bar, phase)) {
    }
}

// One CTA = one tile of T rows of up to two [np, 3] arrays.  Usage:
//     __shared__ RowTile<T, 2> tile;
//     tile.load(a, b, p0, np);              // all threads; b may be nullptr
//     const double* r = tile.row(0, a, p);   // the thread's row p of array 0 (shared or global)
template <int T, int NARR>
struct RowTile {
    alignas(16) double buf[NARR][3 * T];
    alignas(8) uint64_t bar;
    int staged;         // 1: the tile sits in buf (written by thread 0 before the barrier)

    // stage rows [p0, p0 + T) of the arrays (nullptr entries are skipped) when the tile is whole
    // and 16-byte aligned; must be called by every thread of the CTA exactly once
    __device__ __forceinline__ void load(const double* const (&arr)[NARR], long long p0, long long np) {
        if (threadIdx.x == 0) {
            bool ok = p0 + T <= np;
#pragma unroll
            for (int a = 0; a < NARR; ++a)
                if (arr[a]) ok = ok && ((reinterpret_cast<uintptr_t>(arr[a] + 3 * p0) & 15) == 0);
            staged = ok ? 1 : 0;
            if (ok) {
                mbar_init(&bar, 1);
                uint32_t bytes = 0;
#pragma unroll
                for (int a = 0; a < NARR; ++a)
                    if (arr[a]) bytes += 3 * T * (uint32_t)sizeof(double);
                mbar_expect_tx(&bar, bytes);
#pragma unroll
                for (int a = 0; a < NARR; ++a)
                    if (arr[a]) tma_load_1d(buf[a], arr[a] + 3 * p0, 3 * T * (uint32_t)sizeof(double), &bar);
            }
        }
        __syncthreads();
        if (staged) mbar_wait(&bar, 0);
    }

    // row p (global index) of array a: from the staged tile (LDS), else straight from global memory
    __device__ __forceinline__ void get(int a, const double* __restrict__ global, long long p0, long long p,
                                        double& v0, double& v1, double& v2) const {
        if (staged) {
            const double* r = buf[a] + 3 * (int)(p - p0);
            v0 = r[0]; v1 = r[1]; v2 = r[2];
        } else {
            const double* r = global + 3 * p;
            v0 = __ldg(r); v1 = __ldg(r + 1); v2 = __ldg(r + 2);
        }
    }

    // ---- the way back: rows written by their threads into buf[a], stored as ONE bulk copy -------
    // put(): every thread with a row writes it (threads without a row do nothing);
    // flush(): called by every thread of the CTA once, after all put()s of the arrays flushed.
    // The tile must be whole and the destination 16-byte aligned (`can_flush`), else the caller
    // stores its row itself.
    __device__ __forceinline__ static bool can_flush(const double* dst, long long p0, long long np) {
        return p0 + T <= np && (reinterpret_cast<uintptr_t>(dst + 3 * p0) & 15) == 0;
    }
    __device__ __forceinline__ void put(int a, int r, double v0, double v1, double v2) {
        double* q = buf[a] + 3 * r;
        q[0] = v0; q[1] = v1; q[2] = v2;
    }
    __device__ __forceinline__ void flush(int a, double* dst, long long p0) {
        // generic-proxy writes to shared memory -> visible to the async proxy, then one thread issues
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + 3 * p0),
                         "r"(smem_u32(buf[a])), "r"(3 * T * (uint32_t)sizeof(double))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the source may be released (CTA exit / re-use) once it has been read
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
};

// ---- scattered rows of an [n, 3] array (home order reached through a permutation) -----------
// A 24-byte row is 16-byte aligned at either its first or its second element: one 16-byte and
// one 8-byte access instead of three 8-byte ones (a third fewer L1/L2 transactions; for stores,
// which L1 does not merge, a third fewer partial-sector writes).
__device__ __forceinline__ void load_row3(const double* __restrict__ base, long long o, double& a, double& b,
                                          double& c) {
    const double* r = base + 3 * o;
    if ((reinterpret_cast<uintptr_t>(r) & 15) == 0) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(r));
        a = t.x; b = t.y; c = __ldg(r + 2);
    } else {
        a = __ldg(r);
        const double2 t = __ldg(reinterpret_cast<const double2*>(r + 1));
        b = t.x; c = t.y;
    }
}

// the same with streaming (evict-first) hints: rows that are touched exactly once should not push
// the mesh lines, which every neighbouring particle re-reads, out of L2
__device__ __forceinline__ void load_row3_cs(const double* __restrict__ base, long long o, double& a, double& b,
                                             double& c) {
    const double* r = base + 3 * o;
    a = __ldcs(r); b = __ldcs(r + 1); c = __ldcs(r + 2);
}

__device__ __forceinline__ void store_row3_cs(double* __restrict__ base, long long o, double a, double b, double c) {
    double* r = base + 3 * o;
    __stcs(r, a); __stcs(r + 1, b); __stcs(r + 2, c);
}

__device__ __forceinline__ void store_row3(double* __restrict__ base, long long o, double a, double b, double c) {
    double* r = base + 3 * o;
    if ((reinterpret_cast<uintptr_t>(r) & 15) == 0) {
        *reinterpret_cast<double2*>(r) = make_double2(a, b);
        r[2] = c;
    } else {
        r[0] = a;
        *reinterpret_cast<double2*>(r + 1) = make_double2(b, c);
    }
}

}  // namespace vpm
