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

    // row p (global index) of array a: from the staged tile, else straight from global memory
    __device__ __forceinline__ const double* row(int a, const double* global, long long p0, long long p) const {
        return staged ? buf[a] + 3 * (p - p0) : global + 3 * p;
    }
};

}  // namespace vpm
