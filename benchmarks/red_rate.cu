// Reduction-rate microbenchmark behind the paint design (sm_100a):
//   how many fp64 reductions per second does the memory system take, as a function of how the
//   32 addresses of one warp instruction are laid out, and what does a shared-memory fp64
//   atomicAdd (a CAS loop on this architecture) cost?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o benchmarks/_bin/red_rate benchmarks/red_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// every element of out[0..n) receives exactly `rounds` reductions; a warp instruction covers
// 32 * stride consecutive elements, lane l -> base + l * stride (+ phase over the stride)
__global__ void k_red(double* out, long long n, int stride, int rounds) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long warp = tid >> 5, nwarps = nthreads >> 5;
    const long long span = 32LL * stride;
    for (int r = 0; r < rounds; ++r)
        for (long long b = warp * span; b + span <= n; b += nwarps * span)
            for (int ph = 0; ph < stride; ++ph) atomicAdd(out + b + (long long)lane * stride + ph, 1.0);
}

__global__ void k_store(double* out, long long n) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) out[i] = 1.0;
}

// CIC-like: thread per "cell" (sorted, one particle per cell), 8 corner reductions
__global__ void k_red_cic(double* out, int n0, int n1, int n2, int corners) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long nc = (long long)n0 * n1 * n2;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += nthreads) {
        const int iz = (int)(c % n2), i1 = (int)((c / n2) % n1), i0 = (int)(c / ((long long)n1 * n2));
        const int jz = iz + 1 < n2 ? iz + 1 : 0, j1 = i1 + 1 < n1 ? i1 + 1 : 0, j0 = i0 + 1 < n0 ? i0 + 1 : 0;
        const long long ra = (long long)i0 * n1 * n2, rb = (long long)j0 * n1 * n2;
        const long long ca = (long long)i1 * n2, cb = (long long)j1 * n2;
        atomicAdd(out + ra + ca + iz, 1.0);
        if (corners > 1) atomicAdd(out + ra + ca + jz, 1.0);
        if (corners > 2) { atomicAdd(out + ra + cb + iz, 1.0); atomicAdd(out + ra + cb + jz, 1.0); }
        if (corners > 4) {
            atomicAdd(out + rb + ca + iz, 1.0); atomicAdd(out + rb + ca + jz, 1.0);
            atomicAdd(out + rb + cb + iz, 1.0); atomicAdd(out + rb + cb + jz, 1.0);
        }
    }
}

// shared-memory fp64 atomicAdd: each thread performs `iters` x 8 adds into a tile of `tile`
// doubles; pattern 0: lane-consecutive addresses (conflict-free), 1: pairs of lanes share an
// address (run length 2), 2: pseudo-random within the tile
__global__ void k_smem_atomic(double* out, int tile, int iters, int pattern) {
    extern __shared__ double s[];
    for (int i = threadIdx.x; i < tile; i += blockDim.x) s[i] = 0.0;
    __syncthreads();
    unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            int idx;
            if (pattern == 0) idx = (w * 67 + it * 32 + lane + c * 129) % tile;
            else if (pattern == 1) idx = (w * 67 + it * 16 + (lane >> 1) + c * 129) % tile;
            else { h = h * 1664525u + 1013904223u; idx = (h >> 8) % tile; }
            atomicAdd(&s[idx], 1.0);
        }
    }
    __syncthreads();
    double acc = 0.0;
    for (int i = threadIdx.x; i < tile; i += blockDim.x) acc += s[i];
    if (acc == -1.0) out[0] = acc;
}

template <class F>
static float timeit(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 512;
    const long long n = (long long)N * N * N;
    double* out;
    CK(cudaMalloc(&out, n * sizeof(double)));
    CK(cudaMemset(out, 0, n * sizeof(double)));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 8, block = 256;
    printf("mesh %d^3 (%.1f MB), %d SMs\n", N, n * 8e-6, sms);
    float ms = timeit([&] { k_store<<<grid, block>>>(out, n); });
    printf("plain store            : %8.3f ms  %7.1f G elem/s  %7.1f GB/s\n", ms, n / ms * 1e-6, n * 8 / ms * 1e-6);
    ms = timeit([&] { CK(cudaMemsetAsync(out, 0, n * sizeof(double))); });
    printf("memset                 : %8.3f ms  %7.1f GB/s\n", ms, n * 8 / ms * 1e-6);
    for (int stride : {1, 2, 4, 8, 16}) {
        ms = timeit([&] { k_red<<<grid, block>>>(out, n, stride, 1); });
        printf("RED lane stride %2d     : %8.3f ms  %7.1f G RED/s\n", stride, ms, n / ms * 1e-6);
    }
    ms = timeit([&] { k_red<<<grid, block>>>(out, n, 1, 4); });
    printf("RED stride 1, 4 rounds : %8.3f ms  %7.1f G RED/s\n", ms, 4 * n / ms * 1e-6);
    for (int corners : {1, 2, 4, 8}) {
        ms = timeit([&] { k_red_cic<<<grid, block>>>(out, N, N, N, corners); });
        printf("CIC-like %d corners/cell: %8.3f ms  %7.1f G RED/s\n", corners, ms, corners * (double)n / ms * 1e-6);
    }
    const int tile = 8192, iters = 256;
    CK(cudaFuncSetAttribute(k_smem_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, tile * 8));
    for (int pattern = 0; pattern < 3; ++pattern) {
        const int g2 = sms * 3;
        ms = timeit([&] { k_smem_atomic<<<g2, 512, tile * 8>>>(out, tile, iters, pattern); });
        printf("smem atomicAdd(f64) pattern %d: %8.3f ms  %7.1f G adds/s\n", pattern, ms,
               (double)g2 * 512 * iters * 8 / ms * 1e-6);
    }
    return 0;
}
