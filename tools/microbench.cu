// microbench.cu - B200 numbers that decide the table design (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Measures, with CUDA events: random u32 RED and random u32 gather throughput against
// tables of several sizes (HBM-resident vs L2-resident), and shared-memory atomic
// throughput under the bin distributions the featurize kernel sees.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t h)
{
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
    return h;
}

template <int PER>
__global__ void __launch_bounds__(256) red_kernel(uint32_t* table, uint64_t mask, uint64_t n)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i * PER < n; i += stride) {
#pragma unroll
        for (int j = 0; j < PER; ++j) atomicAdd(table + (mix64(i * PER + j) & mask), 1u);
    }
}

template <int PER>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ table, uint64_t mask, uint64_t n, uint32_t* out)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i * PER < n; i += stride) {
        uint32_t v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] = __ldg(table + (mix64(i * PER + j) & mask));
#pragma unroll
        for (int j = 0; j < PER; ++j) acc += v[j];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// mode 0: uniform over 136 bins; 1: 4 hot bins; 2: conflict-free (bin = lane); 3: 4 hot bins, warp-aggregated with match_any
template <int MODE>
__global__ void __launch_bounds__(256) smem_atomic_kernel(uint32_t* out, int iters)
{
    __shared__ uint32_t bins[544];
    for (int i = threadIdx.x; i < 544; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    uint64_t h = mix64(blockIdx.x * 256 + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
        h = h * 6364136223846793005ull + 1442695040888963407ull;
        uint32_t r = (uint32_t)(h >> 33);
        uint32_t b = MODE == 0 ? r % 136u : MODE == 2 ? (threadIdx.x & 31) : (r & 3u);
        if (MODE == 3) {
            uint32_t peers = __match_any_sync(0xffffffffu, b);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&bins[b], __popc(peers));
        } else {
            atomicAdd(&bins[b], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 136 && bins[threadIdx.x] == 0xFFFFFFFFu) out[0] = 1;
}

int main()
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const uint64_t n = 1ull << 31;
    uint32_t* out; CK(cudaMalloc(&out, 4));
    const uint64_t sizes_mb[] = { 16, 32, 64, 128, 512, 2048, 4096 };
    for (uint64_t mb : sizes_mb) {
        uint64_t entries = mb * 1024 * 1024 / 4;
        uint32_t* t; CK(cudaMalloc(&t, entries * 4)); CK(cudaMemset(t, 0, entries * 4));
        float ms;
        red_kernel<8><<<148 * 8, 256>>>(t, entries - 1, n / 4); // warm
        CK(cudaEventRecord(a)); red_kernel<8><<<148 * 8, 256>>>(t, entries - 1, n); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("RED    table %5llu MiB: %7.2f G updates/s (%.2f ms for 2^31)\n", (unsigned long long)mb, n / ms / 1e6, ms);
        gather_kernel<32><<<148 * 8, 256>>>(t, entries - 1, n / 4, out);
        CK(cudaEventRecord(a)); gather_kernel<32><<<148 * 8, 256>>>(t, entries - 1, n, out); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("GATHER table %5llu MiB: %7.2f G loads/s   (%.2f ms for 2^31)\n", (unsigned long long)mb, n / ms / 1e6, ms);
        CK(cudaFree(t));
    }
    const int iters = 4096;
    const double total = 148.0 * 8 * 256 * iters;
    float ms;
#define SM(MODE, label) \
    smem_atomic_kernel<MODE><<<148 * 8, 256>>>(out, 16); \
    CK(cudaEventRecord(a)); smem_atomic_kernel<MODE><<<148 * 8, 256>>>(out, iters); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); \
    CK(cudaEventElapsedTime(&ms, a, b)); printf("SMEM atomics %-34s: %8.2f G/s\n", label, total / ms / 1e6);
    SM(0, "uniform over 136 bins");
    SM(1, "4 hot bins");
    SM(2, "conflict-free (bin = lane)");
    SM(3, "4 hot bins, match_any aggregated");
    CK(cudaDeviceSynchronize());
    return 0;
}
