// microbench2.cu - second round of B200 numbers behind the table design (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
// 1. random u32 gather from an L2-resident 32 MiB slice through: __ldg, ld.global.cg, ld.global.nc.L1::no_allocate,
//    tex1Dfetch (texture path) - is the 1 sector / clk / SM limit of LDG shared by the texture pipe?
// 2. shared-memory atomics over a 2^15-counter table (the smem-resident sub-slice of a two-level count): returning and not
// 3. shared-memory returning atomics over 65 bins (the scatter kernels' slot hand-out), plain and replicated x4
// 4. random LDS over 2^15 words
// 5. DSMEM: red.shared::cluster into the 8 x 128 KB of an 8-CTA cluster
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t h)
{
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
    return h;
}
__device__ __forceinline__ uint32_t lcg(uint64_t& h) { h = h * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(h >> 33); }

template <int MODE, int PER>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ table, cudaTextureObject_t tex, uint32_t mask, uint64_t n, uint32_t* out)
{
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i * PER < n; i += stride) {
        uint32_t v[PER];
        uint64_t h = mix64(i);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const uint32_t idx = lcg(h) & mask;
            if (MODE == 0) v[j] = __ldg(table + idx);
            else if (MODE == 1) v[j] = __ldcg(table + idx);
            else if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v[j]) : "l"(table + idx));
            else v[j] = tex1Dfetch<uint32_t>(tex, (int)idx);
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) acc += v[j];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// MODE 0: non-returning atomics over 2^15 counters, 1: returning, 2: random LDS
template <int MODE>
__global__ void __launch_bounds__(1024) smem_table_kernel(uint32_t* out, int iters)
{
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = i;
    __syncthreads();
    uint64_t h = mix64(blockIdx.x * 1024 + threadIdx.x);
    uint32_t acc = 0;
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t idx = lcg(h) & 32767u;
            if (MODE == 0) atomicAdd(tab + idx, 1u);
            else if (MODE == 1) acc += atomicAdd(tab + idx, 1u);
            else acc += tab[idx];
        }
    }
    __syncthreads();
    if (acc == 0x12345678u || tab[threadIdx.x] == 0xFFFFFFFFu) out[0] = acc;
}

// returning atomics over 65 bins; REP copies of the counters selected by lane
template <int REP>
__global__ void __launch_bounds__(256) smem_slot_kernel(uint32_t* out, int iters)
{
    __shared__ uint32_t cnt[65 * REP];
    __shared__ uint32_t stage[65 * 176];
    for (int i = threadIdx.x; i < 65 * REP; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    uint64_t h = mix64(blockIdx.x * 256 + threadIdx.x);
    const uint32_t rep = (threadIdx.x & (REP - 1)) * 65;
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t r = lcg(h);
            const uint32_t b = (r & 127u) < 20u ? 64u : (r >> 26); // ~15 % of the windows go to the dummy
            const uint32_t slot = atomicAdd(cnt + rep + b, 1u) % 176u;
            stage[b * 176 + slot] = r;
        }
    }
    __syncthreads();
    if (stage[threadIdx.x] == 0x12345678u) out[0] = 1;
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(512) dsmem_kernel(uint32_t* out, int iters)
{
    extern __shared__ uint32_t tab[];
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = 0;
    cluster.sync();
    uint64_t h = mix64(blockIdx.x * 512 + threadIdx.x);
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t r = lcg(h);
            uint32_t* remote = cluster.map_shared_rank(tab, (r >> 15) & 7u);
            atomicAdd(remote + (r & 32767u), 1u);
        }
    }
    cluster.sync();
    if (tab[threadIdx.x] == 0xFFFFFFFFu) out[0] = 1;
}

int main()
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    uint32_t* out; CK(cudaMalloc(&out, 4));
    float ms;
    {
        const uint64_t n = 1ull << 31;
        const uint64_t entries = 32ull * 1024 * 1024 / 4;
        uint32_t* t; CK(cudaMalloc(&t, entries * 4)); CK(cudaMemset(t, 0, entries * 4));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = t;
        rd.res.linear.desc = cudaCreateChannelDesc<uint32_t>(); rd.res.linear.sizeInBytes = entries * 4;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
#define GA(MODE, label) \
        gather_kernel<MODE, 32><<<148 * 8, 256>>>(t, tex, (uint32_t)entries - 1, n / 4, out); \
        CK(cudaEventRecord(a)); gather_kernel<MODE, 32><<<148 * 8, 256>>>(t, tex, (uint32_t)entries - 1, n, out); CK(cudaEventRecord(b)); \
        CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); printf("GATHER 32 MiB %-28s: %7.2f G loads/s\n", label, n / ms / 1e6);
        GA(0, "__ldg");
        GA(1, "ld.global.cg");
        GA(2, "ld.global.nc.L1::no_allocate");
        GA(3, "tex1Dfetch");
        CK(cudaDestroyTextureObject(tex)); CK(cudaFree(t));
    }
    {
        const int iters = 8192;
        const double total = 148.0 * 1024 * iters;
        CK(cudaFuncSetAttribute(smem_table_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        CK(cudaFuncSetAttribute(smem_table_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        CK(cudaFuncSetAttribute(smem_table_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
#define ST(MODE, label) \
        smem_table_kernel<MODE><<<148, 1024, 131072>>>(out, 64); \
        CK(cudaEventRecord(a)); smem_table_kernel<MODE><<<148, 1024, 131072>>>(out, iters); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); \
        CK(cudaEventElapsedTime(&ms, a, b)); printf("SMEM 2^15-word table %-22s: %8.2f G/s\n", label, total / ms / 1e6);
        ST(0, "atomicAdd (no return)");
        ST(1, "atomicAdd (returning)");
        ST(2, "random LDS");
    }
    {
        const int iters = 4096;
        const double total = 148.0 * 4 * 256 * iters;
#define SL(REP, label) \
        smem_slot_kernel<REP><<<148 * 4, 256>>>(out, 64); \
        CK(cudaEventRecord(a)); smem_slot_kernel<REP><<<148 * 4, 256>>>(out, iters); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); \
        CK(cudaEventElapsedTime(&ms, a, b)); printf("SMEM slot hand-out (ATOMS ret + STS), 65 bins %-12s: %8.2f G/s\n", label, total / ms / 1e6);
        SL(1, "x1");
        SL(2, "x2");
        SL(4, "x4");
    }
    {
        const int iters = 4096;
        const double total = 144.0 * 512 * iters;
        CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        dsmem_kernel<<<144, 512, 131072>>>(out, 64);
        CK(cudaEventRecord(a)); dsmem_kernel<<<144, 512, 131072>>>(out, iters); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
        printf("DSMEM red over an 8-CTA cluster (8 x 128 KB)      : %8.2f G/s\n", total / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
