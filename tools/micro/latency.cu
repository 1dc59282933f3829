// Development microbenchmarks (not part of the product): latencies that bound the persistent detection kernel.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)

__device__ __forceinline__ unsigned ld_acq(const unsigned* p){unsigned r; asm volatile("ld.acquire.gpu.global.u32 %0,[%1];":"=r"(r):"l"(p):"memory"); return r;}

__global__ void barrier_kernel(unsigned* ctr, int iters, int mode, long long* out)
{
    long long t0 = clock64();
    for (int g = 1; g <= iters; g++)
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            if (mode & 1) __threadfence();
            atomicAdd(ctr, 1u);
            unsigned target = g * gridDim.x;
            while (ld_acq(ctr) < target) { if (mode & 2) __nanosleep(32); }
            if (mode & 1) __threadfence();
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = clock64() - t0;
}

__global__ void chase_kernel(const unsigned* next, int iters, long long* out, unsigned* sink)
{
    unsigned p = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) p = __ldcg(next + p);
    *out = clock64() - t0;
    *sink = p;
}
__global__ void atomic_chain_kernel(unsigned* a, int iters, long long* out, unsigned* sink)
{
    unsigned v = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) v = atomicAdd(a + (v & 1023u), 1u);
    *out = clock64() - t0;
    *sink = v;
}
__global__ void fence_kernel(unsigned* a, int iters, long long* out)
{
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a[threadIdx.x + i] = i; __threadfence(); }
    *out = clock64() - t0;
}
__global__ void empty_kernel() {}

int main()
{
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    printf("%s SMs=%d clock=%d kHz L2=%d MB\n", prop.name, prop.multiProcessorCount, clk_khz, prop.l2CacheSize >> 20);
    unsigned *ctr, *next, *sink; long long* out;
    CK(cudaMalloc(&ctr, 4096 * 4)); CK(cudaMalloc(&sink, 4)); CK(cudaMallocManaged(&out, 8));
    const int N = 1 << 20;
    CK(cudaMalloc(&next, N * 4));
    unsigned* h = new unsigned[N];
    for (int i = 0; i < N; i++) h[i] = (unsigned)(((long long)i * 40503 + 12345) % N);
    CK(cudaMemcpy(next, h, N * 4, cudaMemcpyHostToDevice));
    for (int threads : {32, 1024})
        for (int mode = 0; mode < 4; mode++)
        {
            CK(cudaMemset(ctr, 0, 4));
            int iters = 200; int grid = prop.multiProcessorCount;
            void* args[] = {&ctr, &iters, &mode, &out};
            CK(cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(grid), dim3(threads), args, 0, 0));
            CK(cudaDeviceSynchronize());
            printf("grid barrier: %d CTAs x %d thr, fence=%d sleep=%d : %.0f cycles/barrier\n", grid, threads, mode & 1, (mode >> 1) & 1, (double)*out / iters);
        }
    chase_kernel<<<1, 1>>>(next, 2000, out, sink); CK(cudaDeviceSynchronize());
    printf("dependent ld.cg chain (4 MB footprint, L2): %.0f cycles/load\n", (double)*out / 2000);
    CK(cudaMemset(ctr, 0, 4096 * 4));
    atomic_chain_kernel<<<1, 1>>>(ctr, 2000, out, sink); CK(cudaDeviceSynchronize());
    printf("dependent atomicAdd chain: %.0f cycles/atomic\n", (double)*out / 2000);
    fence_kernel<<<1, 32>>>(ctr, 500, out); CK(cudaDeviceSynchronize());
    printf("store + __threadfence: %.0f cycles\n", (double)*out / 500);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 10; i++) empty_kernel<<<1, 32>>>();
    cudaEventRecord(e0); for (int i = 0; i < 1000; i++) empty_kernel<<<148, 256>>>(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("back-to-back empty kernel launches: %.2f us each\n", ms);
    return 0;
}
