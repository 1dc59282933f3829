// Does a global line fetched ahead of time hit in L1 on sm_100a? One warp per SM; for each trial a fresh 128-byte line
// (never touched before) is (a) not prefetched, (b) prefetched with prefetch.global.L1, (c) touched with a real
// ld.global.ca / (d) ld.global.nc whose result is not used, (e) prefetched with prefetch.global.L2; after 3000 cycles
// the line is read with a generic load (what the traversal does), an LDG (.ca) and an LDG.CONSTANT (.nc), and the load
// latency is measured with clock64. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_prefetch l1_prefetch.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void spin(long long cycles)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
}
template <int AHEAD, int LOADK, int SPIN>
__global__ void probe(const float2* base, size_t stride_f2, int trials, long long* out)
{
    if (threadIdx.x >= 32) return;
    long long sum = 0;
    float acc = 0.f;
    for (int t = 0; t < trials; t++)
    {
        const float2* p = base + ((size_t)(blockIdx.x * trials + t) * stride_f2) + (threadIdx.x & 7) * 3; // 8 boxes x 24 B
        if (AHEAD == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        if (AHEAD == 2) { float s; asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(s) : "l"(p) : "memory"); }
        if (AHEAD == 3) { float s; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(s) : "l"(p) : "memory"); }
        if (AHEAD == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        if (AHEAD == 5)
        {
            // a real load that IS waited for, by the whole warp, right before the timed one
            float2 w;
            asm volatile("ld.global.ca.v2.f32 {%0, %1}, [%2];" : "=f"(w.x), "=f"(w.y) : "l"(p) : "memory");
            acc += w.x * 1e-30f + (float)(__ballot_sync(0xffffffffu, w.y == 3.f) & 0);
        }
        spin(SPIN);
        const long long t0 = clock64();
        float2 v;
        if (LOADK == 0)
        {
            const float2* q; // laundered: the compiler no longer knows the address space -> generic LD
            asm volatile("mov.u64 %0, %1;" : "=l"(q) : "l"(p));
            v = *q;
        }
        if (LOADK == 1) asm volatile("ld.global.ca.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
        if (LOADK == 2) asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
        if (LOADK == 3) { v.x = 1.f; v.y = (float)threadIdx.x; } // no load at all: the overhead of the measurement
        acc += v.x + v.y;
        const unsigned done = __ballot_sync(0xffffffffu, acc == 12345.f); // waits for the load
        sum += clock64() - t0 + (done & 0);
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sum / trials;
    if (acc == 98765.f) out[0] = 0;
}
template <int AHEAD, int LOADK, int SPIN = 3000>
static void run(const char* name, const float2* buf, size_t stride, long long* d_out, bool cold)
{
    const int sms = 148, trials = 64;
    if (cold) cudaMemset((void*)buf, 0, stride * 8 * sms * trials); // lines in L2 (just written) -- or evicted below
    probe<AHEAD, LOADK, SPIN><<<sms, 32>>>(buf, stride, trials, d_out);
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long s = 0;
    for (int i = 0; i < sms; i++) s += h[i];
    printf("%-44s %6lld cycles\n", name, s / sms);
}
int main()
{
    const size_t stride = 512; // float2 per trial: 4 KB apart
    float2* buf;
    long long* d_out;
    cudaMalloc(&buf, stride * 8 * 148 * 64);
    cudaMalloc(&d_out, 148 * sizeof(long long));
    printf("lines resident in L2 (written by a memset just before):\n");
    run<0, 0>("no prefetch, generic load", buf, stride, d_out, true);
    run<0, 1>("no prefetch, ld.global.ca", buf, stride, d_out, true);
    run<1, 0>("prefetch.global.L1, generic load", buf, stride, d_out, true);
    run<1, 1>("prefetch.global.L1, ld.global.ca", buf, stride, d_out, true);
    run<1, 2>("prefetch.global.L1, ld.global.nc", buf, stride, d_out, true);
    run<2, 0>("touch ld.global.ca, generic load", buf, stride, d_out, true);
    run<2, 1>("touch ld.global.ca, ld.global.ca", buf, stride, d_out, true);
    run<3, 2>("touch ld.global.nc, ld.global.nc", buf, stride, d_out, true);
    run<3, 0>("touch ld.global.nc, generic load", buf, stride, d_out, true);
    run<4, 0>("prefetch.global.L2, generic load", buf, stride, d_out, true);
    run<5, 1, 200>("line loaded + waited for, 200 cycles later: ld.global.ca", buf, stride, d_out, true);
    run<5, 0, 200>("line loaded + waited for, 200 cycles later: generic", buf, stride, d_out, true);
    run<5, 0, 3000>("line loaded + waited for, 3000 cycles later: generic", buf, stride, d_out, true);
    run<5, 0, 20000>("line loaded + waited for, 20000 cycles later: generic", buf, stride, d_out, true);
    run<1, 0, 200>("prefetch.global.L1, 200 cycles later: generic", buf, stride, d_out, true);
    run<1, 0, 600>("prefetch.global.L1, 600 cycles later: generic", buf, stride, d_out, true);
    run<1, 0, 1000>("prefetch.global.L1, 1000 cycles later: generic", buf, stride, d_out, true);
    run<2, 0, 600>("touch ld.global.ca, 600 cycles later: generic", buf, stride, d_out, true);
    run<2, 0, 1000>("touch ld.global.ca, 1000 cycles later: generic", buf, stride, d_out, true);
    run<3, 0, 1000>("touch ld.global.nc, 1000 cycles later: generic", buf, stride, d_out, true);
    run<0, 3>("no load (measurement overhead)", buf, stride, d_out, true);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
