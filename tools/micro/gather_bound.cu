// Development microbenchmark (not part of the product): what bounds the key / emit kernels at T = 2^20?
// Each thread reads `per` consecutive indices (coalesced, 128-bit) and gathers one 16-byte vertex per index from a
// V x float4 array (8.4 MB at V = 2^19: L2-resident), in three index orders:
//   random      -- shuffled faces (the key kernel's case)
//   local       -- lane-consecutive faces reference vertices within a few 128-byte lines (sorted faces on a mesh whose
//                  vertex numbering follows the surface)
//   sequential  -- perfectly coalesced (the lower bound of the LSU path)
// Reports ns per launch and SM cycles per gathered vertex per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)

template <int PER>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ idx4, const float4* __restrict__ pos,
                                                     float* __restrict__ out, unsigned n4)
{
    float acc = 0.f;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
    {
        uint4 q[PER / 4];
#pragma unroll
        for (int k = 0; k < PER / 4; k++) q[k] = idx4[(size_t)i * (PER / 4) + k];
        float4 v[PER];
#pragma unroll
        for (int k = 0; k < PER / 4; k++)
        {
            v[4 * k] = __ldg(pos + q[k].x); v[4 * k + 1] = __ldg(pos + q[k].y);
            v[4 * k + 2] = __ldg(pos + q[k].z); v[4 * k + 3] = __ldg(pos + q[k].w);
        }
#pragma unroll
        for (int k = 0; k < PER; k++) acc += v[k].x + v[k].y * v[k].z;
    }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const unsigned V = 525312, G = 3u << 20; // gathers per launch = 3 vertices x 2^20 faces
    std::vector<unsigned> h(G);
    float4* pos; unsigned* idx; float* out;
    CK(cudaMalloc(&pos, sizeof(float4) * V)); CK(cudaMalloc(&idx, 4 * G)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(pos, 0, sizeof(float4) * V));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[3] = {"random", "local", "sequential"};
    for (int mode = 0; mode < 3; mode++)
    {
        srand(1);
        for (unsigned i = 0; i < G; i++)
        {
            if (mode == 0) h[i] = ((unsigned)rand() * 32768u + (unsigned)rand()) % V;
            else if (mode == 1) h[i] = ((i / 96) * 40 + (unsigned)rand() % 48) % V; // 96 gathers of a warp-step within ~6 lines
            else h[i] = i % V;
        }
        CK(cudaMemcpy(idx, h.data(), 4 * G, cudaMemcpyHostToDevice));
        for (int ctas = 2; ctas <= 8; ctas *= 2)
        {
            auto run = [&](int per) {
                const unsigned n4 = G / per;
                if (per == 4) gather_kernel<4><<<sms * ctas, 256>>>((const uint4*)idx, pos, out, n4);
                else if (per == 12) gather_kernel<12><<<sms * ctas, 256>>>((const uint4*)idx, pos, out, n4);
                else gather_kernel<24><<<sms * ctas, 256>>>((const uint4*)idx, pos, out, n4);
            };
            for (int per : {4, 12, 24})
            {
                for (int w = 0; w < 3; w++) run(per);
                CK(cudaEventRecord(e0));
                const int it = 20;
                for (int w = 0; w < it; w++) run(per);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double us = ms * 1e3 / it;
                printf("%-10s ctas/SM=%d gathers/thread=%2d : %7.2f us  %5.2f cycles per gather per SM\n", names[mode], ctas,
                       per, us, us * 1e-6 * clk * 1e3 / ((double)G / sms));
            }
        }
    }
    return 0;
}
