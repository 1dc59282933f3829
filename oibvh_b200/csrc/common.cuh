// Shared device/host helpers for the oibvh_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oibvh
{

constexpr int kNumSMsB200 = 148;

// ------------------------------------------------------------------------------------------------
// Exact scalar semantics of the reference (so node AABBs are bit-identical, including signed zeros):
//   glm::min(x, y) = (y < x) ? y : x ; glm::max(x, y) = (x < y) ? y : x   (third/glm/detail/func_common.inl:17-30)
// fminf/fmaxf differ on (+0, -0) and NaN, so they are NOT used for boxes.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float gmin(float x, float y) { return (y < x) ? y : x; }
__host__ __device__ __forceinline__ float gmax(float x, float y) { return (x < y) ? y : x; }

struct Box
{
    float lx, ly, lz, hx, hy, hz;
};

// merge_aabb(left, right) = (glm::min(left.min, right.min), glm::max(left.max, right.max))  src/cuda/oibvh.cu:72-78
__device__ __forceinline__ Box box_merge(const Box& l, const Box& r)
{
    Box o;
    o.lx = gmin(l.lx, r.lx);
    o.ly = gmin(l.ly, r.ly);
    o.lz = gmin(l.lz, r.lz);
    o.hx = gmax(l.hx, r.hx);
    o.hy = gmax(l.hy, r.hy);
    o.hz = gmax(l.hz, r.hz);
    return o;
}

// aabb_box_t::overlap, inclusive on every axis (include/utils/utils.h:39-44, src/cuda/collide.cu:12-17)
__device__ __forceinline__ bool box_overlap(const Box& a, const Box& b)
{
    return (a.lx <= b.hx && a.hx >= b.lx) && (a.ly <= b.hy && a.hy >= b.ly) && (a.lz <= b.hz && a.hz >= b.lz);
}

// ------------------------------------------------------------------------------------------------
// Implicit-layout arithmetic, re-derived (not the reference's implicit<->real mapping functions):
// the ostensibly-implicit tree over T leaves is a complete binary tree of leaf level L = ceil(log2 T)
// whose nodes lying wholly to the right of leaf T-1 are dropped. Level l therefore keeps its first
//     cnt(l) = ceil(T / 2^(L-l))
// nodes, and because the array is the BFS order of the kept nodes, level l is the contiguous slice
//     [off(l), off(l) + cnt(l)),   off(l) = sum_{j<l} cnt(j).
// A node is addressed as (level, pos); its children are (level+1, 2pos) -- always kept -- and
// (level+1, 2pos+1), kept iff 2pos+1 < cnt(level+1). Equivalent to include/cuda/oibvh.cuh:56-182
// (tests/test_layout.py checks off/cnt against the reference's implicit_to_real for every node).
// Closed form used on the device: with vl = 2^L - T, v(l) = vl >> (L-l) nodes are dropped at level l and
// 2 v(l-1) - popc(v(l-1)) in the levels above l, so off(l) = 2^l - 1 - 2 v(l-1) + popc(v(l-1)).
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ceil_log2_u32(uint32_t t)
{
#ifdef __CUDA_ARCH__
    return t <= 1 ? 0u : 32u - __clz(t - 1);
#else
    uint32_t l = 0;
    while ((1ull << l) < t) l++;
    return l;
#endif
}

__host__ __device__ __forceinline__ uint32_t level_count(uint32_t T, uint32_t L, uint32_t l)
{
    const uint32_t s = L - l;
    return (uint32_t)(((uint64_t)T + ((1ull << s) - 1)) >> s);
}

__host__ __device__ __forceinline__ uint32_t level_offset(uint32_t T, uint32_t L, uint32_t l)
{
    if (l == 0) return 0;
    const uint32_t vl = (uint32_t)((1ull << L) - T);
    const uint32_t v = vl >> (L - l + 1);
#ifdef __CUDA_ARCH__
    const uint32_t pc = __popc(v);
#else
    const uint32_t pc = (uint32_t)__builtin_popcount(v);
#endif
    return (uint32_t)((1ull << l) - 1) - 2 * v + pc;
}

__host__ __device__ __forceinline__ uint32_t tree_size(uint32_t T)
{
    const uint32_t L = ceil_log2_u32(T);
    return level_offset(T, L, L) + T;
}

// ------------------------------------------------------------------------------------------------
// Memory helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming (read-once) loads that do not pollute L1
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// relaxed/acquire-release accesses at gpu scope for inter-CTA flags
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p)
{
    uint32_t r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p)
{
    uint32_t r;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// A 24-byte node is three 8-byte words; node arrays are 8-byte aligned at every index.
__device__ __forceinline__ Box load_box(const float2* __restrict__ nodes, uint32_t idx)
{
    const float2* p = nodes + 3ull * idx;
    const float2 a = p[0], b = p[1], c = p[2];
    Box o;
    o.lx = a.x; o.ly = a.y; o.lz = b.x; o.hx = b.y; o.hy = c.x; o.hz = c.y;
    return o;
}
__device__ __forceinline__ void store_box(float2* __restrict__ nodes, uint32_t idx, const Box& o)
{
    float2* p = nodes + 3ull * idx;
    p[0] = make_float2(o.lx, o.ly);
    p[1] = make_float2(o.lz, o.hx);
    p[2] = make_float2(o.hy, o.hz);
}

// ------------------------------------------------------------------------------------------------
// Block / grid collectives
// ------------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_totals /* THREADS/32 */)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) warp_totals[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++)
        if ((uint32_t)w < warp) base += warp_totals[w];
    __syncthreads();
    return base + inc - v;
}


// Grid-wide barrier for cooperatively launched kernels: monotonically increasing arrival counter (zeroed before the
// launch), `generation` = number of barriers including this one. One thread per CTA fences, arrives and spins;
// the fence is cumulative over the CTA barrier, so every thread's earlier global writes are visible to every
// thread of the grid afterwards (readers must bypass L1: __ldcg / ld.relaxed.gpu). A bounded spin turns a would-be
// hang into a failure flag.
// `participants` = number of CTAs that arrive at this counter (the whole grid, or one job's CTA range).
__device__ __forceinline__ void grid_sync(uint32_t* counter, uint32_t generation, uint32_t* fail_flag,
                                          uint32_t participants)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        // release-arrive / acquire-poll: the release is cumulative over the CTA barrier above, the acquire orders
        // everything after the barrier below -- no separate membar round trips
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        const uint32_t target = generation * participants;
        uint32_t spins = 0;
        // relaxed polls (an acquire load invalidates the SM's L1 on EVERY iteration, under the feet of the CTA that
        // shares the SM), one acquire fence once the count is reached
        while (ld_relaxed_gpu(counter) < target)
        {
            if (++spins > (1u << 26))
            {
                atomicOr(fail_flag, 1u); // surfaced by the host (oibvh_ctx_synchronize / downloads / scene_add_tree)
                break;
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

} // namespace oibvh
