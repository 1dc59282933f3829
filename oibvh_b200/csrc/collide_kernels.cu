// Broad phase (BVTT front expansion) and narrow phase (17-axis SAT) kernels for sm_100a.
//
// Reference behaviour being reproduced (not its code):
//   seeds + per-level traversal  src/cuda/scene.cu:192-223, 272-311 ; src/cuda/collide.cu:76-246
//   candidate = both leaves with inclusively overlapping boxes       src/cuda/collide.cu:155-162
//   narrow phase = SAT over n1, m1, e_i x f_j, e_i x n1, f_j x m1     src/utils/utils.cpp:71-169
//                                                                    (== third/gProximity/cuda_intersect_tritri.h:350-434)
// Design: the front lives in device memory as 16-byte (objA, objB, nodeA, nodeB) records with nodes addressed as
// (level, position) so no implicit<->real conversion is needed; every warp tests 32 front nodes, prefix-sums the
// fan-outs with shuffles, claims ONE range of the next front with a single atomic, and then writes that range
// cooperatively (32 consecutive 16-byte records per store instruction). Front sizes never visit the host.
#include "common.cuh"
#include "kernels.h"

namespace oibvh
{

constexpr int kColThreads = 256;
constexpr int kColWarps = kColThreads / 32;

__device__ __forceinline__ ObjDesc load_obj(const ObjDesc* __restrict__ objs, uint32_t i)
{
    // 32-byte descriptor as two 128-bit read-only loads
    const uint4* p = reinterpret_cast<const uint4*>(objs + i);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    ObjDesc d;
    d.nodes = reinterpret_cast<const float*>(((uint64_t)a.y << 32) | a.x);
    d.faces = reinterpret_cast<const uint32_t*>(((uint64_t)a.w << 32) | a.z);
    d.pos = reinterpret_cast<const float*>(((uint64_t)b.y << 32) | b.x);
    d.T = b.z;
    d.L = b.w;
    return d;
}

// ---------------------------------------------------------------------------------------------------
// Seeds: one (root, root) BVTT node per object pair i < j, at a deterministic slot so that shards agree.
// Round 0 of the expansion turns them into the reference's entry-level seed rectangle (scene.cu:192-223)
// and prunes object pairs whose root boxes do not overlap.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seed_kernel(uint32_t n_obj, uint4* __restrict__ front, uint32_t front_cap,
                                                   uint32_t* __restrict__ counters)
{
    const uint64_t n_pairs = (uint64_t)n_obj * (n_obj - 1) / 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += stride)
    {
        // p -> (i, j), i < j, row-major over the strict upper triangle
        const double nd = (double)n_obj;
        uint32_t i = (uint32_t)floor((2.0 * nd - 1.0 - sqrt((2.0 * nd - 1.0) * (2.0 * nd - 1.0) - 8.0 * (double)p)) * 0.5);
        // fix rounding of the closed form
        while ((uint64_t)i * (2ull * n_obj - i - 1) / 2 > p) i--;
        while ((uint64_t)(i + 1) * (2ull * n_obj - i - 2) / 2 <= p) i++;
        const uint32_t j = (uint32_t)(p - (uint64_t)i * (2ull * n_obj - i - 1) / 2) + i + 1;
        if (p < front_cap) front[p] = make_uint4(i, j, 0u, 0u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        counters[CTR_FRONT0] = (uint32_t)min(n_pairs, (uint64_t)0xffffffffu);
        if (n_pairs > front_cap) atomicOr(counters + CTR_OVERFLOW, 1u);
    }
}

// ---------------------------------------------------------------------------------------------------
// One round: test every node of the front, emit candidates (leaf, leaf) or the children rectangle
// `levels` levels further down on each side (clamped to the leaf level and to the kept nodes of the level).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kColThreads)
    expand_kernel(const ObjDesc* __restrict__ objs, const uint4* __restrict__ in, uint4* __restrict__ out,
                  uint32_t front_cap, uint4* __restrict__ cand, uint32_t cand_cap, uint32_t* __restrict__ counters,
                  uint32_t round, uint32_t levels, uint32_t rank, uint32_t world)
{
    __shared__ uint32_t s_prefix[kColWarps][33];
    __shared__ uint4 s_item[kColWarps][32]; // objA, objB, first child of A (packed), first child of B (packed)
    __shared__ uint32_t s_nb[kColWarps][32];    // width of the children rectangle
    __shared__ uint32_t s_first[kColWarps][32]; // shard striding: child c = first + m * step
    __shared__ uint32_t s_step[kColWarps][32];

    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t n = min(counters[CTR_FRONT0 + round], front_cap);
    uint32_t* next_count = counters + CTR_FRONT0 + round + 1;
    const uint32_t total_warps = gridDim.x * kColWarps;

    for (uint32_t base = (blockIdx.x * kColWarps + warp) * 32; base < n; base += total_warps * 32)
    {
        const uint32_t i = base + lane;
        const bool active = i < n;
        uint32_t n_children = 0;
        bool is_cand = false;
        uint4 it = make_uint4(0, 0, 0, 0);
        uint32_t child_a = 0, child_b = 0, nB = 1, first = 0, step = 1;
        if (active)
        {
            it = in[i];
            const ObjDesc A = load_obj(objs, it.x), B = load_obj(objs, it.y);
            const uint32_t la = it.z >> kNodeLevelShift, pa = it.z & kNodePosMask;
            const uint32_t lb = it.w >> kNodeLevelShift, pb = it.w & kNodePosMask;
            const Box a = load_box(reinterpret_cast<const float2*>(A.nodes), level_offset(A.T, A.L, la) + pa);
            const Box b = load_box(reinterpret_cast<const float2*>(B.nodes), level_offset(B.T, B.L, lb) + pb);
            if (box_overlap(a, b))
            {
                if (la == A.L && lb == B.L)
                    is_cand = true;
                else
                {
                    const uint32_t da = min(levels, A.L - la), db = min(levels, B.L - lb);
                    const uint32_t fa = pa << da, fb = pb << db;
                    const uint32_t nA = min(1u << da, level_count(A.T, A.L, la + da) - fa);
                    nB = min(1u << db, level_count(B.T, B.L, lb + db) - fb);
                    child_a = ((la + da) << kNodeLevelShift) | fa;
                    child_b = ((lb + db) << kNodeLevelShift) | fb;
                    const uint32_t all = nA * nB;
                    if (world > 1 && round == 0)
                    {
                        // shard the seed rectangle round-robin; (i + c) % world == rank keeps child c
                        first = (rank + world - (i % world)) % world;
                        step = world;
                        n_children = first < all ? (all - first + world - 1) / world : 0;
                    }
                    else
                        n_children = all;
                }
            }
        }

        // ---- candidates: one atomic per warp ----
        const uint32_t cmask = __ballot_sync(0xffffffffu, is_cand);
        if (cmask)
        {
            uint32_t cbase = 0;
            if (lane == 0) cbase = atomicAdd(counters + CTR_CANDIDATES, (uint32_t)__popc(cmask));
            cbase = __shfl_sync(0xffffffffu, cbase, 0);
            if (is_cand)
            {
                const uint32_t dst = cbase + __popc(cmask & lanemask_lt());
                if (dst < cand_cap)
                    cand[dst] = make_uint4(it.x, it.y, it.z & kNodePosMask, it.w & kNodePosMask);
                else
                    atomicOr(counters + CTR_OVERFLOW, 2u);
            }
        }

        // ---- children: warp prefix sum, one atomic, cooperative coalesced emission ----
        uint32_t inc = n_children;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (total == 0) continue;
        s_prefix[warp][lane] = inc - n_children;
        if (lane == 31) s_prefix[warp][32] = total;
        s_item[warp][lane] = make_uint4(it.x, it.y, child_a, child_b);
        s_nb[warp][lane] = nB;
        s_first[warp][lane] = first;
        s_step[warp][lane] = step;
        uint32_t obase = 0;
        if (lane == 0) obase = atomicAdd(next_count, total);
        obase = __shfl_sync(0xffffffffu, obase, 0);
        __syncwarp();
        if (obase + total > front_cap && lane == 0) atomicOr(counters + CTR_OVERFLOW, 1u);
        for (uint32_t j = lane; j < total; j += 32)
        {
            // owner lane: largest s with prefix[s] <= j
            uint32_t s = 0;
#pragma unroll
            for (int bit = 16; bit > 0; bit >>= 1)
                if (s_prefix[warp][s + bit] <= j) s += bit;
            const uint32_t m = j - s_prefix[warp][s];
            const uint32_t nb = s_nb[warp][s];
            const uint32_t c = s_first[warp][s] + m * s_step[warp][s];
            const uint4 src = s_item[warp][s];
            if (obase + j < front_cap) out[obase + j] = make_uint4(src.x, src.y, src.z + c / nb, src.w + c % nb);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// Narrow phase: separating-axis test, literal operation order of the reference CPU code, IEEE fp32 with
// explicit round-to-nearest intrinsics (nvcc would otherwise contract a*b - c*d into FMAs and flip
// touching cases relative to the CPU path).
// ---------------------------------------------------------------------------------------------------
struct V3
{
    float x, y, z;
};
__device__ __forceinline__ V3 vsub(const V3& a, const V3& b)
{
    return V3{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)};
}
// glm::cross(x, y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
__device__ __forceinline__ V3 vcross(const V3& x, const V3& y)
{
    return V3{__fsub_rn(__fmul_rn(x.y, y.z), __fmul_rn(y.y, x.z)), __fsub_rn(__fmul_rn(x.z, y.x), __fmul_rn(y.z, x.x)),
              __fsub_rn(__fmul_rn(x.x, y.y), __fmul_rn(y.x, x.y))};
}
// glm::dot(a, b) = (a.x*b.x + a.y*b.y) + a.z*b.z
__device__ __forceinline__ float vdot(const V3& a, const V3& b)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}

struct TriPair
{
    V3 p1, p2, p3, q1, q2, q3;
};
// project6: true = the axis does NOT separate (strict comparisons: touching counts as intersecting, a zero
// axis never separates)
__device__ __forceinline__ bool axis_overlaps(const V3& ax, const TriPair& t)
{
    const float P1 = vdot(ax, t.p1), P2 = vdot(ax, t.p2), P3 = vdot(ax, t.p3);
    const float Q1 = vdot(ax, t.q1), Q2 = vdot(ax, t.q2), Q3 = vdot(ax, t.q3);
    const float mx1 = fmaxf(fmaxf(P1, P2), P3), mn1 = fminf(fminf(P1, P2), P3);
    const float mx2 = fmaxf(fmaxf(Q1, Q2), Q3), mn2 = fminf(fminf(Q1, Q2), Q3);
    if (mn1 > mx2) return false;
    if (mn2 > mx1) return false;
    return true;
}

__device__ bool triangles_intersect(const V3& P1, const V3& P2, const V3& P3, const V3& Q1, const V3& Q2, const V3& Q3)
{
    TriPair t;
    t.p1 = V3{0.0f, 0.0f, 0.0f};
    t.p2 = vsub(P2, P1);
    t.p3 = vsub(P3, P1);
    t.q1 = vsub(Q1, P1);
    t.q2 = vsub(Q2, P1);
    t.q3 = vsub(Q3, P1);
    V3 e[3], f[3];
    e[0] = vsub(t.p2, t.p1);
    e[1] = vsub(t.p3, t.p2);
    e[2] = vsub(t.p1, t.p3);
    f[0] = vsub(t.q2, t.q1);
    f[1] = vsub(t.q3, t.q2);
    f[2] = vsub(t.q1, t.q3);
    const V3 n1 = vcross(e[0], e[1]);
    const V3 m1 = vcross(f[0], f[1]);
    if (!axis_overlaps(n1, t)) return false;
    if (!axis_overlaps(m1, t)) return false;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (!axis_overlaps(vcross(e[i], f[j]), t)) return false;
#pragma unroll
    for (int i = 0; i < 3; i++)
        if (!axis_overlaps(vcross(e[i], n1), t)) return false;
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (!axis_overlaps(vcross(f[j], m1), t)) return false;
    return true;
}

__device__ __forceinline__ V3 load_v3(const float* __restrict__ pos, uint32_t v)
{
    const float* p = pos + 3ull * v;
    return V3{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}

__global__ void __launch_bounds__(kColThreads)
    narrow_kernel(const ObjDesc* __restrict__ objs, const uint4* __restrict__ cand, uint32_t cand_cap,
                  uint4* __restrict__ pairs, uint32_t pair_cap, uint32_t* __restrict__ counters)
{
    const uint32_t lane = lane_id();
    const uint32_t n = min(counters[CTR_CANDIDATES], cand_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounded = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += stride)
    {
        bool hit = false;
        uint4 c = make_uint4(0, 0, 0, 0);
        if (i < n)
        {
            c = cand[i];
            const ObjDesc A = load_obj(objs, c.x), B = load_obj(objs, c.y);
            const uint32_t* fa = A.faces + 3ull * c.z;
            const uint32_t* fb = B.faces + 3ull * c.w;
            const V3 P1 = load_v3(A.pos, __ldg(fa)), P2 = load_v3(A.pos, __ldg(fa + 1)), P3 = load_v3(A.pos, __ldg(fa + 2));
            const V3 Q1 = load_v3(B.pos, __ldg(fb)), Q2 = load_v3(B.pos, __ldg(fb + 1)), Q3 = load_v3(B.pos, __ldg(fb + 2));
            hit = triangles_intersect(P1, P2, P3, Q1, Q2, Q3);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
        if (mask)
        {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(counters + CTR_PAIRS, (uint32_t)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit)
            {
                const uint32_t dst = base + __popc(mask & lanemask_lt());
                if (dst < pair_cap)
                    pairs[dst] = c; // {bvhA, bvhB, triA, triB} == int_tri_pair_node_t
                else
                    atomicOr(counters + CTR_OVERFLOW, 4u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------------
static inline uint32_t persistent_grid(uint64_t work_items, uint32_t per_block, uint32_t blocks_per_sm)
{
    uint64_t b = (work_items + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t)kNumSMsB200 * blocks_per_sm;
    if (b > cap) b = cap;
    if (b == 0) b = 1;
    return (uint32_t)b;
}

cudaError_t launch_seed(uint32_t n_obj, uint4* front, uint32_t front_cap, uint32_t* counters, cudaStream_t s)
{
    const uint64_t n_pairs = (uint64_t)n_obj * (n_obj - 1) / 2;
    seed_kernel<<<persistent_grid(n_pairs, 256, 8), 256, 0, s>>>(n_obj, front, front_cap, counters);
    return cudaGetLastError();
}

cudaError_t launch_expand(const ObjDesc* objs, const uint4* in, uint4* out, uint32_t front_cap, uint4* cand,
                          uint32_t cand_cap, uint32_t* counters, uint32_t round, uint32_t levels, uint32_t rank,
                          uint32_t world, uint32_t grid_hint, cudaStream_t s)
{
    expand_kernel<<<persistent_grid(grid_hint, kColThreads, 8), kColThreads, 0, s>>>(
        objs, in, out, front_cap, cand, cand_cap, counters, round, levels, rank, world);
    return cudaGetLastError();
}

cudaError_t launch_narrow(const ObjDesc* objs, const uint4* cand, uint32_t cand_cap, uint4* pairs,
                          uint32_t pair_cap, uint32_t* counters, uint32_t grid_hint, cudaStream_t s)
{
    narrow_kernel<<<persistent_grid(grid_hint, kColThreads, 8), kColThreads, 0, s>>>(objs, cand, cand_cap, pairs,
                                                                                      pair_cap, counters);
    return cudaGetLastError();
}

} // namespace oibvh
