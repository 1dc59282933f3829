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

// One persistent cooperative kernel runs the whole detection: seeds -> every expansion round -> narrow phase, with a
// grid-wide barrier between phases, so a detection costs one launch and the per-round latency is one barrier
// (~1 us) instead of a kernel boundary. One CTA of 1024 threads per SM (148 arrivals per barrier).
constexpr int kColThreads = 1024;
constexpr int kColWarps = kColThreads / 32;

__device__ __forceinline__ ObjDesc load_obj(const ObjDesc* __restrict__ objs, uint32_t i)
{
    // 32-byte descriptor as two 128-bit read-only loads
    const uint4* p = reinterpret_cast<const uint4*>(objs + i);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    ObjDesc d;
    d.nodes = reinterpret_cast<const float*>(((uint64_t)a.y << 32) | a.x);
    d.faces = reinterpret_cast<const uint32_t*>(((uint64_t)a.w << 32) | a.z);
    d.pos = reinterpret_cast<const float4*>(((uint64_t)b.y << 32) | b.x);
    d.T = b.z;
    d.L = b.w;
    return d;
}

// ---------------------------------------------------------------------------------------------------
// Seeds: one (root, root) BVTT node per object pair i < j, at a deterministic slot so that shards agree.
// Round 0 of the expansion turns them into the reference's entry-level seed rectangle (scene.cu:192-223)
// and prunes object pairs whose root boxes do not overlap.
// ---------------------------------------------------------------------------------------------------
// p-th object pair (i < j), row-major over the strict upper triangle of the n_obj x n_obj matrix
__device__ __forceinline__ uint4 seed_entry(uint32_t n_obj, uint64_t p)
{
    const double nd = (double)n_obj;
    uint32_t i = (uint32_t)floor((2.0 * nd - 1.0 - sqrt((2.0 * nd - 1.0) * (2.0 * nd - 1.0) - 8.0 * (double)p)) * 0.5);
    while ((uint64_t)i * (2ull * n_obj - i - 1) / 2 > p) i--; // fix the rounding of the closed form
    while ((uint64_t)(i + 1) * (2ull * n_obj - i - 2) / 2 <= p) i++;
    const uint32_t j = (uint32_t)(p - (uint64_t)i * (2ull * n_obj - i - 1) / 2) + i + 1;
    return make_uint4(i, j, 0u, 0u);
}

// Many-body seeding = the top-level pass: one THREAD per object pair i < j tests the two root boxes, and only
// overlapping pairs enter the front (warp-aggregated append). The reference seeds every pair on the host, O(n^2)
// entries (src/cuda/scene.cu:192-223); here 4096 objects are 8.4 M cheap root tests and a front of a few thousand.
__device__ void seed_phase(const ObjDesc* __restrict__ objs, uint32_t n_obj, uint4* __restrict__ front,
                           uint32_t front_cap, uint32_t* __restrict__ counters)
{
    const uint64_t n_pairs = (uint64_t)n_obj * (n_obj - 1) / 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounded = (n_pairs + 31) & ~31ull;
    const uint32_t lane = lane_id();
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < rounded; p += stride)
    {
        bool hit = false;
        uint4 e = make_uint4(0, 0, 0, 0);
        if (p < n_pairs)
        {
            e = seed_entry(n_obj, p);
            const ObjDesc A = load_obj(objs, e.x), B = load_obj(objs, e.y);
            hit = box_overlap(load_box(reinterpret_cast<const float2*>(A.nodes), 0),
                              load_box(reinterpret_cast<const float2*>(B.nodes), 0));
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, hit);
        if (mask)
        {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(counters + CTR_FRONT0, (uint32_t)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit)
            {
                const uint32_t dst = base + __popc(mask & lanemask_lt());
                if (dst < front_cap)
                    front[dst] = e;
                else
                    atomicOr(counters + CTR_OVERFLOW, 1u);
            }
        }
    }
}

// linear index of the object pair (i, j), i < j: the inverse of seed_entry (deterministic shard key)
__device__ __forceinline__ uint32_t pair_linear(uint32_t n_obj, uint32_t i, uint32_t j)
{
    return (uint32_t)((uint64_t)i * (2ull * n_obj - i - 1) / 2) + (j - i - 1);
}

// Grid-wide barrier on a monotonically increasing arrival counter (zeroed with the counter block before the
// launch). `generation` counts the barriers passed so far. The kernel is launched cooperatively, so every CTA is
// resident and the spin terminates; a bounded spin turns a would-be hang into a reported failure.
__device__ __forceinline__ uint32_t grid_barrier(uint32_t* __restrict__ counters, uint32_t generation,
                                                 const uint32_t* read_after)
{
    // returns *read_after as settled after the barrier, read ONCE per CTA and broadcast through shared memory
    // (every warp of the chip polling the same word would serialise on one L2 slice)
    __shared__ uint32_t s_value;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        // release-arrive / acquire-poll (see grid_sync in common.cuh)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counters + CTR_BARRIER) : "memory");
        const uint32_t target = generation * gridDim.x;
        uint32_t spins = 0;
        while (ld_acquire_gpu(counters + CTR_BARRIER) < target)
        {
            if (++spins > (1u << 26))
            {
                atomicOr(counters + CTR_OVERFLOW, 8u);
                break;
            }
        }
        s_value = __ldcg(read_after);
    }
    __syncthreads();
    return s_value;
}

// ---------------------------------------------------------------------------------------------------
// One round of BVTT expansion, warp-cooperative.
//
// A front entry is a node pair whose boxes are KNOWN to overlap (round 0: the untested root pairs). One warp takes
// one pair, addresses the rectangle of descendants `levels` levels further down on each side (clamped to the leaf
// level and to the nodes the level keeps), and its 32 lanes test the nA x nB descendant box pairs -- the 2^k + 2^k
// boxes are two contiguous slices of the level arrays, so the loads are broadcast / L1 hits. Only overlapping
// descendant pairs are emitted: to the next front, or to the candidate list when both sides reached the leaf level
// (a candidate is by definition a leaf pair with overlapping boxes, src/cuda/collide.cu:155-162).
// Emission is staged per warp in shared memory and flushed with ONE global atomic per ~200 records and
// fully coalesced 16-byte stores; the grid is persistent (grid-stride over the front).
// ---------------------------------------------------------------------------------------------------
constexpr int kStageCap = 256; // records per warp staging buffer (4 KB)

constexpr int kObjCache = 64; // object descriptors + level tables kept in shared memory (more: L1/L2 + arithmetic)

__device__ __forceinline__ ObjDesc get_obj(const ObjDesc* s_objs, const ObjDesc* __restrict__ objs, uint32_t i)
{
    return i < (uint32_t)kObjCache ? s_objs[i] : load_obj(objs, i);
}

// level geometry of an object: from the shared-memory tables for cached objects, from arithmetic otherwise
struct LevelView
{
    const uint32_t* off; // null -> compute
    const uint32_t* cnt;
    uint32_t T, L;
    __device__ __forceinline__ uint32_t offset(uint32_t l) const { return off ? off[l] : level_offset(T, L, l); }
    __device__ __forceinline__ uint32_t count(uint32_t l) const { return cnt ? cnt[l] : level_count(T, L, l); }
};
__device__ __forceinline__ LevelView level_view(const uint32_t* s_lv, uint32_t obj, const ObjDesc& d)
{
    LevelView v;
    v.T = d.T;
    v.L = d.L;
    v.off = obj < (uint32_t)kObjCache ? s_lv + obj * 64 : nullptr;
    v.cnt = obj < (uint32_t)kObjCache ? s_lv + obj * 64 + 32 : nullptr;
    return v;
}

__device__ void expand_phase(uint4* s_stage, const ObjDesc* s_objs, const uint32_t* s_lv,
                             const ObjDesc* __restrict__ objs, const uint4* in, uint4* out, uint32_t front_cap,
                             uint4* cand, uint32_t cand_cap, uint32_t* counters, uint32_t round, uint32_t front_size,
                             uint32_t levels, uint32_t rank, uint32_t world, uint32_t n_obj)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    // round 0 of a scene with few object pairs computes its entries (root pairs) instead of reading a seeded front
    const bool computed_seeds = (round == 0 && in == nullptr);
    const uint32_t n = computed_seeds ? front_size : min(front_size, front_cap);
    uint32_t* next_count = counters + CTR_FRONT0 + round + 1;
    const uint32_t total_warps = gridDim.x * kColWarps;
    uint4* stage = s_stage + warp * kStageCap;
    uint32_t staged = 0;       // warp-uniform
    bool staged_cand = false;  // warp-uniform: what the staged records are

    auto flush = [&]() {
        if (staged == 0) return;
        __syncwarp();
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(staged_cand ? counters + CTR_CANDIDATES : next_count, staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint4* dst = staged_cand ? cand : out;
        const uint32_t cap = staged_cand ? cand_cap : front_cap;
        if (base + staged > cap && lane == 0) atomicOr(counters + CTR_OVERFLOW, staged_cand ? 2u : 1u);
        for (uint32_t j = lane; j < staged; j += 32)
            if (base + j < cap) dst[base + j] = stage[j];
        __syncwarp();
        staged = 0;
    };

    // A work item is one 64-combination group of one pair's descendant rectangle (4^levels combinations at most):
    // a 3-level rectangle is one item, the seed round (4..5 levels) is spread over several warps. Each lane tests
    // two combinations per item, so four box loads per lane are in flight together.
    const uint32_t gshift = 2 * levels > 6 ? 2 * levels - 6 : 0; // log2(groups per pair)
    const uint32_t items = n << gshift;                          // n < 2^26, gshift <= 4
    uint32_t w = blockIdx.x * kColWarps + warp;
    // fronts are produced by other SMs inside this same launch: read them through L2 (ld.cg); the entry of the
    // NEXT item is fetched before the current one is processed (one round trip off the critical path)
    uint4 it_next = make_uint4(0, 0, 0, 0);
    auto entry = [&](uint32_t p) { return computed_seeds ? seed_entry(n_obj, p) : __ldcg(in + p); };
    if (w < items) it_next = entry(w >> gshift);
    for (; w < items; w += total_warps)
    {
        const uint4 it = it_next;
        const uint32_t p = w >> gshift, group = w & ((1u << gshift) - 1);
        if (w + total_warps < items) it_next = entry((w + total_warps) >> gshift);
        const ObjDesc A = get_obj(s_objs, objs, it.x), B = get_obj(s_objs, objs, it.y);
        const LevelView va = level_view(s_lv, it.x, A), vb = level_view(s_lv, it.y, B);
        const uint32_t la = it.z >> kNodeLevelShift, pa = it.z & kNodePosMask;
        const uint32_t lb = it.w >> kNodeLevelShift, pb = it.w & kNodePosMask;
        const float2* nodesA = reinterpret_cast<const float2*>(A.nodes);
        const float2* nodesB = reinterpret_cast<const float2*>(B.nodes);
        const uint32_t da = min(levels, A.L - la), db = min(levels, B.L - lb);
        const uint32_t lca = la + da, lcb = lb + db;
        const uint32_t fa = pa << da, fb = pb << db;
        const uint32_t nA = min(1u << da, va.count(lca) - fa);
        const uint32_t nB = min(1u << db, vb.count(lcb) - fb);
        const uint32_t combos = nA * nB;
        if (group * 64 >= combos) continue; // warp-uniform: clamped rectangle, nothing in this group
        const uint32_t baseA = va.offset(lca) + fa, baseB = vb.offset(lcb) + fb;
        const uint32_t c0 = group * 64 + lane, c1 = c0 + 32;
        bool hit0 = c0 < combos, hit1 = c1 < combos;
        // round 0: the seed rectangle is dealt round-robin to the shards
        if (world > 1 && round == 0)
        {
            // keyed by the object pair's linear index, not by the (nondeterministic) slot of a seeded entry
            const uint32_t key = computed_seeds ? p : pair_linear(n_obj, it.x, it.y);
            hit0 = hit0 && ((key + c0) % world) == rank;
            hit1 = hit1 && ((key + c1) % world) == rank;
        }
        // nB is a power of two unless the rectangle is clamped by the end of the level (warp-uniform either way)
        uint32_t ia0, ib0, ia1, ib1;
        if (nB == (1u << db))
        {
            ia0 = c0 >> db;
            ib0 = c0 & (nB - 1);
            ia1 = c1 >> db;
            ib1 = c1 & (nB - 1);
        }
        else
        {
            ia0 = c0 / nB;
            ib0 = c0 - ia0 * nB;
            ia1 = c1 / nB;
            ib1 = c1 - ia1 * nB;
        }
        Box a0, b0, a1, b1;
        if (hit0)
        {
            a0 = load_box(nodesA, baseA + ia0);
            b0 = load_box(nodesB, baseB + ib0);
        }
        if (hit1)
        {
            a1 = load_box(nodesA, baseA + ia1);
            b1 = load_box(nodesB, baseB + ib1);
        }
        if (computed_seeds)
        {
            // computed root pairs are untested: prune object pairs whose root boxes are disjoint (seeded fronts
            // were tested by the seeding pass; the loads overlap the ones above)
            const Box ra = load_box(nodesA, va.offset(la) + pa);
            const Box rb = load_box(nodesB, vb.offset(lb) + pb);
            if (!box_overlap(ra, rb)) continue; // warp-uniform
        }
        if (hit0) hit0 = box_overlap(a0, b0);
        if (hit1) hit1 = box_overlap(a1, b1);
        const bool to_cand = (lca == A.L) && (lcb == B.L);
        if (staged && to_cand != staged_cand) flush();
        staged_cand = to_cand;
        const uint32_t mask0 = __ballot_sync(0xffffffffu, hit0), mask1 = __ballot_sync(0xffffffffu, hit1);
        const uint32_t cnt0 = __popc(mask0), cnt = cnt0 + __popc(mask1);
        if (staged + cnt > (uint32_t)kStageCap) flush();
        const uint32_t na = lca << kNodeLevelShift, nb = lcb << kNodeLevelShift;
        if (hit0)
            stage[staged + __popc(mask0 & lanemask_lt())] =
                to_cand ? make_uint4(it.x, it.y, fa + ia0, fb + ib0) : make_uint4(it.x, it.y, na | (fa + ia0), nb | (fb + ib0));
        if (hit1)
            stage[staged + cnt0 + __popc(mask1 & lanemask_lt())] =
                to_cand ? make_uint4(it.x, it.y, fa + ia1, fb + ib1) : make_uint4(it.x, it.y, na | (fa + ia1), nb | (fb + ib1));
        staged += cnt;
    }
    flush();
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

__device__ __forceinline__ V3 load_v3(const float4* __restrict__ pos, uint32_t v)
{
    const float4 p = __ldg(pos + v);
    return V3{p.x, p.y, p.z};
}

__device__ void narrow_phase(const ObjDesc* s_objs, const ObjDesc* __restrict__ objs, const uint4* cand, uint32_t cand_cap,
                             uint4* __restrict__ pairs, uint32_t pair_cap, uint32_t* counters, uint32_t n_cand)
{
    const uint32_t lane = lane_id();
    const uint32_t n = min(n_cand, cand_cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounded = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += stride)
    {
        bool hit = false;
        uint4 c = make_uint4(0, 0, 0, 0);
        if (i < n)
        {
            c = __ldcg(cand + i);
            const ObjDesc A = get_obj(s_objs, objs, c.x), B = get_obj(s_objs, objs, c.y);
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
// The persistent detection kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kColThreads, 1)
    collide_kernel(const ObjDesc* __restrict__ objs, uint32_t n_obj, uint4* front0, uint4* front1, uint32_t front_cap,
                   uint4* cand, uint32_t cand_cap, uint4* pairs, uint32_t pair_cap, uint32_t* counters,
                   uint32_t rounds, uint32_t levels0, uint32_t levels, uint32_t rank, uint32_t world)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* s_stage = reinterpret_cast<uint4*>(smem_raw); // kColWarps x kStageCap records
    __shared__ __align__(16) ObjDesc s_objs[kObjCache];
    __shared__ uint32_t s_lv[kObjCache * 64]; // per cached object: off[32], cnt[32]
    for (uint32_t i = threadIdx.x; i < min(n_obj, (uint32_t)kObjCache) * 32; i += blockDim.x)
    {
        const uint32_t o = i >> 5, l = i & 31u;
        const ObjDesc d = load_obj(objs, o);
        if (l == 0) s_objs[o] = d;
        s_lv[o * 64 + l] = l <= d.L ? level_offset(d.T, d.L, l) : 0u;
        s_lv[o * 64 + 32 + l] = l <= d.L ? level_count(d.T, d.L, l) : 0u;
    }
    __syncthreads();
    uint32_t gen = 0;
    uint32_t n_stamps = 0;
    auto stamp = [&](uint32_t) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && n_stamps < CTR_WORDS - CTR_TIME0)
            counters[CTR_TIME0 + n_stamps] = (uint32_t)clock64();
        n_stamps++;
    };
    stamp(0);
    // few object pairs (the common two-body scene): round 0 computes the root pairs on the fly, which saves the seeding
    // pass and its grid barrier; many-body scenes seed the front in parallel first
    const uint64_t n_pairs = (uint64_t)n_obj * (n_obj - 1) / 2;
    const bool computed_seeds = n_pairs <= 4096;
    uint32_t front_size;
    if (computed_seeds)
    {
        front_size = (uint32_t)n_pairs;
        if (blockIdx.x == 0 && threadIdx.x == 0) counters[CTR_FRONT0] = front_size;
    }
    else
    {
        seed_phase(objs, n_obj, front0, front_cap, counters);
        front_size = grid_barrier(counters, ++gen, counters + CTR_FRONT0);
    }
    stamp(1);
    for (uint32_t r = 0; r < rounds && front_size != 0; r++) // front_size is uniform over the grid
    {
        uint4* in = (r & 1) ? front1 : front0;
        uint4* out = (r & 1) ? front0 : front1;
        if (r == 0 && computed_seeds) in = nullptr;
        const uint32_t k = r == 0 ? levels0 : levels; // schedule chosen by the host (see scene_enqueue)
        expand_phase(s_stage, s_objs, s_lv, objs, in, out, front_cap, cand, cand_cap, counters, r, front_size, k, rank,
                     world, n_obj);
        front_size = grid_barrier(counters, ++gen, counters + CTR_FRONT0 + r + 1);
        stamp(gen);
    }
    const uint32_t n_cand = grid_barrier(counters, ++gen, counters + CTR_CANDIDATES);
    stamp(gen);
    narrow_phase(s_objs, objs, cand, cand_cap, pairs, pair_cap, counters, n_cand);
    __syncthreads();
    stamp(gen + 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CTR_TIME0 - 1] = n_stamps; // number of stamps
}

constexpr size_t kColSmemBytes = (size_t)kColWarps * kStageCap * sizeof(uint4); // 128 KB

cudaError_t collide_configure(int* grid_blocks)
{
    cudaError_t e = cudaFuncSetAttribute(collide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColSmemBytes);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0, dev = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, collide_kernel, kColThreads, kColSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid_blocks = sms; // one CTA per SM: the fewest barrier arrivals
    return cudaSuccess;
}

cudaError_t launch_collide(int grid_blocks, const ObjDesc* objs, uint32_t n_obj, uint4* front0, uint4* front1,
                           uint32_t front_cap, uint4* cand, uint32_t cand_cap, uint4* pairs, uint32_t pair_cap,
                           uint32_t* counters, uint32_t rounds, uint32_t levels0, uint32_t levels, uint32_t rank,
                           uint32_t world, cudaStream_t s)
{
    void* args[] = {&objs, &n_obj, &front0, &front1, &front_cap, &cand, &cand_cap, &pairs, &pair_cap,
                    &counters, &rounds, &levels0, &levels, &rank, &world};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(collide_kernel), dim3(grid_blocks),
                                       dim3(kColThreads), args, kColSmemBytes, s);
}

} // namespace oibvh
