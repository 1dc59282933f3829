// Broad phase (BVTT front expansion) and narrow phase (17-axis SAT) kernels for sm_100a.
//
// Reference behaviour being reproduced (not its code):
//   seeds + per-level traversal  src/cuda/scene.cu:192-223, 272-311 ; src/cuda/collide.cu:76-246
//   candidate = both leaves with inclusively overlapping boxes       src/cuda/collide.cu:155-162
//   narrow phase = SAT over n1, m1, e_i x f_j, e_i x n1, f_j x m1     src/utils/utils.cpp:71-169
//                                                                    (== third/gProximity/cuda_intersect_tritri.h:350-434)
// Design: BVTT nodes are 16-byte (objA, objB, nodeA, nodeB) records with tree nodes addressed as (level, position), so no
// implicit<->real conversion is needed. The front is ONE work queue in device memory that the persistent warps of a
// single cooperative launch fill and drain without grid barriers (see traverse_queue); leaf pairs whose boxes overlap go
// to a candidate list that dedicated warps of the same launch test while the traversal is still running (aux_loop).
// Nothing visits the host between the seeds and the pair list.
#include "common.cuh"
#include "kernels.h"

namespace oibvh
{

// One persistent cooperative kernel runs the whole detection: seeds -> queue-driven traversal with the narrow phase
// beside it; no grid barrier anywhere (cooperative launch only guarantees that every CTA is resident, which the
// queue's polling consumers rely on). One CTA of 768 threads per SM at 80 registers per thread.
#ifndef OIBVH_COL_THREADS
#define OIBVH_COL_THREADS 768
#endif
constexpr int kColThreads = OIBVH_COL_THREADS;
constexpr int kColWarps = kColThreads / 32;
// Warp roles during the traversal: the first kTravWarps warps of a CTA walk the BVTT; one CONTROL warp keeps the CTA's
// shared copy of the queue's control words fresh; the remaining kAuxWarps - 1 NARROW warps test candidates while the
// traversal is still producing them. Seeding before and the rest of the narrow phase after are done by all warps.
#ifndef OIBVH_COL_AUX_WARPS
#define OIBVH_COL_AUX_WARPS 3
#endif
constexpr int kAuxWarps = OIBVH_COL_AUX_WARPS;
constexpr int kTravWarps = kColWarps - kAuxWarps;
constexpr int kNarrowWarps = kAuxWarps - 1;
static_assert(kAuxWarps >= 2 && kTravWarps >= 4, "one control warp, at least one narrow warp, traversal warps");
static_assert(kColWarps % kNarrowWarps == 0, "the narrow warps' backlog is dealt evenly to the CTA's warps");

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

// Queue records are published and consumed with single-copy-atomic 128-bit accesses (PTX .b128, STG/LDG.E.128.STRONG.GPU
// on sm_100a): a consumer polling a slot sees either the empty marker or the whole record, so no per-record flag or
// fence is needed. A slot is EMPTY while its first word is kQEmpty (objects are numbered below 2^32 - 1).
constexpr uint32_t kQEmpty = 0xffffffffu;
__device__ __forceinline__ void st_rec(uint4* p, const uint4& v)
{
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p),
                 "l"(((uint64_t)v.y << 32) | v.x), "l"(((uint64_t)v.w << 32) | v.z)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_rec(const uint4* p)
{
    uint64_t lo, hi;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}"
                 : "=l"(lo), "=l"(hi)
                 : "l"(p)
                 : "memory");
    return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

// Queue state = ONE 64-bit word: pushed records in the high half (the tail: the next free slot), finished items in the
// low half. One atomic reserves slots, one atomic retires items, and because both hit the same address every
// observer sees an exact snapshot of (pushed, finished) -- the termination test needs no fence.
// Atomics on one address are served one at a time by its L2 slice (~1-2 cycles each): with a BVTT of 10^5 nodes, one
// reserve and one retire per node would cost more than the traversal itself (measured: 264 K cycles). So a warp
// keeps the items it has finished in a register and retires them TOGETHER with its next reservation (one atomic does
// both), or when it runs out of work; retiring late only delays the detection of the end, never fakes it.
__device__ __forceinline__ uint32_t queue_reserve(uint32_t* counters, uint32_t n, uint32_t retire = 0)
{
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(counters + CTR_Q_STATE),
                                             ((unsigned long long)n << 32) | retire);
    return (uint32_t)(old >> 32);
}
// retire n items; true when every record ever pushed has been retired (nothing in flight, nothing can be pushed).
// Seeding counts as one item in flight per CTA (a token that is never pushed: warp 0 of every CTA retires it after the
// CTA's seeding), so the traversal needs no grid barrier between seeding and walking: finished == pushed + gridDim.x.
__device__ __forceinline__ bool queue_retire(uint32_t* counters, uint32_t n)
{
    const unsigned long long now =
        atomicAdd(reinterpret_cast<unsigned long long*>(counters + CTR_Q_STATE), (unsigned long long)n) + n;
    return (uint32_t)(now >> 32) + gridDim.x == (uint32_t)now;
}

// ---------------------------------------------------------------------------------------------------
// Seeds: one (root, root) BVTT node per object pair i < j, at a deterministic slot so that shards agree.
// Round 0 of the expansion turns them into the reference's entry-level seed rectangle (scene.cu:192-223)
// and prunes object pairs whose root boxes do not overlap.
// ---------------------------------------------------------------------------------------------------
// p-th object pair (i < j), row-major over the strict upper triangle of the n_obj x n_obj matrix; with self-collision
// the n pairs (i, i) follow the n (n - 1) / 2 pairs of distinct objects
__device__ __forceinline__ uint4 seed_entry(uint32_t n_obj, uint64_t p)
{
    const uint64_t distinct = (uint64_t)n_obj * (n_obj - 1) / 2;
    if (p >= distinct) return make_uint4((uint32_t)(p - distinct), (uint32_t)(p - distinct), 0u, 0u);
    const double nd = (double)n_obj;
    uint32_t i = (uint32_t)floor((2.0 * nd - 1.0 - sqrt((2.0 * nd - 1.0) * (2.0 * nd - 1.0) - 8.0 * (double)p)) * 0.5);
    while ((uint64_t)i * (2ull * n_obj - i - 1) / 2 > p) i--; // fix the rounding of the closed form
    while ((uint64_t)(i + 1) * (2ull * n_obj - i - 2) / 2 <= p) i++;
    const uint32_t j = (uint32_t)(p - (uint64_t)i * (2ull * n_obj - i - 1) / 2) + i + 1;
    return make_uint4(i, j, 0u, 0u);
}

// Many-body seeding = the top-level pass: every object pair i < j has its two root boxes tested, and only
// overlapping pairs enter the front (warp-aggregated append). The reference seeds every pair on the host, O(n^2)
// entries (src/cuda/scene.cu:192-223); here 4096 objects are 8.4 M cheap root tests and a front of a few thousand.
// The n x n triangle is cut into B x B tiles (B = 256, smaller when that would leave CTAs without a tile); a CTA
// stages the 2B root boxes of its tile in shared memory; a thread keeps one column object in registers and walks the
// rows, so a pair costs two broadcast 128-bit shared-memory reads and the test, and a root box is fetched n / B
// times in total instead of once per pair.
__device__ void seed_phase(float* __restrict__ s_roots, const ObjDesc* __restrict__ objs, uint32_t n_obj,
                           uint4* __restrict__ front, uint32_t front_cap, uint32_t* __restrict__ counters, uint32_t rank,
                           uint32_t world)
{
    // multi-GPU: the TILES are dealt to the ranks (tile t belongs to rank t mod world), so the top-level pass itself
    // shrinks with the number of GPUs and an object pair lives on exactly one rank from the root down
    // (`front` is the work queue: a slot is claimed from its tail counter and filled with one 128-bit store)
    uint32_t bshift = 8;
    auto tiles_of = [&](uint32_t sh) {
        const uint32_t nb = (n_obj + (1u << sh) - 1) >> sh;
        return nb * (nb + 1) / 2;
    };
    while (bshift > 5 && tiles_of(bshift) < gridDim.x * world) bshift--;
    const uint32_t B = 1u << bshift, nb = (n_obj + B - 1) >> bshift, tiles = nb * (nb + 1) / 2;
    const uint32_t lane = lane_id();
    float4* sa = reinterpret_cast<float4*>(s_roots);          // [B][2] row-block roots: (lx ly lz hx) (hy hz - -)
    float4* sb = reinterpret_cast<float4*>(s_roots) + 2 * B;  // [B][2] column-block roots
    uint32_t bi = 0, row_first = 0; // tile t = row_first(bi) + (bj - bi), rows of nb - bi tiles
    for (uint32_t t = blockIdx.x * world + rank; t < tiles; t += gridDim.x * world)
    {
        while (t >= row_first + (nb - bi))
        {
            row_first += nb - bi;
            bi++;
        }
        const uint32_t bj = bi + (t - row_first);
        const uint32_t i0 = bi << bshift, j0 = bj << bshift;
        const uint32_t ni = min(B, n_obj - i0), nj = min(B, n_obj - j0);
        __syncthreads(); // the previous tile is done with the staging area
        for (uint32_t k = threadIdx.x; k < 2 * B; k += blockDim.x)
        {
            const bool col = k >= B;
            const uint32_t local = col ? k - B : k, o = (col ? j0 : i0) + local;
            if (local < (col ? nj : ni))
            {
                const ObjDesc d = load_obj(objs, o);
                const Box b = load_box(reinterpret_cast<const float2*>(d.nodes), 0);
                float4* dst = (col ? sb : sa) + 2 * local;
                dst[0] = make_float4(b.lx, b.ly, b.lz, b.hx);
                dst[1] = make_float4(b.hy, b.hz, 0.0f, 0.0f);
            }
        }
        __syncthreads();
        // a thread keeps ONE column object in registers and walks the rows: per pair two broadcast 128-bit reads
        const uint32_t jj = threadIdx.x & (B - 1);
        const bool col_ok = jj < nj;
        // (volatile: the column box must stay in registers instead of being re-read inside the loop)
        float4 b0, b1;
        {
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sb + 2 * jj);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w)
                         : "r"(addr));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+16];" : "=f"(b1.x), "=f"(b1.y) : "r"(addr));
        }
        const uint32_t row_step = blockDim.x >> bshift;
#pragma unroll 2
        for (uint32_t ii = threadIdx.x >> bshift; ii < B; ii += row_step) // CTA-uniform trip count
        {
            const float4 a0 = sa[2 * ii]; // warp-uniform addresses: broadcast reads
            const float2 a1 = *reinterpret_cast<const float2*>(sa + 2 * ii + 1);
            // diagonal tiles keep i < j only; no short-circuit, the loop body stays branch-free
            const bool hit = col_ok & (ii < ni) & ((bi != bj) | (ii < jj)) & (a0.x <= b0.w) & (a0.w >= b0.x) &
                             (a0.y <= b1.x) & (a1.x >= b0.y) & (a0.z <= b1.y) & (a1.y >= b0.z);
            const uint32_t mask = __ballot_sync(0xffffffffu, hit);
            if (mask)
            {
                uint32_t slot = 0;
                if (lane == 0) slot = queue_reserve(counters, (uint32_t)__popc(mask));
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (hit)
                {
                    const uint32_t dst = slot + __popc(mask & lanemask_lt());
                    if (dst < front_cap)
                        st_rec(front + dst, make_uint4(i0 + ii, j0 + jj, 0u, 0u));
                    else
                        atomicOr(counters + CTR_OVERFLOW, 1u);
                }
            }
        }
    }
}

// linear index of the object pair (i, j), i < j: the inverse of seed_entry (deterministic shard key)
__device__ __forceinline__ uint32_t pair_linear(uint32_t n_obj, uint32_t i, uint32_t j)
{
    if (i == j) return (uint32_t)((uint64_t)n_obj * (n_obj - 1) / 2) + i; // self-collision pair
    return (uint32_t)((uint64_t)i * (2ull * n_obj - i - 1) / 2) + (j - i - 1);
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

// system-scope accesses for the multi-GPU protocol words and the gathering rank's pair counter (peer memory)
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t r;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

constexpr int kObjCache = 64; // object descriptors + level tables kept in shared memory (more: L1/L2 + arithmetic)

__device__ __forceinline__ ObjDesc get_obj(const ObjDesc* s_objs, const ObjDesc* __restrict__ objs, uint32_t i)
{
    // (a compact shared-memory table of all objects was measured slower than these L1/L2 hits: it costs L1 capacity)
    return i < (uint32_t)kObjCache ? s_objs[i] : load_obj(objs, i);
}

// Everything a warp needs to emit: the queue, the pair list (local, or the gathering rank's through the peer mapping),
// the counters, and its two staging areas in shared memory.
#ifndef OIBVH_STAGE_CAP
#define OIBVH_STAGE_CAP 256
#endif
constexpr int kStageCap = OIBVH_STAGE_CAP; // records per warp staging buffer (4 KB), half per record kind
constexpr uint32_t kHalf = kStageCap / 2;
static_assert(kStageCap / 2 >= 64, "one item emits up to 64 records of one kind");

// CTA-uniform part, kept in SHARED memory (one copy per CTA, read by broadcast loads) so that it does not sit in ~14
// registers of every thread across the traversal loop -- the kernel runs at the 64-register cap of a 1024-thread CTA
struct EmitShared
{
    uint4* queue;        // work queue (BVTT nodes known to overlap)
    uint4* pairs;        // pair list: local, or rank 0's (remote)
    uint32_t* counters;  // this scene's counter block
    uint32_t* pair_ctr;  // where pair slots are claimed: counters (local) or rank 0's counter block (remote)
    const ObjDesc* objs;
    uint32_t queue_cap;
    uint32_t pair_cap;
    uint32_t remote;
    uint32_t record;     // 1 = this detection records the BVTT cut (temporal coherence, see record_cut)
    uint4* cut;
    uint32_t cut_cap;
    uint32_t cut_depth;  // levels above the leaves at which the cut lies
    uint4* cand;         // candidate list (leaf pairs with overlapping boxes): narrow_owned_blocks
    uint32_t cand_cap;
    ObjDesc s_objs[kObjCache];
};
struct Emit
{
    const EmitShared& sh;
    uint4* stage;                // this warp's staging: [0, kHalf) queue records, [kHalf, 2 kHalf) candidates
    uint4* params;               // this warp's item parameters: 4 x 32 uint4, [q][item] (traverse_queue)
    uint32_t staged_f, staged_c; // warp-uniform
    uint32_t retire;             // finished items not yet retired (lane 0)
    uint32_t tail_seen;          // latest queue tail this warp has seen (lane 0): sizes its next claim
    __device__ __forceinline__ Emit(const EmitShared& s, uint4* st)
        : sh(s), stage(st), params(nullptr), staged_f(0), staged_c(0), retire(0), tail_seen(0)
    {
    }
};

// Narrow phase of the staged candidates (leaf pairs with overlapping boxes, src/cuda/collide.cu:155-162), by the warp
// that found them: no candidate list in global memory, no barrier, no second pass. Hits are compacted with a ballot
// and one atomic per 32 candidates; a remote rank claims its slots in rank 0's list with a system-scope atomic and
// writes them over NVLink.
#ifndef OIBVH_SAT_INLINE
#define OIBVH_SAT_INLINE __noinline__
#endif
__device__ OIBVH_SAT_INLINE void narrow_lanes(const EmitShared& sh, uint4 c, bool valid, uint32_t lane)
{
    bool hit = false;
    if (valid)
    {
        const ObjDesc A = get_obj(sh.s_objs, sh.objs, c.x), B = get_obj(sh.s_objs, sh.objs, c.y);
        const uint32_t* fa = A.faces + 3ull * c.z;
        const uint32_t* fb = B.faces + 3ull * c.w;
        const uint32_t a0 = __ldg(fa), a1 = __ldg(fa + 1), a2 = __ldg(fa + 2);
        const uint32_t b0 = __ldg(fb), b1 = __ldg(fb + 1), b2 = __ldg(fb + 2);
        // self-collision: two triangles of ONE mesh that share a vertex always touch there -- they are neighbours,
        // not a collision (decided before the test so that the six indices are dead during it)
        const bool neighbours = c.x == c.y && (a0 == b0 || a0 == b1 || a0 == b2 || a1 == b0 || a1 == b1 || a1 == b2 ||
                                               a2 == b0 || a2 == b1 || a2 == b2);
        const V3 P1 = load_v3(A.pos, a0), P2 = load_v3(A.pos, a1), P3 = load_v3(A.pos, a2);
        const V3 Q1 = load_v3(B.pos, b0), Q2 = load_v3(B.pos, b1), Q3 = load_v3(B.pos, b2);
        hit = !neighbours && triangles_intersect(P1, P2, P3, Q1, Q2, Q3);
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (mask)
    {
        uint32_t slot = 0;
        if (lane == 0)
        {
            if (sh.remote)
            {
                atomicAdd(sh.counters + CTR_PAIRS, (uint32_t)__popc(mask)); // this rank's own count
                slot = atomicAdd_system(sh.pair_ctr + CTR_PAIRS, (uint32_t)__popc(mask));
            }
            else
                slot = atomicAdd(sh.counters + CTR_PAIRS, (uint32_t)__popc(mask));
        }
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (hit)
        {
            const uint32_t dst = slot + __popc(mask & lanemask_lt());
            if (dst < sh.pair_cap)
                sh.pairs[dst] = c; // {bvhA, bvhB, triA, triB} == int_tri_pair_node_t
            else if (sh.remote)
                atomicOr_system(sh.pair_ctr + CTR_OVERFLOW, 4u);
            else
                atomicOr(sh.counters + CTR_OVERFLOW, 4u);
        }
    }
    const uint32_t n = __popc(__ballot_sync(0xffffffffu, valid));
    if (lane == 0) atomicAdd(sh.counters + CTR_CANDIDATES, n);
}

// The CTA's copy of (stop flag, queue tail, candidate tail) is written by the control warp (and by a warp that raises the
// stop flag) and read by all the others with no barrier in between -- single words, any interleaving is fine.
// Shared-memory atomics make that explicit (to the reader and to racecheck); they are issued by one lane, a few times
// per thousand cycles.
__device__ __forceinline__ uint32_t ctl_load(volatile uint32_t* p) { return atomicOr(const_cast<uint32_t*>(p), 0u); }
__device__ __forceinline__ void ctl_store(volatile uint32_t* p, uint32_t v) { atomicExch(const_cast<uint32_t*>(p), v); }

// The candidate list is cut into groups of 32 records that the NARROW warps of the grid own round-robin (group g belongs
// to narrow warp g mod NW), like the slots of the work queue: no claim counter. (Measured: claiming groups from one head
// counter with compare-and-swap serialised the claims at one per L2 round trip and, sharing a cache line with the
// counters the producers hit, slowed the traversal by a quarter.) A narrow warp tests its groups as the tail passes them;
// what it has not reached when the traversal ends is shared out among all the warps of its CTA. A group below the tail
// has been reserved by its producers; a record that has not been written yet still holds the empty marker
// (single-copy-atomic 128-bit records, as in the queue) and is waited for. A consumed record is set back to the empty
// marker by its consumer, so the kernel leaves the list empty for the next launch.
// test the candidates [base, base + n), n <= 32, one per lane
__device__ __forceinline__ void narrow_range(const EmitShared& sh, uint32_t base, uint32_t n, uint32_t lane)
{
    const bool valid = lane < n;
    uint4 c = make_uint4(0, 0, 0, 0);
    if (valid)
    {
        uint32_t spins = 0;
        while ((c = ld_rec(sh.cand + base + lane)).x == kQEmpty && ++spins < (1u << 20)) {}
        if (c.x == kQEmpty)
            atomicOr(sh.counters + CTR_OVERFLOW, 8u); // the producer never wrote it: report, do not hang
        else
            sh.cand[base + lane] = make_uint4(kQEmpty, kQEmpty, kQEmpty, kQEmpty); // consumed
    }
    narrow_lanes(sh, c, valid && c.x != kQEmpty, lane);
}
// What the auxiliary warps of a CTA do while the others traverse. The CONTROL warp reads the queue's control words
// (stop flag, queue tail, candidate tail) from global memory about once per thousand cycles and publishes them in
// shared memory, where the CTA's other warps read them -- thousands of idle warps must not poll those words in global
// memory, and a refresh that depends on some idle traversal warp happening to be on duty left gaps of several thousand
// cycles. The NARROW warps test candidates as the traversal produces them, so that most of the narrow phase runs under
// the traversal instead of after it.
__device__ __forceinline__ void aux_loop(const EmitShared& sh, volatile uint32_t* s_ctl, uint32_t* s_backlog, uint32_t warp,
                                         uint32_t lane)
{
    uint32_t* const ctr = sh.counters;
    if (warp == (uint32_t)kTravWarps)
    {
        if (lane == 0)
        {
            uint32_t rounds = 0;
            for (;;)
            {
                ctl_store(s_ctl + 1, ld_relaxed_gpu(ctr + CTR_Q_TAIL));
                ctl_store(s_ctl + 2, ld_relaxed_gpu(ctr + CTR_CAND_TAIL));
                if (ld_relaxed_gpu(ctr + CTR_Q_STOP)) ctl_store(s_ctl, 1u);
                if (ctl_load(s_ctl)) break;
                if (++rounds > (1u << 22)) // seconds: report instead of hanging
                {
                    atomicOr(ctr + CTR_OVERFLOW, 8u);
                    st_relaxed_gpu(ctr + CTR_Q_STOP, 1u);
                    ctl_store(s_ctl, 1u);
                    break;
                }
                __nanosleep(200);
            }
        }
        __syncwarp();
        return;
    }
    // narrow warp `a` of this CTA = narrow warp blockIdx.x * kNarrowWarps + a of the grid
    const uint32_t a = warp - (uint32_t)kTravWarps - 1u;
    const uint32_t nw = blockIdx.x * kNarrowWarps + a, NW = gridDim.x * kNarrowWarps;
    uint32_t k = 0; // this warp's next group is nw + k * NW
    for (;;)
    {
        uint32_t stop = 0, tail = 0;
        if (lane == 0)
        {
            stop = ctl_load(s_ctl);
            tail = ctl_load(s_ctl + 2);
        }
        stop = __shfl_sync(0xffffffffu, stop, 0);
        tail = min(__shfl_sync(0xffffffffu, tail, 0), sh.cand_cap);
        if (stop) break; // the traversal is over: what is left is shared out among all the warps (narrow_rest)
        const uint64_t g = (uint64_t)nw + (uint64_t)k * NW;
        if ((g + 1) * 32u <= tail)
        {
            narrow_range(sh, (uint32_t)(g * 32u), 32u, lane);
            k++;
            continue;
        }
        __nanosleep(400);
    }
    if (lane == 0) s_backlog[a] = k; // where the CTA's warps take over (narrow_rest)
}
// after the traversal: the groups the CTA's narrow warps have not reached, dealt to all of its warps
__device__ __forceinline__ void narrow_rest(const EmitShared& sh, const uint32_t* s_backlog, uint32_t n_cand, uint32_t warp,
                                            uint32_t lane)
{
    n_cand = min(n_cand, sh.cand_cap);
    constexpr uint32_t kStep = kColWarps / kNarrowWarps;
    const uint32_t a = warp % kNarrowWarps;
    const uint32_t nw = blockIdx.x * kNarrowWarps + a, NW = gridDim.x * kNarrowWarps;
    for (uint32_t k = s_backlog[a] + warp / kNarrowWarps;; k += kStep)
    {
        const uint64_t first = ((uint64_t)nw + (uint64_t)k * NW) * 32u;
        if (first >= n_cand) break;
        narrow_range(sh, (uint32_t)first, min(32u, n_cand - (uint32_t)first), lane);
    }
}

// Staged candidates go to the candidate list in global memory (the narrow warps test them: aux_loop, narrow_rest). Testing them right here,
// by the warp that found them, was measured: the call in the middle of the traversal's state costs registers -- the
// kernel needed 128 (16 warps per SM) and large scenes are bound by exactly that; without it the kernel fits 80 registers
// and 24 warps per SM (4096-body scene 283 -> 245 us, terrain 171 -> 140 us; the two-body bench scene is bound by its
// hop latencies and does not care: 47-51 us either way).
__device__ __forceinline__ void flush_candidates(Emit& e, uint32_t lane)
{
    const uint32_t n = e.staged_c;
    if (n == 0) return;
    __syncwarp();
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(e.sh.counters + CTR_CAND_TAIL, n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + n > e.sh.cand_cap && lane == 0)
    {
        atomicOr(e.sh.counters + CTR_OVERFLOW, 32u); // the host regrows the list and repeats the detection
        st_relaxed_gpu(e.sh.counters + CTR_Q_STOP, 1u);
    }
    for (uint32_t j = lane; j < n; j += 32)
        if (base + j < e.sh.cand_cap) st_rec(e.sh.cand + base + j, e.stage[kHalf + j]);
    __syncwarp();
    e.staged_c = 0;
}

// Push the staged BVTT nodes: ONE atomic claims the slots, each record leaves as one 128-bit store. A queue that is
// full sets the overflow flag AND stops the traversal (the host regrows the queue and repeats the detection).
__device__ __forceinline__ void flush_queue(Emit& e, uint32_t lane)
{
    const uint32_t n = e.staged_f;
    if (n == 0) return;
    __syncwarp();
    uint32_t base = 0;
    if (lane == 0)
    {
        base = queue_reserve(e.sh.counters, n, e.retire); // the pending retirements ride along
        e.retire = 0;
        e.tail_seen = base + n;
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + n > e.sh.queue_cap && lane == 0)
    {
        atomicOr(e.sh.counters + CTR_OVERFLOW, 1u);
        st_relaxed_gpu(e.sh.counters + CTR_Q_STOP, 1u);
    }
    for (uint32_t j = lane; j < n; j += 32)
        if (base + j < e.sh.queue_cap) st_rec(e.sh.queue + base + j, e.stage[j]);
    __syncwarp();
    e.staged_f = 0;
}

// Temporal coherence (the reference restarts from entryLevel every frame, src/cuda/scene.cu:236). A
// recording detection writes down a complete CUT of the BVTT: every tested node pair that did NOT overlap above the
// cut depth (its whole subtree was pruned) and every tested pair, overlapping or not, at the cut depth. While the trees
// are only refitted the topology is fixed, so the next detections need not walk down from the roots: they re-test the
// pairs of the cut (cut_seed_phase) and descend below those that overlap now -- exact for any motion, cheapest when
// little has changed. `tested`/`hit` per lane, records with (level << 26 | position) on both sides.
__device__ __forceinline__ void record_cut(const EmitShared& sh, uint32_t lane, bool misses, bool hits, bool t0, bool t1,
                                           bool hit0, bool hit1, const uint4& r0, const uint4& r1)
{
    const bool w0 = t0 && ((hit0 && hits) || (!hit0 && misses)), w1 = t1 && ((hit1 && hits) || (!hit1 && misses));
    const uint32_t m0 = __ballot_sync(0xffffffffu, w0), m1 = __ballot_sync(0xffffffffu, w1);
    const uint32_t n0 = __popc(m0), n = n0 + __popc(m1);
    if (n == 0) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(sh.counters + CTR_CUT, n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + n > sh.cut_cap)
    {
        if (lane == 0) atomicOr(sh.counters + CTR_OVERFLOW, 2u);
        return;
    }
    if (w0) sh.cut[base + __popc(m0 & lanemask_lt())] = r0;
    if (w1) sh.cut[base + n0 + __popc(m1 & lanemask_lt())] = r1;
}

// stage up to 64 overlapping descendant pairs of one item (two per lane) as queue records or candidates
__device__ __forceinline__ void stage_hits(Emit& e, uint32_t lane, bool to_cand, bool hit0, bool hit1, const uint4& r0,
                                           const uint4& r1)
{
    const uint32_t mask0 = __ballot_sync(0xffffffffu, hit0), mask1 = __ballot_sync(0xffffffffu, hit1);
    const uint32_t cnt0 = __popc(mask0), cnt = cnt0 + __popc(mask1);
    if (cnt == 0) return; // warp-uniform
    if (to_cand)
    {
        if (e.staged_c + cnt > kHalf) flush_candidates(e, lane);
    }
    else if (e.staged_f + cnt > kHalf)
        flush_queue(e, lane);
    uint4* st = e.stage + (to_cand ? kHalf + e.staged_c : e.staged_f);
    if (hit0) st[__popc(mask0 & lanemask_lt())] = r0;
    if (hit1) st[cnt0 + __popc(mask1 & lanemask_lt())] = r1;
    if (to_cand)
        e.staged_c += cnt;
    else
        e.staged_f += cnt;
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

// ---------------------------------------------------------------------------------------------------
// Seeding of scenes with few object pairs
// ---------------------------------------------------------------------------------------------------
// Root pairs: global warp w tests the root boxes of object pairs w, w + W, ... and queues the overlapping ones.
template <bool RECORD>
__device__ void root_seed_phase(Emit& e, const uint32_t* s_lv, uint32_t n_pairs, uint32_t n_obj)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t total = gridDim.x * kColWarps * 32;
    for (uint32_t p0 = (blockIdx.x * kColWarps + warp) * 32; p0 < n_pairs; p0 += total)
    {
        const uint32_t p = p0 + lane;
        bool hit = false, shallow = false;
        uint4 it = make_uint4(0, 0, 0, 0);
        if (p < n_pairs)
        {
            it = seed_entry(n_obj, p);
            const ObjDesc A = get_obj(e.sh.s_objs, e.sh.objs, it.x), B = get_obj(e.sh.s_objs, e.sh.objs, it.y);
            hit = box_overlap(load_box(reinterpret_cast<const float2*>(A.nodes), 0),
                              load_box(reinterpret_cast<const float2*>(B.nodes), 0));
            shallow = max(A.L, B.L) <= e.sh.cut_depth;
        }
        // the cut: a disjoint root pair prunes its whole BVTT; an overlapping one belongs to the cut if the trees are
        // no deeper than the cut depth
        if (RECORD) record_cut(e.sh, lane, true, shallow, p < n_pairs, false, hit, false, it, it);
        stage_hits(e, lane, false, hit, false, it, it);
    }
    flush_queue(e, lane);
}

// Dense seeding for scenes of very few objects (two-body scenes): instead of walking down from the root pair through
// several latency-bound hops of a handful of items, ALL node pairs (a, b) at level `k0` of an object pair are tested
// directly, spread over every warp of the grid: 4^8 = 65 K .. 4^11 = 4 M box tests are less work than the hops they
// replace. A level is a contiguous slice, so a warp's 32 consecutive combinations read one broadcast box of A and 32
// consecutive boxes of B. `rank`/`world` deal the combinations to the shards.
template <bool RECORD, bool SELF>
__device__ void dense_seed_phase(Emit& e, const uint32_t* s_lv, uint32_t n_pairs, uint32_t k0, uint32_t rank,
                                 uint32_t world, uint32_t n_obj)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t total_warps = gridDim.x * kColWarps;
    const uint32_t gw = blockIdx.x * kColWarps + warp;
    for (uint32_t p = 0; p < n_pairs; p++)
    {
        const uint4 it = seed_entry(n_obj, p);
        const ObjDesc A = get_obj(e.sh.s_objs, e.sh.objs, it.x), B = get_obj(e.sh.s_objs, e.sh.objs, it.y);
        const LevelView va = level_view(s_lv, it.x, A), vb = level_view(s_lv, it.y, B);
        const float2* nodesA = reinterpret_cast<const float2*>(A.nodes);
        const float2* nodesB = reinterpret_cast<const float2*>(B.nodes);
        if (!box_overlap(load_box(nodesA, 0), load_box(nodesB, 0)))
        {
            // disjoint roots (warp-uniform): one cut record prunes the whole pair
            if (RECORD && gw == 0) record_cut(e.sh, lane, true, false, lane == 0, false, false, false, it, it);
            continue;
        }
        const bool self = SELF && it.x == it.y;
        const uint32_t ka = min(k0, A.L), kb = min(k0, B.L);
        const uint32_t nA = va.count(ka), nB = vb.count(kb), baseA = va.offset(ka), baseB = vb.offset(kb);
        const bool to_cand = (ka == A.L) && (kb == B.L);
        const uint32_t za = to_cand ? 0u : (ka << kNodeLevelShift), zb = to_cand ? 0u : (kb << kNodeLevelShift);
        const uint32_t total = nA * nB; // <= 4^11
        for (uint32_t c0 = gw * 64; c0 < total; c0 += total_warps * 64)
        {
            const uint32_t c[2] = {c0 + lane, c0 + 32 + lane};
            bool hit[2], tested[2];
            uint32_t ia[2], ib[2];
            Box a[2], b[2];
#pragma unroll
            for (int u = 0; u < 2; u++)
            {
                hit[u] = c[u] < total && (world == 1 || ((p + c[u]) % world) == rank);
                ia[u] = c[u] / nB;
                ib[u] = c[u] - ia[u] * nB;
                // self-collision: the BVTT of a tree with itself is symmetric -- keep a <= b (a < b between leaves)
                if (self) hit[u] = hit[u] && (to_cand ? ia[u] < ib[u] : ia[u] <= ib[u]);
                tested[u] = hit[u];
                if (hit[u])
                {
                    a[u] = load_box(nodesA, baseA + ia[u]);
                    b[u] = load_box(nodesB, baseB + ib[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (hit[u]) hit[u] = box_overlap(a[u], b[u]);
            if (RECORD) // misses always (pruned subtrees), hits if this level is already at / below the cut depth
                record_cut(e.sh, lane, true, max(A.L - ka, B.L - kb) <= e.sh.cut_depth, tested[0], tested[1], hit[0], hit[1],
                           make_uint4(it.x, it.y, (ka << kNodeLevelShift) + ia[0], (kb << kNodeLevelShift) + ib[0]),
                           make_uint4(it.x, it.y, (ka << kNodeLevelShift) + ia[1], (kb << kNodeLevelShift) + ib[1]));
            stage_hits(e, lane, to_cand, hit[0], hit[1], make_uint4(it.x, it.y, za + ia[0], zb + ib[0]),
                       make_uint4(it.x, it.y, za + ia[1], zb + ib[1]));
        }
    }
    flush_queue(e, lane);
    flush_candidates(e, lane);
}

// Replay of a recorded cut (temporal coherence): every warp of the grid takes 32 records at a time; a record whose two
// boxes overlap NOW is queued (or, between two leaves, staged as a candidate) and the traversal descends below it.
__device__ void cut_seed_phase(Emit& e, const uint32_t* s_lv, uint32_t n_cut)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t stride = gridDim.x * kColWarps * 32;
    for (uint32_t i0 = (blockIdx.x * kColWarps + warp) * 32; i0 < n_cut; i0 += stride)
    {
        const uint32_t i = i0 + lane;
        bool hit = false, leaf = false;
        uint4 it = make_uint4(0, 0, 0, 0);
        if (i < n_cut)
        {
            it = e.sh.cut[i];
            const ObjDesc A = get_obj(e.sh.s_objs, e.sh.objs, it.x), B = get_obj(e.sh.s_objs, e.sh.objs, it.y);
            const LevelView va = level_view(s_lv, it.x, A), vb = level_view(s_lv, it.y, B);
            const uint32_t la = it.z >> kNodeLevelShift, pa = it.z & kNodePosMask;
            const uint32_t lb = it.w >> kNodeLevelShift, pb = it.w & kNodePosMask;
            hit = box_overlap(load_box(reinterpret_cast<const float2*>(A.nodes), va.offset(la) + pa),
                              load_box(reinterpret_cast<const float2*>(B.nodes), vb.offset(lb) + pb));
            leaf = la == A.L && lb == B.L;
        }
        stage_hits(e, lane, false, hit && !leaf, false, it, it);
        stage_hits(e, lane, true, hit && leaf, false, make_uint4(it.x, it.y, it.z & kNodePosMask, it.w & kNodePosMask), it);
    }
    flush_queue(e, lane);
    flush_candidates(e, lane);
}

// ---------------------------------------------------------------------------------------------------
// Queue-driven BVTT traversal, warp-cooperative, no grid barriers.
//
// A queue record is a node pair whose boxes are KNOWN to overlap (root pairs: tested by the seeding). A warp owns the
// slots it claims from the head counter -- up to 32 at a time when the backlog is long, one when it is short, so that
// a front of a few hundred items still spreads over every warp of the chip -- and polls them until their records
// appear (slots are filled in claim order by whoever emits next, so an idle warp is handed the very next node that is
// produced). For each item the warp addresses the rectangle of descendants `levels` levels further down on each side
// (clamped to the leaf level and to the nodes the level keeps) and its 32 lanes test the nA x nB descendant box
// pairs; only overlapping pairs are emitted -- back into the queue, or, when both sides reached the leaf level, to
// the warp's candidate staging area, whose triangle pairs it tests itself (flush_candidates).
// Set-up and testing are split: what depends only on the item (descriptor fetch, level geometry, rectangle clamping:
// ~200 instructions) is done by lane l for item l of the batch, then the warp walks the batch broadcasting each
// item's parameters by shuffle.
// Termination: pushed and finished records are the two halves of one 64-bit counter (queue_reserve / queue_retire); a
// warp pushes everything its batch produced BEFORE it retires the batch, so the warp whose retirement makes the halves
// equal knows that nothing is in flight and nothing can be pushed any more: it raises the stop flag. The hop latency is ~4 dependent L2 round trips
// (claim, record, boxes, push) instead of a grid barrier on top of them, and hops of different subtrees overlap.
// ---------------------------------------------------------------------------------------------------
#ifdef OIBVH_PROFILE
// where the traversal's warps spend their time (summed over all warps, lane 0's clock): [0] window, [1] polls that
// found nothing, [2] set-up + box tests + staging, [3] queue pushes, [4] narrow phase, [5] total, [6] polls that found
// records, [7] set-up alone; counts: [9] empty polls, [10] batches, [11] items, [12] candidate flushes
__device__ unsigned long long g_col_prof[16];
extern "C" int oibvh_debug_collide_profile(unsigned long long* out, int reset)
{
    if (reset)
    {
        unsigned long long z[16] = {};
        return (int)cudaMemcpyToSymbol(g_col_prof, z, sizeof(z));
    }
    return (int)cudaMemcpyFromSymbol(out, g_col_prof, sizeof(g_col_prof));
}
// hop timeline: per tree level of side A, wall clock (ns) of the first / last node taken from the queue and of the last
// node finished: [level][0..2]; [32][0] = start of the traversal
__device__ unsigned long long g_col_hops[33][4];
extern "C" int oibvh_debug_collide_hops(unsigned long long* out, int reset)
{
    if (reset)
    {
        static unsigned long long z[33][4];
        for (int i = 0; i < 33; i++) { z[i][0] = ~0ull; z[i][1] = z[i][2] = z[i][3] = 0; }
        return (int)cudaMemcpyToSymbol(g_col_hops, z, sizeof(z));
    }
    return (int)cudaMemcpyFromSymbol(out, g_col_hops, sizeof(g_col_hops));
}
__device__ __forceinline__ unsigned long long col_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
#define COL_T(var) const long long var = clock64()
#define COL_ADD(slot, val)                                                                                         \
    do                                                                                                             \
    {                                                                                                              \
        if (lane == 0) prof[slot] += (unsigned long long)(val);                                                    \
    } while (0)
#else
#define COL_T(var)
#define COL_ADD(slot, val)
#endif

template <bool RECORD, bool SELF>
__device__ void traverse_queue(Emit& e, const uint32_t* s_lv, uint32_t* s_hist, volatile uint32_t* s_ctl,
                               uint32_t levels0, uint32_t levels, uint32_t rank, uint32_t world, uint32_t n_obj,
                               bool shard_roots)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
#ifdef OIBVH_PROFILE
    unsigned long long prof[16] = {};
    const long long t_begin = clock64();
#endif
#ifdef OIBVH_PROFILE_HOPS
    if (threadIdx.x == 0 && blockIdx.x == 0) g_col_hops[32][0] = col_gtime();
    __shared__ unsigned long long s_hops[kColWarps][32][3];
    s_hops[warp][lane][0] = ~0ull;
    s_hops[warp][lane][1] = s_hops[warp][lane][2] = 0ull;
    __syncwarp();
#endif
    e.tail_seen = 0;
    e.retire = warp == 0 ? 1u : 0u; // the CTA's seeding token (queue_retire)
    const uint32_t total_warps = gridDim.x * kTravWarps; // traversal warps of the grid
    uint32_t* const ctr = e.sh.counters;
    // Slots are owned statically, round-robin: global warp w owns slots w, w + W, w + 2 W, ... (W = warps of the grid).
    // No head counter, no claim atomics (measured: thousands of warps claiming from one counter waited 5-7 K cycles per
    // claim -- an L2 slice serves same-address atomics a few cycles apart), and the nodes of every hop are dealt evenly
    // to the warps. A warp looks at a window of its next slots (one while the queue is short, up to 32 once there are
    // several nodes per warp) and takes the filled prefix.
    const uint32_t gw = blockIdx.x * kTravWarps + warp;
    uint32_t next = 0; // this warp's next slot is gw + next * W
    for (;;)
    {
        COL_T(t0);
        // ---- wait for this warp's next slot. Lane 0 spins on the first word of the record with as few instructions
        // as possible: 32 idle warps per SM in a fat polling loop eat the issue slots of the warps that have work
        // (measured: every phase of the traversal 3-5x slower); the other lanes wait at the shuffle below. ----
        uint32_t state = 0; // 0 = a record is there, 1 = the traversal is over
        uint4 it = make_uint4(kQEmpty, 0, 0, 0); // lane 0: the record it has been waiting for (one L2 round trip: the
                                                 // poll that finds the record IS its fetch)
        if (lane == 0)
        {
            const uint64_t slot0 = (uint64_t)gw + (uint64_t)next * total_warps;
            if (slot0 < e.sh.queue_cap)
            {
                const uint4* first = e.sh.queue + (uint32_t)slot0;
                uint32_t spins = 0;
                while ((it = ld_rec(first)).x == kQEmpty)
                {
                    if ((++spins & 3u) != 0) continue;
                    COL_ADD(9, 4);
                    // every fourth poll: housekeeping.
                    // (1) Out of work for a while: retire what this warp has finished (it may be the last one). Not at
                    //     the first empty poll: between two hops every warp is briefly idle, and thousands of
                    //     retirements at once would queue up on the counter the pushes need.
                    if (e.retire)
                    {
                        if (queue_retire(ctr, e.retire))
                        {
                            st_relaxed_gpu(ctr + CTR_Q_STOP, 1u);
                            ctl_store(s_ctl, 1u);
                        }
                        e.retire = 0;
                    }
                    // (2) Thousands of idle warps must NOT poll the stop flag and the tail in global memory: those two
                    //     words (the tail shares its line with the counter every push hits) would saturate their L2
                    //     slice. The CTA's control warp keeps a copy in shared memory fresh (aux_loop); everybody
                    //     else reads that.
                    if (ctl_load(s_ctl))
                    {
                        state = 1;
                        break;
                    }
                    if (spins > (1u << 22))
                    {
                        atomicOr(ctr + CTR_OVERFLOW, 8u); // nothing arrived for seconds: report instead of hanging
                        st_relaxed_gpu(ctr + CTR_Q_STOP, 1u);
                        ctl_store(s_ctl, 1u);
                        state = 1;
                        break;
                    }
                }
            }
            else
            {
                // this warp's slots are exhausted (the queue is full): retire, then wait for the end
                if (e.retire)
                {
                    if (queue_retire(ctr, e.retire))
                    {
                        st_relaxed_gpu(ctr + CTR_Q_STOP, 1u);
                        ctl_store(s_ctl, 1u);
                    }
                    e.retire = 0;
                }
                uint32_t spins = 0;
                while (!ctl_load(s_ctl) && ++spins < (1u << 22)) __nanosleep(500);
                state = 1;
            }
        }
        state = __shfl_sync(0xffffffffu, state, 0);
        COL_T(t1);
        COL_ADD(1, t1 - t0);
        if (state) break;
        // ---- take the filled prefix of a window of this warp's next slots: one while the queue is short, up to 32
        // once there are several nodes per warp (records of different producers may become visible out of order:
        // whatever lies behind a gap waits for the next turn) ----
        uint32_t window = 1;
        if (lane == 0) window = min(32u, max(1u, 2u * max(e.tail_seen, ctl_load(s_ctl + 1)) / total_warps));
        window = __shfl_sync(0xffffffffu, window, 0);
        const uint64_t slot64 = (uint64_t)gw + (uint64_t)(next + lane) * total_warps;
        if (lane != 0 && lane < window && slot64 < e.sh.queue_cap) it = ld_rec(e.sh.queue + (uint32_t)slot64);
        const uint32_t filled = __ballot_sync(0xffffffffu, it.x != kQEmpty);
        const uint32_t take = filled == 0xffffffffu ? 32u : (uint32_t)__ffs(~filled) - 1u; // >= 1: lane 0 saw its record
        const bool valid = lane < take;
        const uint32_t got = take == 32u ? 0xffffffffu : ((1u << take) - 1u);
        next += take;
        COL_ADD(10, 1);
        COL_ADD(11, take);
        COL_T(t1b);
        COL_ADD(6, t1b - t1);

#ifdef OIBVH_PROFILE_HOPS
        uint32_t my_level = 0xffffffffu;
#endif
        // ---- phase 1: lane l prepares its item ----
        // The item's parameters go to the warp's parameter block in shared memory ([q][item], 16-byte columns: the
        // stores are conflict-free and phase 2 reads item k's four columns as broadcasts). Keeping them in registers and
        // broadcasting them with ten shuffles per item cost 13 registers through the whole of phase 2 and ~300 cycles
        // per item.
        if (valid)
        {
            uint32_t meta2 = 0;
            const ObjDesc A = get_obj(e.sh.s_objs, e.sh.objs, it.x), B = get_obj(e.sh.s_objs, e.sh.objs, it.y);
            const LevelView va = level_view(s_lv, it.x, A), vb = level_view(s_lv, it.y, B);
            const uint32_t la = it.z >> kNodeLevelShift, pa = it.z & kNodePosMask;
            const uint32_t lb = it.w >> kNodeLevelShift, pb = it.w & kNodePosMask;
            const bool root = (it.z | it.w) == 0u;
            // (a root pair reaches the queue when the seeding is not dense, or when a replayed cut holds one)
            const uint32_t k = root ? min(levels0, kMaxExpandLevels) : levels;
            const uint32_t da = min(k, A.L - la), db = min(k, B.L - lb);
            const uint32_t lca = la + da, lcb = lb + db;
            const uint32_t fa = pa << da, fb = pb << db;
            const uint32_t nA = min(1u << da, va.count(lca) - fa);
            const uint32_t nB = min(1u << db, vb.count(lcb) - fb);
            const uint32_t combos = nA * nB; // <= 1024
            const bool to_cand = (lca == A.L) && (lcb == B.L);
            const uint64_t ptrA = (uint64_t)A.nodes, ptrB = (uint64_t)B.nodes;
            const uint32_t baseA = va.offset(lca) + fa, baseB = vb.offset(lcb) + fb;
            // Start fetching the two runs of descendant boxes NOW, for all the items of the batch at once (the lanes
            // set their items up in parallel): phase 2 walks the items one after the other, and without this every
            // item would wait its own L2 round trip. No register is held.
            {
                const char* ra = reinterpret_cast<const char*>(A.nodes) + 24ull * baseA;
                const char* rb = reinterpret_cast<const char*>(B.nodes) + 24ull * baseB;
                for (uint32_t o = 0; o < 24u * nA + 127u; o += 128u) asm volatile("prefetch.global.L1 [%0];" ::"l"(ra + o));
                for (uint32_t o = 0; o < 24u * nB + 127u; o += 128u) asm volatile("prefetch.global.L1 [%0];" ::"l"(rb + o));
            }
            // emitted node ids are za + ia / zb + ib (fa, fb have their low da / db bits clear)
            const uint32_t za = to_cand ? fa : ((lca << kNodeLevelShift) | fa);
            const uint32_t zb = to_cand ? fb : ((lcb << kNodeLevelShift) | fb);
            const uint32_t meta = combos | (nB << 11) | (db << 17) | ((nB == (1u << db)) ? 1u << 20 : 0u) | (to_cand ? 1u << 21 : 0u) |
                   ((root && shard_roots) ? 1u << 22 : 0u);
            // the children of a root pair are dealt round-robin to the shards, keyed by the pair's linear index
            const uint32_t key = pair_linear(n_obj, it.x, it.y);
            // recording: levels of the children, and whether their misses / hits belong to the cut (depth above the
            // leaves before and after this hop)
            if (RECORD)
            {
                const uint32_t parent_rem = max(A.L - la, B.L - lb), rem = max(A.L - lca, B.L - lcb);
                meta2 = lca | (lcb << 5) | ((parent_rem > e.sh.cut_depth) ? 1u << 10 : 0u) |
                        ((rem <= e.sh.cut_depth && e.sh.cut_depth < parent_rem) ? 1u << 11 : 0u);
            }
            e.params[lane] = make_uint4(meta, key, it.x, it.y);
            e.params[32 + lane] = make_uint4(za, zb, baseA, baseB);
            e.params[64 + lane] = make_uint4((uint32_t)ptrA, (uint32_t)(ptrA >> 32), (uint32_t)ptrB, (uint32_t)(ptrB >> 32));
            if (RECORD) e.params[96 + lane].x = meta2;
            atomicAdd(s_hist + min(la, 31u), 1u); // items per tree level of side A (oibvh_scene_get_round_stats)
#ifdef OIBVH_PROFILE_HOPS
            {
                const unsigned long long now = col_gtime();
                atomicMin(&s_hops[warp][min(la, 31u)][0], now);
                atomicMax(&s_hops[warp][min(la, 31u)][1], now);
                my_level = min(la, 31u);
            }
#endif
        }
        COL_T(t1c);
        COL_ADD(7, t1c - t1b);
        // ---- phase 2: the warp walks the prepared items, 64 descendant pairs per step ----
        __syncwarp();
        for (uint32_t k = 0; k < take; k++)
        {
            COL_T(p2s);
            const uint4 q0 = e.params[k], q1 = e.params[32 + k], q2 = e.params[64 + k];
            const uint32_t mk = q0.x;
            const uint32_t combos = mk & 0x7ffu, nB = (mk >> 11) & 63u, db = (mk >> 17) & 7u;
            const bool to_cand = (mk >> 21) & 1u, sharded = (mk >> 22) & 1u;
            const float2* nodesA = reinterpret_cast<const float2*>(((uint64_t)q2.y << 32) | q2.x);
            const float2* nodesB = reinterpret_cast<const float2*>(((uint64_t)q2.w << 32) | q2.z);
            const uint32_t bA = q1.z, bB = q1.w;
            const uint32_t exk = q0.z, eyk = q0.w;
            const uint32_t zak = q1.x, zbk = q1.y;
            const uint32_t kk = q0.y;
            const bool self = SELF && exk == eyk;
            COL_T(p2a);
            COL_ADD(14, (p2a - p2s) + (long long)(self & 0));
            for (uint32_t g0 = 0; g0 < combos; g0 += 64) // warp-uniform
            {
                const uint32_t c0 = g0 + lane, c1 = c0 + 32;
                bool hit0 = c0 < combos, hit1 = c1 < combos;
                if (sharded)
                {
                    hit0 = hit0 && ((kk + c0) % world) == rank;
                    hit1 = hit1 && ((kk + c1) % world) == rank;
                }
                // nB is a power of two unless the rectangle is clamped by the end of the level (warp-uniform)
                uint32_t ia0, ib0, ia1, ib1;
                if ((mk >> 20) & 1u)
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
                if (self)
                {
                    // self-collision: the BVTT of a tree with itself is symmetric -- keep a <= b (a < b between leaves);
                    // both sides are at the same level, so the node ids compare like the positions
                    hit0 = hit0 && (to_cand ? zak + ia0 < zbk + ib0 : zak + ia0 <= zbk + ib0);
                    hit1 = hit1 && (to_cand ? zak + ia1 < zbk + ib1 : zak + ia1 <= zbk + ib1);
                }
                const bool t0 = hit0, t1 = hit1;
                Box a0, b0, a1, b1;
                if (hit0)
                {
                    a0 = load_box(nodesA, bA + ia0);
                    b0 = load_box(nodesB, bB + ib0);
                }
                if (hit1)
                {
                    a1 = load_box(nodesA, bA + ia1);
                    b1 = load_box(nodesB, bB + ib1);
                }
                if (hit0) hit0 = box_overlap(a0, b0);
                if (hit1) hit1 = box_overlap(a1, b1);
#ifdef OIBVH_PROFILE
                {
                    const uint32_t any = __ballot_sync(0xffffffffu, hit0 || hit1); // forces the loads to complete
                    COL_T(p2b);
                    COL_ADD(13, (p2b - p2a) + (any & 0));
                }
#endif
                if (RECORD)
                {
                    const uint32_t m2 = e.params[96 + k].x;
                    const uint32_t la_c = (m2 & 31u) << kNodeLevelShift, lb_c = ((m2 >> 5) & 31u) << kNodeLevelShift;
                    const uint32_t pa0 = (zak & kNodePosMask) + ia0, pb0 = (zbk & kNodePosMask) + ib0;
                    const uint32_t pa1 = (zak & kNodePosMask) + ia1, pb1 = (zbk & kNodePosMask) + ib1;
                    record_cut(e.sh, lane, (m2 >> 10) & 1u, (m2 >> 11) & 1u, t0, t1, hit0, hit1,
                               make_uint4(exk, eyk, la_c | pa0, lb_c | pb0), make_uint4(exk, eyk, la_c | pa1, lb_c | pb1));
                }
                stage_hits(e, lane, to_cand, hit0, hit1, make_uint4(exk, eyk, zak + ia0, zbk + ib0),
                           make_uint4(exk, eyk, zak + ia1, zbk + ib1));
            }
        }
        __syncwarp(); // the parameter block is rewritten by the next batch's phase 1
        // ---- the batch is finished only when everything it produced has left the warp ----
        COL_T(t3);
        COL_ADD(2, t3 - t1b);
        flush_queue(e, lane);
        COL_T(t4);
        COL_ADD(3, t4 - t3);
        if (e.staged_c) COL_ADD(12, 1);
        flush_candidates(e, lane);
        COL_T(t5);
        COL_ADD(4, t5 - t4);
#ifdef OIBVH_PROFILE_HOPS
        if (my_level != 0xffffffffu) atomicMax(&s_hops[warp][my_level][2], col_gtime());
#endif
        if (lane == 0) e.retire += (uint32_t)__popc(got); // retired with the next reservation, or when out of work
    }
#ifdef OIBVH_PROFILE
#ifdef OIBVH_PROFILE_HOPS
    __syncwarp();
    if (s_hops[warp][lane][1])
    {
        atomicMin(&g_col_hops[lane][0], s_hops[warp][lane][0]);
        atomicMax(&g_col_hops[lane][1], s_hops[warp][lane][1]);
        atomicMax(&g_col_hops[lane][2], s_hops[warp][lane][2]);
    }
#endif
    if (lane == 0)
    {
        prof[5] = (unsigned long long)(clock64() - t_begin);
        for (int i = 0; i < 16; i++)
            if (prof[i]) atomicAdd(g_col_prof + i, prof[i]);
    }
#endif
}

// ---------------------------------------------------------------------------------------------------
// The persistent detection kernel
// ---------------------------------------------------------------------------------------------------
constexpr size_t kColStageBytes = (size_t)kColWarps * kStageCap * sizeof(uint4); // 64 KB
constexpr size_t kColParamBytes = (size_t)kColWarps * 4 * 32 * sizeof(uint4); // 2 KB per warp
constexpr size_t kColSmemBytes = kColStageBytes + kColParamBytes;
// MODE: 0 = from the roots, 1 = from the roots + write the BVTT cut down, 2 = from the recorded cut (temporal
// coherence); SELF = objects are also tested against themselves: separate instantiations, so that the options do not cost the everyday kernel registers (it runs at
// the 128-register cap without spilling; with the options compiled in it spilled and lost 10 %)
template <int MODE, bool SELF>
__global__ void __launch_bounds__(kColThreads, 1)
    collide_kernel(const ObjDesc* __restrict__ objs, uint32_t n_obj, uint4* queue, uint32_t queue_cap, uint4* pairs,
                   uint32_t pair_cap, uint32_t* counters, uint32_t levels0, uint32_t levels, uint32_t rank, uint32_t world,
                   const MgpuArgs mg, const DetectOpts opt)
{
    constexpr bool RECORD = MODE == 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* s_stage = reinterpret_cast<uint4*>(smem_raw); // kColWarps x kStageCap records
    __shared__ __align__(16) EmitShared s_emit;
    ObjDesc* const s_objs = s_emit.s_objs;
    __shared__ uint32_t s_lv[kObjCache * 64]; // per cached object: off[32], cnt[32]
    __shared__ uint32_t s_hist[32];           // items per tree level (statistics)
    __shared__ uint32_t s_backlog[kNarrowWarps]; // groups of the candidate list each narrow warp has tested (aux_loop)
    __shared__ uint32_t s_ctl[3];             // CTA copy of (stop flag, queue tail, candidate tail), refreshed by one warp at a time
    for (uint32_t i = threadIdx.x; i < min(n_obj, (uint32_t)kObjCache) * 32; i += blockDim.x)
    {
        const uint32_t o = i >> 5, l = i & 31u;
        const ObjDesc d = load_obj(objs, o);
        if (l == 0) s_objs[o] = d;
        s_lv[o * 64 + l] = l <= d.L ? level_offset(d.T, d.L, l) : 0u;
        s_lv[o * 64 + 32 + l] = l <= d.L ? level_count(d.T, d.L, l) : 0u;
    }
    if (threadIdx.x < 32) s_hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0)
    {
        s_ctl[0] = s_ctl[1] = s_ctl[2] = 0u;
        const bool remote = mg.mode == 2;
        s_emit.queue = queue;
        s_emit.queue_cap = queue_cap;
        s_emit.counters = counters;
        s_emit.objs = objs;
        s_emit.remote = remote ? 1u : 0u;
        s_emit.pairs = remote ? mg.root_pairs : pairs;
        s_emit.pair_cap = remote ? mg.root_pair_cap : pair_cap;
        s_emit.pair_ctr = remote ? mg.root_counters : counters;
        s_emit.record = RECORD ? 1u : 0u;
        s_emit.cut = opt.cut;
        s_emit.cand = opt.cand;
        s_emit.cand_cap = opt.cand_cap;
        s_emit.cut_cap = opt.cut_cap;
        s_emit.cut_depth = opt.cut_depth;
    }
    // multi-GPU: this launch is frame MG_FRAME + 1 of the scene. A remote rank may append to the root's list only once
    // the root has zeroed its counter block for this frame (MG_OPEN >= frame): one thread per CTA polls the root's word
    // over NVLink NOW, off the critical path; the answer is picked up from shared memory after the seeding.
    __shared__ uint32_t s_mg_ok;
    uint32_t mg_frame = 0;
    if (mg.mode != 0)
    {
        mg_frame = __ldcg(mg.state + MG_FRAME) + 1; // updated only by the last CTA of the PREVIOUS launch
        if (mg.mode == 2 && threadIdx.x == blockDim.x - 32)
        {
            uint32_t spins = 0, ok = 1;
            while (ld_acquire_sys(mg.root_state + MG_OPEN) < mg_frame)
                if (++spins > (1u << 22))
                {
                    ok = 0;
                    break;
                }
            s_mg_ok = ok;
        }
    }
    __syncthreads();
    if (mg.mode == 2 && !s_mg_ok)
    {
        // the root never opened this frame: report it, then run to completion without touching its memory
        if (threadIdx.x == 0)
        {
            atomicOr(counters + CTR_OVERFLOW, 16u);
            s_emit.remote = 0u;
            s_emit.pairs = pairs;
            s_emit.pair_cap = 0; // nothing is stored
            s_emit.pair_ctr = counters;
        }
        __syncthreads();
    }
    uint32_t n_stamps = 0;
    auto stamp = [&]() {
        if (blockIdx.x == 0 && threadIdx.x == 0 && n_stamps < CTR_WORDS_TIME)
            counters[CTR_TIME0 + n_stamps] = (uint32_t)clock64();
        n_stamps++;
    };
    stamp();

    const uint32_t warp = threadIdx.x >> 5;
    Emit e(s_emit, s_stage + warp * kStageCap);
    e.params = reinterpret_cast<uint4*>(smem_raw + kColStageBytes) + (size_t)warp * 4 * 32;

    // ---- seeds: queue records for the object pairs whose root boxes overlap ----
    // few pairs (the common two-body scene): every node pair of a deep level tested densely, or the root pairs
    // themselves; many-body scenes: the tiled top-level pass over the root boxes
    const uint64_t n_pairs = (uint64_t)n_obj * (n_obj - 1) / 2 + (SELF ? n_obj : 0u);
    const bool dense = n_pairs <= 4096 && levels0 > kMaxExpandLevels;
    if (MODE == 2)
        cut_seed_phase(e, s_lv, __ldcg(opt.cut_state)); // temporal coherence: start from the recorded cut
    else if (dense)
        dense_seed_phase<RECORD, SELF>(e, s_lv, (uint32_t)n_pairs, levels0, rank, world, n_obj);
    else if (n_pairs <= 4096)
        root_seed_phase<RECORD>(e, s_lv, (uint32_t)n_pairs, n_obj);
    else
    {
        seed_phase(reinterpret_cast<float*>(smem_raw), objs, n_obj, queue, queue_cap, counters, rank, world);
        if (SELF)
        {
            // self-collision: every object against itself
            __syncthreads(); // the tile staging area is the warps' staging area again
            const uint32_t lane = lane_id();
            for (uint32_t i0 = (blockIdx.x * kColWarps + warp) * 32; i0 < n_obj; i0 += gridDim.x * kColWarps * 32)
                stage_hits(e, lane, false, i0 + lane < n_obj && (i0 + lane) % world == rank, false,
                           make_uint4(i0 + lane, i0 + lane, 0u, 0u), make_uint4(0, 0, 0, 0));
            flush_queue(e, lane);
        }
    }
    // No grid barrier here: a warp that has nothing (more) to seed starts consuming what the others have pushed; the
    // CTA's seeding token keeps the traversal from being declared over while anybody is still seeding.
    __syncthreads();
    stamp();
    if (warp < (uint32_t)kTravWarps)
        traverse_queue<RECORD, SELF>(e, s_lv, s_hist, s_ctl, levels0, levels, rank, world, n_obj,
                                     world > 1 && n_pairs <= 4096 /* many-body scenes shard the seeding instead */);
    else
        aux_loop(s_emit, s_ctl, s_backlog, warp, lane_id());
    __syncthreads();
    stamp();

    // ---- narrow phase, the rest: the traversal is over (everything pushed has been retired, and a warp retires a batch
    // only after its candidates have their places in the list), so the tail is final; every warp claims what is left ----
    {
        const uint32_t lane = lane_id();
        const uint32_t n_cand = ld_relaxed_gpu(counters + CTR_CAND_TAIL);
        narrow_rest(s_emit, s_backlog, n_cand, warp, lane);
    }
    __syncthreads();
    stamp();

    // ---- leave the queue empty for the next launch: every slot that was filled goes back to the empty marker ----
    // (after the stop flag nobody reads a filled slot any more: they have all been consumed)
    {
        const uint32_t filled = min(__ldcg(counters + CTR_Q_TAIL), queue_cap);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < filled; i += gridDim.x * blockDim.x)
            queue[i] = make_uint4(kQEmpty, kQEmpty, kQEmpty, kQEmpty);
    }
    if (threadIdx.x < 32 && s_hist[threadIdx.x]) atomicAdd(counters + CTR_FRONT0 + threadIdx.x, s_hist[threadIdx.x]);
    // a recording detection leaves the size of the cut where the replaying ones find it (every record was written
    // before the traversal stopped; the slowest CTA may write the word last, they all write the same value)
    if (RECORD && threadIdx.x == 0) st_relaxed_gpu(opt.cut_state, min(__ldcg(counters + CTR_CUT), opt.cut_cap));
    stamp();
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[CTR_TIME0 - 1] = n_stamps; // number of stamps
    if (mg.mode != 0 && threadIdx.x == 0)
    {
        // Leave protocol: every CTA makes its appended pairs visible system-wide and counts itself out; the LAST one
        // closes the frame -- a remote rank tells the root it is done, the root waits until every remote rank has
        // said so, which makes "this kernel has completed on the root" mean "the gathered list is complete".
        __threadfence_system();
        if (atomicAdd(mg.state + MG_EXIT, 1u) == gridDim.x - 1)
        {
            __threadfence_system(); // acquire side of the other CTAs' fences
            st_relaxed_gpu(mg.state + MG_EXIT, 0u);
            st_relaxed_gpu(mg.state + MG_FRAME, mg_frame);
            if (mg.mode == 2)
            {
                if (s_mg_ok) atomicAdd_system(mg.root_state + MG_DONE, 1u);
            }
            else
            {
                const uint32_t want = (mg.world - 1) * mg_frame;
                uint32_t spins = 0;
                while (ld_acquire_sys(mg.state + MG_DONE) < want)
                    if (++spins > (1u << 24))
                    {
                        atomicOr(counters + CTR_OVERFLOW, 16u);
                        atomicOr(mg.state + MG_FAIL, 1u);
                        break;
                    }
            }
        }
    }
}

// root: zero the counter block of the coming frame, then publish MG_OPEN = frame at system scope. A separate
// one-CTA launch so that the opening does not depend on the root's (cooperative, machine-filling) detection kernel
// having started.
__global__ void __launch_bounds__(CTR_WORDS) mgpu_open_kernel(uint32_t* counters, uint32_t* state)
{
    counters[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const uint32_t frame = __ldcg(state + MG_FRAME) + 1;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(state + MG_OPEN), "r"(frame) : "memory");
    }
}

cudaError_t launch_mgpu_open(uint32_t* counters, uint32_t* state, cudaStream_t s)
{
    mgpu_open_kernel<<<1, CTR_WORDS, 0, s>>>(counters, state);
    return cudaGetLastError();
}

// =================================================================================================
// Collided-triangle vertex stream: Scene::convertToVertexArray (src/cuda/scene.cu:68-93) as a device-side gather.
// Pair i contributes six packed float3: the three vertices of the A triangle, then of the B triangle. The pair
// count is read from the counters block, so the launch can follow a detection on the stream (or sit in its graph)
// without a host round trip.
// =================================================================================================
__global__ void __launch_bounds__(256) pair_vertices_kernel(const ObjDesc* __restrict__ objs,
                                                            const uint4* __restrict__ pairs, uint32_t pair_cap,
                                                            const uint32_t* __restrict__ counters,
                                                            float* __restrict__ out, uint32_t out_cap_pairs)
{
    const uint32_t n = min(min(__ldcg(counters + CTR_PAIRS), pair_cap), out_cap_pairs);
    const uint32_t total = n * 6u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    {
        const uint32_t p = i / 6u, j = i - p * 6u;       // j = 0..2: triangle A, 3..5: triangle B
        const uint4 rec = __ldcg(pairs + p);             // (objA, objB, triA, triB), written by the collide kernel
        const uint32_t side = j >= 3u ? 1u : 0u;
        const ObjDesc o = objs[side ? rec.y : rec.x];
        const uint32_t tri = side ? rec.w : rec.z;
        const uint32_t vi = __ldg(o.faces + 3ull * tri + (j - 3u * side));
        const float4 v = __ldg(o.pos + vi);
        float* dst = out + 3ull * i;
        dst[0] = v.x;
        dst[1] = v.y;
        dst[2] = v.z;
    }
}

cudaError_t launch_pair_vertices(const ObjDesc* objs, const uint4* pairs, uint32_t pair_cap, const uint32_t* counters,
                                 float* out, uint32_t out_cap_pairs, cudaStream_t s)
{
    if (out_cap_pairs == 0) return cudaSuccess;
    const uint64_t work = (uint64_t)(pair_cap < out_cap_pairs ? pair_cap : out_cap_pairs) * 6u;
    uint64_t blocks = (work + 255) / 256;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    pair_vertices_kernel<<<blocks ? (uint32_t)blocks : 1u, 256, 0, s>>>(objs, pairs, pair_cap, counters, out, out_cap_pairs);
    return cudaGetLastError();
}

// =================================================================================================
// Node-box wireframes: OibvhTree::convertToVertexArray (src/cuda/oibvhTree.cu:69-124) + makeCube
// (src/utils/utils.cpp:15-70). Node i contributes 8 corners and 12 edges (24 indices, offset by 8 i). The corner
// arithmetic follows the reference literally: h = 0.5f * (max - min); corner = (+-h) + (min - (-h)) per axis.
// =================================================================================================
__global__ void __launch_bounds__(256) box_wireframe_kernel(const float* __restrict__ nodes, uint32_t n,
                                                            float* __restrict__ verts, uint32_t* __restrict__ idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2* p = reinterpret_cast<const float2*>(nodes) + 3ull * i;
    const float2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
    const float mn[3] = {a.x, a.y, b.x}, mx[3] = {b.y, c.x, c.y};
    float h[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        h[k] = __fmul_rn(0.5f, __fsub_rn(mx[k], mn[k]));
        d[k] = __fsub_rn(mn[k], -h[k]); // aabb.m_minimum - cubeVertices[4]
    }
    // makeCube corner signs: front quad z = +h (0..3), back quad z = -h (4..7); x: - + + -, y: - - + +
    const float sx[4] = {-1.f, 1.f, 1.f, -1.f}, sy[4] = {-1.f, -1.f, 1.f, 1.f};
    float* v = verts + 24ull * i;
#pragma unroll
    for (int q = 0; q < 8; q++)
    {
        const float cx = sx[q & 3] < 0.f ? -h[0] : h[0];
        const float cy = sy[q & 3] < 0.f ? -h[1] : h[1];
        const float cz = q < 4 ? h[2] : -h[2];
        v[3 * q] = __fadd_rn(cx, d[0]);
        v[3 * q + 1] = __fadd_rn(cy, d[1]);
        v[3 * q + 2] = __fadd_rn(cz, d[2]);
    }
    const uint32_t e[24] = {0, 1, 1, 2, 2, 3, 3, 0, 4, 5, 5, 6, 6, 7, 7, 4, 1, 5, 0, 4, 3, 7, 2, 6}; // utils.cpp:33-69
    uint32_t* o = idx + 24ull * i;
#pragma unroll
    for (int q = 0; q < 24; q++) o[q] = e[q] + 8u * i;
}

cudaError_t launch_box_wireframe(const float* nodes, uint32_t n, float* verts, uint32_t* idx, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    box_wireframe_kernel<<<(n + 255) / 256, 256, 0, s>>>(nodes, n, verts, idx);
    return cudaGetLastError();
}

static const void* collide_fn(uint32_t mode, bool self)
{
    if (mode == 1) return self ? (const void*)collide_kernel<1, true> : (const void*)collide_kernel<1, false>;
    if (mode == 2) return self ? (const void*)collide_kernel<2, true> : (const void*)collide_kernel<2, false>;
    return self ? (const void*)collide_kernel<0, true> : (const void*)collide_kernel<0, false>;
}

cudaError_t collide_configure(int* grid_blocks)
{
    int per_sm = 1 << 20, sms = 0, dev = 0;
    for (int v = 0; v < 6; v++)
    {
        const void* fn = collide_fn(v >> 1, v & 1);
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kColThreads, kColSmemBytes);
        if (e != cudaSuccess) return e;
        per_sm = n < per_sm ? n : per_sm;
    }
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid_blocks = sms; // one CTA per SM
    return cudaSuccess;
}

cudaError_t launch_collide(int grid_blocks, const ObjDesc* objs, uint32_t n_obj, uint4* queue, uint32_t queue_cap,
                           uint4* pairs, uint32_t pair_cap, uint32_t* counters, uint32_t levels0, uint32_t levels,
                           uint32_t rank, uint32_t world, const MgpuArgs& mg, const DetectOpts& opt, cudaStream_t s)
{
    MgpuArgs mga = mg;
    DetectOpts o = opt;
    void* args[] = {&objs, &n_obj, &queue, &queue_cap, &pairs, &pair_cap, &counters, &levels0, &levels, &rank, &world, &mga, &o};
    return cudaLaunchCooperativeKernel(collide_fn(opt.mode, opt.self != 0), dim3(grid_blocks), dim3(kColThreads), args,
                                       kColSmemBytes, s);
}

} // namespace oibvh
