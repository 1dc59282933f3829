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

// Many-body seeding = the top-level pass: every object pair i < j has its two root boxes tested, and only
// overlapping pairs enter the front (warp-aggregated append). The reference seeds every pair on the host, O(n^2)
// entries (src/cuda/scene.cu:192-223); here 4096 objects are 8.4 M cheap root tests and a front of a few thousand.
// The n x n triangle is cut into B x B tiles (B = 256, smaller when that would leave CTAs without a tile); a CTA
// stages the 2B root boxes of its tile in shared memory; a thread keeps one column object in registers and walks the
// rows, so a pair costs two broadcast 128-bit shared-memory reads and the test, and a root box is fetched n / B
// times in total instead of once per pair.
__device__ void seed_phase(float* __restrict__ s_roots, const ObjDesc* __restrict__ objs, uint32_t n_obj,
                           uint4* __restrict__ front, uint32_t front_cap, uint32_t* __restrict__ counters)
{
    uint32_t bshift = 8;
    auto tiles_of = [&](uint32_t sh) {
        const uint32_t nb = (n_obj + (1u << sh) - 1) >> sh;
        return nb * (nb + 1) / 2;
    };
    while (bshift > 5 && tiles_of(bshift) < gridDim.x) bshift--;
    const uint32_t B = 1u << bshift, nb = (n_obj + B - 1) >> bshift, tiles = nb * (nb + 1) / 2;
    const uint32_t lane = lane_id();
    float4* sa = reinterpret_cast<float4*>(s_roots);          // [B][2] row-block roots: (lx ly lz hx) (hy hz - -)
    float4* sb = reinterpret_cast<float4*>(s_roots) + 2 * B;  // [B][2] column-block roots
    uint32_t bi = 0, row_first = 0; // tile t = row_first(bi) + (bj - bi), rows of nb - bi tiles
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x)
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
                if (lane == 0) slot = atomicAdd(counters + CTR_FRONT0, (uint32_t)__popc(mask));
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (hit)
                {
                    const uint32_t dst = slot + __popc(mask & lanemask_lt());
                    if (dst < front_cap)
                        front[dst] = make_uint4(i0 + ii, j0 + jj, 0u, 0u);
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
#ifndef OIBVH_STAGE_CAP
#define OIBVH_STAGE_CAP 256
#endif
constexpr int kStageCap = OIBVH_STAGE_CAP; // records per warp staging buffer (4 KB), half per record kind
static_assert(kStageCap / 2 >= 64, "one item emits up to 64 records of one kind");

constexpr int kObjCache = 64; // object descriptors + level tables kept in shared memory (more: L1/L2 + arithmetic)

__device__ __forceinline__ ObjDesc get_obj(const ObjDesc* s_objs, const ObjDesc* __restrict__ objs, uint32_t i)
{
    // (a compact shared-memory table of all objects was measured slower than these L1/L2 hits: it costs L1 capacity)
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

// Set-up and testing are split. Everything that depends only on the item (descriptor fetch, level geometry,
// rectangle clamping: ~200 instructions) would be warp-uniform work, issued once per item for 32 identical lanes, and
// made wide fronts instruction-bound. So a warp takes a BATCH of up to 32 items: lane l sets up item l (phase 1, 32
// different items per issued instruction), then the warp walks the batch and broadcasts each item's ten parameters
// with shuffles (phase 2: index arithmetic, four box loads per lane, tests, staging). The batch size shrinks with
// the front so that small fronts still spread over every warp of the grid. Measured: -45 % on fronts of 10^5..10^6
// items, unchanged on latency-bound fronts of a few hundred.
__device__ void expand_phase(uint4* s_stage, const ObjDesc* s_objs, const uint32_t* s_lv,
                             const ObjDesc* __restrict__ objs, const uint4* in,
                                     uint4* out, uint32_t front_cap, uint4* cand, uint32_t cand_cap, uint32_t* counters,
                                     uint32_t round, uint32_t front_size, uint32_t levels, uint32_t rank, uint32_t world,
                                     uint32_t n_obj)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const bool computed_seeds = (round == 0 && in == nullptr);
    const uint32_t n = computed_seeds ? front_size : min(front_size, front_cap);
    uint32_t* next_count = counters + CTR_FRONT0 + round + 1;
    const uint32_t total_warps = gridDim.x * kColWarps;
    constexpr uint32_t kHalf = kStageCap / 2;
    uint4* stage = s_stage + warp * kStageCap;
    uint32_t staged_f = 0, staged_c = 0; // warp-uniform

    auto flush = [&](bool is_cand) {
        const uint32_t staged = is_cand ? staged_c : staged_f;
        if (staged == 0) return;
        __syncwarp();
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(is_cand ? counters + CTR_CANDIDATES : next_count, staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint4* dst = is_cand ? cand : out;
        const uint4* src = stage + (is_cand ? kHalf : 0u);
        const uint32_t cap = is_cand ? cand_cap : front_cap;
        if (base + staged > cap && lane == 0) atomicOr(counters + CTR_OVERFLOW, is_cand ? 2u : 1u);
        for (uint32_t j = lane; j < staged; j += 32)
            if (base + j < cap) dst[base + j] = src[j];
        __syncwarp();
        if (is_cand)
            staged_c = 0;
        else
            staged_f = 0;
    };

    const uint32_t gshift = 2 * levels > 6 ? 2 * levels - 6 : 0; // log2(64-combination groups per pair)
    const uint32_t items = n << gshift;                          // n < 2^26, gshift <= 4
    const uint32_t batch = min(32u, max(1u, (items + total_warps - 1) / total_warps));
    const bool sharded = world > 1 && round == 0;
    // fronts are produced by other SMs inside this same launch: read them through L2 (ld.cg); the entries of the
    // NEXT batch are fetched before the current one is processed (one round trip off the critical path)
    auto entry = [&](uint32_t w) {
        if (!(lane < batch && w < items)) return make_uint4(0, 0, 0, 0);
        return computed_seeds ? seed_entry(n_obj, w >> gshift) : __ldcg(in + (w >> gshift));
    };
    uint32_t wb = (blockIdx.x * kColWarps + warp) * batch;
    uint4 it_next = entry(wb + lane);
    for (; wb < items; wb += total_warps * batch)
    {
        // ---- phase 1: lane l prepares item wb + l ----
        const uint32_t w = wb + lane;
        bool valid = lane < batch && w < items;
        const uint4 it = it_next;
        it_next = entry(w + total_warps * batch);
        uint32_t ex = 0, ey = 0, za = 0, zb = 0, baseA = 0, baseB = 0, meta = 0, key = 0;
        uint64_t ptrA = 0, ptrB = 0;
        if (valid)
        {
            const uint32_t p = w >> gshift, group = w & ((1u << gshift) - 1);
            const ObjDesc A = get_obj(s_objs, objs, it.x), B = get_obj(s_objs, objs, it.y);
            const LevelView va = level_view(s_lv, it.x, A), vb = level_view(s_lv, it.y, B);
            const uint32_t la = it.z >> kNodeLevelShift, pa = it.z & kNodePosMask;
            const uint32_t lb = it.w >> kNodeLevelShift, pb = it.w & kNodePosMask;
            const uint32_t da = min(levels, A.L - la), db = min(levels, B.L - lb);
            const uint32_t lca = la + da, lcb = lb + db;
            const uint32_t fa = pa << da, fb = pb << db;
            const uint32_t nA = min(1u << da, va.count(lca) - fa);
            const uint32_t nB = min(1u << db, vb.count(lcb) - fb);
            const uint32_t combos = nA * nB; // <= 1024
            valid = group * 64 < combos;     // clamped rectangle: nothing in this group
            if (valid && computed_seeds)
            {
                // computed root pairs are untested: prune object pairs whose root boxes are disjoint
                const Box ra = load_box(reinterpret_cast<const float2*>(A.nodes), va.offset(la) + pa);
                const Box rb = load_box(reinterpret_cast<const float2*>(B.nodes), vb.offset(lb) + pb);
                valid = box_overlap(ra, rb);
            }
            const bool to_cand = (lca == A.L) && (lcb == B.L);
            ptrA = (uint64_t)A.nodes;
            ptrB = (uint64_t)B.nodes;
            baseA = va.offset(lca) + fa;
            baseB = vb.offset(lcb) + fb;
            ex = it.x;
            ey = it.y;
            // emitted node ids are za + ia / zb + ib (fa, fb have their low da / db bits clear)
            za = to_cand ? fa : ((lca << kNodeLevelShift) | fa);
            zb = to_cand ? fb : ((lcb << kNodeLevelShift) | fb);
            meta = combos | (nB << 11) | (db << 17) | ((nB == (1u << db)) ? 1u << 20 : 0u) | (to_cand ? 1u << 21 : 0u) |
                   (group << 22);
            // round 0 deals the seed rectangle round-robin to the shards, keyed by the object pair's linear index
            // (not by the nondeterministic slot of a seeded entry)
            if (sharded) key = computed_seeds ? p : pair_linear(n_obj, it.x, it.y);
        }
        // ---- phase 2: the warp walks the prepared items ----
        for (uint32_t todo = __ballot_sync(0xffffffffu, valid); todo; todo &= todo - 1)
        {
            const int k = __ffs(todo) - 1;
            const uint32_t mk = __shfl_sync(0xffffffffu, meta, k);
            const uint32_t combos = mk & 0x7ffu, nB = (mk >> 11) & 63u, db = (mk >> 17) & 7u, group = mk >> 22;
            const bool to_cand = (mk >> 21) & 1u;
            const float2* nodesA = reinterpret_cast<const float2*>(__shfl_sync(0xffffffffu, ptrA, k));
            const float2* nodesB = reinterpret_cast<const float2*>(__shfl_sync(0xffffffffu, ptrB, k));
            const uint32_t bA = __shfl_sync(0xffffffffu, baseA, k), bB = __shfl_sync(0xffffffffu, baseB, k);
            const uint32_t c0 = group * 64 + lane, c1 = c0 + 32;
            bool hit0 = c0 < combos, hit1 = c1 < combos;
            if (sharded)
            {
                const uint32_t kk = __shfl_sync(0xffffffffu, key, k);
                hit0 = hit0 && ((kk + c0) % world) == rank;
                hit1 = hit1 && ((kk + c1) % world) == rank;
            }
            // nB is a power of two unless the rectangle is clamped by the end of the level (warp-uniform either way)
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
            const uint32_t mask0 = __ballot_sync(0xffffffffu, hit0), mask1 = __ballot_sync(0xffffffffu, hit1);
            const uint32_t cnt0 = __popc(mask0), cnt = cnt0 + __popc(mask1);
            if (cnt == 0) continue; // warp-uniform
            if ((to_cand ? staged_c : staged_f) + cnt > kHalf) flush(to_cand);
            const uint32_t exk = __shfl_sync(0xffffffffu, ex, k), eyk = __shfl_sync(0xffffffffu, ey, k);
            const uint32_t zak = __shfl_sync(0xffffffffu, za, k), zbk = __shfl_sync(0xffffffffu, zb, k);
            uint4* st = stage + (to_cand ? kHalf + staged_c : staged_f);
            if (hit0) st[__popc(mask0 & lanemask_lt())] = make_uint4(exk, eyk, zak + ia0, zbk + ib0);
            if (hit1) st[cnt0 + __popc(mask1 & lanemask_lt())] = make_uint4(exk, eyk, zak + ia1, zbk + ib1);
            if (to_cand)
                staged_c += cnt;
            else
                staged_f += cnt;
        }
    }
    flush(false);
    flush(true);
}

// Dense round 0 for scenes of very few objects (two-body scenes): instead of walking down from the root pair through
// several latency-bound rounds of a handful of items (each costs a grid barrier plus ~4 dependent L2 round trips,
// ~4.5 us, whatever its size), ALL node pairs (a, b) at level `k0` of an object pair are tested directly, spread over
// every warp of the grid: 4^8 = 65 K .. 4^11 = 4 M box tests are less work than the rounds they replace. A level is a
// contiguous slice, so a warp's 32 consecutive combinations read one broadcast box of A and 32 consecutive boxes of B.
// Same emission path (per-warp staging, one atomic per flush) and the same shard rule as expand_phase.
__device__ void dense_seed_phase(uint4* s_stage, const ObjDesc* s_objs, const uint32_t* s_lv,
                                 const ObjDesc* __restrict__ objs, uint4* out, uint32_t front_cap, uint4* cand,
                                 uint32_t cand_cap, uint32_t* counters, uint32_t n_pairs, uint32_t k0, uint32_t rank,
                                 uint32_t world, uint32_t n_obj)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t* next_count = counters + CTR_FRONT0 + 1;
    const uint32_t total_warps = gridDim.x * kColWarps;
    constexpr uint32_t kHalf = kStageCap / 2;
    uint4* stage = s_stage + warp * kStageCap;
    uint32_t staged = 0; // warp-uniform; one kind per object pair (flushed before the kind can change)
    auto flush = [&](bool is_cand) {
        if (staged == 0) return;
        __syncwarp();
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(is_cand ? counters + CTR_CANDIDATES : next_count, staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint4* dst = is_cand ? cand : out;
        const uint32_t cap = is_cand ? cand_cap : front_cap;
        if (base + staged > cap && lane == 0) atomicOr(counters + CTR_OVERFLOW, is_cand ? 2u : 1u);
        for (uint32_t j = lane; j < staged; j += 32)
            if (base + j < cap) dst[base + j] = stage[j];
        __syncwarp();
        staged = 0;
    };
    const uint32_t gw = blockIdx.x * kColWarps + warp;
    for (uint32_t p = 0; p < n_pairs; p++)
    {
        const uint4 it = seed_entry(n_obj, p);
        const ObjDesc A = get_obj(s_objs, objs, it.x), B = get_obj(s_objs, objs, it.y);
        const LevelView va = level_view(s_lv, it.x, A), vb = level_view(s_lv, it.y, B);
        const float2* nodesA = reinterpret_cast<const float2*>(A.nodes);
        const float2* nodesB = reinterpret_cast<const float2*>(B.nodes);
        if (!box_overlap(load_box(nodesA, 0), load_box(nodesB, 0))) continue; // disjoint roots (warp-uniform)
        const uint32_t ka = min(k0, A.L), kb = min(k0, B.L);
        const uint32_t nA = va.count(ka), nB = vb.count(kb), baseA = va.offset(ka), baseB = vb.offset(kb);
        const bool to_cand = (ka == A.L) && (kb == B.L);
        const uint32_t za = to_cand ? 0u : (ka << kNodeLevelShift), zb = to_cand ? 0u : (kb << kNodeLevelShift);
        const uint32_t total = nA * nB; // <= 4^11
        for (uint32_t c0 = gw * 64; c0 < total; c0 += total_warps * 64)
        {
            const uint32_t c[2] = {c0 + lane, c0 + 32 + lane};
            bool hit[2];
            uint32_t ia[2], ib[2];
            Box a[2], b[2];
#pragma unroll
            for (int u = 0; u < 2; u++)
            {
                hit[u] = c[u] < total && (world == 1 || ((p + c[u]) % world) == rank);
                ia[u] = c[u] / nB;
                ib[u] = c[u] - ia[u] * nB;
                if (hit[u])
                {
                    a[u] = load_box(nodesA, baseA + ia[u]);
                    b[u] = load_box(nodesB, baseB + ib[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (hit[u]) hit[u] = box_overlap(a[u], b[u]);
            const uint32_t mask0 = __ballot_sync(0xffffffffu, hit[0]), mask1 = __ballot_sync(0xffffffffu, hit[1]);
            const uint32_t cnt0 = __popc(mask0), cnt = cnt0 + __popc(mask1);
            if (cnt == 0) continue; // warp-uniform
            if (staged + cnt > kHalf) flush(to_cand);
            if (hit[0]) stage[staged + __popc(mask0 & lanemask_lt())] = make_uint4(it.x, it.y, za + ia[0], zb + ib[0]);
            if (hit[1]) stage[staged + cnt0 + __popc(mask1 & lanemask_lt())] = make_uint4(it.x, it.y, za + ia[1], zb + ib[1]);
            staged += cnt;
        }
        flush(to_cand);
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

// REMOTE = this rank is not the gathering rank: hits are appended to the ROOT's pair list through the peer mapping
// (one system-scope atomic per warp claims the slots, then 16-byte stores over NVLink); the local counter still
// counts this rank's own hits.
template <bool REMOTE>
__device__ void narrow_phase(const ObjDesc* s_objs, const ObjDesc* __restrict__ objs, const uint4* cand, uint32_t cand_cap,
                             uint4* __restrict__ pairs, uint32_t pair_cap, uint32_t* counters, uint32_t n_cand,
                             uint32_t* root_counters)
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
            // full descriptors (faces, vertices): shared-memory cache for the first objects, global otherwise
            const ObjDesc A = c.x < (uint32_t)kObjCache ? s_objs[c.x] : load_obj(objs, c.x);
            const ObjDesc B = c.y < (uint32_t)kObjCache ? s_objs[c.y] : load_obj(objs, c.y);
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
            if (lane == 0)
            {
                if (REMOTE)
                {
                    atomicAdd(counters + CTR_PAIRS, (uint32_t)__popc(mask)); // this rank's own count
                    base = atomicAdd_system(root_counters + CTR_PAIRS, (uint32_t)__popc(mask));
                }
                else
                    base = atomicAdd(counters + CTR_PAIRS, (uint32_t)__popc(mask));
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit)
            {
                const uint32_t dst = base + __popc(mask & lanemask_lt());
                if (dst < pair_cap)
                    pairs[dst] = c; // {bvhA, bvhB, triA, triB} == int_tri_pair_node_t
                else if (REMOTE)
                    atomicOr_system(root_counters + CTR_OVERFLOW, 4u);
                else
                    atomicOr(counters + CTR_OVERFLOW, 4u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// The persistent detection kernel
// ---------------------------------------------------------------------------------------------------
constexpr size_t kColStageBytes = (size_t)kColWarps * kStageCap * sizeof(uint4); // 128 KB
constexpr size_t kColSmemBytes = kColStageBytes;
__global__ void __launch_bounds__(kColThreads, 1)
    collide_kernel(const ObjDesc* __restrict__ objs, uint32_t n_obj, uint4* front0, uint4* front1, uint32_t front_cap,
                   uint4* cand, uint32_t cand_cap, uint4* pairs, uint32_t pair_cap, uint32_t* counters,
                   uint32_t rounds, uint32_t levels0, uint32_t levels, uint32_t rank, uint32_t world, const MgpuArgs mg)
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
    // multi-GPU: this launch is frame MG_FRAME + 1 of the scene. A remote rank may append to the root's list only once
    // the root has zeroed its counter block for this frame (MG_OPEN >= frame): one thread per CTA polls the root's word
    // over NVLink NOW, off the critical path, and the narrow phase picks the answer up from shared memory.
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
        seed_phase(reinterpret_cast<float*>(smem_raw), objs, n_obj, front0, front_cap, counters);
        front_size = grid_barrier(counters, ++gen, counters + CTR_FRONT0);
    }
    stamp(1);
    for (uint32_t r = 0; r < rounds && front_size != 0; r++) // front_size is uniform over the grid
    {
        uint4* in = (r & 1) ? front1 : front0;
        uint4* out = (r & 1) ? front0 : front1;
        if (r == 0 && computed_seeds) in = nullptr;
        const uint32_t k = r == 0 ? levels0 : levels; // schedule chosen by the host (see scene_enqueue)
        if (r == 0 && computed_seeds && k > kMaxExpandLevels)
            dense_seed_phase(s_stage, s_objs, s_lv, objs, out, front_cap, cand, cand_cap, counters, front_size, k, rank,
                             world, n_obj);
        else
            expand_phase(s_stage, s_objs, s_lv, objs, in, out, front_cap, cand, cand_cap, counters, r, front_size, k,
                         rank, world, n_obj);
        front_size = grid_barrier(counters, ++gen, counters + CTR_FRONT0 + r + 1);
        stamp(gen);
    }
    const uint32_t n_cand = grid_barrier(counters, ++gen, counters + CTR_CANDIDATES);
    stamp(gen);
    if (mg.mode == 2)
    {
        if (s_mg_ok)
            narrow_phase<true>(s_objs, objs, cand, cand_cap, mg.root_pairs, mg.root_pair_cap, counters, n_cand,
                               mg.root_counters);
        else if (threadIdx.x == 0)
            atomicOr(counters + CTR_OVERFLOW, 16u); // the root never opened this frame
    }
    else
        narrow_phase<false>(s_objs, objs, cand, cand_cap, pairs, pair_cap, counters, n_cand, nullptr);
    __syncthreads();
    stamp(gen + 1);
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
                atomicAdd_system(mg.root_state + MG_DONE, 1u);
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
                           uint32_t world, const MgpuArgs& mg, cudaStream_t s)
{
    MgpuArgs mga = mg;
    void* args[] = {&objs, &n_obj, &front0, &front1, &front_cap, &cand, &cand_cap, &pairs, &pair_cap,
                    &counters, &rounds, &levels0, &levels, &rank, &world, &mga};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(collide_kernel), dim3(grid_blocks),
                                       dim3(kColThreads), args, kColSmemBytes, s);
}

} // namespace oibvh
