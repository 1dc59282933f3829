// Build / refit kernels of the ostensibly-implicit BVH for sm_100a.
//
//   pack_* / unpack_*    boundary layout conversion: packed xyz / packed index triples <-> 16-byte records, so that
//                        every gather on the hot path is ONE 128-bit load (one L2 sector) instead of three
//   morton_hist_kernel   faces + positions -> 30-bit Morton keys (+ all radix-digit histograms in the same pass)
//   onesweep_pass_kernel one stable LSD radix pass (key, face id) with decoupled look-back (single sweep per digit)
//   tree_emit_kernel     leaf AABBs + the whole bottom-up AABB reduction of a 1024-leaf subtree per CTA in
//                        registers / warp shuffles / shared memory, coalesced level-slice stores, and the top of
//                        the tree finished by the last CTA to retire (one launch for the whole tree).
//                        BUILD variant also gathers the faces into Morton order.
//
// Reference behaviour being reproduced (not its code): src/cuda/oibvh.cu:6-22, 24-70 (keys, leaf boxes),
// src/cuda/oibvhTree.cu:287-299 (stable sort by key), src/cuda/oibvh.cu:72-219 (parent = left U right | left).
#include "common.cuh"
#include "kernels.h"

namespace oibvh
{

// =================================================================================================
// Layout conversion at the boundary
// =================================================================================================
__global__ void __launch_bounds__(256) pack_pos_kernel(const float* __restrict__ xyz, float4* __restrict__ out, uint32_t V)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float* p = xyz + 3ull * i;
    out[i] = make_float4(p[0], p[1], p[2], 1.0f);
}
__global__ void __launch_bounds__(256) unpack_pos_kernel(const float4* __restrict__ in, float* __restrict__ xyz, uint32_t V)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float4 v = in[i];
    float* p = xyz + 3ull * i;
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
__global__ void __launch_bounds__(256) pack_faces_kernel(const uint32_t* __restrict__ f3, uint4* __restrict__ out, uint32_t T)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint32_t* f = f3 + 3ull * i;
    out[i] = make_uint4(f[0], f[1], f[2], 0u);
}

// =================================================================================================
// Morton keys
// =================================================================================================
// strech_by_3 / morton3D semantics (src/cuda/oibvh.cu:6-22): 10 bits per axis, x in the highest lane.
__device__ __forceinline__ uint32_t spread3(uint32_t x)
{
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x30000ffu;
    x = (x | (x << 8)) & 0x300f00fu;
    x = (x | (x << 4)) & 0x30c30c3u;
    x = (x | (x << 2)) & 0x9249249u;
    return x;
}

// (unsigned)thrust::min(thrust::max(q * 1024.0f, 0.0f), 1023.0f) with thrust's operand order
// (max: lhs < rhs ? rhs : lhs, min: rhs < lhs ? rhs : lhs) so NaN propagates and converts to 0 like
// cvt.rzi.u32.f32 does (planar meshes give 0/0 on the flat axis).
__device__ __forceinline__ uint32_t quantise10(float q)
{
    const float s = __fmul_rn(q, 1024.0f);
    const float a = (s < 0.0f) ? 0.0f : s;
    const float b = (1023.0f < a) ? 1023.0f : a;
    return (b != b) ? 0u : __float2uint_rz(b);
}

// glm::min(glm::min(v0, v1), v2) / glm::max(glm::max(v0, v1), v2)   (src/cuda/oibvh.cu:33-39)
__device__ __forceinline__ Box box_of(const float4& a, const float4& b, const float4& c)
{
    Box o;
    o.lx = gmin(gmin(a.x, b.x), c.x);
    o.ly = gmin(gmin(a.y, b.y), c.y);
    o.lz = gmin(gmin(a.z, b.z), c.z);
    o.hx = gmax(gmax(a.x, b.x), c.x);
    o.hy = gmax(gmax(a.y, b.y), c.y);
    o.hz = gmax(gmax(a.z, b.z), c.z);
    return o;
}

// centroid = (min + max) * 0.5 ; offset = centroid - meshMin ; q = offset / (meshMax - meshMin)  (oibvh.cu:63-69)
// written with explicit round-to-nearest intrinsics so nothing is contracted into an FMA.
__device__ __forceinline__ uint32_t morton_of_box(const Box& b, const MeshAabb& m)
{
    const float cx = __fmul_rn(__fadd_rn(b.lx, b.hx), 0.5f);
    const float cy = __fmul_rn(__fadd_rn(b.ly, b.hy), 0.5f);
    const float cz = __fmul_rn(__fadd_rn(b.lz, b.hz), 0.5f);
    const float qx = __fdiv_rn(__fsub_rn(cx, m.v[0]), __fsub_rn(m.v[3], m.v[0]));
    const float qy = __fdiv_rn(__fsub_rn(cy, m.v[1]), __fsub_rn(m.v[4], m.v[1]));
    const float qz = __fdiv_rn(__fsub_rn(cz, m.v[2]), __fsub_rn(m.v[5], m.v[2]));
    return (spread3(quantise10(qx)) << 2) | (spread3(quantise10(qy)) << 1) | spread3(quantise10(qz));
}

#ifndef OIBVH_MORTON_FPT
#define OIBVH_MORTON_FPT 4
#endif
#ifndef OIBVH_MORTON_CTAS_PER_SM
#define OIBVH_MORTON_CTAS_PER_SM 5
#endif
constexpr int kMortonFacesPerThread = OIBVH_MORTON_FPT;

template <int RADIX_BITS, int PASSES, bool HIST>
__global__ void __launch_bounds__(256) morton_hist_kernel(const uint4* __restrict__ faces4,
                                                          const float4* __restrict__ pos4, uint32_t T, MeshAabb mesh,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ hist)
{
    constexpr int RADIX = 1 << RADIX_BITS;
    __shared__ uint32_t sh[HIST ? PASSES * RADIX : 1];
    if (HIST)
    {
        for (int i = threadIdx.x; i < PASSES * RADIX; i += blockDim.x) sh[i] = 0;
        __syncthreads();
    }

    // block-strided tiles of 256 x 4 faces; lane-consecutive faces are memory-consecutive (one 128-bit load each)
    const uint32_t tile = 256 * kMortonFacesPerThread;
    for (uint64_t base = (uint64_t)blockIdx.x * tile; base < T; base += (uint64_t)gridDim.x * tile)
    {
        uint4 f[kMortonFacesPerThread];
        bool valid[kMortonFacesPerThread];
#pragma unroll
        for (int k = 0; k < kMortonFacesPerThread; k++)
        {
            const uint64_t i = base + k * 256 + threadIdx.x;
            valid[k] = i < T;
            f[k] = valid[k] ? ldg_stream_u4(faces4 + i) : make_uint4(0, 0, 0, 0);
        }
        float4 v[kMortonFacesPerThread][3];
#pragma unroll
        for (int k = 0; k < kMortonFacesPerThread; k++)
        {
            v[k][0] = __ldg(pos4 + f[k].x);
            v[k][1] = __ldg(pos4 + f[k].y);
            v[k][2] = __ldg(pos4 + f[k].z);
        }
#pragma unroll
        for (int k = 0; k < kMortonFacesPerThread; k++)
        {
            if (!valid[k]) continue;
            const uint32_t key = morton_of_box(box_of(v[k][0], v[k][1], v[k][2]), mesh);
            keys[base + k * 256 + threadIdx.x] = key;
            if (HIST)
            {
#pragma unroll
                for (int p = 0; p < PASSES; p++)
                    atomicAdd(sh + p * RADIX + ((key >> (p * RADIX_BITS)) & (RADIX - 1)), 1u);
            }
        }
    }
    if (HIST)
    {
        __syncthreads();
        for (int i = threadIdx.x; i < PASSES * RADIX; i += blockDim.x)
        {
            const uint32_t c = sh[i];
            if (c) atomicAdd(hist + i, c);
        }
    }
}

// =================================================================================================
// Onesweep: one stable LSD pass with chained-scan decoupled look-back.
// status word: [31:30] flag (0 = not ready, 1 = tile aggregate, 2 = inclusive prefix), [29:0] count
// =================================================================================================
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kCountMask = ~kFlagMask;
constexpr int kLookbackBatch = 8; // predecessor status words fetched per round trip

#ifdef OIBVH_PROFILE
__device__ unsigned long long g_sort_prof[4][2][8]; // [pass][first/last tile][stamp]
#define SORT_STAMP(k)                                                                                              \
    do                                                                                                             \
    {                                                                                                              \
        if (threadIdx.x == 0 && (tile == 0 || tile == gridDim.x - 1))                                              \
            g_sort_prof[shift / RADIX_BITS][tile == 0 ? 0 : 1][k] = clock64();                                     \
    } while (0)
#else
#define SORT_STAMP(k)
#endif

template <int RADIX_BITS, int IPT>
__global__ void __launch_bounds__(1 << RADIX_BITS)
    onesweep_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                         uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t T, uint32_t shift,
                         const uint32_t* __restrict__ hist, uint32_t* status, uint32_t* ticket)
{
    constexpr int RADIX = 1 << RADIX_BITS;
    constexpr int THREADS = RADIX; // one thread per digit in the scan / look-back phases
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * IPT;
    constexpr uint32_t MASK = RADIX - 1;

    __shared__ uint32_t s_hist[WARPS][RADIX];
    __shared__ uint32_t s_keys[TILE];
    // per-warp peer masks of the ranking phase live in the (not yet used) key staging area
    static_assert(WARPS * RADIX <= TILE, "peer masks must fit in the key staging area");
    uint32_t(*s_mask)[RADIX] = reinterpret_cast<uint32_t(*)[RADIX]>(s_keys);
    __shared__ uint32_t s_vals[TILE];
    __shared__ uint32_t s_digit_base[RADIX];
    __shared__ uint32_t s_global_base[RADIX];
    __shared__ uint32_t s_scan[WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u); // tiles are claimed in launch order: predecessors are resident
#pragma unroll
    for (int w = 0; w < WARPS; w++)
    {
        s_hist[w][tid] = 0;
        s_mask[w][tid] = 0;
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * TILE;
    const uint32_t tile_valid = min((uint32_t)TILE, T - tile_base);
    SORT_STAMP(0);

    // ---- load (warp-striped: lane-consecutive keys are memory-consecutive) + stable in-warp ranking ----
    uint32_t key[IPT], val[IPT];
    uint16_t rank[IPT];
    const uint32_t warp_base = tile_base + warp * (32 * IPT);
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        key[j] = (i < T) ? ldg_stream_u32(keys_in + i) : 0xffffffffu;
    }
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        val[j] = (vals_in && i < T) ? ldg_stream_u32(vals_in + i) : i; // first pass: the value is the face id
    }
    SORT_STAMP(1);
    // Stable in-warp ranking. Peer groups (lanes holding the same digit) are found with one shared-memory atomicOr
    // per lane into a per-warp, per-digit lane mask: constant cost, unlike match.any whose latency grows with the
    // number of distinct digits in the warp (32 for the high-entropy low digits of a Morton key).
    uint32_t* my_hist = s_hist[warp];
    uint32_t* my_mask = s_mask[warp];
    const uint32_t lane_bit = 1u << lane;
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        const bool valid = i < T;
        const uint32_t d = (key[j] >> shift) & MASK;
        if (valid) atomicOr(my_mask + d, lane_bit);
        __syncwarp();
        uint32_t peers = 0, before = 0;
        if (valid)
        {
            peers = my_mask[d];
            before = my_hist[d];
        }
        const uint32_t lower = __popc(peers & lanemask_lt());
        rank[j] = (uint16_t)(before + lower);
        __syncwarp();
        if (valid && lower == 0) // lowest lane of the group closes it
        {
            my_hist[d] = before + __popc(peers);
            my_mask[d] = 0;
        }
        __syncwarp();
    }
    __syncthreads();

    SORT_STAMP(2);
    // ---- per-digit: warp-exclusive offsets, tile count, tile-local digit base, global base via look-back ----
    {
        const uint32_t d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++)
        {
            const uint32_t c = s_hist[w][d];
            s_hist[w][d] = run;
            run += c;
        }
        const uint32_t tile_count = run;
        uint32_t* my_status = status + (size_t)tile * RADIX + d;
        // publish the aggregate first so successors can make progress while we scan
        st_relaxed_gpu(my_status, (tile == 0 ? kFlagPrefix : kFlagAggregate) | tile_count);

        const uint32_t bin_start = block_exclusive_scan<THREADS>(hist[d], s_scan);
        const uint32_t digit_base = block_exclusive_scan<THREADS>(tile_count, s_scan);

        SORT_STAMP(3);
        uint32_t exclusive = 0;
        if (tile != 0)
        {
            // decoupled look-back, kLookbackBatch predecessors per round trip: consume them in order, stop at the
            // first inclusive prefix, re-issue from the first unpublished one. (Efficient in the streaming regime,
            // tiles >> resident CTAs, where the predecessor is usually complete; single-wave sizes use the
            // cooperative kernel below instead.)
            int t = (int)tile - 1;
            bool done = false;
            while (!done)
            {
                uint32_t w[kLookbackBatch];
#pragma unroll
                for (int b = 0; b < kLookbackBatch; b++)
                    w[b] = (t - b >= 0) ? ld_relaxed_gpu(status + (size_t)(t - b) * RADIX + d) : kFlagPrefix;
                bool stop = false;
#pragma unroll
                for (int b = 0; b < kLookbackBatch; b++)
                {
                    const uint32_t f = w[b] & kFlagMask;
                    if (!stop && !done)
                    {
                        if (f == 0)
                            stop = true; // not published yet (it is resident: ticket order) -> retry from here
                        else
                        {
                            exclusive += w[b] & kCountMask;
                            t--;
                            if (f == kFlagPrefix) done = true;
                        }
                    }
                }
            }
            st_relaxed_gpu(my_status, kFlagPrefix | (exclusive + tile_count));
        }
        s_digit_base[d] = digit_base;
        s_global_base[d] = bin_start + exclusive - digit_base;
    }
    __syncthreads();

    SORT_STAMP(4);
    // ---- reorder inside the tile through shared memory, then write digit runs coalesced ----
    uint16_t slot[IPT];
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        const uint32_t d = (key[j] >> shift) & MASK;
        slot[j] = (uint16_t)(s_digit_base[d] + my_hist[d] + rank[j]);
        if (i < T) s_keys[slot[j]] = key[j];
    }
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        if (i < T) s_vals[slot[j]] = val[j];
    }
    __syncthreads();
    SORT_STAMP(5);
#pragma unroll
    for (int k = 0; k < IPT; k++)
    {
        const uint32_t s = tid + k * THREADS;
        if (s < tile_valid)
        {
            const uint32_t kk = s_keys[s];
            const uint32_t dst = s_global_base[(kk >> shift) & MASK] + s;
            keys_out[dst] = kk;
            vals_out[dst] = s_vals[s];
        }
    }
    SORT_STAMP(6);
}

// =================================================================================================
// Leaf boxes + bottom-up reduction, one CHUNK-leaf subtree per CTA, one 128-leaf subtree per warp.
//
// thread: 4 consecutive leaves -> heights 0..2 in registers
// warp:   heights 3..7 by shuffles; the warp's nodes of every height are a contiguous slice of that level in
//         global memory, so they are staged in a 3 KB per-warp buffer and copied out coalesced
// CTA:    heights 8..10 from the 8 warp roots
// grid:   the last CTA to retire reduces the chunk roots to the tree root (finish_top)
// =================================================================================================
struct LevelTable
{
    uint32_t off[32]; // first node of level l in the node array
    uint32_t cnt[32]; // nodes kept at level l
};

#ifndef OIBVH_EMIT_THREADS
#define OIBVH_EMIT_THREADS 256
#endif
constexpr int kEmitThreads = OIBVH_EMIT_THREADS;
constexpr int kEmitWarps = kEmitThreads / 32;
constexpr int kLeavesPerThread = 4;
constexpr int kWarpLeaves = 32 * kLeavesPerThread;      // 128
constexpr int kChunk = kEmitThreads * kLeavesPerThread; // 1024 leaves
constexpr int kChunkLevels = kEmitThreads == 256 ? 10 : (kEmitThreads == 128 ? 9 : (kEmitThreads == 512 ? 11 : 8)); // log2(kChunk)
static_assert((1 << kChunkLevels) == kChunk, "chunk must be a power of two");
constexpr int kWarpLevels = 7;                          // log2(kWarpLeaves)

// slot of the first height-h node inside a warp's staging area (h = 1..7): 0, 64, 96, 112, 120, 124, 126
__device__ __forceinline__ constexpr int woff(int h) { return kWarpLeaves - (2 * kWarpLeaves >> h); }

__device__ __forceinline__ void stage_box(float2* sm, int slot, const Box& b)
{
    sm[3 * slot] = make_float2(b.lx, b.ly);
    sm[3 * slot + 1] = make_float2(b.lz, b.hx);
    sm[3 * slot + 2] = make_float2(b.hy, b.hz);
}
__device__ __forceinline__ Box unstage_box(const float2* sm, int slot)
{
    const float2 a = sm[3 * slot], b = sm[3 * slot + 1], c = sm[3 * slot + 2];
    Box o;
    o.lx = a.x; o.ly = a.y; o.lz = b.x; o.hx = b.y; o.hy = c.x; o.hz = c.y;
    return o;
}
__device__ __forceinline__ Box shfl_down_box(const Box& b, int delta)
{
    Box o;
    o.lx = __shfl_down_sync(0xffffffffu, b.lx, delta);
    o.ly = __shfl_down_sync(0xffffffffu, b.ly, delta);
    o.lz = __shfl_down_sync(0xffffffffu, b.lz, delta);
    o.hx = __shfl_down_sync(0xffffffffu, b.hx, delta);
    o.hy = __shfl_down_sync(0xffffffffu, b.hy, delta);
    o.hz = __shfl_down_sync(0xffffffffu, b.hz, delta);
    return o;
}
// global nodes written by other CTAs are read through L2 (never the non-coherent path)
__device__ __forceinline__ Box load_box_cg(const float2* nodes, uint32_t idx)
{
    const float2* p = nodes + 3ull * idx;
    const float2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
    Box o;
    o.lx = a.x; o.ly = a.y; o.lz = b.x; o.hx = b.y; o.hy = c.x; o.hz = c.y;
    return o;
}

// warp-cooperative copy of n8 8-byte words from shared to global memory; 16-byte stores wherever dst allows
__device__ __forceinline__ void warp_copy_out(float2* __restrict__ dst, const float2* src, uint32_t n8, uint32_t lane)
{
    uint32_t head = (uint32_t)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1u); // 1 if dst is only 8-byte aligned
    if (head > n8) head = n8;
    if (lane == 0 && head) dst[0] = src[0];
    const uint32_t body = (n8 - head) >> 1; // 16-byte words
    float4* d4 = reinterpret_cast<float4*>(dst + head);
    for (uint32_t i = lane; i < body; i += 32)
    {
        const float2 a = src[head + 2 * i], b = src[head + 2 * i + 1];
        d4[i] = make_float4(a.x, a.y, b.x, b.y);
    }
    if (lane == 0 && ((n8 - head) & 1u)) dst[n8 - 1] = src[n8 - 1];
}

// shared memory: per-warp staging areas for heights 0, 1, 2 (3072 + 1536 + 768 B, each with 16 B of slack for the
// alignment phase), 5440 B per warp, 42.5 KB per CTA
constexpr int kWarpStageBytes = 3072 + 16 + 1536 + 16 + 768 + 16 + 16;
static_assert(kWarpStageBytes % 16 == 0, "per-warp staging must keep 16-byte alignment");
constexpr size_t kEmitSmemBytes = (size_t)kEmitWarps * kWarpStageBytes;

// ---- TMA bulk store of a level slice (full chunks): one elected lane copies the warp's staged nodes to global ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Issue the copy of `bytes` staged bytes to gdst (8-byte aligned). The staging area was filled at offset mis * 8
// (mis = 1 iff gdst is not 16-byte aligned) so that source and destination share their 16-byte phase: the aligned
// interior goes out as one bulk copy, the 8-byte head/tail of a misaligned slice as plain stores. Lane 0 only.
__device__ __forceinline__ void slice_store(float2* gdst, const float2* stage /*16B aligned*/, uint32_t bytes, uint32_t mis)
{
    if (mis == 0)
        bulk_store(gdst, stage, bytes);
    else
    {
        gdst[0] = stage[1];
        bulk_store(gdst + 1, stage + 2, bytes - 16);
        gdst[bytes / 8 - 1] = stage[bytes / 8];
    }
}

// write N floats (N % 4 == 0) held in registers to shared memory at `lane_base + mis * 8` bytes, lane_base 16-byte
// aligned: 128-bit stores, with an 8-byte head and tail when mis = 1
template <int N>
__device__ __forceinline__ void stage_floats(float2* lane_base, const float (&f)[N], uint32_t mis)
{
    if (mis == 0)
    {
        float4* d = reinterpret_cast<float4*>(lane_base);
#pragma unroll
        for (int q = 0; q < N / 4; q++) d[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    }
    else
    {
        lane_base[1] = make_float2(f[0], f[1]);
        float4* d = reinterpret_cast<float4*>(lane_base + 2);
#pragma unroll
        for (int q = 0; q < N / 4 - 1; q++)
            d[q] = make_float4(f[4 * q + 2], f[4 * q + 3], f[4 * q + 4], f[4 * q + 5]);
        lane_base[N / 2] = make_float2(f[N - 2], f[N - 1]);
    }
}
__device__ __forceinline__ void box_floats(const Box& b, float* f)
{
    f[0] = b.lx; f[1] = b.ly; f[2] = b.lz; f[3] = b.hx; f[4] = b.hy; f[5] = b.hz;
}

#ifdef OIBVH_PROFILE
// phase stamps of warp 1 / lane 0 of every CTA (ns, %globaltimer): [0] entry, [1] faces arrived, [2] vertices arrived,
// [3] heights 0..2 staged and bulk stores issued, [4] heights 3..7 done, [5] bulk stores have read shared memory
__device__ unsigned long long g_emit_phase[4096][8];
__device__ __forceinline__ void phase_stamp(int k, uint32_t consume, bool on)
{
    __shared__ uint32_t s_prof_scratch[32];
    if (!on) return;
    asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_prof_scratch + (k & 31))), "r"(consume) : "memory");
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    if (blockIdx.x < 4096) g_emit_phase[blockIdx.x][k] = t;
}
extern "C" int oibvh_debug_emit_phases(unsigned long long* out /* 4096*8 */)
{
    return (int)cudaMemcpyFromSymbol(out, g_emit_phase, sizeof(g_emit_phase));
}
#define PHASE_STAMP(k, x) phase_stamp(k, (uint32_t)(x), lane == 0 && ((warp_leaf0 / kWarpLeaves) % kEmitWarps) == 1)
#else
#define PHASE_STAMP(k, x)
#endif

// One warp's 128 leaves: faces -> leaf boxes -> heights 1..7, all stored. FULL = every leaf of the chunk exists
// (compile-time true folds every bounds predicate away; the partial last chunk takes the generic instantiation).
// Returns the warp root (height 7) in lane 0.
template <bool BUILD, bool FULL>
__device__ __forceinline__ Box emit_warp(const uint4* __restrict__ faces_in4, const uint32_t* __restrict__ perm,
                                         uint32_t* __restrict__ faces_sorted, const float4* __restrict__ pos4,
                                         float2* __restrict__ nodes, uint32_t T, uint32_t L, const LevelTable& lv,
                                         float2* wsm, uint32_t warp_leaf0, uint32_t lane)
{
    const uint32_t leaf0 = warp_leaf0 + lane * kLeavesPerThread;
    auto in = [&](uint32_t i) { return FULL || i < T; };
    PHASE_STAMP(0, lane);
    // ---- faces of this thread's 4 consecutive leaves ----
    uint32_t idx[12];
    const bool full = FULL || leaf0 + 4 <= T;
    if (BUILD)
    {
        uint32_t id[4];
        if (full)
        {
            const uint4 q = ldg_stream_u4(reinterpret_cast<const uint4*>(perm + leaf0));
            id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 4; k++) id[k] = in(leaf0 + k) ? perm[leaf0 + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const uint4 f = in(leaf0 + k) ? __ldg(faces_in4 + id[k]) : make_uint4(0, 0, 0, 0);
            idx[3 * k] = f.x; idx[3 * k + 1] = f.y; idx[3 * k + 2] = f.z;
        }
        if (full)
        {
            uint4* o = reinterpret_cast<uint4*>(faces_sorted + 3ull * leaf0);
            o[0] = make_uint4(idx[0], idx[1], idx[2], idx[3]);
            o[1] = make_uint4(idx[4], idx[5], idx[6], idx[7]);
            o[2] = make_uint4(idx[8], idx[9], idx[10], idx[11]);
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 12; k++)
                if (in(leaf0 + k / 3)) faces_sorted[3ull * leaf0 + k] = idx[k];
        }
    }
    else
    {
        if (full)
        {
            const uint4* f4 = reinterpret_cast<const uint4*>(faces_sorted + 3ull * leaf0);
            const uint4 a = ldg_stream_u4(f4), b = ldg_stream_u4(f4 + 1), c = ldg_stream_u4(f4 + 2);
            idx[0] = a.x; idx[1] = a.y; idx[2] = a.z; idx[3] = a.w;
            idx[4] = b.x; idx[5] = b.y; idx[6] = b.z; idx[7] = b.w;
            idx[8] = c.x; idx[9] = c.y; idx[10] = c.z; idx[11] = c.w;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 12; k++) idx[k] = in(leaf0 + k / 3) ? faces_sorted[3ull * leaf0 + k] : 0u;
        }
    }

    PHASE_STAMP(1, idx[11]);
    // ---- heights 0..2 in registers. A node at height h, position p exists iff p * 2^h < T. ----
    float4 v[12];
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = __ldg(pos4 + idx[k]); // idx = 0 for missing leaves: harmless load
    PHASE_STAMP(2, __float_as_uint(v[11].x) ^ __float_as_uint(v[5].y) ^ __float_as_uint(v[0].z));
    Box leaf[4];
#pragma unroll
    for (int k = 0; k < 4; k++) leaf[k] = box_of(v[3 * k], v[3 * k + 1], v[3 * k + 2]);
    Box h1a, h1b, cur;
    if (FULL)
    {
        // Heights 0..2 (87.5 % of the bytes): staged with 128-bit shared stores in three per-warp areas and sent
        // to global memory by TMA bulk copies (one elected lane, no per-lane copy loop). Level slices come from
        // the host-built table; a slice is 8- or 16-byte aligned depending on the parity of its first node.
        float2* st0 = wsm;                                   // 128 leaves   3072 B (+16)
        float2* st1 = wsm + (3072 + 16) / 8;                 //  64 nodes    1536 B (+16)
        float2* st2 = wsm + (3072 + 16 + 1536 + 16) / 8;     //  32 nodes     768 B (+16)
        float2* g0 = nodes + 3ull * (lv.off[L] + warp_leaf0);
        float2* g1 = nodes + 3ull * (lv.off[L - 1] + (warp_leaf0 >> 1));
        float2* g2 = nodes + 3ull * (lv.off[L - 2] + (warp_leaf0 >> 2));
        const uint32_t m0 = lv.off[L] & 1u, m1 = lv.off[L - 1] & 1u, m2 = lv.off[L - 2] & 1u;
        float f[24];
#pragma unroll
        for (int k = 0; k < 4; k++) box_floats(leaf[k], f + 6 * k);
        stage_floats<24>(st0 + lane * 12, f, m0);
        h1a = box_merge(leaf[0], leaf[1]);
        h1b = box_merge(leaf[2], leaf[3]);
        float g[12];
        box_floats(h1a, g);
        box_floats(h1b, g + 6);
        stage_floats<12>(st1 + lane * 6, g, m1);
        cur = box_merge(h1a, h1b);
        st2[m2 + lane * 3] = make_float2(cur.lx, cur.ly);
        st2[m2 + lane * 3 + 1] = make_float2(cur.lz, cur.hx);
        st2[m2 + lane * 3 + 2] = make_float2(cur.hy, cur.hz);
        fence_proxy_async_smem(); // generic-proxy writes above -> visible to the async proxy
        __syncwarp();
#ifndef OIBVH_EXP_NOSTORE // timing diagnostic only: skip the bulk stores of heights 0..2 (87.5 % of the written bytes)
        if (lane == 0)
        {
            slice_store(g0, st0, 3072, m0);
            slice_store(g1, st1, 1536, m1);
            slice_store(g2, st2, 768, m2);
            bulk_commit();
        }
#endif
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (in(leaf0 + k)) stage_box(wsm, lane * 4 + k, leaf[k]);
        const uint32_t warp_valid = (warp_leaf0 < T) ? min((uint32_t)kWarpLeaves, T - warp_leaf0) : 0u;
        __syncwarp();
        if (warp_valid) warp_copy_out(nodes + 3ull * (lv.off[L] + warp_leaf0), wsm, 3 * warp_valid, lane);
        __syncwarp();
        h1a = in(leaf0 + 1) ? box_merge(leaf[0], leaf[1]) : leaf[0];
        h1b = in(leaf0 + 3) ? box_merge(leaf[2], leaf[3]) : leaf[2];
        if (in(leaf0)) stage_box(wsm, lane * 2, h1a);     // height 1: slots [0, 64)
        if (in(leaf0 + 2)) stage_box(wsm, lane * 2 + 1, h1b);
        cur = in(leaf0 + 2) ? box_merge(h1a, h1b) : h1a;
        if (in(leaf0)) stage_box(wsm, 64 + lane, cur);    // height 2: slots [64, 96)
        __syncwarp();
        if (L >= 1 && warp_valid)
            warp_copy_out(nodes + 3ull * (lv.off[L - 1] + (warp_leaf0 >> 1)), wsm, 3 * ((warp_valid + 1) >> 1), lane);
        if (L >= 2 && warp_valid)
            warp_copy_out(nodes + 3ull * (lv.off[L - 2] + (warp_leaf0 >> 2)), wsm + 3 * 64, 3 * ((warp_valid + 3) >> 2),
                          lane);
    }
    PHASE_STAMP(3, __float_as_uint(cur.lx));
#pragma unroll
    for (int h = 3; h <= kWarpLevels; h++)
    {
        const int delta = 1 << (h - 3);
        const Box right = shfl_down_box(cur, delta);
        // right child = height h-1 node owned by lane + delta, first leaf = leaf0 + delta * 4
        if (in(leaf0 + (uint32_t)delta * 4)) cur = box_merge(cur, right);
        if ((lane & (2 * delta - 1)) == 0 && in(leaf0) && (FULL || (uint32_t)h <= L))
            store_box(nodes, lv.off[L - h] + (leaf0 >> h), cur);
    }
    PHASE_STAMP(4, __float_as_uint(cur.hx));
    if (FULL && lane == 0) bulk_wait_read(); // the staging areas may be reused / released after this
    PHASE_STAMP(5, lane);
    return cur;
}

#ifndef OIBVH_EMIT_MINB
#define OIBVH_EMIT_MINB 4
#endif
#ifndef OIBVH_EMIT_MINB_BUILD
#define OIBVH_EMIT_MINB_BUILD 3
#endif
// measured on B200: refit is fastest at 64 registers (4 CTAs/SM, no spills), build (one more gather level in
// flight) at 85 registers (3 CTAs/SM); 48 registers / 5 CTAs spills and loses 15-25 %
// Largest number of chunks whose upper levels go to the finisher CTA: one CTA of 8 warps handles a group of 32 chunks in
// ~2 dependent L2 round trips, which keeps pace with up to a few thousand chunks; beyond that (measured at 16 392
// chunks: 285 vs 251 us) the distributed scheme below (the last arriver of a group reduces it) scales better and its
// tail no longer matters.
#ifndef OIBVH_EMIT_FINISHER_MAX_CHUNKS
#define OIBVH_EMIT_FINISHER_MAX_CHUNKS 4096
#endif
// Levels above the chunk roots, done by ONE extra "finisher" CTA of the same launch (block 0).
// A chunk CTA only publishes: the lane that stored the chunk root adds 1 to the arrival counter of its group of 32
// chunks with a release-RED (fire and forget: no return value to wait for, the warp exits, the CTA slot is free for the
// next chunk). The finisher's warps poll the group counters (acquire), reduce every completed group of 32 nodes by
// shuffles (5 levels), and walk the remaining stages inside the CTA (CTA barrier between stages). Measured at
// T = 2^20 against the previous scheme (every CTA's warp 0: fence + atomicAdd with return, the last arriver of a group
// reduces it and arrives one stage up): a chunk CTA held its slot a median 1.5 us after its chunk was done, and the
// launch ended 4.5 us after the last chunk (two dependent fence/atomic/load stages).
constexpr uint32_t kFinisherSmemNodes = 512; // stage results handed to the next stage through shared memory
__device__ __noinline__ void emit_finisher(float2* __restrict__ nodes, uint32_t L, const LevelTable& lv, uint32_t* ctr,
                                              float2* scratch /* the CTA's (otherwise unused) staging area */,
                                              uint32_t* status)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t level = L - kChunkLevels; // level of the completed nodes that feed this stage
    bool first_stage = true;
    const float2* buf_in = nullptr;    // null: inputs come from global memory (stage 1: written by the chunk CTAs)
    float2* buf_out = scratch;
    while (level > 0)
    {
        const uint32_t n = lv.cnt[level];
        const uint32_t groups = (n + 31) >> 5;
        const uint32_t next_level = level > 5 ? level - 5 : 0;
        // a stage's outputs (one node per group, level - 5) stay in shared memory for the next stage: no L2 round trip
        const bool out_smem = next_level > 0 && groups <= kFinisherSmemNodes;
        for (uint32_t g = warp; g < groups; g += kEmitWarps)
        {
            const uint32_t first = g << 5;
            const uint32_t members = min(32u, n - first);
            if (first_stage)
            {
                if (lane == 0)
                {
                    uint32_t spins = 0;
                    while (ld_acquire_gpu(ctr + g) < members)
                        if (++spins > (1u << 25))
                        {
                            // a chunk CTA never arrived: report it through the context's status word (the host turns
                            // it into an error at the next download / synchronize) instead of trapping, which would
                            // poison the whole CUDA context of the process
                            atomicOr(status, 2u);
                            break;
                        }
                    ctr[g] = 0; // every member has arrived: re-arm for the next launch on this tree
                }
                __syncwarp();
            }
            Box cur = Box{0, 0, 0, 0, 0, 0};
            if (lane < members)
                cur = buf_in ? unstage_box(buf_in, first + lane) : load_box_cg(nodes, lv.off[level] + first + lane);
            uint32_t l = level;
#pragma unroll
            for (int s5 = 0; s5 < 5; s5++)
            {
                if (l == 0) break; // warp-uniform
                const Box right = shfl_down_box(cur, 1 << s5);
                const uint32_t child = (first >> s5) + (lane >> s5); // position at level l of this lane's node
                const bool owner = (lane & ((2u << s5) - 1)) == 0;
                if (owner && child < lv.cnt[l])
                {
                    if (child + 1 < lv.cnt[l]) cur = box_merge(cur, right);
                    store_box(nodes, lv.off[l - 1] + (child >> 1), cur);
                }
                l--;
            }
            if (out_smem && lane == 0) stage_box(buf_out, g, cur); // lane 0 holds node g of level - 5
        }
        __syncthreads(); // this stage's nodes are visible to the CTA's warps (shared memory, or global via ld.cg)
        buf_in = out_smem ? buf_out : nullptr;
        buf_out = (buf_out == scratch) ? scratch + 3 * kFinisherSmemNodes : scratch;
        level = next_level;
        first_stage = false;
    }
}

#ifdef OIBVH_PROFILE
// per-CTA wall-clock stamps (ns, %globaltimer): [0] CTA start, [1] warp 1 past the chunk, [2] warp 0 before the
// top-of-tree stages, [3] warp 0 done
__device__ unsigned long long g_emit_prof[4096][4];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define EMIT_STAMP(slot, cond)                                                                                     \
    do                                                                                                             \
    {                                                                                                              \
        if ((cond) && blockIdx.x < 4096) g_emit_prof[blockIdx.x][slot] = gtime();                                  \
    } while (0)
extern "C" int oibvh_debug_emit_profile(unsigned long long* out /* 4096*4 */)
{
    return (int)cudaMemcpyFromSymbol(out, g_emit_prof, sizeof(g_emit_prof));
}
#else
#define EMIT_STAMP(slot, cond)
#endif

template <bool BUILD>
__global__ void __launch_bounds__(kEmitThreads, BUILD ? OIBVH_EMIT_MINB_BUILD : OIBVH_EMIT_MINB)
    tree_emit_kernel(const uint4* __restrict__ faces_in4,     // BUILD: input-order faces (16-byte records)
                     const uint32_t* __restrict__ perm,       // BUILD: sorted position -> input face id
                     uint32_t* __restrict__ faces_sorted,     // BUILD: output ; else: input (packed triples)
                     const float4* __restrict__ pos4, float2* __restrict__ nodes, uint32_t T, uint32_t L,
                     const __grid_constant__ LevelTable lv, uint32_t* done_counter, uint32_t* status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    __shared__ float2 s_top[3 * 16]; // heights 7..10 of the chunk: 8 + 4 + 2 + 1 nodes

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    EMIT_STAMP(0, tid == 0);
    // block 0 is the finisher when the launch has one (the grid is then one CTA larger than the number of chunks). It is
    // resident from the start and reduces groups as they complete; as the LAST block it would start in the last wave and
    // then walk its groups one after the other: measured +13 us at T = 2^22.
    const uint32_t has_finisher = gridDim.x * (uint32_t)kChunk >= T + (uint32_t)kChunk ? 1u : 0u;
    if (has_finisher && blockIdx.x == 0)
    {
        emit_finisher(nodes, L, lv, done_counter, sm, status);
        EMIT_STAMP(3, tid == 0);
        return;
    }
    const uint32_t chunk = blockIdx.x - has_finisher;
    const uint32_t warp_leaf0 = chunk * kChunk + warp * kWarpLeaves;
    float2* wsm = sm + warp * (kWarpStageBytes / 8); // this warp's staging areas

    Box cur;
    if ((chunk + 1) * (uint32_t)kChunk <= T)
        cur = emit_warp<BUILD, true>(faces_in4, perm, faces_sorted, pos4, nodes, T, L, lv, wsm, warp_leaf0, lane);
    else
        cur = emit_warp<BUILD, false>(faces_in4, perm, faces_sorted, pos4, nodes, T, L, lv, wsm, warp_leaf0, lane);
    if (lane == 0 && warp_leaf0 < T) stage_box(s_top, warp, cur); // warp root = height 7
    __syncthreads();

    // ---- heights 8..10 across the 8 warps (first lanes of warp 0; tiny) ----
    if (warp == 0)
    {
        // s_top slots: height 7: [0,8), 8: [8,12), 9: [12,14), 10: [14,15)
        int src = 0, dstb = kEmitWarps;
#pragma unroll
        for (int h = kWarpLevels + 1; h <= kChunkLevels; h++)
        {
            const uint32_t n = kChunk >> h;
            if ((uint32_t)h <= L && lane < n)
            {
                const uint32_t first_leaf = chunk * kChunk + (lane << h);
                if (first_leaf < T)
                {
                    Box b = unstage_box(s_top, src + 2 * lane);
                    if (first_leaf + (1u << (h - 1)) < T) b = box_merge(b, unstage_box(s_top, src + 2 * lane + 1));
                    stage_box(s_top, dstb + lane, b);
                    store_box(nodes, lv.off[L - h] + (chunk * kChunk >> h) + lane, b);
                }
            }
            src = dstb;
            dstb += n;
            __syncwarp();
        }
    }

    // ---- levels above the chunk roots: hierarchical completion, 32 nodes (5 levels) per stage ----
    // Every group of 32 consecutive completed nodes has an arrival counter; the warp whose arrival completes a
    // group reduces it with shuffles, which completes one node 5 levels up, and so on to the root. No CTA waits:
    // warps 1..7 are done, warp 0 pays one fence + one atomic, and the tail is a few warp-level stages.
    if (L <= (uint32_t)kChunkLevels) return; // the single chunk already holds the root
#ifdef OIBVH_EXP_NOTOP
    return;
#endif
    if (has_finisher)
    {
        // warp 0 lane 0 stored the chunk root (height kChunkLevels) in the loop above: its release orders that store
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done_counter + (chunk >> 5)) : "memory");
        EMIT_STAMP(1, tid == 32);
        return;
    }
    __syncthreads(); // every store of this CTA has been issued
    EMIT_STAMP(1, tid == 32);
    if (warp != 0) return;
    EMIT_STAMP(2, lane == 0);
    EMIT_STAMP(3, lane == 0);
    uint32_t level = L - kChunkLevels, pos = chunk;
    uint32_t* ctr = done_counter;
    while (level > 0)
    {
        const uint32_t n = lv.cnt[level];
        const uint32_t group = pos >> 5, first = group << 5;
        const uint32_t members = min(32u, n - first);
        uint32_t last = 0;
        if (lane == 0)
        {
            __threadfence(); // cumulative over the CTA barrier / the previous stage's stores
            const uint32_t prev = atomicAdd(ctr + group, 1u);
            last = (prev == members - 1);
            if (last)
            {
                ctr[group] = 0; // re-arm for the next launch on this tree
                __threadfence();
            }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        EMIT_STAMP(3, lane == 0);
        if (!last) return;
        Box cur = Box{0, 0, 0, 0, 0, 0};
        if (lane < members) cur = load_box_cg(nodes, lv.off[level] + first + lane);
        uint32_t l = level;
#pragma unroll
        for (int s5 = 0; s5 < 5; s5++)
        {
            if (l == 0) break; // warp-uniform
            const Box right = shfl_down_box(cur, 1 << s5);
            const uint32_t child = (first >> s5) + (lane >> s5); // position at level l of this lane's node
            const bool owner = (lane & ((2u << s5) - 1)) == 0;
            if (owner && child < lv.cnt[l])
            {
                if (child + 1 < lv.cnt[l]) cur = box_merge(cur, right);
                store_box(nodes, lv.off[l - 1] + (child >> 1), cur);
            }
            l--;
        }
        ctr += (n + 31) >> 5; // counters of the next stage follow this stage's
        pos = group;
        level = l;
    }
}

// =================================================================================================
// Device-side Mesh::transform: p = M * (p, 1) with glm's mat4*vec4 order (m0*x + m1*y) + (m2*z + m3*w)
// (third/glm/detail/type_mat4x4.inl:561-572), never contracted. Replaces transform_vec4_kernel
// (src/cuda/transform.cu:35-40) and its H2D/D2H round trip.
// =================================================================================================
__global__ void __launch_bounds__(256) transform_kernel(float4* __restrict__ pos4, uint32_t V, Mat4 M)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float4 p = pos4[i];
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const float add0 = __fadd_rn(__fmul_rn(M.m[0 + k], p.x), __fmul_rn(M.m[4 + k], p.y));
        const float add1 = __fadd_rn(__fmul_rn(M.m[8 + k], p.z), __fmul_rn(M.m[12 + k], 1.0f));
        r[k] = __fadd_rn(add0, add1);
    }
    pos4[i] = make_float4(r[0], r[1], r[2], 1.0f);
}

// =================================================================================================
// Small trees (T <= kSmallTreeMax): keys, stable sort, face gather, leaves and every inner level by ONE CTA; a launch
// takes a table of trees (one CTA each), so a many-body scene of thousands of small objects builds or refits in a
// single launch instead of three (one) launches per object. Same results as the large-tree path: the sort key is
// (Morton key, input index), i.e. thrust::stable_sort_by_key order (src/cuda/oibvhTree.cu:295-296); a parent with a
// dropped right child copies its left child.
// =================================================================================================
constexpr int kSmallThreads = 128;
constexpr int kSmallSmemLevel = 512; // levels of at most this many nodes are kept in shared memory (6 x 2 KB)
constexpr int kSmallBitonicMax = 2048; // static-shared-memory variant (bitonic network); must hold 6 x kSmallSmemLevel floats

__device__ __forceinline__ Box smem_box(const float* s_box, uint32_t i)
{
    Box b;
    b.lx = s_box[i];
    b.ly = s_box[kSmallSmemLevel + i];
    b.lz = s_box[2 * kSmallSmemLevel + i];
    b.hx = s_box[3 * kSmallSmemLevel + i];
    b.hy = s_box[4 * kSmallSmemLevel + i];
    b.hz = s_box[5 * kSmallSmemLevel + i];
    return b;
}

// Stable sort for 512 < T <= 4096 inside one CTA: LSD radix, four 8-bit digits. The keys stay where they were
// computed (key[input index]); what moves is the 16-bit input index, ping-pong between two shared arrays (48 KB in
// all, four CTAs per SM). Each warp owns a contiguous quarter of the sequence and ranks it 32 entries per step with
// shared-memory peer masks (the ranking of onesweep_pass_kernel); ranks are parked in shared memory, a 256-entry scan
// turns the per-warp digit counts into bases, and the indices move to the other array. About 110 instructions per key
// for the whole sort, against ~550 for a bitonic network at T = 4096.
struct SmallRadixSmem
{
    uint32_t key[kSmallTreeMax];    // by input index
    uint16_t id[kSmallTreeMax];     // current order (result)
    uint16_t rank[kSmallTreeMax];   // \ after the sort these two arrays (16 KB, contiguous)
    uint16_t id_alt[kSmallTreeMax]; // / are reused for the shared tree levels
    uint2 tab[kSmallThreads / 32][256]; // (.x running count -> slot base, .y peer mask)
    uint32_t scan[kSmallThreads / 32];
};
static_assert(2 * sizeof(uint16_t) * kSmallTreeMax >= sizeof(float) * 6 * kSmallSmemLevel, "room for the shared levels");

__device__ __forceinline__ void small_radix_sort(SmallRadixSmem& sm, uint32_t T)
{
    constexpr int WARPS = kSmallThreads / 32;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint32_t chunk = (((T + WARPS - 1) / WARPS) + 31) & ~31u; // entries per warp, whole steps of 32
    const uint32_t lane_bit = 1u << lane;
    uint2* my_tab = sm.tab[warp];
    for (uint32_t i = tid; i < T; i += kSmallThreads) sm.id[i] = (uint16_t)i;
    for (int pass = 0; pass < 4; pass++)
    {
        const uint16_t* src = (pass & 1) ? sm.id_alt : sm.id;
        uint16_t* dst = (pass & 1) ? sm.id : sm.id_alt;
        const uint32_t shift = 8 * pass;
        for (uint32_t i = tid; i < WARPS * 256; i += kSmallThreads) (&sm.tab[0][0])[i] = make_uint2(0u, 0u);
        __syncthreads();
        for (uint32_t s = 0; s < chunk; s += 32)
        {
            const uint32_t i = warp * chunk + s + lane;
            const bool valid = i < T;
            const uint32_t dgt = valid ? (sm.key[src[i]] >> shift) & 255u : 0u;
            if (valid) atomicOr(&my_tab[dgt].y, lane_bit);
            __syncwarp();
            uint2 e = make_uint2(0u, 0u);
            if (valid) e = my_tab[dgt]; // (count before this step, peers of this step)
            const uint32_t lower = __popc(e.y & lanemask_lt());
            if (valid) sm.rank[i] = (uint16_t)(e.x + lower);
            __syncwarp();
            if (valid && lower == 0) my_tab[dgt] = make_uint2(e.x + __popc(e.y), 0u); // lowest lane closes the group
            __syncwarp();
        }
        __syncthreads();
        // digit bases: exclusive over (digit, warp); two digits per thread
        {
            static_assert(kSmallThreads * 2 == 256, "two digits per thread");
            uint32_t tot[2];
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                const uint32_t dg = 2 * tid + q;
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < WARPS; w++)
                {
                    const uint32_t c = sm.tab[w][dg].x;
                    sm.tab[w][dg].x = run;
                    run += c;
                }
                tot[q] = run;
            }
            const uint32_t pair = tot[0] + tot[1];
            const uint32_t base = block_exclusive_scan<kSmallThreads>(pair, sm.scan);
#pragma unroll
            for (int w = 0; w < WARPS; w++)
            {
                sm.tab[w][2 * tid].x += base;
                sm.tab[w][2 * tid + 1].x += base + tot[0];
            }
        }
        __syncthreads();
        for (uint32_t s = 0; s < chunk; s += 32)
        {
            const uint32_t i = warp * chunk + s + lane;
            if (i < T)
            {
                const uint16_t id = src[i];
                dst[my_tab[(sm.key[id] >> shift) & 255u].x + sm.rank[i]] = id;
            }
        }
        __syncthreads();
    }
    // four passes: the result is back in sm.id
}

// RADIX selects the sort of a build: false = bitonic network in static shared memory (T <= kSmallBitonicMax; the
// host sends trees of up to kSmallBitonicSplit triangles here), true = small_radix_sort in 80 KB of dynamic shared
// memory (T <= kSmallTreeMax). Refits ignore it.
template <bool BUILD, bool RADIX>
__global__ void __launch_bounds__(kSmallThreads) small_tree_kernel(const SmallTreeDesc* __restrict__ descs)
{
    extern __shared__ __align__(16) unsigned char small_dyn[];
    __shared__ unsigned long long s_kv_static[(BUILD && !RADIX) ? kSmallBitonicMax : 1];
    __shared__ float s_box_refit[BUILD ? 1 : 6 * kSmallSmemLevel];
    SmallRadixSmem& rs = *reinterpret_cast<SmallRadixSmem*>(small_dyn);
    unsigned long long* s_kv = s_kv_static;
    // a build is done with the sort buffers by the time the levels start: the shared levels reuse them
    float* s_box = !BUILD ? s_box_refit : (RADIX ? reinterpret_cast<float*>(rs.rank) : reinterpret_cast<float*>(s_kv_static));
    static_assert(sizeof(unsigned long long) * kSmallBitonicMax >= sizeof(float) * 6 * kSmallSmemLevel,
                  "the bitonic buffer is reused for the shared levels");
    const SmallTreeDesc d = descs[blockIdx.x];
    const uint32_t T = d.T, L = d.L, tid = threadIdx.x;
    float2* nodes = reinterpret_cast<float2*>(d.nodes);
    const uint32_t leaf_off = level_offset(T, L, L);
    if (BUILD)
    {
        const uint32_t P = RADIX ? T : 1u << L; // bitonic: padded to a power of two; padding sorts last
#pragma unroll 4
        for (uint32_t i = tid; i < P; i += kSmallThreads) // independent iterations: several gathers in flight
        {
            unsigned long long kv = ~0ull;
            if (i < T)
            {
                const uint4 f = d.faces_in[i];
                const uint32_t key = morton_of_box(box_of(d.pos[f.x], d.pos[f.y], d.pos[f.z]), d.mesh);
                kv = ((unsigned long long)key << 32) | i;
                if (RADIX) rs.key[i] = key;
            }
            if (!RADIX) s_kv[i] = kv;
        }
        __syncthreads();
        if (RADIX)
            small_radix_sort(rs, T);
        else
        {
            // bitonic network on the 64-bit composites (all distinct, so the result is the stable order)
            for (uint32_t k = 2; k <= P; k <<= 1)
                for (uint32_t j = k >> 1; j > 0; j >>= 1)
                {
                    for (uint32_t t = tid; t < (P >> 1); t += kSmallThreads)
                    {
                        const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const uint32_t hi = lo | j;
                        const unsigned long long a = s_kv[lo], b = s_kv[hi];
                        if ((a > b) == ((lo & k) == 0))
                        {
                            s_kv[lo] = b;
                            s_kv[hi] = a;
                        }
                    }
                    __syncthreads();
                }
        }
#pragma unroll 4
        for (uint32_t i = tid; i < T; i += kSmallThreads)
        {
            uint32_t id, key;
            if (RADIX)
            {
                id = rs.id[i];
                key = rs.key[id];
            }
            else
            {
                const unsigned long long kv = s_kv[i];
                id = (uint32_t)kv;
                key = (uint32_t)(kv >> 32);
            }
            d.keys[i] = key;
            d.perm[i] = id;
            const uint4 f = d.faces_in[id];
            d.faces[3 * i] = f.x;
            d.faces[3 * i + 1] = f.y;
            d.faces[3 * i + 2] = f.z;
            store_box(nodes, leaf_off + i, box_of(d.pos[f.x], d.pos[f.y], d.pos[f.z]));
        }
    }
    else
    {
#pragma unroll 4
        for (uint32_t i = tid; i < T; i += kSmallThreads) // independent iterations: several gathers in flight
        {
            const uint32_t a = d.faces[3 * i], b = d.faces[3 * i + 1], c = d.faces[3 * i + 2];
            store_box(nodes, leaf_off + i, box_of(d.pos[a], d.pos[b], d.pos[c]));
        }
    }
    // inner levels, bottom-up. Wide levels read their children back through L2 (written by this CTA); from the
    // first level of at most kSmallSmemLevel nodes on, the level just produced also stays in shared memory, so the
    // ~10 top levels cost two CTA barriers each instead of an L2 round trip.
    uint32_t child_off = leaf_off, child_cnt = T;
    bool child_in_smem = false;
    for (int l = (int)L - 1; l >= 0; l--)
    {
        const uint32_t cnt = level_count(T, L, (uint32_t)l), off = level_offset(T, L, (uint32_t)l);
        __syncthreads(); // the children are complete
        if (cnt > (uint32_t)kSmallSmemLevel)
        {
#pragma unroll 2
            for (uint32_t p = tid; p < cnt; p += kSmallThreads)
            {
                Box b = load_box_cg(nodes, child_off + 2 * p);
                if (2 * p + 1 < child_cnt) b = box_merge(b, load_box_cg(nodes, child_off + 2 * p + 1));
                store_box(nodes, off + p, b);
            }
        }
        else
        {
            constexpr int PER = kSmallSmemLevel / kSmallThreads;
            Box mine[PER];
#pragma unroll
            for (int k = 0; k < PER; k++)
            {
                const uint32_t p = tid + k * kSmallThreads;
                if (p < cnt)
                {
                    if (child_in_smem)
                    {
                        mine[k] = smem_box(s_box, 2 * p);
                        if (2 * p + 1 < child_cnt) mine[k] = box_merge(mine[k], smem_box(s_box, 2 * p + 1));
                    }
                    else
                    {
                        mine[k] = load_box_cg(nodes, child_off + 2 * p);
                        if (2 * p + 1 < child_cnt)
                            mine[k] = box_merge(mine[k], load_box_cg(nodes, child_off + 2 * p + 1));
                    }
                }
            }
            __syncthreads(); // every child has been read: the parents may overwrite the shared level
#pragma unroll
            for (int k = 0; k < PER; k++)
            {
                const uint32_t p = tid + k * kSmallThreads;
                if (p < cnt)
                {
                    store_box(nodes, off + p, mine[k]);
                    float* o = s_box + p;
                    o[0] = mine[k].lx;
                    o[kSmallSmemLevel] = mine[k].ly;
                    o[2 * kSmallSmemLevel] = mine[k].lz;
                    o[3 * kSmallSmemLevel] = mine[k].hx;
                    o[4 * kSmallSmemLevel] = mine[k].hy;
                    o[5 * kSmallSmemLevel] = mine[k].hz;
                }
            }
            child_in_smem = true;
        }
        child_off = off;
        child_cnt = cnt;
    }
}

// Rigid transforms of many trees in one launch: block -> tree through the block map that follows the n descriptors.
__global__ void __launch_bounds__(256) transform_many_kernel(const XformDesc* __restrict__ descs, uint32_t n,
                                                             const float* __restrict__ mats)
{
    const uint32_t* block_tree = reinterpret_cast<const uint32_t*>(descs + n);
    const uint32_t lo = block_tree[blockIdx.x];
    const XformDesc d = descs[lo];
    const uint32_t i = (blockIdx.x - d.block0) * 256 + threadIdx.x;
    if (i >= d.V) return;
    const float* M = mats + 16ull * lo;
    const float4 p = d.pos[i];
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        const float add0 = __fadd_rn(__fmul_rn(M[0 + k], p.x), __fmul_rn(M[4 + k], p.y));
        const float add1 = __fadd_rn(__fmul_rn(M[8 + k], p.z), __fmul_rn(M[12 + k], 1.0f));
        r[k] = __fadd_rn(add0, add1);
    }
    d.pos[i] = make_float4(r[0], r[1], r[2], 1.0f);
}

// =================================================================================================
// Launchers
// =================================================================================================
cudaError_t launch_pack_positions(const float* xyz, float4* pos4, uint32_t V, cudaStream_t s)
{
    pack_pos_kernel<<<(V + 255) / 256, 256, 0, s>>>(xyz, pos4, V);
    return cudaGetLastError();
}
cudaError_t launch_unpack_positions(const float4* pos4, float* xyz, uint32_t V, cudaStream_t s)
{
    unpack_pos_kernel<<<(V + 255) / 256, 256, 0, s>>>(pos4, xyz, V);
    return cudaGetLastError();
}
cudaError_t launch_pack_faces(const uint32_t* f3, uint4* faces4, uint32_t T, cudaStream_t s)
{
    pack_faces_kernel<<<(T + 255) / 256, 256, 0, s>>>(f3, faces4, T);
    return cudaGetLastError();
}

cudaError_t launch_morton_hist(const uint4* faces4, const float4* pos4, uint32_t T, const MeshAabb& mesh,
                               uint32_t* keys, uint32_t* hist, cudaStream_t s)
{
    const uint32_t tile = 256 * kMortonFacesPerThread;
    uint32_t blocks = (T + tile - 1) / tile;
    const uint32_t cap = kNumSMsB200 * OIBVH_MORTON_CTAS_PER_SM; // persistent: one wave of resident CTAs
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    if (hist)
        morton_hist_kernel<kRadixBits, kRadixPasses, true><<<blocks, 256, 0, s>>>(faces4, pos4, T, mesh, keys, hist);
    else
        morton_hist_kernel<kRadixBits, kRadixPasses, false><<<blocks, 256, 0, s>>>(faces4, pos4, T, mesh, keys, hist);
    return cudaGetLastError();
}

uint32_t onesweep_tiles(uint32_t T) { return (T + kSortTile - 1) / kSortTile; }

cudaError_t launch_onesweep_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                                 uint32_t* vals_out, uint32_t T, int pass, const uint32_t* hist, uint32_t* status,
                                 uint32_t* ticket, cudaStream_t s)
{
    const uint32_t tiles = onesweep_tiles(T);
    onesweep_pass_kernel<kRadixBits, kSortItemsPerThread><<<tiles, 1 << kRadixBits, 0, s>>>(
        keys_in, vals_in, keys_out, vals_out, T, (uint32_t)(pass * kRadixBits), hist + (size_t)pass * (1 << kRadixBits),
        status + (size_t)pass * tiles * (1 << kRadixBits), ticket + pass);
    return cudaGetLastError();
}

cudaError_t tree_emit_configure()
{
    cudaError_t e = cudaFuncSetAttribute(tree_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kEmitSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tree_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)kEmitSmemBytes);
}

size_t emit_counter_words(uint32_t T)
{
    size_t n = (T + kChunk - 1) / kChunk, words = 0;
    while (n > 1)
    {
        n = (n + 31) / 32;
        words += n;
    }
    return words + 32;
}

cudaError_t launch_tree_emit(bool build, const uint4* faces_in4, const uint32_t* perm, uint32_t* faces_sorted,
                             const float4* pos4, float* nodes, uint32_t T, uint32_t* done_counter, uint32_t* status,
                             cudaStream_t s)
{
    const uint32_t L = ceil_log2_u32(T);
    uint32_t chunks = (T + kChunk - 1) / kChunk;
    if (L > (uint32_t)kChunkLevels && chunks <= (uint32_t)OIBVH_EMIT_FINISHER_MAX_CHUNKS)
        chunks += 1; // block 0 = the finisher CTA (levels above the chunk roots); the kernel sees the larger grid
    LevelTable lv;
    for (uint32_t l = 0; l < 32; l++)
    {
        lv.off[l] = l <= L ? level_offset(T, L, l) : 0u;
        lv.cnt[l] = l <= L ? level_count(T, L, l) : 0u;
    }
    if (build)
        tree_emit_kernel<true><<<chunks, kEmitThreads, kEmitSmemBytes, s>>>(
            faces_in4, perm, faces_sorted, pos4, reinterpret_cast<float2*>(nodes), T, L, lv, done_counter, status);
    else
        tree_emit_kernel<false><<<chunks, kEmitThreads, kEmitSmemBytes, s>>>(
            nullptr, nullptr, faces_sorted, pos4, reinterpret_cast<float2*>(nodes), T, L, lv, done_counter, status);
    return cudaGetLastError();
}

#ifdef OIBVH_PROFILE
extern "C" int oibvh_debug_sort_profile(unsigned long long* out /* 4*2*8 */)
{
    return (int)cudaMemcpyFromSymbol(out, g_sort_prof, sizeof(g_sort_prof));
}
#endif

cudaError_t launch_transform(float4* pos4, uint32_t V, const Mat4& M, cudaStream_t s)
{
    if (V == 0) return cudaSuccess;
    transform_kernel<<<(V + 255) / 256, 256, 0, s>>>(pos4, V, M);
    return cudaGetLastError();
}

cudaError_t small_trees_configure()
{
    return cudaFuncSetAttribute(small_tree_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(SmallRadixSmem));
}

// descs[0, n_bitonic) have T <= kSmallBitonicSplit (bitonic build), descs[n_bitonic, n) the rest (radix build)
cudaError_t launch_small_trees(bool build, const SmallTreeDesc* descs, uint32_t n_bitonic, uint32_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    if (!build)
    {
        small_tree_kernel<false, false><<<n, kSmallThreads, 0, s>>>(descs);
        return cudaGetLastError();
    }
    if (n_bitonic) small_tree_kernel<true, false><<<n_bitonic, kSmallThreads, 0, s>>>(descs);
    if (n > n_bitonic)
        small_tree_kernel<true, true><<<n - n_bitonic, kSmallThreads, sizeof(SmallRadixSmem), s>>>(descs + n_bitonic);
    return cudaGetLastError();
}

cudaError_t launch_transform_many(const XformDesc* descs, uint32_t n, uint32_t total_blocks, const float* mats,
                                  cudaStream_t s)
{
    if (n == 0 || total_blocks == 0) return cudaSuccess;
    transform_many_kernel<<<total_blocks, 256, 0, s>>>(descs, n, mats);
    return cudaGetLastError();
}
} // namespace oibvh
