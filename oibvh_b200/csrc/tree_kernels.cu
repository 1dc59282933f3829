// Build / refit kernels of the ostensibly-implicit BVH for sm_100a.
//
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

struct Vec3
{
    float x, y, z;
};

__device__ __forceinline__ Vec3 load_pos(const float* __restrict__ pos, uint32_t v)
{
    const float* p = pos + 3ull * v;
    Vec3 r;
    r.x = __ldg(p);
    r.y = __ldg(p + 1);
    r.z = __ldg(p + 2);
    return r;
}

// glm::min(glm::min(v0, v1), v2) / glm::max(glm::max(v0, v1), v2)   (src/cuda/oibvh.cu:33-39)
__device__ __forceinline__ Box face_box(const float* __restrict__ pos, uint32_t i0, uint32_t i1, uint32_t i2)
{
    const Vec3 a = load_pos(pos, i0), b = load_pos(pos, i1), c = load_pos(pos, i2);
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

// warp-aggregated shared-memory histogram update: one atomic per distinct digit per warp
__device__ __forceinline__ void hist_add(uint32_t* h, uint32_t digit, bool valid)
{
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : 0xffffffffu);
    if (valid && lane_id() == (uint32_t)(__ffs(peers) - 1)) atomicAdd(h + digit, (uint32_t)__popc(peers));
}

template <int RADIX_BITS, int PASSES>
__global__ void __launch_bounds__(256) morton_hist_kernel(const uint32_t* __restrict__ faces,
                                                          const float* __restrict__ pos, uint32_t T, MeshAabb mesh,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ hist)
{
    constexpr int RADIX = 1 << RADIX_BITS;
    __shared__ uint32_t sh[PASSES * RADIX];
    for (int i = threadIdx.x; i < PASSES * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();

    const uint32_t groups = (T + 3) / 4; // 4 faces = 48 B = three 128-bit loads
    const uint4* f4 = reinterpret_cast<const uint4*>(faces);
    // whole warps iterate together so the match in hist_add always sees 32 lanes
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounded = (groups + 31u) & ~31u;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < rounded; g += stride)
    {
        uint32_t idx[12];
        const uint32_t f0 = g * 4;
        const bool full = g < groups && f0 + 4 <= T;
        if (full)
        {
            const uint4 a = ldg_stream_u4(f4 + 3ull * g), b = ldg_stream_u4(f4 + 3ull * g + 1),
                        c = ldg_stream_u4(f4 + 3ull * g + 2);
            idx[0] = a.x; idx[1] = a.y; idx[2] = a.z; idx[3] = a.w;
            idx[4] = b.x; idx[5] = b.y; idx[6] = b.z; idx[7] = b.w;
            idx[8] = c.x; idx[9] = c.y; idx[10] = c.z; idx[11] = c.w;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 12; k++)
            {
                const uint64_t e = 12ull * g + k;
                idx[k] = (g < groups && e < 3ull * T) ? faces[e] : 0u;
            }
        }
        uint32_t key[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const bool valid = g < groups && f0 + k < T;
            key[k] = 0;
            if (valid) key[k] = morton_of_box(face_box(pos, idx[3 * k], idx[3 * k + 1], idx[3 * k + 2]), mesh);
#pragma unroll
            for (int p = 0; p < PASSES; p++)
                hist_add(sh + p * RADIX, (key[k] >> (p * RADIX_BITS)) & (RADIX - 1), valid);
        }
        if (full)
            *reinterpret_cast<uint4*>(keys + f0) = make_uint4(key[0], key[1], key[2], key[3]);
        else
        {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (g < groups && f0 + k < T) keys[f0 + k] = key[k];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PASSES * RADIX; i += blockDim.x)
    {
        const uint32_t c = sh[i];
        if (c) atomicAdd(hist + i, c);
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
    __shared__ uint32_t s_vals[TILE];
    __shared__ uint32_t s_digit_base[RADIX];
    __shared__ uint32_t s_global_base[RADIX];
    __shared__ uint32_t s_scan[WARPS];
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u); // tiles are claimed in launch order: predecessors are resident
#pragma unroll
    for (int w = 0; w < WARPS; w++) s_hist[w][tid] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * TILE;
    const uint32_t tile_valid = min((uint32_t)TILE, T - tile_base);

    // ---- load (warp-striped: lane-consecutive keys are memory-consecutive) + stable in-warp ranking ----
    uint32_t key[IPT];
    uint16_t rank[IPT];
    const uint32_t warp_base = tile_base + warp * (32 * IPT);
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        key[j] = (i < T) ? ldg_stream_u32(keys_in + i) : 0xffffffffu;
    }
    uint32_t* my_hist = s_hist[warp];
#pragma unroll
    for (int j = 0; j < IPT; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        const bool valid = i < T;
        const uint32_t d = (key[j] >> shift) & MASK;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : (uint32_t)RADIX);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader)
        {
            old = my_hist[d];
            my_hist[d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = (uint16_t)(old + __popc(peers & lanemask_lt()));
        __syncwarp();
    }
    __syncthreads();

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
        if (tile == 0)
            st_relaxed_gpu(my_status, kFlagPrefix | tile_count);
        else
            st_relaxed_gpu(my_status, kFlagAggregate | tile_count);

        const uint32_t bin_start = block_exclusive_scan<THREADS>(hist[d], s_scan);
        const uint32_t digit_base = block_exclusive_scan<THREADS>(tile_count, s_scan);

        uint32_t exclusive = 0;
        if (tile != 0)
        {
            int t = (int)tile - 1;
            while (true)
            {
                const uint32_t s = ld_relaxed_gpu(status + (size_t)t * RADIX + d);
                const uint32_t f = s & kFlagMask;
                if (f == 0) continue; // predecessor not published yet (it is resident: ticket order)
                exclusive += s & kCountMask;
                if (f == kFlagPrefix) break;
                t--;
            }
            st_relaxed_gpu(my_status, kFlagPrefix | (exclusive + tile_count));
        }
        s_digit_base[d] = digit_base;
        s_global_base[d] = bin_start + exclusive - digit_base;
    }
    __syncthreads();

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
        if (i < T) s_vals[slot[j]] = vals_in ? ldg_stream_u32(vals_in + i) : i;
    }
    __syncthreads();
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
}

// =================================================================================================
// Leaf boxes + bottom-up reduction, one CHUNK-leaf subtree per CTA.
// =================================================================================================
constexpr int kEmitThreads = 256;
constexpr int kLeavesPerThread = 4;
constexpr int kChunk = kEmitThreads * kLeavesPerThread; // 1024 leaves
constexpr int kChunkLevels = 10;                        // log2(kChunk)
// staging: nodes of local height h (0 = leaves) occupy slots [sm_off(h), sm_off(h) + (kChunk >> h))
__device__ __forceinline__ constexpr int sm_off(int h) { return 2 * kChunk - (2 * kChunk >> h); }
constexpr int kStageNodes = 2 * kChunk - 1;
constexpr size_t kEmitSmemBytes = (size_t)kStageNodes * 24 + 16;

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

// Levels [0, top_level) of the tree, given that level `top_level` is complete in global memory.
// Run by one CTA (the last one to finish its chunk). Uses the staging buffer as two ping-pong arrays.
__device__ void finish_top(float2* __restrict__ nodes, float2* sm, uint32_t T, uint32_t L, uint32_t top_level)
{
    uint32_t l = top_level;
    // wide levels: straight through L2
    while (l > 0 && level_count(T, L, l) > (uint32_t)kChunk)
    {
        const uint32_t cnt_c = level_count(T, L, l), cnt_p = level_count(T, L, l - 1);
        const uint32_t off_c = level_offset(T, L, l), off_p = level_offset(T, L, l - 1);
        for (uint32_t p = threadIdx.x; p < cnt_p; p += blockDim.x)
        {
            Box b = load_box_cg(nodes, off_c + 2 * p);
            if (2 * p + 1 < cnt_c) b = box_merge(b, load_box_cg(nodes, off_c + 2 * p + 1));
            store_box(nodes, off_p + p, b);
        }
        __threadfence_block();
        __syncthreads();
        l--;
    }
    if (l == 0) return;
    // narrow levels: shared memory
    float2* cur = sm;
    float2* nxt = sm + 3 * kChunk;
    uint32_t cnt_c = level_count(T, L, l);
    {
        const float2* src = nodes + 3ull * level_offset(T, L, l);
        for (uint32_t i = threadIdx.x; i < 3 * cnt_c; i += blockDim.x) cur[i] = __ldcg(src + i);
    }
    __syncthreads();
    while (l > 0)
    {
        const uint32_t cnt_p = level_count(T, L, l - 1);
        const uint32_t off_p = level_offset(T, L, l - 1);
        for (uint32_t p = threadIdx.x; p < cnt_p; p += blockDim.x)
        {
            Box b = unstage_box(cur, 2 * p);
            if (2 * p + 1 < cnt_c) b = box_merge(b, unstage_box(cur, 2 * p + 1));
            stage_box(nxt, p, b);
            store_box(nodes, off_p + p, b);
        }
        __syncthreads();
        float2* t = cur; cur = nxt; nxt = t;
        cnt_c = cnt_p;
        l--;
    }
}

template <bool BUILD>
__global__ void __launch_bounds__(kEmitThreads)
    tree_emit_kernel(const uint32_t* __restrict__ faces_in,   // BUILD: input-order faces ; else: sorted faces
                     const uint32_t* __restrict__ perm,       // BUILD: sorted position -> input face id
                     uint32_t* __restrict__ faces_sorted,     // BUILD: output
                     const float* __restrict__ pos, float2* __restrict__ nodes, uint32_t T, uint32_t L,
                     uint32_t* done_counter)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    __shared__ bool s_last;

    const uint32_t tid = threadIdx.x, lane = lane_id();
    const uint32_t chunk = blockIdx.x;
    const uint32_t leaf0 = chunk * kChunk + tid * kLeavesPerThread;

    // ---- faces of this thread's 4 consecutive leaves ----
    uint32_t idx[12];
    const bool full = leaf0 + 4 <= T;
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
            for (int k = 0; k < 4; k++) id[k] = (leaf0 + k < T) ? perm[leaf0 + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const uint32_t* f = faces_in + 3ull * id[k];
            const bool v = leaf0 + k < T;
            idx[3 * k] = v ? __ldg(f) : 0u;
            idx[3 * k + 1] = v ? __ldg(f + 1) : 0u;
            idx[3 * k + 2] = v ? __ldg(f + 2) : 0u;
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
                if (leaf0 + k / 3 < T) faces_sorted[3ull * leaf0 + k] = idx[k];
        }
    }
    else
    {
        if (full)
        {
            const uint4* f4 = reinterpret_cast<const uint4*>(faces_in + 3ull * leaf0);
            const uint4 a = ldg_stream_u4(f4), b = ldg_stream_u4(f4 + 1), c = ldg_stream_u4(f4 + 2);
            idx[0] = a.x; idx[1] = a.y; idx[2] = a.z; idx[3] = a.w;
            idx[4] = b.x; idx[5] = b.y; idx[6] = b.z; idx[7] = b.w;
            idx[8] = c.x; idx[9] = c.y; idx[10] = c.z; idx[11] = c.w;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < 12; k++) idx[k] = (leaf0 + k / 3 < T) ? faces_in[3ull * leaf0 + k] : 0u;
        }
    }

    // ---- heights 0..2 in registers. A node at height h, position p exists iff p * 2^h < T. ----
    Box leaf[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        leaf[k] = Box{0, 0, 0, 0, 0, 0};
        if (leaf0 + k < T)
        {
            leaf[k] = face_box(pos, idx[3 * k], idx[3 * k + 1], idx[3 * k + 2]);
            stage_box(sm, sm_off(0) + tid * 4 + k, leaf[k]);
        }
    }
    Box h1[2];
    h1[0] = (leaf0 + 1 < T) ? box_merge(leaf[0], leaf[1]) : leaf[0];
    h1[1] = (leaf0 + 3 < T) ? box_merge(leaf[2], leaf[3]) : leaf[2];
    if (leaf0 < T) stage_box(sm, sm_off(1) + tid * 2, h1[0]);
    if (leaf0 + 2 < T) stage_box(sm, sm_off(1) + tid * 2 + 1, h1[1]);
    Box cur = (leaf0 + 2 < T) ? box_merge(h1[0], h1[1]) : h1[0];
    if (leaf0 < T) stage_box(sm, sm_off(2) + tid, cur);

    // ---- heights 3..7 by warp shuffles: lane with (lane % 2^(h-2)) == 0 owns the height-h node ----
#pragma unroll
    for (int h = 3; h <= 7; h++)
    {
        const int delta = 1 << (h - 3);
        const Box right = shfl_down_box(cur, delta);
        // right child = height h-1 node of thread tid+delta, first leaf = leaf0 + delta*4
        if (leaf0 + (uint32_t)delta * 4 < T) cur = box_merge(cur, right);
        if ((lane & (2 * delta - 1)) == 0 && leaf0 < T) stage_box(sm, sm_off(h) + (tid >> (h - 2)), cur);
    }
    __syncthreads();

    // ---- heights 8..10 across the 8 warps (warp 0 only; tiny) ----
    if (tid < 32)
    {
#pragma unroll
        for (int h = 8; h <= kChunkLevels; h++)
        {
            const uint32_t n = kChunk >> h; // nodes at this height in the chunk
            if (tid < n)
            {
                const uint32_t first_leaf = chunk * kChunk + (tid << h);
                if (first_leaf < T)
                {
                    Box b = unstage_box(sm, sm_off(h - 1) + 2 * tid);
                    if (first_leaf + (1u << (h - 1)) < T) b = box_merge(b, unstage_box(sm, sm_off(h - 1) + 2 * tid + 1));
                    stage_box(sm, sm_off(h) + tid, b);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- coalesced stores: the chunk's nodes of one height are one contiguous slice of that level ----
    const uint32_t hmax = min((uint32_t)kChunkLevels, L);
    for (uint32_t h = 0; h <= hmax; h++)
    {
        const uint32_t l = L - h;
        const uint32_t per = kChunk >> h;
        const uint32_t first = chunk * per;
        const uint32_t cnt = level_count(T, L, l);
        if (first >= cnt) break;
        const uint32_t n = min(per, cnt - first);
        float2* dst = nodes + 3ull * (level_offset(T, L, l) + first);
        const float2* src = sm + 3 * sm_off(h);
        for (uint32_t i = tid; i < 3 * n; i += kEmitThreads) dst[i] = src[i];
    }

    // ---- the last CTA to retire finishes levels above the chunk roots ----
    if (L <= (uint32_t)kChunkLevels) return; // the single chunk already holds the root
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
        const uint32_t prev = atomicAdd(done_counter, 1u);
        s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    finish_top(nodes, sm, T, L, L - kChunkLevels);
    if (tid == 0) *done_counter = 0; // re-arm for the next launch on this tree
}

// =================================================================================================
// Device-side Mesh::transform: p = M * (p, 1) with glm's mat4*vec4 order (m0*x + m1*y) + (m2*z + m3*w)
// (third/glm/detail/type_mat4x4.inl:561-572), never contracted. Replaces transform_vec4_kernel
// (src/cuda/transform.cu:35-40) and its H2D/D2H round trip.
// =================================================================================================
__global__ void __launch_bounds__(256) transform_kernel(float* __restrict__ pos, uint32_t V, Mat4 M)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    float* p = pos + 3ull * i;
    const float x = p[0], y = p[1], z = p[2];
#pragma unroll
    for (int r = 0; r < 3; r++)
    {
        const float add0 = __fadd_rn(__fmul_rn(M.m[0 + r], x), __fmul_rn(M.m[4 + r], y));
        const float add1 = __fadd_rn(__fmul_rn(M.m[8 + r], z), __fmul_rn(M.m[12 + r], 1.0f));
        p[r] = __fadd_rn(add0, add1);
    }
}

// =================================================================================================
// Launchers
// =================================================================================================
cudaError_t launch_morton_hist(const uint32_t* faces, const float* pos, uint32_t T, const MeshAabb& mesh,
                               uint32_t* keys, uint32_t* hist, cudaStream_t s)
{
    const uint32_t groups = (T + 3) / 4;
    uint32_t blocks = (groups + 255) / 256;
    const uint32_t cap = kNumSMsB200 * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    morton_hist_kernel<kRadixBits, kRadixPasses><<<blocks, 256, 0, s>>>(faces, pos, T, mesh, keys, hist);
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

cudaError_t launch_tree_emit(bool build, const uint32_t* faces_in, const uint32_t* perm, uint32_t* faces_sorted,
                             const float* pos, float* nodes, uint32_t T, uint32_t* done_counter, cudaStream_t s)
{
    const uint32_t L = ceil_log2_u32(T);
    const uint32_t chunks = (T + kChunk - 1) / kChunk;
    if (build)
        tree_emit_kernel<true><<<chunks, kEmitThreads, kEmitSmemBytes, s>>>(
            faces_in, perm, faces_sorted, pos, reinterpret_cast<float2*>(nodes), T, L, done_counter);
    else
        tree_emit_kernel<false><<<chunks, kEmitThreads, kEmitSmemBytes, s>>>(
            faces_in, nullptr, nullptr, pos, reinterpret_cast<float2*>(nodes), T, L, done_counter);
    return cudaGetLastError();
}

cudaError_t launch_transform(float* pos, uint32_t V, const Mat4& M, cudaStream_t s)
{
    if (V == 0) return cudaSuccess;
    transform_kernel<<<(V + 255) / 256, 256, 0, s>>>(pos, V, M);
    return cudaGetLastError();
}

} // namespace oibvh
