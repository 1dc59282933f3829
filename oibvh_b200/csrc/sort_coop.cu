// Cooperative stable LSD radix sort of (30-bit Morton key, face id) for single-wave sizes: all passes in ONE launch.
//
// Decoupled look-back (tree_kernels.cu) is a streaming algorithm: it is fast when tiles outnumber the resident CTAs
// and a tile's predecessor has usually finished. When the whole input fits one wave (<= ~1.2 M keys on 148 SMs) every
// tile starts at the same instant and the look-back degenerates into a serial chain / quadratic polling. Here the
// input is statically partitioned instead -- CTA c owns the contiguous chunk c of every pass -- and the cross-CTA
// prefix is computed with grid barriers:
//     rank chunk in shared memory -> counts[digit][cta] -> barrier -> one warp scans each digit row -> barrier ->
//     global base = (prefix over digit totals) + (row prefix at this cta) -> reorder in smem -> coalesced stores -> barrier
// Stability: chunks are ordered by cta, keys inside a chunk by (warp, item, lane), and ranks are assigned in
// exactly that order, so equal keys keep their input order (thrust::stable_sort_by_key semantics,
// src/cuda/oibvhTree.cu:295-296).
#include "common.cuh"
#include "kernels.h"

namespace oibvh
{

constexpr int kCoopThreads = 512;
constexpr int kCoopWarps = kCoopThreads / 32;
constexpr int kCoopIptMax = 8;
constexpr int kCoopRadix = 1 << kRadixBits; // 256
constexpr int kCoopTileMax = kCoopThreads * kCoopIptMax; // 4096 keys per CTA
constexpr int kCoopRowSeg = 10;                          // row scan: entries per lane -> grids up to 320 CTAs
// two CTAs of 512 threads per SM measured faster than one of 1024 (shorter block-level phases)

// control block (uint32 words): [0] barrier counter, [1] failure flag, [64, 64+256) digit totals, [512, ...) counts
constexpr int kCoopCtlTotals = 64;
constexpr int kCoopCtlMat = 512;

#ifdef OIBVH_PROFILE
__device__ unsigned long long g_coop_prof[4][2][12]; // [pass][first/last cta][stamp]
#define COOP_STAMP(k)                                                                                              \
    do                                                                                                             \
    {                                                                                                              \
        if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                                  \
            g_coop_prof[pass][blockIdx.x == 0 ? 0 : 1][k] = clock64();                                             \
    } while (0)
extern "C" int oibvh_debug_coop_profile(unsigned long long* out)
{
    return (int)cudaMemcpyFromSymbol(out, g_coop_prof, sizeof(g_coop_prof));
}
#else
#define COOP_STAMP(k)
#endif

struct CoopSmem
{
    uint2 rank_tab[kCoopWarps][kCoopRadix]; // ranking: (.x running count, .y peer mask); afterwards .x = slot base
    uint2 kv[kCoopTileMax];                 // reorder phase: (key, value) by slot
    uint32_t global_base[kCoopRadix];
    uint32_t scan[16];
};

__global__ void __launch_bounds__(kCoopThreads, 2)
    coop_sort_kernel(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t T, uint32_t ipt,
                     uint32_t* ctl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoopSmem& sm = *reinterpret_cast<CoopSmem*>(smem_raw);

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint32_t cta = blockIdx.x, G = gridDim.x;
    const uint32_t chunk = kCoopThreads * ipt;
    const uint32_t cta_base = cta * chunk;
    const uint32_t cta_valid = cta_base < T ? min(chunk, T - cta_base) : 0u;
    const uint32_t warp_base = cta_base + warp * (32 * ipt);
    uint32_t* mat = ctl + kCoopCtlMat;
    uint32_t* totals = ctl + kCoopCtlTotals;
    uint32_t gen = 0;

    uint32_t *kin = keys_a, *kout = keys_b, *vin = nullptr, *vout = vals_b;
    for (int pass = 0; pass < kRadixPasses; pass++)
    {
        const uint32_t shift = pass * kRadixBits;
        constexpr uint32_t MASK = kCoopRadix - 1;
        COOP_STAMP(0);
        // ---- load this CTA's chunk (written by other SMs in the previous pass: through L2) ----
        uint32_t key[kCoopIptMax], val[kCoopIptMax];
        uint16_t rank[kCoopIptMax];
#pragma unroll
        for (int j = 0; j < kCoopIptMax; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            const bool valid = (uint32_t)j < ipt && i < T;
            key[j] = valid ? __ldcg(kin + i) : 0xffffffffu;
            val[j] = (valid && vin) ? __ldcg(vin + i) : i; // first pass: the value is the face id
        }
        for (int i = tid; i < kCoopWarps * kCoopRadix; i += kCoopThreads) (&sm.rank_tab[0][0])[i] = make_uint2(0u, 0u);
        __syncthreads();

        COOP_STAMP(1);
        // ---- stable in-warp ranking with shared-memory peer masks (see onesweep_pass_kernel) ----
        uint2* my_tab = sm.rank_tab[warp];
        const uint32_t lane_bit = 1u << lane;
#pragma unroll
        for (int j = 0; j < kCoopIptMax; j++)
        {
            if ((uint32_t)j < ipt) // warp-uniform
            {
                const uint32_t i = warp_base + j * 32 + lane;
                const bool valid = i < T;
                const uint32_t d = (key[j] >> shift) & MASK;
                if (valid) atomicOr(&my_tab[d].y, lane_bit);
                __syncwarp();
                uint2 e = make_uint2(0u, 0u);
                if (valid) e = my_tab[d]; // (count before this step, peers of this step)
                const uint32_t lower = __popc(e.y & lanemask_lt());
                rank[j] = (uint16_t)(e.x + lower);
                __syncwarp();
                if (valid && lower == 0) my_tab[d] = make_uint2(e.x + __popc(e.y), 0u); // lowest lane closes the group
                __syncwarp();
            }
        }
        __syncthreads();

        COOP_STAMP(2);
        // ---- per digit: warp-exclusive offsets, CTA count -> counts[digit][cta]; local slot bases ----
        // Everything that needs only this CTA's data happens BEFORE the first grid barrier (it overlaps with waiting
        // for the slowest CTA): the exclusive scan of the CTA's digit counts and the reorder into shared memory.
        uint32_t cta_count = 0, digit_base = 0;
        if (tid < kCoopRadix)
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kCoopWarps; w++)
            {
                const uint32_t c = sm.rank_tab[w][tid].x;
                sm.rank_tab[w][tid].x = run;
                run += c;
            }
            cta_count = run;
            mat[(size_t)tid * G + cta] = cta_count;
            uint32_t inc = cta_count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += n;
            }
            if (lane == 31) sm.scan[warp] = inc;
            digit_base = inc - cta_count; // + totals of the lower warps, added below
        }
        __syncthreads();
        if (tid < kCoopRadix)
        {
#pragma unroll
            for (int w = 0; w < kCoopRadix / 32; w++)
                if ((uint32_t)w < warp) digit_base += sm.scan[w];
            // fold the CTA-level digit base into the per-warp offsets: slot = rank_tab[warp][d].x + rank
#pragma unroll
            for (int w = 0; w < kCoopWarps; w++) sm.rank_tab[w][tid].x += digit_base;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kCoopIptMax; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            if ((uint32_t)j < ipt && i < T)
            {
                const uint32_t d = (key[j] >> shift) & MASK;
                sm.kv[my_tab[d].x + rank[j]] = make_uint2(key[j], val[j]);
            }
        }
        COOP_STAMP(3);
        grid_sync(ctl, ++gen, ctl + 1);
        COOP_STAMP(4);

        // ---- one warp scans each digit row (exclusive prefix over CTAs) and records the row total ----
        {
            // rows cta, cta + G, ... : one warp each
            const uint32_t r = cta + warp * G;
            if (r < (uint32_t)kCoopRadix)
            {
                uint32_t* row = mat + (size_t)r * G;
                uint32_t v[kCoopRowSeg];
                uint32_t sum = 0;
#pragma unroll
                for (int k = 0; k < kCoopRowSeg; k++)
                {
                    const uint32_t c = lane * kCoopRowSeg + k;
                    v[k] = c < G ? __ldcg(row + c) : 0u;
                    sum += v[k];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += n;
                }
                uint32_t run = inc - sum;
#pragma unroll
                for (int k = 0; k < kCoopRowSeg; k++)
                {
                    const uint32_t c = lane * kCoopRowSeg + k;
                    if (c < G) row[c] = run;
                    run += v[k];
                }
                if (lane == 31) totals[r] = inc;
            }
        }
        COOP_STAMP(5);
        grid_sync(ctl, ++gen, ctl + 1);
        COOP_STAMP(6);

        // ---- global base of every digit for this CTA: (scan of the row totals) + (row prefix at this CTA) ----
        if (tid < kCoopRadix)
        {
            const uint32_t col = __ldcg(mat + (size_t)tid * G + cta);
            const uint32_t tot = __ldcg(totals + tid);
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += n;
            }
            if (lane == 31) sm.scan[warp] = inc;
            sm.global_base[tid] = inc - tot + col - digit_base; // + totals of the lower warps, added below
        }
        __syncthreads();
        if (tid < kCoopRadix)
        {
            uint32_t base = 0;
#pragma unroll
            for (int w = 0; w < kCoopRadix / 32; w++)
                if ((uint32_t)w < warp) base += sm.scan[w];
            sm.global_base[tid] += base;
        }
        __syncthreads();

        COOP_STAMP(7);
        // ---- write digit runs coalesced ----
#pragma unroll
        for (int k = 0; k < kCoopIptMax; k++)
        {
            const uint32_t s = tid + k * kCoopThreads;
            if (s < cta_valid)
            {
                const uint2 e = sm.kv[s];
                const uint32_t dst = sm.global_base[(e.x >> shift) & MASK] + s;
                kout[dst] = e.x;
                vout[dst] = e.y;
            }
        }
        COOP_STAMP(8);
        if (pass + 1 < kRadixPasses) grid_sync(ctl, ++gen, ctl + 1);
        COOP_STAMP(9);
        // ping-pong
        uint32_t* nk = kout;
        uint32_t* nv = vout;
        kout = (nk == keys_b) ? keys_a : keys_b;
        vout = (nv == vals_b) ? vals_a : vals_b;
        kin = nk;
        vin = nv;
    }
}

static int g_coop_sort_grid = 0;

cudaError_t coop_sort_configure()
{
    cudaError_t e = cudaFuncSetAttribute(coop_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(CoopSmem));
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0, dev = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_sort_kernel, kCoopThreads, sizeof(CoopSmem));
    if (e != cudaSuccess) return e;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    int grid = sms * (per_sm >= 2 ? 2 : 1);
    if (grid > 32 * kCoopRowSeg) grid = 32 * kCoopRowSeg; // row scan covers at most 320 CTAs
    g_coop_sort_grid = grid;
    return cudaSuccess;
}

uint32_t coop_sort_capacity() { return (uint32_t)g_coop_sort_grid * kCoopTileMax; }
size_t coop_sort_ctl_words() { return (size_t)kCoopCtlMat + (size_t)kCoopRadix * 32 * kCoopRowSeg; }

cudaError_t launch_coop_sort(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t T,
                             uint32_t* ctl, cudaStream_t s)
{
    const uint32_t G = (uint32_t)g_coop_sort_grid;
    uint32_t ipt = (T + G * kCoopThreads - 1) / (G * kCoopThreads);
    if (ipt == 0) ipt = 1;
    if (ipt > (uint32_t)kCoopIptMax) return cudaErrorInvalidValue;
    void* args[] = {&keys_a, &keys_b, &vals_a, &vals_b, &T, &ipt, &ctl};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(coop_sort_kernel), dim3(G), dim3(kCoopThreads),
                                       args, sizeof(CoopSmem), s);
}

} // namespace oibvh
