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
static_assert(kCoopWarps * kCoopRadix <= kCoopTileMax, "peer masks alias the key staging area");

// control block (uint32 words): [0] barrier counter, [1] failure flag, [64, 64+256) digit totals, [512, ...) counts
constexpr int kCoopCtlTotals = 64;
constexpr int kCoopCtlMat = 512;

struct CoopSmem
{
    uint32_t hist[kCoopWarps][kCoopRadix]; // per-warp digit counts, then warp-exclusive offsets
    uint32_t keys[kCoopTileMax];           // ranking phase: per-warp peer masks ; reorder phase: keys by slot
    uint32_t vals[kCoopTileMax];
    uint32_t digit_base[kCoopRadix];
    uint32_t global_base[kCoopRadix];
    uint32_t scan[kCoopWarps];
};

__global__ void __launch_bounds__(kCoopThreads, 2)
    coop_sort_kernel(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t T, uint32_t ipt,
                     uint32_t* ctl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoopSmem& sm = *reinterpret_cast<CoopSmem*>(smem_raw);
    uint32_t(*s_mask)[kCoopRadix] = reinterpret_cast<uint32_t(*)[kCoopRadix]>(sm.keys);

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
        for (int i = tid; i < kCoopWarps * kCoopRadix; i += kCoopThreads)
        {
            (&sm.hist[0][0])[i] = 0;
            (&s_mask[0][0])[i] = 0;
        }
        __syncthreads();

        // ---- stable in-warp ranking with shared-memory peer masks (see onesweep_pass_kernel) ----
        uint32_t* my_hist = sm.hist[warp];
        uint32_t* my_mask = s_mask[warp];
        const uint32_t lane_bit = 1u << lane;
#pragma unroll
        for (int j = 0; j < kCoopIptMax; j++)
        {
            if ((uint32_t)j < ipt) // warp-uniform
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
                if (valid && lower == 0)
                {
                    my_hist[d] = before + __popc(peers);
                    my_mask[d] = 0;
                }
                __syncwarp();
            }
        }
        __syncthreads();

        // ---- per digit: warp-exclusive offsets, CTA count -> counts[digit][cta] ----
        uint32_t cta_count = 0;
        if (tid < kCoopRadix)
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kCoopWarps; w++)
            {
                const uint32_t c = sm.hist[w][tid];
                sm.hist[w][tid] = run;
                run += c;
            }
            cta_count = run;
            mat[(size_t)tid * G + cta] = cta_count;
        }
        grid_sync(ctl, ++gen, ctl + 1);

        // ---- one warp scans each digit row (exclusive prefix over CTAs) and records the row total ----
        if (warp == 0)
        {
            for (uint32_t r = cta; r < (uint32_t)kCoopRadix; r += G)
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
        grid_sync(ctl, ++gen, ctl + 1);

        // ---- global base of every digit for this CTA ----
        uint32_t col = 0, tot = 0;
        if (tid < kCoopRadix)
        {
            col = __ldcg(mat + (size_t)tid * G + cta);
            tot = __ldcg(totals + tid);
        }
        const uint32_t bin_start = block_exclusive_scan<kCoopThreads>(tot, sm.scan);
        const uint32_t digit_base = block_exclusive_scan<kCoopThreads>(cta_count, sm.scan);
        if (tid < kCoopRadix)
        {
            sm.digit_base[tid] = digit_base;
            sm.global_base[tid] = bin_start + col - digit_base;
        }
        __syncthreads();

        // ---- reorder inside the chunk through shared memory, then write digit runs coalesced ----
#pragma unroll
        for (int j = 0; j < kCoopIptMax; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            if ((uint32_t)j < ipt && i < T)
            {
                const uint32_t d = (key[j] >> shift) & MASK;
                const uint32_t slot = sm.digit_base[d] + my_hist[d] + rank[j];
                sm.keys[slot] = key[j];
                sm.vals[slot] = val[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kCoopIptMax; k++)
        {
            const uint32_t s = tid + k * kCoopThreads;
            if (s < cta_valid)
            {
                const uint32_t kk = sm.keys[s];
                const uint32_t dst = sm.global_base[(kk >> shift) & MASK] + s;
                kout[dst] = kk;
                vout[dst] = sm.vals[s];
            }
        }
        if (pass + 1 < kRadixPasses) grid_sync(ctl, ++gen, ctl + 1);
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
