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

#include <algorithm>

namespace oibvh
{

constexpr int kCoopThreads = 512;
constexpr int kCoopWarps = kCoopThreads / 32;
constexpr int kCoopIptSingle = 8;  // one tree over the whole grid
constexpr int kCoopIptMulti = 16;  // several trees side by side: fewer CTAs each, longer chunks
constexpr int kCoopRadix = 1 << kRadixBits; // 256
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

template <int IPT_MAX>
struct CoopSmem
{
    uint2 rank_tab[kCoopWarps][kCoopRadix]; // ranking: (.x running count, .y peer mask); afterwards .x = slot base
    uint2 kv[kCoopThreads * IPT_MAX];       // reorder phase: (key, value) by slot
    uint32_t global_base[kCoopRadix];
    uint32_t scan[16];
};

// One sort job inside a launch: a contiguous range of CTAs works on one (key, value) array. Several trees are
// sorted by ONE launch (CTA ranges side by side, shared grid barriers), which costs the barriers of one sort.
struct SortSeg
{
    uint32_t *keys_a, *keys_b, *vals_a, *vals_b;
    uint32_t* ctl; // this job's totals / counts block
    uint32_t T, ipt, cta0, ncta;
};
constexpr int kMaxSortSegs = 4;
struct SortSegs
{
    SortSeg s[kMaxSortSegs];
    uint32_t n;
};

template <int IPT_MAX>
__global__ void __launch_bounds__(kCoopThreads, 2) coop_sort_kernel(const SortSegs segs, uint32_t* sync_ctl)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoopSmem<IPT_MAX>& sm = *reinterpret_cast<CoopSmem<IPT_MAX>*>(smem_raw);

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    uint32_t si = 0;
#pragma unroll
    for (int i = 1; i < kMaxSortSegs; i++)
        if ((uint32_t)i < segs.n && blockIdx.x >= segs.s[i].cta0) si = i;
    const SortSeg& seg = segs.s[si];
    uint32_t* const keys_a = seg.keys_a;
    uint32_t* const keys_b = seg.keys_b;
    uint32_t* const vals_a = seg.vals_a;
    uint32_t* const vals_b = seg.vals_b;
    uint32_t* const ctl = seg.ctl;
    const uint32_t T = seg.T, ipt = seg.ipt;
    const uint32_t cta = blockIdx.x - seg.cta0, G = seg.ncta;
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
        uint32_t key[IPT_MAX];
        uint16_t rank[IPT_MAX];
#pragma unroll
        for (int j = 0; j < IPT_MAX; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            const bool valid = (uint32_t)j < ipt && i < T;
            key[j] = valid ? __ldcg(kin + i) : 0xffffffffu;
        }
        for (int i = tid; i < kCoopWarps * kCoopRadix; i += kCoopThreads) (&sm.rank_tab[0][0])[i] = make_uint2(0u, 0u);
        __syncthreads();

        COOP_STAMP(1);
        // ---- stable in-warp ranking with shared-memory peer masks (see onesweep_pass_kernel) ----
        uint2* my_tab = sm.rank_tab[warp];
        const uint32_t lane_bit = 1u << lane;
#pragma unroll
        for (int j = 0; j < IPT_MAX; j++)
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
        // values are fetched only now (registers stay free during ranking); first pass: the value is the face id
#pragma unroll
        for (int j = 0; j < IPT_MAX; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            if ((uint32_t)j < ipt && i < T)
            {
                const uint32_t val = vin ? __ldcg(vin + i) : i;
                const uint32_t d = (key[j] >> shift) & MASK;
                sm.kv[my_tab[d].x + rank[j]] = make_uint2(key[j], val);
            }
        }
        COOP_STAMP(3);
        grid_sync(ctl, ++gen, ctl + 1, G);
        COOP_STAMP(4);

        // ---- one warp scans each digit row (exclusive prefix over CTAs) and records the row total ----
        {
            // digit rows are dealt to the warps of this job's CTAs (a job with few CTAs takes several per warp)
            for (uint32_t r = cta + warp * G; r < (uint32_t)kCoopRadix; r += G * kCoopWarps)
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
        grid_sync(ctl, ++gen, ctl + 1, G);
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
        for (int k = 0; k < IPT_MAX; k++)
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
        if (pass + 1 < kRadixPasses) grid_sync(ctl, ++gen, ctl + 1, G);
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
    cudaError_t e = cudaFuncSetAttribute(coop_sort_kernel<kCoopIptSingle>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(CoopSmem<kCoopIptSingle>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(coop_sort_kernel<kCoopIptMulti>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CoopSmem<kCoopIptMulti>));
    if (e != cudaSuccess) return e;
    int per_sm = 0, per_sm_multi = 0, sms = 0, dev = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_sort_kernel<kCoopIptSingle>, kCoopThreads,
                                                      sizeof(CoopSmem<kCoopIptSingle>));
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_multi, coop_sort_kernel<kCoopIptMulti>, kCoopThreads,
                                                      sizeof(CoopSmem<kCoopIptMulti>));
    if (e != cudaSuccess) return e;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (per_sm < 1 || per_sm_multi < 1) return cudaErrorLaunchOutOfResources;
    int grid = sms * (std::min(per_sm, per_sm_multi) >= 2 ? 2 : 1);
    if (grid > 32 * kCoopRowSeg) grid = 32 * kCoopRowSeg; // row scan covers at most 320 CTAs
    g_coop_sort_grid = grid;
    return cudaSuccess;
}

uint32_t coop_sort_capacity() { return (uint32_t)g_coop_sort_grid * kCoopThreads * kCoopIptSingle; }
uint32_t coop_sort_capacity_multi() { return (uint32_t)g_coop_sort_grid * kCoopThreads * kCoopIptMulti; }
size_t coop_sort_ctl_words() { return (size_t)kCoopCtlMat + (size_t)kCoopRadix * 32 * kCoopRowSeg; }

// Sort n <= 4 arrays in one cooperative launch. CTAs are dealt in proportion to the sizes. Returns
// cudaErrorInvalidValue when the arrays do not fit one wave (the caller then sorts them one by one / streams).
cudaError_t launch_coop_sort_many(uint32_t n, uint32_t* const* keys_a, uint32_t* const* keys_b, uint32_t* const* vals_a,
                                  uint32_t* const* vals_b, const uint32_t* T, uint32_t* const* ctl, cudaStream_t s)
{
    if (n == 0 || n > (uint32_t)kMaxSortSegs) return cudaErrorInvalidValue;
    const uint32_t G = (uint32_t)g_coop_sort_grid;
    SortSegs segs;
    segs.n = n;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += T[i];
    uint32_t next = 0, max_ipt = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        uint32_t ncta = (i + 1 == n) ? G - next : (uint32_t)std::max<uint64_t>(1, (uint64_t)G * T[i] / total);
        if (next + ncta > G || ncta == 0) return cudaErrorInvalidValue;
        uint32_t ipt = (T[i] + ncta * kCoopThreads - 1) / (ncta * kCoopThreads);
        if (ipt == 0) ipt = 1;
        max_ipt = std::max(max_ipt, ipt);
        segs.s[i] = SortSeg{keys_a[i], keys_b[i], vals_a[i], vals_b[i], ctl[i], T[i], ipt, next, ncta};
        next += ncta;
    }
    uint32_t* sync_ctl = ctl[0];
    void* args[] = {&segs, &sync_ctl};
    if (max_ipt <= (uint32_t)kCoopIptSingle)
        return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(coop_sort_kernel<kCoopIptSingle>), dim3(G),
                                           dim3(kCoopThreads), args, sizeof(CoopSmem<kCoopIptSingle>), s);
    if (max_ipt <= (uint32_t)kCoopIptMulti)
        return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(coop_sort_kernel<kCoopIptMulti>), dim3(G),
                                           dim3(kCoopThreads), args, sizeof(CoopSmem<kCoopIptMulti>), s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_coop_sort(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t T,
                             uint32_t* ctl, cudaStream_t s)
{
    return launch_coop_sort_many(1, &keys_a, &keys_b, &vals_a, &vals_b, &T, &ctl, s);
}

} // namespace oibvh
