// EXPERIMENTAL -- off by default (OIBVH_SORT_MSD=1 selects it in oibvh_tree_build / oibvh_tree_build_many for
// single-wave sizes). Written against the executable model tools/sort_model.py and the thread-level transcription
// tools/msd_kernel_emulation.py; it has been compiled for sm_100a but NOT yet run on a GPU: the
// shipped path is the cooperative 4-pass LSD sort in sort_coop.cu.
//
// Stable sort of (30-bit Morton key, face id) with ONE global data movement instead of four:
//   1. msd_fine_hist_kernel   histogram of the top 16 key bits (65 536 bins)
//   2. msd_plan_kernel        bins -> ranges of at most kMsdCap keys that are contiguous in the output: a bin opens a
//                             range when it starts in another kMsdWindow-key output window than the non-empty bin
//                             before it, or when it or that bin is heavy (> kMsdCap - kMsdWindow keys); writes the
//                             bin -> range table, the range starts, and a fallback flag (a bin above kMsdCap keys, or
//                             more than kMsdMaxRanges ranges) that makes the sort kernel take its built-in 4-pass
//                             LSD path instead (msd_lsd4_fallback: the algorithm of sort_coop.cu), so no host
//                             decision is needed and the whole build stays graph-capturable
//   3. msd_sort_kernel        cooperative: stable partition of the input by range id (per-CTA ranking, counts matrix,
//                             row scan between two grid barriers, as in sort_coop.cu but the "digit" is the range id),
//                             then every range is sorted stably by its full key inside one CTA's shared memory
//                             (four 8-bit passes over 16-bit local indices, the keys stay put) and written in place.
// Stability: the partition keeps input order inside a range (chunks ordered by CTA, keys inside a chunk ranked in
// (warp, item, lane) order), the range-local LSD passes are stable, and ranges are ordered by key: the result is the
// stable order of thrust::stable_sort_by_key (src/cuda/oibvhTree.cu:295-296).
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace oibvh
{

constexpr int kMsdThreads = 512;
constexpr int kMsdWarps = kMsdThreads / 32;
constexpr int kMsdIpt = 16;
constexpr int kMsdCap = kMsdThreads * kMsdIpt; // 8192 keys: one CTA's chunk of the partition, and the largest range
constexpr int kMsdWindow = 4096;
constexpr int kMsdFineBits = 16;
constexpr int kMsdFineBins = 1 << kMsdFineBits;
constexpr int kMsdFineShift = 30 - kMsdFineBits;
constexpr int kMsdMaxRanges = 1024;
constexpr int kMsdRangeBits = 10;
constexpr int kMsdRowSeg = 10;      // row scan: entries per lane -> grids up to 320 CTAs
constexpr int kMsdPlanThreads = 1024;
constexpr int kMsdValBits = 21;     // input positions are packed with the range id in one word: T < 2^21

// control block (uint32 words)
constexpr size_t kMsdCtlBarrier = 0, kMsdCtlFail = 1, kMsdCtlRanges = 2, kMsdCtlFallback = 3;
constexpr size_t kMsdCtlRangeStart = 64;                                       // kMsdMaxRanges + 1 words
constexpr size_t kMsdCtlRangeOfBin = kMsdCtlRangeStart + kMsdMaxRanges + 64;   // 65 536 x uint16
constexpr size_t kMsdCtlHist = kMsdCtlRangeOfBin + kMsdFineBins / 2;           // 65 536 words
constexpr size_t kMsdCtlMat = kMsdCtlHist + kMsdFineBins;                      // kMsdMaxRanges x grid words

// ---------------------------------------------------------------------------------------------------------------
// 1. fine histogram
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msd_fine_hist_kernel(const uint32_t* __restrict__ keys, uint32_t T,
                                                            uint32_t* __restrict__ hist)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x)
        atomicAdd(hist + (__ldg(keys + i) >> kMsdFineShift), 1u);
}

// ---------------------------------------------------------------------------------------------------------------
// block-wide exclusive scan (THREADS a multiple of 32, <= 1024); s_warp holds >= 33 words; returns the exclusive
// prefix of v over the thread index and the block total
// ---------------------------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ uint32_t msd_block_scan(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        const uint32_t w = lane < (uint32_t)(THREADS / 32) ? s_warp[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += n;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const uint32_t r = inc - v + s_warp[warp];
    *total = s_warp[32];
    __syncthreads(); // s_warp may be reused by the caller
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// 2. plan: one CTA, thread t owns the 64 consecutive bins [64 t, 64 t + 64)
// ---------------------------------------------------------------------------------------------------------------
struct MsdPlanJobs
{
    uint32_t* ctl[4];
    uint32_t T[4];
};
__global__ void __launch_bounds__(kMsdPlanThreads) msd_plan_kernel(const MsdPlanJobs pj)
{
    uint32_t* __restrict__ ctl = pj.ctl[blockIdx.x];
    const uint32_t T = pj.T[blockIdx.x];
    constexpr int PER = kMsdFineBins / kMsdPlanThreads; // 64
    __shared__ uint32_t s_warp[40];
    __shared__ uint32_t s_last_start[kMsdPlanThreads]; // start of the thread's last non-empty bin
    __shared__ uint32_t s_last_count[kMsdPlanThreads]; // its count (0: the thread has no non-empty bin)
    const uint32_t tid = threadIdx.x;
    const uint32_t* hist = ctl + kMsdCtlHist;
    uint16_t* range_of_bin = reinterpret_cast<uint16_t*>(ctl + kMsdCtlRangeOfBin);
    uint32_t* range_start = ctl + kMsdCtlRangeStart;
    const uint32_t b0 = tid * PER;

    // The thread's 64 counts are read ONCE (sixteen 128-bit loads) and parked in shared memory as 16-bit values,
    // transposed (count k of thread t at [k][t]: conflict-free), clamped to 16 bits (a count that does not fit is far
    // above kMsdCap and raises the fallback through max_bin): the walks below never touch global memory again.
    // (Four walks of 64 scalar global loads each were estimated at ~13 us for the one CTA; keeping the counts in
    // registers spilled 0.5 KB per thread.)
    extern __shared__ __align__(16) unsigned char plan_smem[];
    uint16_t* s_cnt = reinterpret_cast<uint16_t*>(plan_smem); // [PER][kMsdPlanThreads] = 128 KB
    uint32_t sum = 0, mx = 0;
    {
        const uint4* h4 = reinterpret_cast<const uint4*>(hist + b0);
#pragma unroll 4
        for (int q = 0; q < PER / 4; q++)
        {
            const uint4 v = __ldcg(h4 + q);
            sum += v.x + v.y + v.z + v.w;
            mx = max(max(mx, v.x), max(max(v.y, v.z), v.w));
            s_cnt[(4 * q + 0) * kMsdPlanThreads + tid] = (uint16_t)min(v.x, 0xffffu);
            s_cnt[(4 * q + 1) * kMsdPlanThreads + tid] = (uint16_t)min(v.y, 0xffffu);
            s_cnt[(4 * q + 2) * kMsdPlanThreads + tid] = (uint16_t)min(v.z, 0xffffu);
            s_cnt[(4 * q + 3) * kMsdPlanThreads + tid] = (uint16_t)min(v.w, 0xffffu);
        }
    }
    auto count_of = [&](int k) { return (uint32_t)s_cnt[k * kMsdPlanThreads + tid]; }; // own entries only: no barrier needed

    // ---- bin starts: exclusive prefix of the histogram ----
    uint32_t total = 0;
    const uint32_t seg_start = msd_block_scan<kMsdPlanThreads>(sum, s_warp, &total);
    const uint32_t block_max = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31u) == 0) s_warp[tid >> 5] = block_max;
    __syncthreads();
    uint32_t max_bin = 0;
    for (int w = 0; w < kMsdPlanThreads / 32; w++) max_bin = max(max_bin, s_warp[w]);
    __syncthreads();

    // ---- the thread's last non-empty bin (carried to the threads behind it) ----
    {
        uint32_t run = seg_start, ls = 0, lc = 0;
#pragma unroll 8
        for (int k = 0; k < PER; k++)
        {
            const uint32_t c = count_of(k);
            if (c)
            {
                ls = run;
                lc = c;
            }
            run += c;
        }
        s_last_start[tid] = ls;
        s_last_count[tid] = lc;
    }
    __syncthreads();
    // nearest non-empty bin before this thread's segment
    bool have_prev = false, prev_heavy = false;
    uint32_t prev_win = 0;
    for (int t = (int)tid - 1; t >= 0; t--)
        if (s_last_count[t])
        {
            have_prev = true;
            prev_heavy = s_last_count[t] > (uint32_t)(kMsdCap - kMsdWindow);
            prev_win = s_last_start[t] / kMsdWindow;
            break;
        }

    // ---- how many ranges open inside this segment ----
    uint32_t opens = 0;
    {
        uint32_t run = seg_start;
        bool hp = have_prev, ph = prev_heavy;
        uint32_t pw = prev_win;
#pragma unroll 8
        for (int k = 0; k < PER; k++)
        {
            const uint32_t c = count_of(k);
            if (c)
            {
                const bool heavy = c > (uint32_t)(kMsdCap - kMsdWindow);
                const uint32_t win = run / kMsdWindow;
                if (!hp || heavy || ph || win != pw) opens++;
                hp = true;
                ph = heavy;
                pw = win;
            }
            run += c;
        }
    }
    uint32_t n_ranges = 0;
    const uint32_t first_range = msd_block_scan<kMsdPlanThreads>(opens, s_warp, &n_ranges);

    // ---- range ids and range starts ----
    {
        uint32_t run = seg_start, next = first_range; // id of the next range to open
        bool hp = have_prev, ph = prev_heavy;
        uint32_t pw = prev_win;
        uint32_t out2 = 0; // two 16-bit range ids per 32-bit store
#pragma unroll 8
        for (int k = 0; k < PER; k++)
        {
            const uint32_t c = count_of(k);
            uint32_t id = next ? next - 1 : 0; // empty bins: never looked up
            if (c)
            {
                const bool heavy = c > (uint32_t)(kMsdCap - kMsdWindow);
                const uint32_t win = run / kMsdWindow;
                if (!hp || heavy || ph || win != pw)
                {
                    if (next < (uint32_t)kMsdMaxRanges) range_start[next] = run;
                    next++;
                }
                id = next - 1;
                hp = true;
                ph = heavy;
                pw = win;
            }
            id = min(id, (uint32_t)kMsdMaxRanges - 1u);
            if (k & 1)
                reinterpret_cast<uint32_t*>(range_of_bin)[(b0 + k) >> 1] = out2 | (id << 16);
            else
                out2 = id;
            run += c;
        }
    }
    if (tid == 0)
    {
        const bool fallback = max_bin > (uint32_t)kMsdCap || n_ranges > (uint32_t)kMsdMaxRanges || total != T || n_ranges == 0;
        ctl[kMsdCtlRanges] = n_ranges;
        ctl[kMsdCtlFallback] = fallback ? 1u : 0u;
        if (n_ranges <= (uint32_t)kMsdMaxRanges) range_start[n_ranges] = T;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. partition + range-local sorts
// ---------------------------------------------------------------------------------------------------------------
struct MsdSmem
{
    union
    {
        uint2 kv[kMsdCap]; // partition: (key, input position | range id << 21) by slot
        struct
        {
            uint32_t key[kMsdCap];    // range-local sort: keys by position in the range (they stay put)
            uint16_t id[kMsdCap];     // current order: positions in the range
            uint16_t id_alt[kMsdCap];
        } loc;
    };
    union
    {
        uint16_t cnt[kMsdWarps][kMsdMaxRanges]; // partition: per-warp running counts -> warp-exclusive offsets
        uint2 tab[kMsdWarps][256];              // range-local passes: (.x running count -> offset, .y peer mask)
    };
    uint32_t base[kMsdMaxRanges]; // partition: slot base of a range in this CTA, then its global base; passes: digit bases
    uint32_t scan[40];
};

struct MsdJob
{
    uint32_t *keys_a, *keys_b, *vals_a, *vals_b; // input keys in keys_a; result in (keys_a, vals_a)
    uint32_t* ctl;
    uint32_t T;
    uint32_t cta0, ncta; // this job's CTA range inside the launch (several trees side by side, per-job barriers)
};
constexpr int kMsdMaxJobs = 4;
struct MsdJobs
{
    MsdJob j[kMsdMaxJobs];
    uint32_t n;
};

// ---------------------------------------------------------------------------------------------------------------
// Built-in fallback: the stable 4-pass LSD sort of sort_coop.cu (8-bit digits, peer-mask ranking, counts matrix + row
// scan between grid barriers) on this job's CTAs, for inputs whose plan asks for it. Uses the same shared arrays:
// kv = reorder buffer, tab = ranking tables, base[0..255] = global digit bases. Counts matrix: ctl + kMsdCtlMat
// (256 x G words), digit totals: ctl + kMsdCtlRangeStart (256 words). Result in (keys_a, vals_a).
// ---------------------------------------------------------------------------------------------------------------
__device__ void msd_lsd4_fallback(MsdSmem& sm, const MsdJob& job, uint32_t cta, uint32_t G, uint32_t& gen)
{
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint32_t T = job.T;
    uint32_t* const ctl = job.ctl;
    uint32_t* const mat = ctl + kMsdCtlMat;
    uint32_t* const totals = ctl + kMsdCtlRangeStart;
    const uint32_t ipt = (T + G * kMsdThreads - 1) / (G * kMsdThreads);
    const uint32_t chunk = kMsdThreads * ipt;
    const uint32_t cta_base = cta * chunk;
    const uint32_t cta_valid = cta_base < T ? min(chunk, T - cta_base) : 0u;
    const uint32_t warp_base = cta_base + warp * (32 * ipt);
    uint32_t *kin = job.keys_a, *kout = job.keys_b, *vin = nullptr, *vout = job.vals_b;
    for (int pass = 0; pass < 4; pass++)
    {
        const uint32_t shift = pass * 8;
        uint32_t key[kMsdIpt];
        uint16_t rank[kMsdIpt];
#pragma unroll
        for (int j = 0; j < kMsdIpt; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            const bool valid = (uint32_t)j < ipt && i < T;
            key[j] = valid ? __ldcg(kin + i) : 0xffffffffu;
        }
        for (uint32_t i = tid; i < (uint32_t)(kMsdWarps * 256); i += kMsdThreads) (&sm.tab[0][0])[i] = make_uint2(0u, 0u);
        __syncthreads();
        uint2* my_tab = sm.tab[warp];
        const uint32_t lane_bit = 1u << lane;
#pragma unroll
        for (int j = 0; j < kMsdIpt; j++)
        {
            if ((uint32_t)j < ipt) // warp-uniform
            {
                const uint32_t i = warp_base + j * 32 + lane;
                const bool valid = i < T;
                const uint32_t d = (key[j] >> shift) & 255u;
                if (valid) atomicOr(&my_tab[d].y, lane_bit);
                __syncwarp();
                uint2 e = make_uint2(0u, 0u);
                if (valid) e = my_tab[d];
                const uint32_t lower = __popc(e.y & lanemask_lt());
                rank[j] = (uint16_t)(e.x + lower);
                __syncwarp();
                if (valid && lower == 0) my_tab[d] = make_uint2(e.x + __popc(e.y), 0u);
                __syncwarp();
            }
        }
        __syncthreads();
        uint32_t cta_count = 0;
        if (tid < 256)
        {
#pragma unroll
            for (int w = 0; w < kMsdWarps; w++)
            {
                const uint32_t c = sm.tab[w][tid].x;
                sm.tab[w][tid].x = cta_count;
                cta_count += c;
            }
            mat[(size_t)tid * G + cta] = cta_count;
        }
        uint32_t all = 0;
        const uint32_t digit_base = msd_block_scan<kMsdThreads>(cta_count, sm.scan, &all); // threads >= 256 add 0
        if (tid < 256) sm.base[256 + tid] = digit_base; // local slot base of every digit
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kMsdIpt; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            if ((uint32_t)j < ipt && i < T)
            {
                const uint32_t val = vin ? __ldcg(vin + i) : i;
                const uint32_t d = (key[j] >> shift) & 255u;
                sm.kv[sm.base[256 + d] + my_tab[d].x + rank[j]] = make_uint2(key[j], val);
            }
        }
        grid_sync(ctl + kMsdCtlBarrier, ++gen, ctl + kMsdCtlFail, G);
        for (uint32_t r = cta + warp * G; r < 256u; r += G * kMsdWarps)
        {
            uint32_t* row = mat + (size_t)r * G;
            uint32_t v[kMsdRowSeg];
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < kMsdRowSeg; k++)
            {
                const uint32_t c = lane * kMsdRowSeg + k;
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
            for (int k = 0; k < kMsdRowSeg; k++)
            {
                const uint32_t c = lane * kMsdRowSeg + k;
                if (c < G) row[c] = run;
                run += v[k];
            }
            if (lane == 31) totals[r] = inc;
        }
        grid_sync(ctl + kMsdCtlBarrier, ++gen, ctl + kMsdCtlFail, G);
        {
            const uint32_t tot = tid < 256 ? __ldcg(totals + tid) : 0u;
            const uint32_t col = tid < 256 ? __ldcg(mat + (size_t)tid * G + cta) : 0u;
            uint32_t sum_all = 0;
            const uint32_t ex = msd_block_scan<kMsdThreads>(tot, sm.scan, &sum_all);
            if (tid < 256) sm.base[tid] = ex + col - sm.base[256 + tid]; // global base - local slot base
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kMsdIpt; k++)
        {
            const uint32_t s = tid + k * kMsdThreads;
            if (s < cta_valid)
            {
                const uint2 e = sm.kv[s];
                const uint32_t dst = sm.base[(e.x >> shift) & 255u] + s;
                kout[dst] = e.x;
                vout[dst] = e.y;
            }
        }
        grid_sync(ctl + kMsdCtlBarrier, ++gen, ctl + kMsdCtlFail, G);
        uint32_t* nk = kout;
        uint32_t* nv = vout;
        kout = (nk == job.keys_b) ? job.keys_a : job.keys_b;
        vout = (nv == job.vals_b) ? job.vals_a : job.vals_b;
        kin = nk;
        vin = nv;
    }
}

__global__ void __launch_bounds__(kMsdThreads, 2) msd_sort_kernel(const MsdJobs jobs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MsdSmem& sm = *reinterpret_cast<MsdSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    uint32_t ji = 0;
#pragma unroll
    for (int i = 1; i < kMsdMaxJobs; i++)
        if ((uint32_t)i < jobs.n && blockIdx.x >= jobs.j[i].cta0) ji = i;
    const MsdJob& job = jobs.j[ji];
    const uint32_t cta = blockIdx.x - job.cta0, G = job.ncta, T = job.T;
    uint32_t* const ctl = job.ctl;
    if (ctl[kMsdCtlFallback] != 0) // the plan (an earlier launch) asks for the 4-pass sort: uniform over the job's CTAs
    {
        uint32_t gen0 = 0;
        msd_lsd4_fallback(sm, job, cta, G, gen0);
        return;
    }
    const uint32_t n_ranges = ctl[kMsdCtlRanges]; // written by the plan kernel (an earlier launch)
    const uint16_t* __restrict__ range_of_bin = reinterpret_cast<const uint16_t*>(ctl + kMsdCtlRangeOfBin);
    const uint32_t* __restrict__ range_start = ctl + kMsdCtlRangeStart;
    uint32_t* const mat = ctl + kMsdCtlMat;
    uint32_t gen = 0;

    const uint32_t ipt = (T + G * kMsdThreads - 1) / (G * kMsdThreads); // <= kMsdIpt (checked by the host)
    const uint32_t chunk = kMsdThreads * ipt;
    const uint32_t cta_base = cta * chunk;
    const uint32_t cta_valid = cta_base < T ? min(chunk, T - cta_base) : 0u;
    const uint32_t warp_base = cta_base + warp * (32 * ipt);

    // ================= stable partition by range id =================
    uint32_t key[kMsdIpt];
    uint16_t rid[kMsdIpt], rank[kMsdIpt];
#pragma unroll
    for (int j = 0; j < kMsdIpt; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        const bool valid = (uint32_t)j < ipt && i < T;
        key[j] = valid ? __ldg(job.keys_a + i) : 0xffffffffu;
        rid[j] = valid ? range_of_bin[key[j] >> kMsdFineShift] : (uint16_t)0;
    }
    for (uint32_t i = tid; i < (uint32_t)(kMsdWarps * kMsdMaxRanges / 2); i += kMsdThreads)
        reinterpret_cast<uint32_t*>(&sm.cnt[0][0])[i] = 0u;
    __syncthreads();
    // in-warp stable ranking by ballots over the bits of the range id: constant cost, no shared-memory atomics
    uint16_t* my_cnt = sm.cnt[warp];
#pragma unroll
    for (int j = 0; j < kMsdIpt; j++)
    {
        if ((uint32_t)j < ipt) // warp-uniform
        {
            const uint32_t i = warp_base + j * 32 + lane;
            const bool valid = i < T;
            const uint32_t r = rid[j];
            uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int b = 0; b < kMsdRangeBits; b++)
            {
                const bool bit = (r >> b) & 1u;
                const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? bal : ~bal;
            }
            const uint32_t lower = __popc(peers & lanemask_lt());
            const int leader = peers ? __ffs(peers) - 1 : 0;
            uint32_t prev = 0;
            if (valid && (int)lane == leader)
            {
                prev = my_cnt[r];
                my_cnt[r] = (uint16_t)(prev + __popc(peers));
            }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            rank[j] = (uint16_t)(prev + lower);
            __syncwarp();
        }
    }
    __syncthreads();
    // per-range CTA totals -> counts matrix; warp-exclusive offsets; slot bases (two consecutive ranges per thread)
    {
        uint32_t tot[2];
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            const uint32_t r = 2 * tid + q;
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kMsdWarps; w++)
            {
                const uint32_t c = sm.cnt[w][r];
                sm.cnt[w][r] = (uint16_t)run;
                run += c;
            }
            tot[q] = run;
            if (r < n_ranges) mat[(size_t)r * G + cta] = run;
        }
        uint32_t all = 0;
        const uint32_t ex = msd_block_scan<kMsdThreads>(tot[0] + tot[1], sm.scan, &all);
        sm.base[2 * tid] = ex;
        sm.base[2 * tid + 1] = ex + tot[0];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMsdIpt; j++)
    {
        const uint32_t i = warp_base + j * 32 + lane;
        if ((uint32_t)j < ipt && i < T)
        {
            const uint32_t r = rid[j];
            sm.kv[sm.base[r] + my_cnt[r] + rank[j]] = make_uint2(key[j], i | (r << kMsdValBits));
        }
    }
    grid_sync(ctl + kMsdCtlBarrier, ++gen, ctl + kMsdCtlFail, G);
    // one warp scans each range row of the counts matrix (exclusive prefix over CTAs)
    for (uint32_t r = cta + warp * G; r < n_ranges; r += G * kMsdWarps)
    {
        uint32_t* row = mat + (size_t)r * G;
        uint32_t v[kMsdRowSeg];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < kMsdRowSeg; k++)
        {
            const uint32_t c = lane * kMsdRowSeg + k;
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
        for (int k = 0; k < kMsdRowSeg; k++)
        {
            const uint32_t c = lane * kMsdRowSeg + k;
            if (c < G) row[c] = run;
            run += v[k];
        }
    }
    grid_sync(ctl + kMsdCtlBarrier, ++gen, ctl + kMsdCtlFail, G);
    // global base of every range for this CTA's keys: range start + keys of the CTAs before it - local slot base
    {
        uint32_t gb[2];
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            const uint32_t r = 2 * tid + q;
            gb[q] = r < n_ranges ? range_start[r] + __ldcg(mat + (size_t)r * G + cta) - sm.base[r] : 0u;
        }
        __syncthreads();
        sm.base[2 * tid] = gb[0];
        sm.base[2 * tid + 1] = gb[1];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMsdIpt; k++)
    {
        const uint32_t s = tid + k * kMsdThreads;
        if (s < cta_valid)
        {
            const uint2 e = sm.kv[s];
            const uint32_t dst = sm.base[e.y >> kMsdValBits] + s;
            job.keys_b[dst] = e.x;
            job.vals_b[dst] = e.y & ((1u << kMsdValBits) - 1u);
        }
    }
    grid_sync(ctl + kMsdCtlBarrier, ++gen, ctl + kMsdCtlFail, G);

    // ================= range-local stable sorts =================
    const uint32_t lane_bit = 1u << lane;
    for (uint32_t r = cta; r < n_ranges; r += G)
    {
        const uint32_t a = range_start[r];
        const uint32_t n = min(range_start[r + 1] - a, (uint32_t)kMsdCap);
        __syncthreads(); // the previous range (or the partition) is done with the shared arrays
        for (uint32_t e = tid; e < n; e += kMsdThreads)
        {
            sm.loc.key[e] = __ldcg(job.keys_b + a + e);
            sm.loc.id[e] = (uint16_t)e;
        }
        const uint32_t wchunk = (((n + kMsdWarps - 1) / kMsdWarps) + 31u) & ~31u; // entries per warp, whole steps of 32
        uint2* my_tab = sm.tab[warp];
        for (int pass = 0; pass < 4; pass++)
        {
            const uint16_t* src = (pass & 1) ? sm.loc.id_alt : sm.loc.id;
            uint16_t* dst = (pass & 1) ? sm.loc.id : sm.loc.id_alt;
            const uint32_t shift = 8u * pass;
            for (uint32_t i = tid; i < (uint32_t)(kMsdWarps * 256); i += kMsdThreads) (&sm.tab[0][0])[i] = make_uint2(0u, 0u);
            __syncthreads();
            uint16_t rk[kMsdIpt];
#pragma unroll
            for (int s = 0; s < kMsdIpt; s++)
            {
                if ((uint32_t)s * 32u < wchunk) // warp-uniform
                {
                    const uint32_t i = warp * wchunk + s * 32 + lane;
                    const bool valid = i < n && s * 32u + lane < wchunk;
                    const uint32_t dgt = valid ? (sm.loc.key[src[i]] >> shift) & 255u : 0u;
                    if (valid) atomicOr(&my_tab[dgt].y, lane_bit);
                    __syncwarp();
                    uint2 e = make_uint2(0u, 0u);
                    if (valid) e = my_tab[dgt]; // (count before this step, peers of this step)
                    const uint32_t lower = __popc(e.y & lanemask_lt());
                    rk[s] = (uint16_t)(e.x + lower);
                    __syncwarp();
                    if (valid && lower == 0) my_tab[dgt] = make_uint2(e.x + __popc(e.y), 0u); // lowest lane closes the group
                    __syncwarp();
                }
            }
            __syncthreads();
            // digit bases: exclusive over (digit, warp)
            {
                uint32_t tot = 0;
                if (tid < 256)
                {
#pragma unroll
                    for (int w = 0; w < kMsdWarps; w++)
                    {
                        const uint32_t c = sm.tab[w][tid].x;
                        sm.tab[w][tid].x = tot;
                        tot += c;
                    }
                }
                uint32_t all = 0;
                const uint32_t ex = msd_block_scan<kMsdThreads>(tot, sm.scan, &all);
                if (tid < 256) sm.base[tid] = ex;
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < kMsdIpt; s++)
            {
                if ((uint32_t)s * 32u < wchunk)
                {
                    const uint32_t i = warp * wchunk + s * 32 + lane;
                    if (i < n && s * 32u + lane < wchunk)
                    {
                        const uint16_t idx = src[i];
                        const uint32_t dgt = (sm.loc.key[idx] >> shift) & 255u;
                        dst[sm.base[dgt] + my_tab[dgt].x + rk[s]] = idx;
                    }
                }
            }
            __syncthreads();
        }
        // four passes: the order is back in loc.id; keys from shared memory, values from the partition output
        for (uint32_t e = tid; e < n; e += kMsdThreads)
        {
            const uint32_t idx = sm.loc.id[e];
            job.keys_a[a + e] = sm.loc.key[idx];
            job.vals_a[a + e] = __ldcg(job.vals_b + a + idx);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int g_msd_grid = 0;
constexpr size_t kMsdPlanSmemBytes = (size_t)kMsdFineBins * sizeof(uint16_t); // the counts, 16 bits each

cudaError_t msd_sort_configure()
{
    cudaError_t e = cudaFuncSetAttribute(msd_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MsdSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(msd_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsdPlanSmemBytes);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0, dev = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msd_sort_kernel, kMsdThreads, sizeof(MsdSmem));
    if (e != cudaSuccess) return e;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    g_msd_grid = std::min(sms * std::min(per_sm, 2), 32 * kMsdRowSeg);
    return cudaSuccess;
}

uint32_t msd_sort_capacity()
{
    const uint64_t cap = (uint64_t)g_msd_grid * kMsdCap;
    return (uint32_t)std::min<uint64_t>(cap, (1u << kMsdValBits) - 1u);
}
size_t msd_sort_ctl_words() { return kMsdCtlMat + (size_t)kMsdMaxRanges * 32 * kMsdRowSeg; }

// Enqueue, for n <= 4 trees, the histogram (one launch per tree, on `s`) and the plan (one launch, one CTA per tree).
// ctl[i]: msd_sort_ctl_words() words per tree. Capturable: nothing is read back.
cudaError_t launch_msd_plan_many(uint32_t n, const uint32_t* const* keys, const uint32_t* T, uint32_t* const* ctl,
                                 cudaStream_t s)
{
    if (n == 0 || n > (uint32_t)kMsdMaxJobs) return cudaErrorInvalidValue;
    MsdPlanJobs pj{};
    for (uint32_t i = 0; i < n; i++)
    {
        cudaError_t e = cudaMemsetAsync(ctl[i], 0, (kMsdCtlHist + kMsdFineBins) * sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
        msd_fine_hist_kernel<<<148 * 4, 256, 0, s>>>(keys[i], T[i], ctl[i] + kMsdCtlHist);
        pj.ctl[i] = ctl[i];
        pj.T[i] = T[i];
    }
    msd_plan_kernel<<<n, kMsdPlanThreads, kMsdPlanSmemBytes, s>>>(pj);
    return cudaGetLastError();
}

// Sort n <= 4 arrays (keys_a[i], identity values) -> (keys_a[i], vals_a[i]) with the plans in ctl[i], in ONE
// cooperative launch: CTA ranges side by side in proportion to the sizes, per-job grid barriers. Returns
// cudaErrorInvalidValue when an array does not fit its CTA range (the caller then uses the 4-pass path).
// CTA ranges of n jobs in proportion to their sizes; false when a job does not fit its range (chunk of at most kMsdCap
// keys per CTA) or the packed position field
static bool msd_split(uint32_t n, const uint32_t* T, uint32_t* cta0, uint32_t* ncta)
{
    if (g_msd_grid == 0 || n == 0 || n > (uint32_t)kMsdMaxJobs) return false;
    const uint32_t G = (uint32_t)g_msd_grid;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += T[i];
    if (total == 0) return false;
    uint32_t next = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t c = (i + 1 == n) ? G - next : (uint32_t)std::max<uint64_t>(1, (uint64_t)G * T[i] / total);
        if (next + c > G || c == 0) return false;
        if (T[i] == 0 || T[i] > (uint64_t)c * kMsdCap || T[i] >= (1u << kMsdValBits)) return false;
        cta0[i] = next;
        ncta[i] = c;
        next += c;
    }
    return true;
}
bool msd_sort_fits(uint32_t n, const uint32_t* T)
{
    uint32_t a[kMsdMaxJobs], b[kMsdMaxJobs];
    return msd_split(n, T, a, b);
}

cudaError_t launch_msd_sort_many(uint32_t n, uint32_t* const* keys_a, uint32_t* const* keys_b, uint32_t* const* vals_a,
                                 uint32_t* const* vals_b, const uint32_t* T, uint32_t* const* ctl, cudaStream_t s)
{
    uint32_t cta0[kMsdMaxJobs], ncta[kMsdMaxJobs];
    if (!msd_split(n, T, cta0, ncta)) return cudaErrorInvalidValue;
    const uint32_t G = (uint32_t)g_msd_grid;
    MsdJobs jobs{};
    jobs.n = n;
    for (uint32_t i = 0; i < n; i++)
        jobs.j[i] = MsdJob{keys_a[i], keys_b[i], vals_a[i], vals_b[i], ctl[i], T[i], cta0[i], ncta[i]};
    void* args[] = {&jobs};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(msd_sort_kernel), dim3(G), dim3(kMsdThreads), args,
                                       sizeof(MsdSmem), s);
}

} // namespace oibvh
