// Cooperative stable LSD radix sort of (30-bit Morton key, face id) for single-wave sizes: all passes in ONE launch.
//
// Replaces thrust::stable_sort_by_key x2 of the reference (src/cuda/oibvhTree.cu:287-299; same result: ascending by
// key, ties in input order) for every tree whose keys fit one wave of the machine. Larger trees stream through
// onesweep_pass_kernel (tree_kernels.cu).
//
// What shapes this kernel on sm_100a (measured in round 1, B300_MICROARCH.md "Atomics"): shared-memory atomics cost
// ~2 cycles PER LANE, plain LDS/STS one cycle per conflict-free warp access, and a fully scattered global store one
// L1 wavefront per lane. So
//   * ranking uses NO shared-memory atomics: the lanes of a warp that hold the same digit find each other with
//     BITS warp ballots (one per digit bit), the lowest lane of each group bumps a warp-private 16-bit counter with
//     a plain load + store, and the others get the old value by shuffle;
//   * (key, value) travel as ONE 8-byte record, reordered through shared memory so that digit runs leave as
//     contiguous 8-byte stores;
//   * 10-bit digits: three passes cover the 30 key bits (8 grid barriers instead of the 12 of four 8-bit passes).
// Cross-CTA prefix: the input is statically partitioned (CTA c owns chunk c of every pass; when every tile is
// resident at once a decoupled look-back degenerates into a serial chain), counts[digit][cta] -> barrier -> one warp
// scans each digit row -> barrier -> scatter -> barrier.
// Stability: chunks are ordered by cta, keys inside a chunk by (warp, round, lane), and ranks are handed out in
// exactly that order.
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace oibvh
{

constexpr int kLsdThreads = 512;
constexpr int kLsdWarps = kLsdThreads / 32;
constexpr int kLsdRowPitch = 320;  // 16-bit entries per digit row = most CTAs one job can have (multiple of 64)
constexpr int kLsdRowSeg = kLsdRowPitch / 32; // row scan: entries per lane
constexpr int kLsdKeyBits = 30;
constexpr int kLsdKSingle = 8;  // records per thread, one tree over the whole grid
constexpr int kLsdKMulti = 14;  // several trees side by side: fewer CTAs each, longer chunks

// control block of one job (uint32 words), sized for the widest digit (10 bits)
constexpr int kLsdMaxRadix = 1024;
constexpr size_t kLsdCtlBarrier = 0;  // arrival counter of this job's barriers (own 128-byte line)
constexpr size_t kLsdCtlExit = 32;    // exit counter: the last CTA to leave re-arms the block for the next launch
constexpr size_t kLsdCtlTotals = 64;  // [radix] digit totals of the current pass
constexpr size_t kLsdCtlCounts = kLsdCtlTotals + kLsdMaxRadix;                          // u16 [radix][pitch]
constexpr size_t kLsdCtlPrefix = kLsdCtlCounts + (size_t)kLsdMaxRadix * kLsdRowPitch / 2; // u32 [cta][radix]
constexpr size_t kLsdCtlWords = kLsdCtlPrefix + (size_t)kLsdRowPitch * kLsdMaxRadix;

struct LsdJob
{
    const uint32_t* keys_in; // pass 0 reads the keys, value = index
    uint2 *rec_a, *rec_b;    // ping-pong (key, value) records
    uint32_t *keys_out, *vals_out; // the last pass writes the sorted keys and the permutation separately
    uint32_t* ctl;
    uint32_t T, ipt, cta0, ncta;
};
struct LsdJobs
{
    LsdJob j[kMaxLsdJobs];
    uint32_t n;
    uint32_t* status; // the context's device status word: bit 0 = a barrier timed out (read back by the host)
};

#ifdef OIBVH_PROFILE
__device__ unsigned long long g_lsd_prof[4][2][12]; // [pass][first/last cta][stamp]
#define LSD_STAMP(k)                                                                                               \
    do                                                                                                             \
    {                                                                                                              \
        if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                                  \
            g_lsd_prof[pass][blockIdx.x == 0 ? 0 : 1][k] = clock64();                                              \
    } while (0)
extern "C" int oibvh_debug_lsd_profile(unsigned long long* out)
{
    return (int)cudaMemcpyFromSymbol(out, g_lsd_prof, sizeof(g_lsd_prof));
}
// wall-clock (ns) of every CTA at [pass][barrier A/B/C][arrive, leave] + kernel entry in [.][3][0]
__device__ unsigned long long g_lsd_bar[320][4][4][2];
__device__ __forceinline__ unsigned long long lsd_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
#define LSD_BAR(b, w)                                                                                              \
    do                                                                                                             \
    {                                                                                                              \
        if (threadIdx.x == 0 && blockIdx.x < 320) g_lsd_bar[blockIdx.x][pass][b][w] = lsd_gtime();                 \
    } while (0)
extern "C" int oibvh_debug_lsd_barriers(unsigned long long* out)
{
    return (int)cudaMemcpyFromSymbol(out, g_lsd_bar, sizeof(g_lsd_bar));
}
#else
#define LSD_STAMP(k)
#define LSD_BAR(b, w)
#endif

template <int K, int BITS>
struct LsdSmem
{
    static constexpr int RADIX = 1 << BITS;
    uint16_t tab[kLsdWarps][RADIX]; // ranking: running count per (warp, digit); afterwards the warp's exclusive offset
    uint2 kv[kLsdThreads * K];      // records in digit order
    uint32_t gbase[RADIX];          // global position of slot 0 of every digit run, minus the run's first slot
    uint32_t scan[kLsdWarps];
};

// Peer group of this lane: the valid lanes of the warp whose digit equals d. BITS ballots, no shared memory.
// Per bit: the ballot, a sign-extending bit-field extract (all ones iff this lane's bit is set) and one LOP3 folding
// `diff |= ballot ^ own` (~4 instructions per bit; predicated and select-based forms compile to 4-6).
template <int BITS>
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, uint32_t valid_mask)
{
    uint32_t diff = 0;
#pragma unroll
    for (int b = 0; b < BITS; b++)
    {
        uint32_t bal;
        int own; // all ones iff this lane's bit b is set
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     ".reg .b32 t;\n\t"
                     "and.b32 t, %2, %3;\n\t"
                     "setp.ne.b32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                     "bfe.s32 %1, %2, %4, 1;\n\t"
                     "}"
                     : "=r"(bal), "=r"(own)
                     : "r"(d), "r"(1u << b), "r"(b));
        diff |= bal ^ (uint32_t)own; // lanes whose bit b differs from mine
    }
    return valid_mask & ~diff;
}

template <int K, int BITS>
__global__ void __launch_bounds__(kLsdThreads, 2) lsd_sort_kernel(const LsdJobs jobs)
{
    using Smem = LsdSmem<K, BITS>;
    constexpr int RADIX = 1 << BITS;
    constexpr int PAIRS = RADIX / 2; // a thread owns the digits (2t, 2t + 1): two 16-bit counters in one word
    constexpr int PASSES = (kLsdKeyBits + BITS - 1) / BITS;
    constexpr uint32_t MASK = RADIX - 1;
    static_assert(PAIRS <= kLsdThreads, "one thread per digit pair");
    static_assert(kLsdThreads * K < 65536, "slots are 16-bit");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    uint32_t ji = 0;
#pragma unroll
    for (int i = 1; i < kMaxLsdJobs; i++)
        if ((uint32_t)i < jobs.n && blockIdx.x >= jobs.j[i].cta0) ji = i;
    const LsdJob& job = jobs.j[ji];
    uint32_t* const ctl = job.ctl;
    const uint32_t T = job.T, ipt = job.ipt;
    const uint32_t cta = blockIdx.x - job.cta0, G = job.ncta;
    const uint32_t chunk = kLsdThreads * ipt;
    const uint32_t cta_base = cta * chunk;
    const uint32_t cta_valid = cta_base < T ? min(chunk, T - cta_base) : 0u;
    const uint32_t warp_base = cta_base + warp * (32 * ipt);
    const bool full = cta_valid == chunk; // every slot of this CTA holds a key
    uint32_t* const totals = ctl + kLsdCtlTotals;
    uint16_t* const counts = reinterpret_cast<uint16_t*>(ctl + kLsdCtlCounts);
    uint32_t* const prefix = ctl + kLsdCtlPrefix;
    uint32_t gen = 0;

    const uint2* rin = nullptr;
    uint2* rout = job.rec_a;
    for (int pass = 0; pass < PASSES; pass++)
    {
        const uint32_t shift = pass * BITS;
        const bool last = pass + 1 == PASSES;
        LSD_STAMP(0);
        LSD_BAR(3, 0);
        // ---- load this CTA's keys (written by other SMs in the previous pass: through L2) ----
        uint32_t key[K];
#pragma unroll
        for (int j = 0; j < K; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            const bool valid = (uint32_t)j < ipt && i < T;
            key[j] = valid ? (rin ? __ldcg(&rin[i].x) : __ldcg(job.keys_in + i)) : 0xffffffffu;
        }
        {
            uint32_t* t32 = reinterpret_cast<uint32_t*>(&sm.tab[0][0]);
            for (int i = tid; i < kLsdWarps * PAIRS; i += kLsdThreads) t32[i] = 0u;
        }
        __syncthreads();

        LSD_STAMP(1);
        // ---- stable ranking inside the warp's slice: ballots find the peers, the lowest peer bumps the counter ----
        uint16_t* my_tab = sm.tab[warp];
        uint32_t rank2[(K + 1) / 2]; // two 16-bit ranks per register
#pragma unroll
        for (int j = 0; j < (K + 1) / 2; j++) rank2[j] = 0;
        if (full && ipt == (uint32_t)K)
        {
            // every slot of the CTA holds a key (all CTAs but the last of a job): no per-round bounds checks, no valid
            // mask -- straight-line code, ~20 % fewer instructions in the phase that is issue-bound
#pragma unroll
            for (int j = 0; j < K; j++)
            {
                const uint32_t d = (key[j] >> shift) & MASK;
                const uint32_t peers = digit_peers<BITS>(d, 0xffffffffu);
                const uint32_t lower = __popc(peers & lanemask_lt());
                const uint32_t before = my_tab[d];
                __syncwarp();
                if (lower == 0) my_tab[d] = (uint16_t)(before + __popc(peers));
                rank2[j >> 1] |= (before + lower) << (16 * (j & 1));
                __syncwarp();
            }
        }
        else
        {
    #pragma unroll
            for (int j = 0; j < K; j++)
            {
                if ((uint32_t)j < ipt && warp_base + j * 32 < T) // warp-uniform: rounds wholly past the end are skipped
                {
                    const uint32_t d = (key[j] >> shift) & MASK;
                    uint32_t vm = 0xffffffffu;
                    bool valid = true;
                    if (!full)
                    {
                        valid = warp_base + j * 32 + lane < T;
                        vm = __ballot_sync(0xffffffffu, valid);
                    }
                    const uint32_t peers = digit_peers<BITS>(d, vm);
                    const uint32_t lower = __popc(peers & lanemask_lt());
                    // every lane reads the running count of its digit (peers read the same word: broadcast), then the
                    // lowest peer bumps it
                    const uint32_t before = my_tab[d];
                    __syncwarp();
                    if (valid && lower == 0) my_tab[d] = (uint16_t)(before + __popc(peers));
                    rank2[j >> 1] |= (before + lower) << (16 * (j & 1));
                    __syncwarp();
                }
            }
        }
        // values are fetched now (registers were free during ranking) and used after the digit scan below, which
        // hides their L2 round trip; first pass: the value is the face id
        uint32_t val[K];
#pragma unroll
        for (int j = 0; j < K; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            val[j] = i;
            if (rin && (uint32_t)j < ipt && i < T) val[j] = __ldcg(&rin[i].y);
        }
        __syncthreads();

        LSD_STAMP(2);
        // ---- per digit pair: warp-exclusive offsets, CTA counts -> counts[digit][cta]; local run starts ----
        uint32_t cnt_lo = 0, cnt_hi = 0, run_lo = 0, run_hi = 0;
        if (tid < PAIRS)
        {
            uint32_t* t32 = reinterpret_cast<uint32_t*>(&sm.tab[0][0]);
            uint32_t c[kLsdWarps];
#pragma unroll
            for (int w = 0; w < kLsdWarps; w++) c[w] = t32[w * PAIRS + tid]; // independent loads, then a register scan
            uint32_t run = 0; // two running sums, one per half word (a CTA holds < 65536 keys: no carry)
#pragma unroll
            for (int w = 0; w < kLsdWarps; w++)
            {
                t32[w * PAIRS + tid] = run;
                run += c[w];
            }
            cnt_lo = run & 0xffffu;
            cnt_hi = run >> 16;
            counts[(size_t)(2 * tid) * kLsdRowPitch + cta] = (uint16_t)cnt_lo;
            counts[(size_t)(2 * tid + 1) * kLsdRowPitch + cta] = (uint16_t)cnt_hi;
        }
        {
            const uint32_t ex = block_exclusive_scan<kLsdThreads>(cnt_lo + cnt_hi, sm.scan);
            run_lo = ex;
            run_hi = ex + cnt_lo;
            if (tid < PAIRS)
            {
                // fold the first slot of the digit run into the per-warp offsets: slot = tab[warp][d] + rank
                // (one random shared load per key instead of two; these 16 accesses are conflict-free)
                uint32_t* t32 = reinterpret_cast<uint32_t*>(&sm.tab[0][0]);
                const uint32_t add = run_lo | (run_hi << 16);
#pragma unroll
                for (int w = 0; w < kLsdWarps; w++) t32[w * PAIRS + tid] += add;
            }
        }
        __syncthreads();
        // Everything that needs only this CTA's data happens BEFORE the first grid barrier: the reorder into shared
        // memory overlaps with waiting for the slowest CTA.
#pragma unroll
        for (int j = 0; j < K; j++)
        {
            const uint32_t i = warp_base + j * 32 + lane;
            if ((uint32_t)j < ipt && i < T)
            {
                const uint32_t d = (key[j] >> shift) & MASK;
                const uint32_t r = (rank2[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                sm.kv[(uint32_t)my_tab[d] + r] = make_uint2(key[j], val[j]);
            }
        }
        LSD_STAMP(3);
        LSD_BAR(0, 0);
        grid_sync(ctl + kLsdCtlBarrier, ++gen, jobs.status, G);
        LSD_BAR(0, 1);
        LSD_STAMP(4);

        // ---- one warp scans each digit row (exclusive prefix over CTAs) and records the row total ----
        for (uint32_t r = cta + warp * G; r < (uint32_t)RADIX; r += G * kLsdWarps)
        {
            const uint32_t* row32 = reinterpret_cast<const uint32_t*>(counts + (size_t)r * kLsdRowPitch);
            uint32_t v[kLsdRowSeg];
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < kLsdRowSeg / 2; k++)
            {
                const uint32_t w = __ldcg(row32 + lane * (kLsdRowSeg / 2) + k);
                const uint32_t c = lane * kLsdRowSeg + 2 * k;
                v[2 * k] = c < G ? (w & 0xffffu) : 0u;
                v[2 * k + 1] = c + 1 < G ? (w >> 16) : 0u;
                sum += v[2 * k] + v[2 * k + 1];
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
            for (int k = 0; k < kLsdRowSeg; k++)
            {
                const uint32_t c = lane * kLsdRowSeg + k;
                if (c < G) prefix[(size_t)c * RADIX + r] = run; // [cta][digit]: every CTA reads its row coalesced
                run += v[k];
            }
            if (lane == 31) totals[r] = inc;
        }
        LSD_STAMP(5);
        LSD_BAR(1, 0);
        grid_sync(ctl + kLsdCtlBarrier, ++gen, jobs.status, G);
        LSD_BAR(1, 1);
        LSD_STAMP(6);
        if (last && tid == 0)
        {
            // nobody touches the barrier counter any more: the last CTA to get here re-arms the control block
            if (atomicAdd(ctl + kLsdCtlExit, 1u) == G - 1)
            {
                st_relaxed_gpu(ctl + kLsdCtlBarrier, 0u);
                st_relaxed_gpu(ctl + kLsdCtlExit, 0u);
            }
        }

        // ---- global position of every digit run of this CTA: (scan of the row totals) + (row prefix at this CTA) ----
        {
            uint32_t tot_lo = 0, tot_hi = 0, pre_lo = 0, pre_hi = 0;
            if (tid < PAIRS)
            {
                const uint2 t2 = __ldcg(reinterpret_cast<const uint2*>(totals) + tid);
                const uint2 p2 = __ldcg(reinterpret_cast<const uint2*>(prefix + (size_t)cta * RADIX) + tid);
                tot_lo = t2.x; tot_hi = t2.y; pre_lo = p2.x; pre_hi = p2.y;
            }
            const uint32_t ex = block_exclusive_scan<kLsdThreads>(tot_lo + tot_hi, sm.scan);
            if (tid < PAIRS)
            {
                sm.gbase[2 * tid] = ex + pre_lo - run_lo;
                sm.gbase[2 * tid + 1] = ex + tot_lo + pre_hi - run_hi;
            }
        }
        __syncthreads();

        LSD_STAMP(7);
        // ---- write the digit runs: consecutive slots of one run are consecutive in global memory ----
        if (!last)
        {
#pragma unroll
            for (int k = 0; k < K; k++)
            {
                const uint32_t s = tid + k * kLsdThreads;
                if (s < cta_valid)
                {
                    const uint2 e = sm.kv[s];
                    rout[sm.gbase[(e.x >> shift) & MASK] + s] = e;
                }
            }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < K; k++)
            {
                const uint32_t s = tid + k * kLsdThreads;
                if (s < cta_valid)
                {
                    const uint2 e = sm.kv[s];
                    const uint32_t dst = sm.gbase[(e.x >> shift) & MASK] + s;
                    job.keys_out[dst] = e.x;
                    job.vals_out[dst] = e.y;
                }
            }
        }
        LSD_STAMP(8);
        LSD_BAR(2, 0);
        if (!last) grid_sync(ctl + kLsdCtlBarrier, ++gen, jobs.status, G);
        LSD_BAR(2, 1);
        LSD_STAMP(9);
        rin = rout;
        rout = (rout == job.rec_a) ? job.rec_b : job.rec_a;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int g_lsd_grid = 0;
constexpr int kLsdBits = 10; // measured on B200 (two 2^20-key sorts side by side): 3 x 10 bits 76 us, 4 x 8 bits 85 us,
                             // 10 bits ranked with match.any 91 us, the round-1 kernel (4 x 8 bits, atomicOr ranking) 87 us;
                             // starting the second job 3-6 us late (so that the two CTAs of an SM are in different
                             // phases) cost 7 us

template <int K>
static cudaError_t lsd_configure_one(int* per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(lsd_sort_kernel<K, kLsdBits>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(LsdSmem<K, kLsdBits>));
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lsd_sort_kernel<K, kLsdBits>, kLsdThreads,
                                                      sizeof(LsdSmem<K, kLsdBits>));
    *per_sm = std::min(*per_sm, n);
    return e;
}

// per device (function attributes are per device); the grid size is the same on every B200
cudaError_t lsd_sort_configure()
{
    int per_sm = 1 << 20, sms = 0, dev = 0;
    cudaError_t e;
    if ((e = lsd_configure_one<kLsdKSingle>(&per_sm)) != cudaSuccess) return e;
    if ((e = lsd_configure_one<kLsdKMulti>(&per_sm)) != cudaSuccess) return e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    int grid = sms * (per_sm >= 2 ? 2 : 1);
    if (grid > kLsdRowPitch) grid = kLsdRowPitch;
    g_lsd_grid = grid;
    return cudaSuccess;
}

uint32_t lsd_sort_capacity() { return (uint32_t)g_lsd_grid * kLsdThreads * kLsdKSingle; }
uint32_t lsd_sort_capacity_multi() { return (uint32_t)g_lsd_grid * kLsdThreads * kLsdKMulti; }
size_t lsd_sort_ctl_words() { return kLsdCtlWords; }

// Sort n <= kMaxLsdJobs arrays in one cooperative launch: keys[i] (T[i] keys, value = index) -> sorted keys in
// keys[i] (in place) and the permutation in vals[i]. rec[i] = scratch of 2 * T[i] records. CTAs are dealt in
// proportion to the sizes. ctl = n consecutive control blocks of lsd_sort_ctl_words() words, zeroed once when
// allocated (the kernel re-arms them). Returns cudaErrorInvalidValue when the arrays do not fit one wave.
cudaError_t launch_lsd_sort_many(uint32_t n, uint32_t* const* keys, uint32_t* const* vals, uint2* const* rec,
                                 const uint32_t* T, uint32_t* ctl, uint32_t* status, cudaStream_t s)
{
    if (n == 0 || n > (uint32_t)kMaxLsdJobs || g_lsd_grid == 0) return cudaErrorInvalidValue;
    const uint32_t G = (uint32_t)g_lsd_grid;
    LsdJobs jobs;
    jobs.n = n;
    jobs.status = status;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += T[i];
    if (total == 0) return cudaErrorInvalidValue;
    uint32_t next = 0, max_ipt = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        uint32_t ncta = (i + 1 == n) ? G - next : (uint32_t)std::max<uint64_t>(1, (uint64_t)G * T[i] / total);
        if (next + ncta > G || ncta == 0) return cudaErrorInvalidValue;
        uint32_t ipt = (T[i] + ncta * kLsdThreads - 1) / (ncta * kLsdThreads);
        if (ipt == 0) ipt = 1;
        max_ipt = std::max(max_ipt, ipt);
        jobs.j[i] = LsdJob{keys[i], rec[i], rec[i] + T[i], keys[i], vals[i], ctl + (size_t)i * kLsdCtlWords,
                           T[i], ipt, next, ncta};
        next += ncta;
    }
    void* args[] = {&jobs};
    const void* fn = nullptr;
    size_t smem = 0;
    if (max_ipt <= (uint32_t)kLsdKSingle)
    {
        fn = (const void*)lsd_sort_kernel<kLsdKSingle, kLsdBits>;
        smem = sizeof(LsdSmem<kLsdKSingle, kLsdBits>);
    }
    else if (max_ipt <= (uint32_t)kLsdKMulti)
    {
        fn = (const void*)lsd_sort_kernel<kLsdKMulti, kLsdBits>;
        smem = sizeof(LsdSmem<kLsdKMulti, kLsdBits>);
    }
    else
        return cudaErrorInvalidValue;
    return cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kLsdThreads), args, smem, s);
}

} // namespace oibvh
