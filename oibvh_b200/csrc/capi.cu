// C ABI of oibvh_b200 (see include/oibvh_b200.h): handle management, device memory, kernel schedules.
// No CPU fallback lives here: every compute entry point enqueues CUDA kernels or fails.
#include "../../include/oibvh_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>
#include <unistd.h>

using namespace oibvh;

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

// which sort a tree gets: the cooperative single-wave LSD sort (sort_lsd.cu) when all its keys fit one wave of the
// machine, the streaming onesweep passes (tree_kernels.cu) above. OIBVH_STREAMING_SORT=1 forces the streaming sort for
// every size (measurement only: DESIGN.md section 3.3 quotes both at T = 2^20).
static bool single_wave_sort(uint32_t T)
{
    static const bool force_streaming = getenv("OIBVH_STREAMING_SORT") != nullptr;
    return !force_streaming && T <= lsd_sort_capacity();
}

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                                 \
    do                                                                                                           \
    {                                                                                                            \
        cudaError_t e__ = (call);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            return fail(e__ == cudaErrorMemoryAllocation ? OIBVH_ERR_NOMEM : OIBVH_ERR_CUDA, "%s:%d %s -> %s",   \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e__));                                     \
    } while (0)

#define REQUIRE(cond, msg)                                                                                       \
    do                                                                                                           \
    {                                                                                                            \
        if (!(cond)) return fail(OIBVH_ERR_INVALID, "%s: %s", __func__, msg);                                    \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------------
struct StageEvent
{
    int stage;
    cudaEvent_t a, b;
};

struct oibvh_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr; // host->device uploads run here and overlap compute on `stream`
    // independent launches of one call (the emit kernels of a multi-tree build) alternate between `stream` and this
    // one, forked / joined with events, so the partially filled last wave of one kernel is topped up by the next
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint64_t launches = 0;
    bool timing = false;
    bool capturing = false;
    uint64_t capture_launches = 0;
    int collide_grid = 0;    // CTAs of the persistent detection kernel (cooperative launch)
    uint32_t dense_seed_level = 8; // two-body scenes: level tested densely in round 0 (OIBVH_DENSE_SEED_LEVEL, 0 = off)
    uint64_t generation = 0; // bumped whenever device buffers referenced by enqueued work are reallocated
    std::vector<StageEvent> events;
    float stage_ms[OIBVH_STAGE_COUNT] = {};
    // device tables of the *_many entry points, kept while the same list of trees is passed again (per-frame calls
    // then upload nothing and can be captured into a graph)
    struct BatchTable
    {
        std::vector<oibvh_tree*> key;
        void* dev = nullptr;
        size_t cap_bytes = 0;
        uint32_t total_blocks = 0;
    };
    BatchTable small_table, xform_table;
    float* d_mats = nullptr;
    size_t d_mats_cap = 0;
    // control blocks of the cooperative sort (kMaxLsdJobs jobs per launch). One set per context: the sorts of a
    // context are ordered on `stream`, and every launch leaves the blocks re-armed.
    uint32_t* lsd_ctl = nullptr;
    // device status word the build kernels OR their bounded-wait failures into (bit 0: sort barrier, bit 1: emit
    // finisher); checked by oibvh_ctx_synchronize and oibvh_tree_download
    uint32_t* d_status = nullptr;
};

struct oibvh_graph
{
    oibvh_ctx* ctx = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;
    uint64_t generation = 0;
};

struct oibvh_tree
{
    oibvh_ctx* ctx = nullptr;
    uint32_t T = 0, V = 0, N = 0, L = 0;
    MeshAabb mesh;
    bool built = false;
    // persistent state
    float4* pos = nullptr;         // V x (x, y, z, 1): one 128-bit load per gathered vertex
    float* pos_stage = nullptr;    // V x 3 packed: landing / take-off buffer for host transfers
    uint4* faces_in = nullptr;     // T x (i0, i1, i2, 0), input order
    uint32_t* faces = nullptr;     // T x 3, Morton order
    float* nodes = nullptr;        // N x 6
    // sort state: (keys_a, vals_a) hold the sorted keys / permutation after a build
    uint32_t *keys_a = nullptr, *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr;
    uint2* sort_rec = nullptr; // 2 T (key, value) records: ping-pong buffers of the cooperative sort
    uint32_t* sort_ctl = nullptr; // [hist: passes*radix][ticket: passes (padded to 64)][status: passes*tiles*radix]
    size_t sort_ctl_words = 0;
    uint32_t* done_counter = nullptr;
    // T <= kSmallTreeMax: built / refitted by one CTA (small_tree_kernel); no second sort buffers, no control blocks
    bool small = false;
    SmallTreeDesc* d_small = nullptr; // this tree's entry for single-tree launches
    // upload pipeline: ev_uploaded = staging buffer filled (copy stream), ev_consumed = staging buffer packed (compute)
    cudaEvent_t ev_uploaded = nullptr, ev_consumed = nullptr;
    bool consumed_recorded = false;
    bool upload_pending = false; // staging buffer holds positions that have not been packed yet
    uint64_t epoch = 0;          // bumped by everything that changes what a detection would read (positions, nodes)
    uint64_t build_serial = 0;   // bumped by every build: a recorded BVTT cut is tied to the face order it was made on
};

struct oibvh_scene
{
    oibvh_ctx* ctx = nullptr;
    std::vector<oibvh_tree*> trees;
    ObjDesc* d_objs = nullptr;
    size_t d_objs_cap = 0;
    bool objs_dirty = true;
    uint4* queue = nullptr; // BVTT work queue: every slot holds the empty marker (all bits set) between detections
    uint4* pairs = nullptr;
    uint32_t front_cap = 0, cand_cap = 0, pair_cap = 0;
    // Device counter block (CTR_WORDS). Once the work queues exist it is the HEAD of the pair-list allocation:
    // [counters: CTR_WORDS x 4 bytes = 32 records][pairs: pair_cap records], so a multi-GPU caller moves the counts
    // and the pair list with ONE collective (oibvh_scene_device_counters). counters_own is the stand-alone block a
    // scene starts with (and falls back to while the queues are being re-allocated).
    uint32_t* counters = nullptr;
    uint32_t* counters_own = nullptr;
    uint4* pair_block = nullptr;
    uint32_t* h_counters = nullptr; // pinned host mirror
    uint32_t rank = 0, world = 1;
    uint32_t last_entry = 0, last_expand = 0, last_rounds = 0; // parameters of the last enqueued detection
    // multi-GPU (oibvh_mgpu_*): protocol block of this scene + the gathering rank's buffers as seen from this device
    uint32_t* mg_state = nullptr; // MG_WORDS, persistent (never zeroed per frame)
    int mg_mode = 0;              // 0 = single GPU, 1 = gathering rank, 2 = remote rank
    uint32_t* mg_root_state = nullptr;
    uint32_t* mg_root_counters = nullptr;
    uint4* mg_root_pairs = nullptr;
    uint32_t mg_root_pair_cap = 0;
    void *mg_ipc_block = nullptr, *mg_ipc_state = nullptr; // mappings opened with cudaIpcOpenMemHandle
    bool mg_opened_ahead = false;
    // the trees' modification counters when the last detection was enqueued (an overflow re-run is only valid if
    // nothing has touched them since)
    std::vector<uint64_t> enq_epochs;
    // opt-in extensions
    bool self_collision = false; // also test every object against itself
    bool coherent = false;       // temporal coherence: detections start from a recorded BVTT cut while it is valid
    uint32_t cut_depth = 6;      // the cut lies this many levels above the leaves
    uint4* cut = nullptr;        // cand_cap records
    uint4* cand = nullptr;       // candidate list, cand_cap records
    bool cut_valid = false;
    uint32_t last_mode = 0;      // mode of the last enqueued detection (0 plain, 1 recording, 2 from the cut)
    std::vector<uint64_t> cut_sig; // what the cut was recorded on: scene set-up + the build serial of every tree
};

namespace
{

struct DeviceGuard
{
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

void count_launch(oibvh_ctx* c, uint64_t n = 1)
{
    if (c->capturing)
        c->capture_launches += n;
    else
        c->launches += n;
}

struct StageScope
{
    oibvh_ctx* c;
    int idx = -1;
    StageScope(oibvh_ctx* ctx, int stage) : c(ctx)
    {
        if (!c->timing || c->capturing) return;
        StageEvent ev{stage, nullptr, nullptr};
        if (cudaEventCreate(&ev.a) != cudaSuccess || cudaEventCreate(&ev.b) != cudaSuccess) return;
        cudaEventRecord(ev.a, c->stream);
        c->events.push_back(ev);
        idx = (int)c->events.size() - 1;
    }
    ~StageScope()
    {
        if (idx >= 0) cudaEventRecord(c->events[idx].b, c->stream);
    }
};

template <typename Tp>
int dev_alloc(Tp** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    CU(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(Tp)));
    return OIBVH_OK;
}

void tree_free(oibvh_tree* t)
{
    cudaFree(t->pos);
    cudaFree(t->pos_stage);
    cudaFree(t->faces_in);
    cudaFree(t->faces);
    cudaFree(t->nodes);
    cudaFree(t->keys_a);
    cudaFree(t->keys_b);
    cudaFree(t->vals_a);
    cudaFree(t->vals_b);
    cudaFree(t->sort_rec);
    cudaFree(t->d_small);
    cudaFree(t->sort_ctl);
    cudaFree(t->done_counter);
    if (t->ev_uploaded) cudaEventDestroy(t->ev_uploaded);
    if (t->ev_consumed) cudaEventDestroy(t->ev_consumed);
}

// Positions uploaded by oibvh_tree_set_positions land in the staging buffer on the copy stream; the compute stream
// picks them up (wait + widen to 16-byte records) lazily, right before the first operation that reads them, so the
// upload overlaps everything enqueued in between.
int tree_flush_upload(oibvh_tree* t)
{
    if (!t->upload_pending) return OIBVH_OK;
    oibvh_ctx* ctx = t->ctx;
    CU(cudaStreamWaitEvent(ctx->stream, t->ev_uploaded, 0));
    CU(launch_pack_positions(t->pos_stage, t->pos, t->V, ctx->stream));
    count_launch(ctx);
    CU(cudaEventRecord(t->ev_consumed, ctx->stream));
    t->consumed_recorded = true;
    t->upload_pending = false;
    return OIBVH_OK;
}

SmallTreeDesc small_desc_of(const oibvh_tree* t)
{
    SmallTreeDesc h;
    h.faces_in = t->faces_in;
    h.pos = t->pos;
    h.faces = t->faces;
    h.nodes = t->nodes;
    h.keys = t->keys_a;
    h.perm = t->vals_a;
    h.mesh = t->mesh;
    h.T = t->T;
    h.L = t->L;
    return h;
}

int tree_alloc(oibvh_ctx* ctx, uint32_t V, uint32_t T, const float mesh_aabb[6], oibvh_tree** out)
{
    oibvh_tree* t = new (std::nothrow) oibvh_tree;
    if (!t) return fail(OIBVH_ERR_NOMEM, "host allocation failed");
    t->ctx = ctx;
    t->T = T;
    t->V = V;
    t->L = ceil_log2_u32(T);
    t->N = tree_size(T);
    memcpy(t->mesh.v, mesh_aabb, sizeof(float) * 6);
    const uint32_t tiles = onesweep_tiles(T);
    const size_t radix = (size_t)1 << kRadixBits;
    // control block of the streaming sort (trees beyond the single-wave capacity): digit histograms, tickets, tile status
    t->sort_ctl_words = single_wave_sort(T) ? 64 : kRadixPasses * radix + 64 + (size_t)kRadixPasses * tiles * radix;
    t->small = T <= kSmallTreeMax;
    int rc = OIBVH_OK;
    // round the index buffers up to whole 16-byte groups so that 128-bit accesses of the last group stay in bounds
    const size_t T4 = ((size_t)T + 3) / 4 * 4;
    if ((rc = dev_alloc(&t->pos, (size_t)V)) || (rc = dev_alloc(&t->pos_stage, (size_t)V * 3)) ||
        (rc = dev_alloc(&t->faces_in, T4)) ||
        (rc = dev_alloc(&t->faces, T4 * 3)) || (rc = dev_alloc(&t->nodes, (size_t)t->N * 6)) ||
        (rc = dev_alloc(&t->keys_a, T4)) || (rc = dev_alloc(&t->vals_a, T4)) ||
        (t->small ? (rc = dev_alloc(&t->d_small, 1))
                  : ((single_wave_sort(T)
                          ? (rc = dev_alloc(&t->sort_rec, 2 * T4))                                       // single-wave sort
                          : ((rc = dev_alloc(&t->keys_b, T4)) || (rc = dev_alloc(&t->vals_b, T4)))) ||   // streaming sort
                     (rc = dev_alloc(&t->sort_ctl, t->sort_ctl_words)) ||
                     (rc = dev_alloc(&t->done_counter, emit_counter_words(T))))))
    {
        tree_free(t);
        delete t;
        return rc;
    }
    cudaError_t e = cudaSuccess;
    if (t->small)
    {
        const SmallTreeDesc h = small_desc_of(t);
        e = cudaMemcpyAsync(t->d_small, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream); // pageable: staged on return
    }
    else
        e = cudaMemsetAsync(t->done_counter, 0, sizeof(uint32_t) * emit_counter_words(T), ctx->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_uploaded, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_consumed, cudaEventDisableTiming);
    if (e != cudaSuccess)
    {
        tree_free(t);
        delete t;
        return fail(OIBVH_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
    }
    *out = t;
    return OIBVH_OK;
}

void scene_free_buffers(oibvh_scene* s)
{
    cudaFree(s->queue);
    cudaFree(s->pair_block);
    cudaFree(s->cut);
    cudaFree(s->cand);
    s->cut = s->cand = nullptr;
    s->cut_valid = false;
    s->queue = s->pairs = s->pair_block = nullptr;
    s->counters = s->counters_own;
}

int scene_alloc_buffers(oibvh_scene* s, uint32_t front_cap, uint32_t cand_cap, uint32_t pair_cap)
{
    scene_free_buffers(s);
    s->ctx->generation++;
    int rc;
    static_assert(CTR_WORDS * sizeof(uint32_t) % sizeof(uint4) == 0, "the counter block is a whole number of records");
    constexpr size_t kCtrRecords = CTR_WORDS * sizeof(uint32_t) / sizeof(uint4);
    // (cand_cap sizes the candidate list and, for a scene with temporal coherence, the BVTT cut)
    if ((rc = dev_alloc(&s->queue, front_cap)) || (rc = dev_alloc(&s->pair_block, kCtrRecords + (size_t)pair_cap)) ||
        (rc = dev_alloc(&s->cand, cand_cap)) ||
        (s->coherent && (rc = dev_alloc(&s->cut, cand_cap))))
        return rc;
    {
        cudaError_t e = cudaMemsetAsync(s->queue, 0xff, sizeof(uint4) * (size_t)front_cap, s->ctx->stream);
        // (the candidate list uses the same empty marker)
        if (e == cudaSuccess) e = cudaMemsetAsync(s->cand, 0xff, sizeof(uint4) * (size_t)cand_cap, s->ctx->stream);
        if (e != cudaSuccess) return fail(OIBVH_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
    }
    s->counters = reinterpret_cast<uint32_t*>(s->pair_block);
    s->pairs = s->pair_block + kCtrRecords;
    {
        cudaError_t e = cudaMemsetAsync(s->counters, 0, sizeof(uint32_t) * CTR_WORDS, s->ctx->stream);
        if (e != cudaSuccess) return fail(OIBVH_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
    }
    s->front_cap = front_cap;
    s->cand_cap = cand_cap;
    s->pair_cap = pair_cap;
    return OIBVH_OK;
}

uint32_t grow_to(uint32_t cap, uint64_t needed)
{
    uint64_t n = std::max<uint64_t>((uint64_t)cap * 2, needed + needed / 2 + 1024);
    n = std::min<uint64_t>(n, 0xfffffff0ull);
    return (uint32_t)n;
}

} // namespace

// the stream has been synchronised: did a build kernel give up a bounded wait since the last check?
static int ctx_check_status(oibvh_ctx* ctx)
{
    uint32_t st = 0;
    CU(cudaMemcpy(&st, ctx->d_status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st == 0) return OIBVH_OK;
    cudaMemset(ctx->d_status, 0, sizeof(st));
    if (ctx->lsd_ctl) cudaMemset(ctx->lsd_ctl, 0, sizeof(uint32_t) * kMaxLsdJobs * lsd_sort_ctl_words());
    return fail(OIBVH_ERR_INTERNAL, "a build kernel timed out waiting for other thread blocks (%s%s): the trees built or "
                                    "refitted since the last check are invalid -- rebuild them",
                (st & 1u) ? "radix sort barrier " : "", (st & 2u) ? "top-of-tree reduction" : "");
}

// ---------------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------------
extern "C" const char* oibvh_last_error(void) { return g_last_error.c_str(); }
extern "C" int oibvh_version(void) { return 100; }
extern "C" int oibvh_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
static int ctx_create_impl(int device, void* stream, bool use_given, oibvh_ctx** out)
{
    REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(OIBVH_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    DeviceGuard g(device);
    if (!g.ok) return fail(OIBVH_ERR_CUDA, "cannot select device %d", device);
    oibvh_ctx* c = new (std::nothrow) oibvh_ctx;
    if (!c) return fail(OIBVH_ERR_NOMEM, "host allocation failed");
    c->device = device;
    if (use_given)
    {
        c->stream = static_cast<cudaStream_t>(stream);
        c->own_stream = false;
    }
    else
    {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess)
        {
            delete c;
            return fail(OIBVH_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = tree_emit_configure();
    if (e == cudaSuccess) e = collide_configure(&c->collide_grid);
    if (const char* env = getenv("OIBVH_DENSE_SEED_LEVEL")) // tuning knob: 0 disables the dense round 0
    {
        const long v = strtol(env, nullptr, 10);
        c->dense_seed_level = v <= 0 ? 1000u : (uint32_t)std::min<long>(std::max<long>(v, 6), 9);
    }
    if (e == cudaSuccess) e = lsd_sort_configure();
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->lsd_ctl), sizeof(uint32_t) * kMaxLsdJobs * lsd_sort_ctl_words());
    if (e == cudaSuccess) e = cudaMemset(c->lsd_ctl, 0, sizeof(uint32_t) * kMaxLsdJobs * lsd_sort_ctl_words());
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->d_status), 64);
    if (e == cudaSuccess) e = cudaMemset(c->d_status, 0, 64);
    if (e == cudaSuccess) e = small_trees_configure();
    if (e != cudaSuccess)
    {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return fail(OIBVH_ERR_CUDA, "kernel configuration failed (is this an sm_100a device?): %s",
                    cudaGetErrorString(e));
    }
    *out = c;
    return OIBVH_OK;
}

extern "C" int oibvh_ctx_create(int device, oibvh_ctx** out) { return ctx_create_impl(device, nullptr, false, out); }
extern "C" int oibvh_ctx_create_on_stream(int device, void* cuda_stream, oibvh_ctx** out)
{
    return ctx_create_impl(device, cuda_stream, true, out);
}

extern "C" int oibvh_ctx_destroy(oibvh_ctx* ctx)
{
    if (!ctx) return OIBVH_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& ev : ctx->events)
    {
        cudaEventDestroy(ev.a);
        cudaEventDestroy(ev.b);
    }
    if (ctx->copy_stream)
    {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->aux_stream)
    {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaFree(ctx->small_table.dev);
    cudaFree(ctx->xform_table.dev);
    cudaFree(ctx->d_mats);
    cudaFree(ctx->lsd_ctl);
    cudaFree(ctx->d_status);
    delete ctx;
    return OIBVH_OK;
}

extern "C" int oibvh_ctx_synchronize(oibvh_ctx* ctx)
{
    REQUIRE(ctx != nullptr, "ctx is NULL");
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return ctx_check_status(ctx);
}

extern "C" int oibvh_ctx_get_stream(oibvh_ctx* ctx, void** cuda_stream)
{
    REQUIRE(ctx && cuda_stream, "NULL argument");
    *cuda_stream = ctx->stream;
    return OIBVH_OK;
}

extern "C" int oibvh_ctx_launch_count(oibvh_ctx* ctx, uint64_t* launches)
{
    REQUIRE(ctx && launches, "NULL argument");
    *launches = ctx->launches;
    return OIBVH_OK;
}

extern "C" int oibvh_ctx_enable_timing(oibvh_ctx* ctx, int enable)
{
    REQUIRE(ctx != nullptr, "ctx is NULL");
    ctx->timing = enable != 0;
    return OIBVH_OK;
}

extern "C" int oibvh_ctx_stage_ms(oibvh_ctx* ctx, float ms[OIBVH_STAGE_COUNT])
{
    REQUIRE(ctx && ms, "NULL argument");
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < OIBVH_STAGE_COUNT; i++) ctx->stage_ms[i] = 0.f;
    for (auto& ev : ctx->events)
    {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev.a, ev.b) == cudaSuccess) ctx->stage_ms[ev.stage] += t;
        cudaEventDestroy(ev.a);
        cudaEventDestroy(ev.b);
    }
    ctx->events.clear();
    memcpy(ms, ctx->stage_ms, sizeof(float) * OIBVH_STAGE_COUNT);
    return OIBVH_OK;
}

// ---- CUDA-graph capture of whatever the caller enqueues between begin/end (a whole frame, typically) ----
extern "C" int oibvh_ctx_capture_begin(oibvh_ctx* ctx)
{
    REQUIRE(ctx != nullptr, "ctx is NULL");
    REQUIRE(!ctx->capturing, "capture already in progress");
    DeviceGuard g(ctx->device);
    CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    ctx->capture_launches = 0;
    return OIBVH_OK;
}

// a call that failed between capture_begin and capture_end leaves the stream in capture mode: abandon the capture
extern "C" int oibvh_ctx_capture_abort(oibvh_ctx* ctx)
{
    REQUIRE(ctx != nullptr, "ctx is NULL");
    if (!ctx->capturing) return OIBVH_OK;
    DeviceGuard g(ctx->device);
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaStreamEndCapture(ctx->stream, &graph); // an invalidated capture returns an error and a null graph
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return OIBVH_OK;
}

extern "C" int oibvh_ctx_capture_end(oibvh_ctx* ctx, oibvh_graph** out)
{
    REQUIRE(ctx && out, "NULL argument");
    REQUIRE(ctx->capturing, "no capture in progress");
    DeviceGuard g(ctx->device);
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamEndCapture(ctx->stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess)
    {
        cudaGraphDestroy(graph);
        return fail(OIBVH_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    oibvh_graph* gr = new (std::nothrow) oibvh_graph;
    if (!gr)
    {
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        return fail(OIBVH_ERR_NOMEM, "host allocation failed");
    }
    gr->ctx = ctx;
    gr->graph = graph;
    gr->exec = exec;
    gr->launches = ctx->capture_launches;
    gr->generation = ctx->generation;
    *out = gr;
    return OIBVH_OK;
}

extern "C" int oibvh_graph_launch(oibvh_graph* graph)
{
    REQUIRE(graph != nullptr, "graph is NULL");
    oibvh_ctx* ctx = graph->ctx;
    if (graph->generation != ctx->generation)
        return fail(OIBVH_ERR_INVALID, "graph is stale: device buffers were reallocated after capture; re-capture");
    DeviceGuard g(ctx->device);
    CU(cudaGraphLaunch(graph->exec, ctx->stream));
    ctx->launches += graph->launches;
    return OIBVH_OK;
}

extern "C" int oibvh_graph_destroy(oibvh_graph* graph)
{
    if (!graph) return OIBVH_OK;
    DeviceGuard g(graph->ctx->device);
    cudaGraphExecDestroy(graph->exec);
    cudaGraphDestroy(graph->graph);
    delete graph;
    return OIBVH_OK;
}

// ---------------------------------------------------------------------------------------------------
// tree
// ---------------------------------------------------------------------------------------------------
static int tree_create_impl(oibvh_ctx* ctx, const float* positions, uint32_t V, const uint32_t* indices, uint32_t T,
                            const float mesh_aabb[6], bool from_device, oibvh_tree** out)
{
    REQUIRE(ctx && positions && indices && mesh_aabb && out, "NULL argument");
    *out = nullptr;
    REQUIRE(T >= 2, "a tree needs at least 2 triangles (the reference's schedule underflows for 1)");
    REQUIRE(T <= (1u << kNodeLevelShift), "too many triangles for one tree (max 2^26)");
    REQUIRE(V >= 1, "no vertices");
    if (!from_device)
    {
        const size_t n = (size_t)T * 3;
        uint32_t mx = 0;
        for (size_t i = 0; i < n; i++) mx = std::max(mx, indices[i]);
        if (mx >= V) return fail(OIBVH_ERR_INVALID, "triangle index %u out of range (V = %u)", mx, V);
    }
    DeviceGuard g(ctx->device);
    oibvh_tree* t = nullptr;
    int rc = tree_alloc(ctx, V, T, mesh_aabb, &t);
    if (rc) return rc;
    // packed triples arrive (H2D, or straight from the caller's device buffers) and are widened to 16-byte records
    cudaError_t e = cudaSuccess;
    uint32_t* faces_stage = nullptr;
    if (from_device)
    {
        e = launch_pack_positions(positions, t->pos, V, ctx->stream);
        if (e == cudaSuccess) e = launch_pack_faces(indices, t->faces_in, T, ctx->stream);
    }
    else
    {
        e = cudaMalloc(reinterpret_cast<void**>(&faces_stage), sizeof(uint32_t) * 3 * (size_t)T);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(t->pos_stage, positions, sizeof(float) * 3 * (size_t)V, cudaMemcpyHostToDevice,
                                ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(faces_stage, indices, sizeof(uint32_t) * 3 * (size_t)T, cudaMemcpyHostToDevice,
                                ctx->stream);
        if (e == cudaSuccess) e = launch_pack_positions(t->pos_stage, t->pos, V, ctx->stream);
        if (e == cudaSuccess) e = launch_pack_faces(faces_stage, t->faces_in, T, ctx->stream);
    }
    count_launch(ctx, 2);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // caller may free its buffers
    cudaFree(faces_stage);
    if (e != cudaSuccess)
    {
        tree_free(t);
        delete t;
        return fail(OIBVH_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    *out = t;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_create(oibvh_ctx* ctx, const float* host_positions, uint32_t V, const uint32_t* host_indices,
                                 uint32_t T, const float mesh_aabb[6], oibvh_tree** out)
{
    return tree_create_impl(ctx, host_positions, V, host_indices, T, mesh_aabb, false, out);
}

extern "C" int oibvh_tree_create_from_device(oibvh_ctx* ctx, const float* dev_positions, uint32_t V,
                                             const uint32_t* dev_indices, uint32_t T, const float mesh_aabb[6],
                                             oibvh_tree** out)
{
    return tree_create_impl(ctx, dev_positions, V, dev_indices, T, mesh_aabb, true, out);
}

extern "C" int oibvh_tree_clone(const oibvh_tree* other, oibvh_tree** out)
{
    REQUIRE(other && out, "NULL argument");
    *out = nullptr;
    oibvh_ctx* ctx = other->ctx;
    DeviceGuard g(ctx->device);
    {
        int frc = tree_flush_upload(const_cast<oibvh_tree*>(other));
        if (frc) return frc;
    }
    oibvh_tree* t = nullptr;
    int rc = tree_alloc(ctx, other->V, other->T, other->mesh.v, &t);
    if (rc) return rc;
    const size_t T = other->T;
    cudaError_t e = cudaSuccess;
    auto cp = [&](void* d, const void* s, size_t bytes) {
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
    };
    cp(t->pos, other->pos, sizeof(float4) * (size_t)other->V);
    cp(t->faces_in, other->faces_in, sizeof(uint4) * T);
    if (other->built)
    {
        cp(t->faces, other->faces, sizeof(uint32_t) * 3 * T);
        cp(t->nodes, other->nodes, sizeof(float) * 6 * (size_t)other->N);
        cp(t->keys_a, other->keys_a, sizeof(uint32_t) * T);
        cp(t->vals_a, other->vals_a, sizeof(uint32_t) * T);
    }
    if (e != cudaSuccess)
    {
        tree_free(t);
        delete t;
        return fail(OIBVH_ERR_CUDA, "clone copy failed: %s", cudaGetErrorString(e));
    }
    t->built = other->built;
    *out = t;
    return OIBVH_OK;
}

// device-to-device copy of everything a built tree consists of, from src's device to dst's, on dst's stream
static int tree_copy_state(oibvh_tree* dst, const oibvh_tree* src, bool with_faces)
{
    oibvh_ctx* dc = dst->ctx;
    const oibvh_ctx* sc = src->ctx;
    const size_t T = src->T;
    cudaError_t e = cudaSuccess;
    auto cp = [&](void* d, const void* s, size_t bytes) {
        if (e != cudaSuccess) return;
        e = dc->device == sc->device ? cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, dc->stream)
                                     : cudaMemcpyPeerAsync(d, dc->device, s, sc->device, bytes, dc->stream);
    };
    cp(dst->pos, src->pos, sizeof(float4) * (size_t)src->V);
    if (with_faces) cp(dst->faces_in, src->faces_in, sizeof(uint4) * T);
    if (src->built)
    {
        cp(dst->nodes, src->nodes, sizeof(float) * 6 * (size_t)src->N);
        if (with_faces)
        {
            cp(dst->faces, src->faces, sizeof(uint32_t) * 3 * T);
            cp(dst->keys_a, src->keys_a, sizeof(uint32_t) * T);
            cp(dst->vals_a, src->vals_a, sizeof(uint32_t) * T);
        }
    }
    if (e != cudaSuccess) return fail(OIBVH_ERR_CUDA, "replica copy failed: %s", cudaGetErrorString(e));
    dst->built = src->built;
    dst->upload_pending = false;
    dst->epoch++;
    dst->build_serial++;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_replicate(const oibvh_tree* src, oibvh_ctx* dst_ctx, oibvh_tree** out)
{
    REQUIRE(src && dst_ctx && out, "NULL argument");
    *out = nullptr;
    REQUIRE(!dst_ctx->capturing && !src->ctx->capturing, "replicate outside a graph capture");
    {
        DeviceGuard gs(src->ctx->device);
        int rc = tree_flush_upload(const_cast<oibvh_tree*>(src));
        if (rc) return rc;
        CU(cudaStreamSynchronize(src->ctx->stream)); // src's pending builds / refits have produced what is copied
    }
    DeviceGuard g(dst_ctx->device);
    oibvh_tree* t = nullptr;
    int rc = tree_alloc(dst_ctx, src->V, src->T, src->mesh.v, &t);
    if (rc) return rc;
    rc = tree_copy_state(t, src, true);
    if (rc)
    {
        tree_free(t);
        delete t;
        return rc;
    }
    *out = t;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_sync_replica(oibvh_tree* replica, const oibvh_tree* src)
{
    REQUIRE(replica && src, "NULL argument");
    REQUIRE(replica->T == src->T && replica->V == src->V, "not a replica of this tree (sizes differ)");
    REQUIRE(!replica->ctx->capturing && !src->ctx->capturing, "sync outside a graph capture");
    {
        DeviceGuard gs(src->ctx->device);
        int rc = tree_flush_upload(const_cast<oibvh_tree*>(src));
        if (rc) return rc;
        CU(cudaStreamSynchronize(src->ctx->stream));
    }
    DeviceGuard g(replica->ctx->device);
    return tree_copy_state(replica, src, true);
}

extern "C" int oibvh_tree_destroy(oibvh_tree* tree)
{
    if (!tree) return OIBVH_OK;
    DeviceGuard g(tree->ctx->device);
    cudaStreamSynchronize(tree->ctx->stream);
    tree->ctx->generation++; // graphs that captured this tree's buffers must not replay
    tree->ctx->small_table.key.clear(); // cached *_many tables may name this tree
    tree->ctx->xform_table.key.clear();
    tree_free(tree);
    delete tree;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_set_positions(oibvh_tree* tree, const float* host_positions)
{
    REQUIRE(tree && host_positions, "NULL argument");
    DeviceGuard g(tree->ctx->device);
    oibvh_ctx* ctx = tree->ctx;
    const size_t bytes = sizeof(float) * 3 * (size_t)tree->V;
    if (ctx->capturing)
    {
        // inside a graph capture everything stays on the one captured stream. The memcpy node keeps the CALLER'S
        // pointer: every replay reads host_positions again, so the buffer must outlive the graph (documented in the
        // header); use oibvh_tree_set_positions_from_device for data that is produced per frame.
        CU(cudaMemcpyAsync(tree->pos_stage, host_positions, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    else
    {
        // The upload runs on the context's copy stream, so it overlaps whatever the compute stream is still doing
        // (e.g. the previous tree's build); the compute stream only waits for it right before it needs the data.
        // The staging buffer is reused: the copy first waits until its previous content has been packed.
        if (tree->upload_pending)
        {
            // a previous upload was never consumed: it is simply superseded (same staging buffer, same stream order)
        }
        else if (tree->consumed_recorded)
            CU(cudaStreamWaitEvent(ctx->copy_stream, tree->ev_consumed, 0));
        CU(cudaMemcpyAsync(tree->pos_stage, host_positions, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaEventRecord(tree->ev_uploaded, ctx->copy_stream));
        tree->upload_pending = true;
        tree->epoch++;
        return OIBVH_OK;
    }
    tree->epoch++;
    CU(launch_pack_positions(tree->pos_stage, tree->pos, tree->V, ctx->stream));
    count_launch(ctx);
    return OIBVH_OK;
}

extern "C" int oibvh_tree_set_positions_from_device(oibvh_tree* tree, const float* dev_positions)
{
    REQUIRE(tree && dev_positions, "NULL argument");
    DeviceGuard g(tree->ctx->device);
    tree->upload_pending = false; // superseded
    tree->epoch++;
    CU(launch_pack_positions(dev_positions, tree->pos, tree->V, tree->ctx->stream));
    count_launch(tree->ctx);
    return OIBVH_OK;
}

extern "C" int oibvh_tree_transform(oibvh_tree* tree, const float M[16])
{
    REQUIRE(tree && M, "NULL argument");
    DeviceGuard g(tree->ctx->device);
    {
        int rc = tree_flush_upload(tree);
        if (rc) return rc;
    }
    Mat4 m;
    memcpy(m.m, M, sizeof(float) * 16);
    tree->epoch++;
    StageScope st(tree->ctx, OIBVH_STAGE_TRANSFORM);
    CU(launch_transform(tree->pos, tree->V, m, tree->ctx->stream));
    count_launch(tree->ctx);
    return OIBVH_OK;
}

extern "C" int oibvh_tree_build(oibvh_tree* tree)
{
    REQUIRE(tree != nullptr, "tree is NULL");
    oibvh_ctx* ctx = tree->ctx;
    DeviceGuard g(ctx->device);
    {
        int rc = tree_flush_upload(tree);
        if (rc) return rc;
    }
    StageScope scope(ctx, OIBVH_STAGE_BUILD);
    cudaStream_t s = ctx->stream;
    if (tree->small)
    {
        // keys, sort, gather, leaves and all levels by one CTA
        CU(launch_small_trees(true, tree->d_small, tree->T <= kSmallBitonicSplit ? 1u : 0u, 1, s));
        count_launch(ctx);
        tree->built = true;
        tree->epoch++;
        tree->build_serial++;
        return OIBVH_OK;
    }
    const size_t radix = (size_t)1 << kRadixBits;
    uint32_t* hist = tree->sort_ctl;
    uint32_t* ticket = tree->sort_ctl + kRadixPasses * radix;
    uint32_t* status = ticket + 64;
    if (tree->sort_rec)
    {
        // single-wave size: keys, then all radix passes in one cooperative launch (sort_lsd.cu)
        {
            StageScope sk(ctx, OIBVH_STAGE_KEYS);
            CU(launch_morton_hist(tree->faces_in, tree->pos, tree->T, tree->mesh, tree->keys_a, nullptr, s));
            count_launch(ctx);
        }
        StageScope ss(ctx, OIBVH_STAGE_SORT);
        CU(launch_lsd_sort_many(1, &tree->keys_a, &tree->vals_a, &tree->sort_rec, &tree->T, ctx->lsd_ctl, ctx->d_status, s));
        count_launch(ctx);
    }
    else
    {
        // streaming size: keys + digit histograms, then one onesweep launch per digit
        {
            StageScope sk(ctx, OIBVH_STAGE_KEYS);
            CU(cudaMemsetAsync(tree->sort_ctl, 0, tree->sort_ctl_words * sizeof(uint32_t), s));
            CU(launch_morton_hist(tree->faces_in, tree->pos, tree->T, tree->mesh, tree->keys_a, hist, s));
            count_launch(ctx);
        }
        StageScope ss(ctx, OIBVH_STAGE_SORT);
        uint32_t *kin = tree->keys_a, *kout = tree->keys_b, *vin = nullptr, *vout = tree->vals_b;
        for (int p = 0; p < kRadixPasses; p++)
        {
            CU(launch_onesweep_pass(kin, vin, kout, vout, tree->T, p, hist, status, ticket, s));
            count_launch(ctx);
            // ping-pong: after pass p the data is in (kout, vout)
            uint32_t* nk = kout;
            uint32_t* nv = vout;
            kout = (nk == tree->keys_b) ? tree->keys_a : tree->keys_b;
            vout = (nv == tree->vals_b) ? tree->vals_a : tree->vals_b;
            kin = nk;
            vin = nv;
        }
    }
    static_assert(kRadixPasses % 2 == 0, "an even number of passes leaves the result in (keys_a, vals_a)");
    StageScope se(ctx, OIBVH_STAGE_EMIT);
    CU(launch_tree_emit(true, tree->faces_in, tree->vals_a, tree->faces, tree->pos, tree->nodes, tree->T,
                        tree->done_counter, ctx->d_status, s));
    count_launch(ctx);
    tree->built = true;
    tree->epoch++;
    tree->build_serial++;
    return OIBVH_OK;
}

// Device table for a list of trees, cached on the context while the same list is passed again.
static int batch_table_upload(oibvh_ctx* ctx, oibvh_ctx::BatchTable& tab, const std::vector<oibvh_tree*>& list,
                              const void* host, size_t bytes)
{
    REQUIRE(!ctx->capturing, "call the *_many entry point once with this list of trees before capturing it");
    tab.key.clear();
    ctx->generation++; // graphs captured with the previous table must not replay
    if (bytes > tab.cap_bytes)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(tab.dev);
        tab.dev = nullptr;
        tab.cap_bytes = 0;
        CU(cudaMalloc(&tab.dev, bytes));
        tab.cap_bytes = bytes;
    }
    CU(cudaMemcpyAsync(tab.dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // the host copy goes out of scope
    tab.key = list;
    return OIBVH_OK;
}

// table order: trees of up to kSmallBitonicSplit triangles first (their builds run the bitonic variant), then the rest
static int small_batch_table(oibvh_ctx* ctx, const std::vector<oibvh_tree*>& list, const SmallTreeDesc** out,
                             uint32_t* n_bitonic)
{
    auto& tab = ctx->small_table;
    if (!(tab.dev && tab.key == list))
    {
        std::vector<SmallTreeDesc> h;
        h.reserve(list.size());
        for (auto* t : list)
            if (t->T <= kSmallBitonicSplit) h.push_back(small_desc_of(t));
        const uint32_t first = (uint32_t)h.size();
        for (auto* t : list)
            if (t->T > kSmallBitonicSplit) h.push_back(small_desc_of(t));
        int rc = batch_table_upload(ctx, tab, list, h.data(), sizeof(SmallTreeDesc) * h.size());
        if (rc) return rc;
        tab.total_blocks = first;
    }
    *out = static_cast<const SmallTreeDesc*>(tab.dev);
    *n_bitonic = tab.total_blocks;
    return OIBVH_OK;
}

static int build_large_many(oibvh_tree* const* trees, uint32_t n);

extern "C" int oibvh_tree_build_many(oibvh_tree* const* trees, uint32_t n)
{
    REQUIRE(trees != nullptr && n >= 1, "no trees");
    oibvh_ctx* ctx = trees[0] ? trees[0]->ctx : nullptr;
    std::vector<oibvh_tree*> small_list, large_list;
    for (uint32_t i = 0; i < n; i++)
    {
        REQUIRE(trees[i] != nullptr, "NULL tree");
        REQUIRE(trees[i]->ctx == ctx, "trees belong to different contexts");
        (trees[i]->small ? small_list : large_list).push_back(trees[i]);
    }
    if (small_list.size() == 1)
    {
        int rc = oibvh_tree_build(small_list[0]);
        if (rc) return rc;
    }
    else if (!small_list.empty())
    {
        // every small tree of the list in ONE launch, one CTA each
        DeviceGuard g(ctx->device);
        const SmallTreeDesc* table = nullptr;
        uint32_t n_bitonic = 0;
        int rc = small_batch_table(ctx, small_list, &table, &n_bitonic);
        if (rc) return rc;
        for (auto* t : small_list)
            if ((rc = tree_flush_upload(t))) return rc;
        StageScope scope(ctx, OIBVH_STAGE_BUILD);
        CU(launch_small_trees(true, table, n_bitonic, (uint32_t)small_list.size(), ctx->stream));
        count_launch(ctx, (n_bitonic ? 1 : 0) + (n_bitonic < small_list.size() ? 1 : 0));
        for (auto* t : small_list)
        {
            t->built = true;
            t->epoch++;
            t->build_serial++;
        }
    }
    if (large_list.empty()) return OIBVH_OK;
    return build_large_many(large_list.data(), (uint32_t)large_list.size());
}

static int build_large_many(oibvh_tree* const* trees, uint32_t n)
{
    oibvh_ctx* ctx = trees[0]->ctx;
    uint64_t total = 0;
    bool batch = n >= 2 && n <= 4;
    for (uint32_t i = 0; i < n; i++)
    {
        REQUIRE(trees[i]->ctx == ctx, "trees belong to different contexts");
        total += trees[i]->T;
    }
    // every tree must fit the single-wave sort and all of them together one launch; the launcher re-checks the split
    if (batch && total > lsd_sort_capacity_multi()) batch = false;
    for (uint32_t i = 0; i < n && batch; i++)
        if (!trees[i]->sort_rec) batch = false;
    if (!batch)
    {
        for (uint32_t i = 0; i < n; i++)
        {
            int rc = oibvh_tree_build(trees[i]);
            if (rc) return rc;
        }
        return OIBVH_OK;
    }
    DeviceGuard g(ctx->device);
    for (uint32_t i = 0; i < n; i++)
    {
        int rc = tree_flush_upload(trees[i]);
        if (rc) return rc;
    }
    StageScope scope(ctx, OIBVH_STAGE_BUILD);
    cudaStream_t s = ctx->stream;
    uint32_t *ka[4], *va[4], T[4];
    uint2* rec[4];
    // the key kernels of different trees are independent and gather-bound: odd trees go to the auxiliary stream like
    // the emit kernels below
    std::unique_ptr<StageScope> sub(new StageScope(ctx, OIBVH_STAGE_KEYS)); // per-kernel clocks inside the stage
    CU(cudaEventRecord(ctx->ev_fork, s));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    for (uint32_t i = 0; i < n; i++)
    {
        oibvh_tree* t = trees[i];
        cudaStream_t st = (i & 1) ? ctx->aux_stream : s;
        CU(launch_morton_hist(t->faces_in, t->pos, t->T, t->mesh, t->keys_a, nullptr, st));
        count_launch(ctx);
        ka[i] = t->keys_a; va[i] = t->vals_a; rec[i] = t->sort_rec; T[i] = t->T;
    }
    CU(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    sub.reset(); // records the end of the key kernels before the next scope opens
    sub.reset(new StageScope(ctx, OIBVH_STAGE_SORT));
    const cudaError_t e = launch_lsd_sort_many(n, ka, va, rec, T, ctx->lsd_ctl, ctx->d_status, s);
    if (e == cudaErrorInvalidValue)
    {
        // very unequal sizes (a tree would need more than kLsdKMulti keys per thread of its CTA share): sort one by one
        cudaGetLastError();
        for (uint32_t i = 0; i < n; i++)
        {
            oibvh_tree* t = trees[i];
            CU(launch_lsd_sort_many(1, &t->keys_a, &t->vals_a, &t->sort_rec, &t->T, ctx->lsd_ctl, ctx->d_status, s));
            count_launch(ctx);
        }
    }
    else
    {
        CU(e);
        count_launch(ctx);
    }
    // the emit kernels are independent: odd trees go to the auxiliary stream, so the kernels share the machine and
    // the wave quantisation of each (1024 CTAs on 444 slots = 2.3 -> 3 waves) is paid once for all of them
    sub.reset();
    sub.reset(new StageScope(ctx, OIBVH_STAGE_EMIT));
    CU(cudaEventRecord(ctx->ev_fork, s));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    for (uint32_t i = 0; i < n; i++)
    {
        oibvh_tree* t = trees[i];
        CU(launch_tree_emit(true, t->faces_in, t->vals_a, t->faces, t->pos, t->nodes, t->T, t->done_counter, ctx->d_status,
                            (i & 1) ? ctx->aux_stream : s));
        count_launch(ctx);
        t->built = true;
        t->epoch++;
        t->build_serial++;
    }
    CU(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    return OIBVH_OK;
}

extern "C" int oibvh_tree_refit(oibvh_tree* tree)
{
    REQUIRE(tree != nullptr, "tree is NULL");
    REQUIRE(tree->built, "refit before build");
    oibvh_ctx* ctx = tree->ctx;
    DeviceGuard g(ctx->device);
    {
        int rc = tree_flush_upload(tree);
        if (rc) return rc;
    }
    StageScope scope(ctx, OIBVH_STAGE_REFIT);
    if (tree->small)
        CU(launch_small_trees(false, tree->d_small, 0, 1, ctx->stream));
    else
        CU(launch_tree_emit(false, nullptr, nullptr, tree->faces, tree->pos, tree->nodes, tree->T, tree->done_counter, ctx->d_status,
                            ctx->stream));
    count_launch(ctx);
    tree->epoch++;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_refit_many(oibvh_tree* const* trees, uint32_t n)
{
    REQUIRE(trees != nullptr && n >= 1, "no trees");
    oibvh_ctx* ctx = trees[0] ? trees[0]->ctx : nullptr;
    std::vector<oibvh_tree*> small_list;
    for (uint32_t i = 0; i < n; i++)
    {
        REQUIRE(trees[i] != nullptr, "NULL tree");
        REQUIRE(trees[i]->ctx == ctx, "trees belong to different contexts");
        REQUIRE(trees[i]->built, "refit before build");
        if (trees[i]->small) small_list.push_back(trees[i]);
    }
    DeviceGuard g(ctx->device);
    if (small_list.size() >= 2)
    {
        const SmallTreeDesc* table = nullptr;
        uint32_t n_bitonic = 0;
        int rc = small_batch_table(ctx, small_list, &table, &n_bitonic);
        if (rc) return rc;
        for (auto* t : small_list)
            if ((rc = tree_flush_upload(t))) return rc;
        StageScope scope(ctx, OIBVH_STAGE_REFIT);
        CU(launch_small_trees(false, table, n_bitonic, (uint32_t)small_list.size(), ctx->stream));
        count_launch(ctx);
        for (auto* t : small_list) t->epoch++;
    }
    // the remaining trees one launch each; independent launches alternate between the two streams (see build_many)
    std::vector<oibvh_tree*> rest;
    for (uint32_t i = 0; i < n; i++)
        if (!trees[i]->small || small_list.size() < 2) rest.push_back(trees[i]);
    if (rest.size() < 2)
    {
        for (auto* t : rest)
        {
            int rc = oibvh_tree_refit(t);
            if (rc) return rc;
        }
        return OIBVH_OK;
    }
    for (auto* t : rest)
    {
        int rc = tree_flush_upload(t);
        if (rc) return rc;
    }
    StageScope scope(ctx, OIBVH_STAGE_REFIT);
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    for (size_t i = 0; i < rest.size(); i++)
    {
        oibvh_tree* t = rest[i];
        cudaStream_t st = (i & 1) ? ctx->aux_stream : ctx->stream;
        if (t->small)
            CU(launch_small_trees(false, t->d_small, 0, 1, st));
        else
            CU(launch_tree_emit(false, nullptr, nullptr, t->faces, t->pos, t->nodes, t->T, t->done_counter, ctx->d_status, st));
        count_launch(ctx);
        t->epoch++;
    }
    CU(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return OIBVH_OK;
}

// Rigid transform + refit of several trees in one call (the transform belongs to the refit). Per tree
// the two launches are ordered on ONE stream and independent trees alternate between the context's two streams, so
// the transform of one body runs under the refit of another instead of in front of both (in the two-body frame: the
// 6 us transform of B disappears under the refit of A). Matrices travel as kernel arguments, so the call can be
// captured into a graph with host matrices. Same kernels, same bits as oibvh_tree_transform + oibvh_tree_refit.
extern "C" int oibvh_tree_transform_refit_many(oibvh_tree* const* trees, uint32_t n, const float* host_mats,
                                               const unsigned char* apply)
{
    REQUIRE(trees != nullptr && n >= 1, "no trees");
    REQUIRE(host_mats != nullptr || apply != nullptr, "matrices are NULL");
    oibvh_ctx* ctx = trees[0] ? trees[0]->ctx : nullptr;
    bool any_small = false;
    for (uint32_t i = 0; i < n; i++)
    {
        REQUIRE(trees[i] != nullptr, "NULL tree");
        REQUIRE(trees[i]->ctx == ctx, "trees belong to different contexts");
        REQUIRE(trees[i]->built, "refit before build");
        REQUIRE(!(apply && apply[i]) || host_mats != nullptr, "a tree is to be transformed but the matrices are NULL");
        any_small = any_small || trees[i]->small;
    }
    if (any_small || n < 2)
    {
        // many-body scenes go through the batched launches (one transform launch, one refit launch for all)
        for (uint32_t i = 0; i < n; i++)
            if (!apply || apply[i])
            {
                int rc = oibvh_tree_transform(trees[i], host_mats + 16 * (size_t)i);
                if (rc) return rc;
            }
        return oibvh_tree_refit_many(trees, n);
    }
    DeviceGuard g(ctx->device);
    for (uint32_t i = 0; i < n; i++)
    {
        int rc = tree_flush_upload(trees[i]);
        if (rc) return rc;
    }
    StageScope scope(ctx, OIBVH_STAGE_REFIT); // (the transforms are inside: they overlap the other trees' refits)
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    // trees WITH a transform first on the auxiliary stream so that the longer chains start first
    for (uint32_t i = 0; i < n; i++)
    {
        oibvh_tree* t = trees[i];
        const bool xf = !apply || apply[i];
        cudaStream_t st = (i & 1) ? ctx->aux_stream : ctx->stream;
        if (xf)
        {
            Mat4 m;
            memcpy(m.m, host_mats + 16 * (size_t)i, sizeof(float) * 16);
            CU(launch_transform(t->pos, t->V, m, st));
            count_launch(ctx);
        }
        CU(launch_tree_emit(false, nullptr, nullptr, t->faces, t->pos, t->nodes, t->T, t->done_counter, ctx->d_status, st));
        count_launch(ctx);
        t->epoch++;
    }
    CU(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return OIBVH_OK;
}

static int transform_many_impl(oibvh_tree* const* trees, uint32_t n, const float* mats, bool from_device)
{
    REQUIRE(trees != nullptr && n >= 1 && mats != nullptr, "NULL argument");
    oibvh_ctx* ctx = trees[0] ? trees[0]->ctx : nullptr;
    std::vector<oibvh_tree*> list(trees, trees + n);
    for (uint32_t i = 0; i < n; i++)
    {
        REQUIRE(trees[i] != nullptr, "NULL tree");
        REQUIRE(trees[i]->ctx == ctx, "trees belong to different contexts");
    }
    DeviceGuard g(ctx->device);
    auto& tab = ctx->xform_table;
    if (!(tab.dev && tab.key == list))
    {
        uint64_t blocks = 0;
        for (uint32_t i = 0; i < n; i++) blocks += (trees[i]->V + 255) / 256;
        REQUIRE(blocks < 0x7fffffffull, "too many vertices for one transform launch");
        // n descriptors, then the tree index of every 256-vertex block
        static_assert(sizeof(XformDesc) == 16, "descriptor layout");
        std::vector<unsigned char> h(sizeof(XformDesc) * (size_t)n + sizeof(uint32_t) * (size_t)blocks);
        XformDesc* hd = reinterpret_cast<XformDesc*>(h.data());
        uint32_t* hb = reinterpret_cast<uint32_t*>(hd + n);
        uint32_t next = 0;
        for (uint32_t i = 0; i < n; i++)
        {
            hd[i].pos = trees[i]->pos;
            hd[i].V = trees[i]->V;
            hd[i].block0 = next;
            for (uint32_t b = 0; b < (trees[i]->V + 255) / 256; b++) hb[next++] = i;
        }
        int rc = batch_table_upload(ctx, tab, list, h.data(), h.size());
        if (rc) return rc;
        tab.total_blocks = (uint32_t)blocks;
    }
    for (uint32_t i = 0; i < n; i++)
    {
        int rc = tree_flush_upload(trees[i]);
        if (rc) return rc;
    }
    const float* d_mats = mats;
    if (!from_device)
    {
        const size_t bytes = sizeof(float) * 16 * (size_t)n;
        if (bytes > ctx->d_mats_cap)
        {
            REQUIRE(!ctx->capturing, "call oibvh_tree_transform_many once with this many trees before capturing it");
            CU(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_mats);
            ctx->d_mats = nullptr;
            ctx->d_mats_cap = 0;
            ctx->generation++;
            CU(cudaMalloc(reinterpret_cast<void**>(&ctx->d_mats), bytes));
            ctx->d_mats_cap = bytes;
        }
        // (a captured graph would re-read the caller's host array on every replay, and callers pass temporaries)
        REQUIRE(!ctx->capturing, "host matrices cannot be captured into a graph: use oibvh_tree_transform_many_from_device");
        CU(cudaMemcpyAsync(ctx->d_mats, mats, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_mats = ctx->d_mats;
    }
    StageScope st(ctx, OIBVH_STAGE_TRANSFORM);
    CU(launch_transform_many(static_cast<const XformDesc*>(tab.dev), n, tab.total_blocks, d_mats, ctx->stream));
    count_launch(ctx);
    for (uint32_t i = 0; i < n; i++) trees[i]->epoch++;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_transform_many(oibvh_tree* const* trees, uint32_t n, const float* host_mats)
{
    return transform_many_impl(trees, n, host_mats, false);
}

extern "C" int oibvh_tree_transform_many_from_device(oibvh_tree* const* trees, uint32_t n, const float* dev_mats)
{
    return transform_many_impl(trees, n, dev_mats, true);
}

extern "C" int oibvh_tree_get_info(const oibvh_tree* tree, uint32_t* T, uint32_t* V, uint32_t* N, uint32_t* depth)
{
    REQUIRE(tree != nullptr, "tree is NULL");
    if (T) *T = tree->T;
    if (V) *V = tree->V;
    if (N) *N = tree->N;
    if (depth)
    {
        // OibvhTree::getDepth() = ilog2(m_aabbTree.size())  (src/cuda/oibvhTree.cu:45-48)
        uint32_t d = 0;
        while ((2ull << d) <= tree->N) d++;
        *depth = d;
    }
    return OIBVH_OK;
}

extern "C" int oibvh_tree_is_built(const oibvh_tree* tree, int* built)
{
    REQUIRE(tree && built, "NULL argument");
    *built = tree->built ? 1 : 0;
    return OIBVH_OK;
}

extern "C" int oibvh_tree_download(oibvh_tree* tree, oibvh_aabb* host_nodes, uint32_t* host_sorted_faces,
                                   uint32_t* host_perm)
{
    REQUIRE(tree != nullptr, "tree is NULL");
    REQUIRE(tree->built, "download before build");
    oibvh_ctx* ctx = tree->ctx;
    DeviceGuard g(ctx->device);
    static_assert(sizeof(oibvh_aabb) == 24, "aabb record is 24 bytes");
    if (host_nodes)
        CU(cudaMemcpyAsync(host_nodes, tree->nodes, sizeof(oibvh_aabb) * (size_t)tree->N, cudaMemcpyDeviceToHost,
                           ctx->stream));
    if (host_sorted_faces)
        CU(cudaMemcpyAsync(host_sorted_faces, tree->faces, sizeof(uint32_t) * 3 * (size_t)tree->T,
                           cudaMemcpyDeviceToHost, ctx->stream));
    if (host_perm)
        CU(cudaMemcpyAsync(host_perm, tree->vals_a, sizeof(uint32_t) * (size_t)tree->T, cudaMemcpyDeviceToHost,
                           ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ctx_check_status(ctx);
}

extern "C" int oibvh_tree_download_positions(oibvh_tree* tree, float* host_positions)
{
    REQUIRE(tree && host_positions, "NULL argument");
    DeviceGuard g(tree->ctx->device);
    {
        int rc = tree_flush_upload(tree);
        if (rc) return rc;
    }
    CU(launch_unpack_positions(tree->pos, tree->pos_stage, tree->V, tree->ctx->stream));
    count_launch(tree->ctx);
    CU(cudaMemcpyAsync(host_positions, tree->pos_stage, sizeof(float) * 3 * (size_t)tree->V, cudaMemcpyDeviceToHost,
                       tree->ctx->stream));
    CU(cudaStreamSynchronize(tree->ctx->stream));
    return OIBVH_OK;
}

extern "C" int oibvh_tree_download_keys(oibvh_tree* tree, uint32_t* host_sorted_keys)
{
    REQUIRE(tree && host_sorted_keys, "NULL argument");
    REQUIRE(tree->built, "download before build");
    DeviceGuard g(tree->ctx->device);
    CU(cudaMemcpyAsync(host_sorted_keys, tree->keys_a, sizeof(uint32_t) * (size_t)tree->T, cudaMemcpyDeviceToHost,
                       tree->ctx->stream));
    CU(cudaStreamSynchronize(tree->ctx->stream));
    return OIBVH_OK;
}

extern "C" int oibvh_tree_device_views(oibvh_tree* tree, const oibvh_aabb** dev_nodes,
                                       const uint32_t** dev_sorted_faces, const float** dev_positions)
{
    REQUIRE(tree != nullptr, "tree is NULL");
    if (dev_nodes) *dev_nodes = reinterpret_cast<const oibvh_aabb*>(tree->nodes);
    if (dev_sorted_faces) *dev_sorted_faces = tree->faces;
    if (dev_positions) *dev_positions = reinterpret_cast<const float*>(tree->pos);
    return OIBVH_OK;
}

// ---------------------------------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------------------------------
extern "C" int oibvh_scene_create(oibvh_ctx* ctx, oibvh_scene** out)
{
    REQUIRE(ctx && out, "NULL argument");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    oibvh_scene* s = new (std::nothrow) oibvh_scene;
    if (!s) return fail(OIBVH_ERR_NOMEM, "host allocation failed");
    s->ctx = ctx;
    int rc = dev_alloc(&s->counters_own, (size_t)CTR_WORDS);
    if (rc)
    {
        delete s;
        return rc;
    }
    s->counters = s->counters_own;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&s->h_counters), sizeof(uint32_t) * CTR_WORDS);
    if (e != cudaSuccess)
    {
        cudaFree(s->counters_own);
        delete s;
        return fail(OIBVH_ERR_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(e));
    }
    memset(s->h_counters, 0, sizeof(uint32_t) * CTR_WORDS);
    rc = dev_alloc(&s->mg_state, (size_t)MG_WORDS); // persistent words: multi-GPU protocol, size of the recorded cut
    if (rc == OIBVH_OK && cudaMemset(s->mg_state, 0, sizeof(uint32_t) * MG_WORDS) != cudaSuccess) rc = OIBVH_ERR_CUDA;
    if (rc)
    {
        cudaFree(s->counters_own);
        cudaFreeHost(s->h_counters);
        delete s;
        return rc;
    }
    *out = s;
    return OIBVH_OK;
}

static void mgpu_unmap(oibvh_scene* s)
{
    if (s->mg_ipc_block) cudaIpcCloseMemHandle(s->mg_ipc_block);
    if (s->mg_ipc_state) cudaIpcCloseMemHandle(s->mg_ipc_state);
    s->mg_ipc_block = s->mg_ipc_state = nullptr;
    s->mg_root_state = s->mg_root_counters = nullptr;
    s->mg_root_pairs = nullptr;
    s->mg_root_pair_cap = 0;
    s->mg_mode = 0;
    s->mg_opened_ahead = false;
}

extern "C" int oibvh_scene_destroy(oibvh_scene* scene)
{
    if (!scene) return OIBVH_OK;
    DeviceGuard g(scene->ctx->device);
    cudaStreamSynchronize(scene->ctx->stream);
    scene->ctx->generation++; // graphs that captured this scene's buffers must not replay
    mgpu_unmap(scene);
    cudaFree(scene->mg_state);
    scene_free_buffers(scene);
    cudaFree(scene->d_objs);
    cudaFree(scene->counters_own);
    cudaFreeHost(scene->h_counters);
    delete scene;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_add_tree(oibvh_scene* scene, oibvh_tree* tree)
{
    REQUIRE(scene && tree, "NULL argument");
    REQUIRE(tree->built, "tree must be built before it is added (Scene::addOibvhTree asserts m_buildDone)");
    REQUIRE(tree->ctx == scene->ctx, "tree and scene belong to different contexts");
    scene->trees.push_back(tree);
    scene->objs_dirty = true;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_set_shard(oibvh_scene* scene, uint32_t rank, uint32_t world)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    REQUIRE(world >= 1 && rank < world, "need rank < world");
    scene->rank = rank;
    scene->world = world;
    return OIBVH_OK;
}

static int scene_upload_objs(oibvh_scene* s)
{
    if (!s->objs_dirty) return OIBVH_OK;
    oibvh_ctx* ctx = s->ctx;
    const size_t n = s->trees.size();
    if (n > s->d_objs_cap)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(s->d_objs);
        s->d_objs = nullptr;
        ctx->generation++;
        int rc = dev_alloc(&s->d_objs, n);
        if (rc) return rc;
        s->d_objs_cap = n;
    }
    std::vector<ObjDesc> h(n);
    for (size_t i = 0; i < n; i++)
    {
        const oibvh_tree* t = s->trees[i];
        h[i].nodes = t->nodes;
        h[i].faces = t->faces;
        h[i].pos = t->pos;
        h[i].T = t->T;
        h[i].L = t->L;
    }
    CU(cudaMemcpyAsync(s->d_objs, h.data(), sizeof(ObjDesc) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream)); // h goes out of scope
    s->objs_dirty = false;
    return OIBVH_OK;
}

// rerun = the overflow path of oibvh_scene_get_counts repeating the SAME detection with larger queues: it must read
// exactly what the first attempt read, so it neither flushes pending uploads nor runs if a tree changed in between
static int scene_enqueue(oibvh_scene* s, uint32_t entry_level, uint32_t expand_levels, bool rerun = false)
{
    const uint32_t requested_expand = expand_levels;
    oibvh_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t n_obj = (uint32_t)s->trees.size();
    uint32_t maxL = 0;
    for (auto* t : s->trees)
    {
        maxL = std::max(maxL, t->L);
        if (rerun) continue;
        int rc = tree_flush_upload(t); // the narrow phase reads the positions
        if (rc) return rc;
    }
    // One warp tests the 4^k descendant pairs of a node pair, 64 per step, so k = 3 costs one step per pair. The
    // front grows geometrically towards the leaves; the cheapest schedule (measured) makes EVERY hop after the first a
    // 3-level hop -- in particular the last, widest one -- and lets the root pairs absorb the remainder.
    // expand_levels == 0 selects that schedule (entry_level is then ignored: both are hints, the pair set does not
    // depend on them); explicit values are honoured up to 5 (round 0) / 4 levels.
    uint32_t k0;
    if (expand_levels == 0)
    {
        expand_levels = 3;
        k0 = maxL <= 5 ? maxL : 3 + (maxL - 3) % 3;
        // two-body scenes skip the first latency-bound rounds: all node pairs of level k0 are tested directly by
        // the whole grid (dense_seed_phase); k0 keeps (maxL - k0) a multiple of 3
        if (n_obj == 2 && maxL >= ctx->dense_seed_level + 3)
            k0 = ctx->dense_seed_level + (maxL - ctx->dense_seed_level) % 3;
    }
    else
    {
        expand_levels = std::min(expand_levels, 4u);
        k0 = entry_level > 0 ? std::min(entry_level, 5u) : expand_levels;
    }
    const uint32_t rounds = CTR_MAX_ROUNDS - 1; // statistics: one counter per tree level

    // temporal coherence: record the cut on the first detection (and whenever what it was recorded on has changed),
    // start from it otherwise. Scenes beyond the few-pairs regime (a cut would hold O(n^2) disjoint root pairs) always
    // start from the roots.
    DetectOpts opt{};
    opt.self = s->self_collision ? 1u : 0u;
    opt.cut = s->cut;
    opt.cut_cap = s->cand_cap;
    opt.cand = s->cand;
    opt.cand_cap = s->cand_cap;
    opt.cut_depth = s->cut_depth;
    opt.cut_state = s->mg_state + MG_CUT;
    const uint64_t n_pairs = (uint64_t)n_obj * (n_obj - 1) / 2 + (s->self_collision ? n_obj : 0);
    if (rerun)
        opt.mode = s->last_mode == 2 ? 2u : (s->last_mode == 1 ? 1u : 0u);
    else if (s->coherent && s->cut && n_pairs <= 4096)
    {
        std::vector<uint64_t> sig = {n_obj, s->rank, s->world, k0, expand_levels, opt.self, s->cut_depth, s->cand_cap};
        for (auto* t : s->trees) sig.push_back(t->build_serial);
        if (s->cut_valid && sig == s->cut_sig)
            opt.mode = 2;
        else
        {
            REQUIRE(!ctx->capturing, "the first detection of a coherent scene records the cut: run it before capturing");
            opt.mode = 1;
            s->cut_sig = sig;
            s->cut_valid = true;
        }
    }
    s->last_mode = opt.mode;
    MgpuArgs mg{};
    mg.mode = (uint32_t)s->mg_mode;
    mg.world = s->world;
    mg.state = s->mg_state;
    mg.root_state = s->mg_root_state;
    mg.root_counters = s->mg_root_counters;
    mg.root_pairs = s->mg_root_pairs;
    mg.root_pair_cap = s->mg_root_pair_cap;
    if (s->mg_mode == 1)
    {
        // the gathering rank zeroes its counter block and THEN lets the other ranks append (MG_OPEN), in a launch of
        // its own so that the opening never waits for this rank's machine-filling detection kernel
        if (!s->mg_opened_ahead)
        {
            CU(launch_mgpu_open(s->counters, s->mg_state, st));
            count_launch(ctx);
        }
        s->mg_opened_ahead = false;
    }
    else
        CU(cudaMemsetAsync(s->counters, 0, sizeof(uint32_t) * CTR_WORDS, st));
    {
        // broad and narrow phase run inside one persistent kernel; the stage clock covers both
        StageScope scope(ctx, OIBVH_STAGE_BROAD);
        CU(launch_collide(ctx->collide_grid, s->d_objs, n_obj, s->queue, s->front_cap, s->pairs, s->pair_cap,
                          s->counters, k0, expand_levels, s->rank, s->world, mg, opt, st));
        count_launch(ctx);
    }
    s->enq_epochs.resize(s->trees.size());
    for (size_t i = 0; i < s->trees.size(); i++) s->enq_epochs[i] = s->trees[i]->epoch;
    CU(cudaMemcpyAsync(s->h_counters, s->counters, sizeof(uint32_t) * CTR_WORDS, cudaMemcpyDeviceToHost, st));
    s->last_entry = entry_level;
    s->last_expand = requested_expand;
    s->last_rounds = rounds;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_set_self_collision(oibvh_scene* scene, int enable)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    scene->self_collision = enable != 0;
    scene->cut_valid = false;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_set_coherence(oibvh_scene* scene, int enable, uint32_t cut_depth)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    oibvh_ctx* ctx = scene->ctx;
    REQUIRE(!ctx->capturing, "set_coherence outside a graph capture");
    REQUIRE(scene->mg_mode == 0 || !enable || scene->cut != nullptr, "enable coherence before oibvh_mgpu_export / attach");
    DeviceGuard g(ctx->device);
    scene->coherent = enable != 0;
    scene->cut_depth = cut_depth ? std::min(cut_depth, 24u) : 6u;
    scene->cut_valid = false;
    if (scene->coherent && !scene->cut && scene->front_cap > 0)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        int rc = dev_alloc(&scene->cut, (size_t)scene->cand_cap);
        if (rc) return rc;
        ctx->generation++;
    }
    return OIBVH_OK;
}

extern "C" int oibvh_scene_reserve(oibvh_scene* scene, uint32_t front_records, uint32_t candidate_records,
                                   uint32_t pair_records)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    REQUIRE(scene->mg_mode == 0, "the queues of a multi-GPU scene are fixed (other ranks map them): detach first");
    oibvh_ctx* ctx = scene->ctx;
    REQUIRE(!ctx->capturing, "reserve outside a graph capture");
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    const uint32_t fc = std::max(std::max(front_records, scene->front_cap), 1u << 16);
    const uint32_t cc = std::max(std::max(candidate_records, scene->cand_cap), 1u << 16);
    const uint32_t pc = std::max(std::max(pair_records, scene->pair_cap), 1u << 16);
    if (fc == scene->front_cap && cc == scene->cand_cap && pc == scene->pair_cap) return OIBVH_OK;
    return scene_alloc_buffers(scene, fc, cc, pc);
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU: peer-mapped pair list of the gathering rank (see include/oibvh_b200.h)
// ---------------------------------------------------------------------------------------------------
namespace
{
struct MgpuHandleLayout
{
    cudaIpcMemHandle_t block; // [counters | pairs] allocation of the root
    cudaIpcMemHandle_t state; // protocol block of the root
    uint32_t pair_cap;
    int32_t device;
    uint64_t pid;             // same process: the raw pointers below are used instead of the IPC handles
    uint64_t block_ptr, state_ptr;
};
static_assert(sizeof(MgpuHandleLayout) <= sizeof(oibvh_mgpu_handle), "handle layout must fit the public struct");

int mgpu_state_alloc(oibvh_scene* s)
{
    s->cut_valid = false; // (the words of the protocol and of the cut share the block)
    CU(cudaMemsetAsync(s->mg_state, 0, sizeof(uint32_t) * MG_WORDS, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    return OIBVH_OK;
}
} // namespace

extern "C" int oibvh_mgpu_export(oibvh_scene* scene, oibvh_mgpu_handle* out)
{
    REQUIRE(scene && out, "NULL argument");
    REQUIRE(scene->rank == 0 && scene->world >= 2, "export from rank 0 of a world of >= 2 (oibvh_scene_set_shard first)");
    oibvh_ctx* ctx = scene->ctx;
    REQUIRE(!ctx->capturing, "export outside a graph capture");
    DeviceGuard g(ctx->device);
    if (scene->front_cap == 0)
    {
        int rc = scene_alloc_buffers(scene, 1u << 20, 1u << 20, 1u << 19);
        if (rc) return rc;
    }
    int rc = mgpu_state_alloc(scene);
    if (rc) return rc;
    MgpuHandleLayout h;
    memset(&h, 0, sizeof(h));
    CU(cudaIpcGetMemHandle(&h.block, scene->pair_block));
    CU(cudaIpcGetMemHandle(&h.state, scene->mg_state));
    h.pair_cap = scene->pair_cap;
    h.device = ctx->device;
    h.pid = (uint64_t)getpid();
    h.block_ptr = (uint64_t)(uintptr_t)scene->pair_block;
    h.state_ptr = (uint64_t)(uintptr_t)scene->mg_state;
    memset(out, 0, sizeof(*out));
    memcpy(out->bytes, &h, sizeof(h));
    scene->mg_mode = 1;
    scene->mg_opened_ahead = false;
    ctx->generation++; // graphs captured in single-GPU mode must be re-captured
    return OIBVH_OK;
}

extern "C" int oibvh_mgpu_attach(oibvh_scene* scene, const oibvh_mgpu_handle* root)
{
    REQUIRE(scene && root, "NULL argument");
    REQUIRE(scene->world >= 2 && scene->rank >= 1, "attach from a rank >= 1 (oibvh_scene_set_shard first)");
    REQUIRE(scene->mg_mode == 0, "scene is already in multi-GPU mode");
    oibvh_ctx* ctx = scene->ctx;
    REQUIRE(!ctx->capturing, "attach outside a graph capture");
    DeviceGuard g(ctx->device);
    MgpuHandleLayout h;
    memcpy(&h, root->bytes, sizeof(h));
    int rc = mgpu_state_alloc(scene);
    if (rc) return rc;
    void *block = nullptr, *state = nullptr;
    if (h.pid == (uint64_t)getpid())
    {
        // same process (one thread per GPU): the root's pointers are valid here once peer access is enabled
        if (h.device != ctx->device)
        {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
            if (!can) return fail(OIBVH_ERR_CUDA, "device %d cannot access device %d as a peer", ctx->device, h.device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                cudaGetLastError();
            else
                CU(e);
        }
        block = reinterpret_cast<void*>((uintptr_t)h.block_ptr);
        state = reinterpret_cast<void*>((uintptr_t)h.state_ptr);
    }
    else
    {
        CU(cudaIpcOpenMemHandle(&block, h.block, cudaIpcMemLazyEnablePeerAccess));
        const cudaError_t e = cudaIpcOpenMemHandle(&state, h.state, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
        {
            cudaIpcCloseMemHandle(block);
            return fail(OIBVH_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        }
        scene->mg_ipc_block = block;
        scene->mg_ipc_state = state;
    }
    static_assert(CTR_WORDS * sizeof(uint32_t) % sizeof(uint4) == 0, "the counter block is a whole number of records");
    scene->mg_root_counters = static_cast<uint32_t*>(block);
    scene->mg_root_pairs = static_cast<uint4*>(block) + CTR_WORDS * sizeof(uint32_t) / sizeof(uint4);
    scene->mg_root_state = static_cast<uint32_t*>(state);
    scene->mg_root_pair_cap = h.pair_cap;
    scene->mg_mode = 2;
    ctx->generation++;
    return OIBVH_OK;
}

extern "C" int oibvh_mgpu_detach(oibvh_scene* scene)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    if (scene->mg_mode == 0) return OIBVH_OK;
    DeviceGuard g(scene->ctx->device);
    CU(cudaStreamSynchronize(scene->ctx->stream));
    mgpu_unmap(scene);
    scene->ctx->generation++;
    return OIBVH_OK;
}

extern "C" int oibvh_mgpu_open_frame(oibvh_scene* scene)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    REQUIRE(scene->mg_mode == 1, "only the gathering rank opens frames (oibvh_mgpu_export first)");
    REQUIRE(!scene->mg_opened_ahead, "the next frame is already open");
    oibvh_ctx* ctx = scene->ctx;
    DeviceGuard g(ctx->device);
    CU(launch_mgpu_open(scene->counters, scene->mg_state, ctx->stream));
    count_launch(ctx);
    scene->mg_opened_ahead = true;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_detect_async(oibvh_scene* scene, uint32_t entry_level, uint32_t expand_levels)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    REQUIRE(scene->trees.size() >= 1, "empty scene");
    oibvh_ctx* ctx = scene->ctx;
    DeviceGuard g(ctx->device);
    if (ctx->capturing)
    {
        REQUIRE(!scene->objs_dirty && scene->front_cap > 0, "run one detection before capturing a graph");
    }
    else
    {
        int rc = scene_upload_objs(scene);
        if (rc) return rc;
        if (scene->front_cap == 0)
        {
            rc = scene_alloc_buffers(scene, 1u << 20, 1u << 20, 1u << 19);
            if (rc) return rc;
        }
    }
    if (scene->trees.size() < 2 && !scene->self_collision)
    {
        // a single object has no pairs i<j (the reference loops over i<j only, scene.cu:195-196)
        CU(cudaMemsetAsync(scene->counters, 0, sizeof(uint32_t) * CTR_WORDS, ctx->stream));
        CU(cudaMemcpyAsync(scene->h_counters, scene->counters, sizeof(uint32_t) * CTR_WORDS, cudaMemcpyDeviceToHost,
                           ctx->stream));
        scene->last_rounds = 0;
        return OIBVH_OK;
    }
    return scene_enqueue(scene, entry_level, expand_levels);
}

extern "C" int oibvh_scene_get_counts(oibvh_scene* scene, uint32_t* n_pairs, uint32_t* n_candidates)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    oibvh_ctx* ctx = scene->ctx;
    DeviceGuard g(ctx->device);
    for (int attempt = 0; attempt < 12; attempt++)
    {
        CU(cudaStreamSynchronize(ctx->stream));
        const uint32_t* h = scene->h_counters;
        const uint32_t max_front = h[CTR_Q_TAIL]; // records pushed (a lower bound when the traversal was cut short)
        if (h[CTR_OVERFLOW] & 8u)
        {
            // the queue may hold stale records: put the empty marker back before anything else runs on it
            if (scene->queue) cudaMemsetAsync(scene->queue, 0xff, sizeof(uint4) * (size_t)scene->front_cap, ctx->stream);
            if (scene->cand) cudaMemsetAsync(scene->cand, 0xff, sizeof(uint4) * (size_t)scene->cand_cap, ctx->stream);
            return fail(OIBVH_ERR_INTERNAL, "a wait timed out in the detection kernel");
        }
        if (h[CTR_OVERFLOW] & 16u)
            return fail(OIBVH_ERR_INTERNAL, "multi-GPU detection: a rank did not open / finish the frame in time");
        if (h[CTR_OVERFLOW] != 0 && scene->queue)
        {
            // an aborted traversal leaves records in the queue: restore the empty marker (regrowing does it too)
            CU(cudaMemsetAsync(scene->queue, 0xff, sizeof(uint4) * (size_t)scene->front_cap, ctx->stream));
            CU(cudaMemsetAsync(scene->cand, 0xff, sizeof(uint4) * (size_t)scene->cand_cap, ctx->stream));
        }
        if (h[CTR_OVERFLOW] != 0 && scene->mg_mode != 0)
            return fail(OIBVH_ERR_OVERFLOW, "multi-GPU detection: a work queue overflowed (flags %u); the queues of a "
                                            "multi-GPU scene are fixed -- oibvh_scene_reserve before oibvh_mgpu_export",
                        h[CTR_OVERFLOW]);
        if (h[CTR_OVERFLOW] == 0)
        {
            if (n_pairs) *n_pairs = h[CTR_PAIRS];
            if (n_candidates) *n_candidates = h[CTR_CANDIDATES];
            return OIBVH_OK;
        }
        // a queue overflowed: grow (counts keep counting past the capacity, so they are lower bounds) and redo
        uint32_t fc = scene->front_cap, cc = scene->cand_cap, pc = scene->pair_cap;
        scene->cut_valid = false; // a detection that overflowed has not recorded a complete cut
        if (h[CTR_OVERFLOW] & 1u) fc = grow_to(fc, max_front); // every BVTT node of the traversal passes through the queue
        if (h[CTR_OVERFLOW] & 2u) cc = grow_to(cc, h[CTR_CUT]);
        if (h[CTR_OVERFLOW] & 32u) cc = grow_to(cc, h[CTR_CAND_TAIL]);
        if (h[CTR_OVERFLOW] & 4u) pc = grow_to(pc, h[CTR_PAIRS]);
        if (fc == scene->front_cap && cc == scene->cand_cap && pc == scene->pair_cap)
            return fail(OIBVH_ERR_OVERFLOW, "work queues cannot grow further");
        int rc = scene_alloc_buffers(scene, fc, cc, pc);
        if (rc) return rc;
        // Re-run only if the trees are exactly as the first attempt saw them. A caller that pipelines (detect_async,
        // then uploads / transforms / refits for the next frame, then reads back) gets OIBVH_ERR_OVERFLOW instead of
        // a pair set computed from a mixture of two frames; the queues are already regrown, so re-issuing the frame
        // succeeds.
        bool same = scene->enq_epochs.size() == scene->trees.size();
        for (size_t i = 0; same && i < scene->trees.size(); i++)
            same = scene->trees[i]->epoch == scene->enq_epochs[i] && !scene->trees[i]->upload_pending;
        if (!same)
            return fail(OIBVH_ERR_OVERFLOW, "a work queue overflowed and the trees have been modified since the "
                                            "detection was enqueued: the queues have been regrown, re-issue the frame");
        rc = scene_enqueue(scene, scene->last_entry, scene->last_expand, true);
        if (rc) return rc;
    }
    return fail(OIBVH_ERR_OVERFLOW, "work queues still overflow after repeated growth");
}

extern "C" int oibvh_scene_detect(oibvh_scene* scene, uint32_t entry_level, uint32_t expand_levels, uint32_t* n_pairs,
                                  uint32_t* n_candidates)
{
    int rc = oibvh_scene_detect_async(scene, entry_level, expand_levels);
    if (rc) return rc;
    return oibvh_scene_get_counts(scene, n_pairs, n_candidates);
}

extern "C" int oibvh_scene_get_pairs(oibvh_scene* scene, oibvh_int_tri_pair* host_pairs)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    REQUIRE(scene->mg_mode != 2, "multi-GPU: the pair list is gathered on rank 0");
    uint32_t n = 0;
    int rc = oibvh_scene_get_counts(scene, &n, nullptr);
    if (rc) return rc;
    if (n == 0) return OIBVH_OK;
    REQUIRE(host_pairs != nullptr, "host_pairs is NULL");
    DeviceGuard g(scene->ctx->device);
    static_assert(sizeof(oibvh_int_tri_pair) == sizeof(uint4), "pair record is 16 bytes");
    CU(cudaMemcpyAsync(host_pairs, scene->pairs, sizeof(oibvh_int_tri_pair) * (size_t)n, cudaMemcpyDeviceToHost,
                       scene->ctx->stream));
    CU(cudaStreamSynchronize(scene->ctx->stream));
    return OIBVH_OK;
}

extern "C" int oibvh_scene_device_pairs(oibvh_scene* scene, const oibvh_int_tri_pair** dev_pairs, uint32_t* n_pairs)
{
    REQUIRE(scene && dev_pairs && n_pairs, "NULL argument");
    REQUIRE(scene->mg_mode != 2, "multi-GPU: the pair list is gathered on rank 0");
    int rc = oibvh_scene_get_counts(scene, n_pairs, nullptr);
    if (rc) return rc;
    *dev_pairs = reinterpret_cast<const oibvh_int_tri_pair*>(scene->pairs);
    return OIBVH_OK;
}

extern "C" int oibvh_scene_pair_vertices_device(oibvh_scene* scene, float* dev_vertices, uint32_t capacity_pairs)
{
    REQUIRE(scene && dev_vertices, "NULL argument");
    REQUIRE(scene->pairs != nullptr && !scene->objs_dirty, "run a detection first");
    oibvh_ctx* ctx = scene->ctx;
    DeviceGuard g(ctx->device);
    CU(launch_pair_vertices(scene->d_objs, scene->pairs, scene->pair_cap, scene->counters, dev_vertices,
                            capacity_pairs, ctx->stream));
    count_launch(ctx);
    return OIBVH_OK;
}

extern "C" int oibvh_scene_pair_vertices(oibvh_scene* scene, float* host_vertices)
{
    REQUIRE(scene != nullptr, "scene is NULL");
    uint32_t n = 0;
    int rc = oibvh_scene_get_counts(scene, &n, nullptr); // also regrows the queues / re-runs after an overflow
    if (rc) return rc;
    if (n == 0) return OIBVH_OK;
    REQUIRE(host_vertices != nullptr, "host_vertices is NULL");
    oibvh_ctx* ctx = scene->ctx;
    DeviceGuard g(ctx->device);
    float* d = nullptr;
    rc = dev_alloc(&d, (size_t)n * 18);
    if (rc) return rc;
    cudaError_t e = launch_pair_vertices(scene->d_objs, scene->pairs, scene->pair_cap, scene->counters, d, n, ctx->stream);
    count_launch(ctx);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(host_vertices, d, sizeof(float) * 18 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(OIBVH_ERR_CUDA, "pair_vertices: %s", cudaGetErrorString(e));
    return OIBVH_OK;
}

extern "C" int oibvh_tree_box_wireframe(oibvh_tree* tree, uint32_t max_nodes, float* host_vertices,
                                        uint32_t* host_indices, uint32_t* n_boxes)
{
    REQUIRE(tree && n_boxes, "NULL argument");
    REQUIRE(tree->built, "tree is not built (OibvhTree::convertToVertexArray asserts m_buildDone)");
    const uint32_t n = std::min(tree->N - tree->T, max_nodes);
    *n_boxes = n;
    if (n == 0) return OIBVH_OK;
    REQUIRE(host_vertices && host_indices, "NULL output buffer");
    oibvh_ctx* ctx = tree->ctx;
    DeviceGuard g(ctx->device);
    float* dv = nullptr;
    uint32_t* di = nullptr;
    int rc = dev_alloc(&dv, (size_t)n * 24);
    if (rc) return rc;
    rc = dev_alloc(&di, (size_t)n * 24);
    if (rc)
    {
        cudaFree(dv);
        return rc;
    }
    cudaError_t e = launch_box_wireframe(tree->nodes, n, dv, di, ctx->stream);
    count_launch(ctx);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(host_vertices, dv, sizeof(float) * 24 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(host_indices, di, sizeof(uint32_t) * 24 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dv);
    cudaFree(di);
    if (e != cudaSuccess) return fail(OIBVH_ERR_CUDA, "box_wireframe: %s", cudaGetErrorString(e));
    return OIBVH_OK;
}

extern "C" int oibvh_scene_get_phase_cycles(oibvh_scene* scene, uint32_t* cycles, uint32_t max_phases,
                                            uint32_t* n_phases)
{
    REQUIRE(scene && cycles && n_phases, "NULL argument");
    int rc = oibvh_scene_get_counts(scene, nullptr, nullptr);
    if (rc) return rc;
    const uint32_t* h = scene->h_counters;
    uint32_t stamps = h[CTR_TIME0 - 1];
    if (stamps > (uint32_t)CTR_WORDS_TIME) stamps = CTR_WORDS_TIME;
    const uint32_t n = stamps ? std::min(max_phases, stamps - 1) : 0;
    for (uint32_t i = 0; i < n; i++) cycles[i] = h[CTR_TIME0 + i + 1] - h[CTR_TIME0 + i]; // wraps correctly
    *n_phases = n;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_device_counters(oibvh_scene* scene, const uint32_t** dev_counters)
{
    REQUIRE(scene && dev_counters, "NULL argument");
    *dev_counters = scene->counters;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_pair_capacity(oibvh_scene* scene, uint32_t* capacity)
{
    REQUIRE(scene && capacity, "NULL argument");
    *capacity = scene->pair_cap;
    return OIBVH_OK;
}

extern "C" int oibvh_scene_get_round_stats(oibvh_scene* scene, uint32_t* tested, uint32_t max_rounds,
                                           uint32_t* n_rounds)
{
    REQUIRE(scene && tested && n_rounds, "NULL argument");
    int rc = oibvh_scene_get_counts(scene, nullptr, nullptr);
    if (rc) return rc;
    const uint32_t n = std::min(max_rounds, (uint32_t)CTR_MAX_ROUNDS); // one entry per tree level of side A
    for (uint32_t r = 0; r < n; r++) tested[r] = scene->h_counters[CTR_FRONT0 + r];
    *n_rounds = n;
    return OIBVH_OK;
}
