/*
 * oibvh_b200 -- C ABI of the B200-native oibvh collision path (build, refit, broad phase, narrow phase).
 *
 * This is the drop-in boundary: plain pointers and sizes, opaque handles, int status codes. The reference
 * (hhhcbw/oibvh) has no FFI layer -- its host classes launch kernels inline -- so each entry point below names
 * the reference host method it replaces (paths relative to the reference checkout). The C++ facade in
 * include/oibvh/ (oibvh.hpp, model.hpp) re-creates the reference's classes (Mesh, OibvhTree, Scene, DeviceType, aabb_box_t,
 * int_tri_pair_node_t) on top of exactly these calls; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns OIBVH_OK (0) or a negative oibvh_status; oibvh_last_error() gives the text of the
 *     last failure on the calling thread. No exceptions cross the boundary.
 *   - a context binds one CUDA device and one stream. Handles are not thread-safe; distinct contexts may be
 *     used from distinct threads.
 *   - "host" pointers are caller-owned host memory (pageable or pinned); "dev" pointers are device memory on the
 *     context's device. Nothing is retained after the call returns unless stated.
 *   - build/refit/transform/detect_async only ENQUEUE work on the context stream. Calls that return data to
 *     the host (download*, detect, get_counts, get_pairs) synchronise the stream first.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with OIBVH_ERR_CUDA.
 *
 * Record layouts (bit-compatible with the reference, include/utils/utils.h:13-61):
 *   oibvh_aabb          24 B  { float min[3]; float max[3]; }                     == aabb_box_t
 *   oibvh_int_tri_pair  16 B  { uint32 bvh_index[2]; uint32 tri_index[2]; }       == int_tri_pair_node_t
 * tri_index values are positions in the tree's Morton-sorted face array, exactly like the reference GPU path
 * (src/cuda/collide.cu:296-297); oibvh_tree_download(perm) maps them back to input face ids.
 */
#ifndef OIBVH_B200_H
#define OIBVH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oibvh_ctx oibvh_ctx;
typedef struct oibvh_tree oibvh_tree;
typedef struct oibvh_scene oibvh_scene;
typedef struct oibvh_graph oibvh_graph;

typedef struct oibvh_aabb
{
    float min[3];
    float max[3];
} oibvh_aabb;

typedef struct oibvh_int_tri_pair
{
    uint32_t bvh_index[2];
    uint32_t tri_index[2];
} oibvh_int_tri_pair;

typedef enum oibvh_status
{
    OIBVH_OK = 0,
    OIBVH_ERR_INVALID = -1,  /* bad argument / wrong state (e.g. scene_add_tree before build) */
    OIBVH_ERR_CUDA = -2,     /* a CUDA runtime call or kernel failed (includes "no device") */
    OIBVH_ERR_NOMEM = -3,    /* device or host allocation failed */
    OIBVH_ERR_OVERFLOW = -4, /* a work queue could not be grown enough to hold the BVTT front / pair list */
    OIBVH_ERR_INTERNAL = -5
} oibvh_status;

/* stage ids for oibvh_ctx_stage_ms */
enum
{
    OIBVH_STAGE_BUILD = 0,
    OIBVH_STAGE_REFIT = 1,
    OIBVH_STAGE_BROAD = 2,
    OIBVH_STAGE_NARROW = 3,
    /* kernels inside the stages above (the per-kernel roofline table of bench.py): */
    OIBVH_STAGE_KEYS = 4,      /* build: Morton keys */
    OIBVH_STAGE_SORT = 5,      /* build: stable radix sort of (key, face id) */
    OIBVH_STAGE_EMIT = 6,      /* build: face gather + leaf boxes + bottom-up reduction */
    OIBVH_STAGE_TRANSFORM = 7, /* oibvh_tree_transform[_many] */
    OIBVH_STAGE_COUNT = 8
};

const char* oibvh_last_error(void);
/* library/ABI version: major*10000 + minor*100 + patch */
int oibvh_version(void);
/* number of CUDA devices visible (0 without a GPU); never fails */
int oibvh_device_count(void);

/* ---- context: replaces the implicit "current device, default stream" of the reference
 *      (Scene::detectCollisionOnGPU cudaSetDevice, src/cuda/scene.cu:229) --------------------------------- */
int oibvh_ctx_create(int device, oibvh_ctx** out);
/* same, but enqueue on a caller-owned cudaStream_t (e.g. the framework's current stream); 0 = legacy default */
int oibvh_ctx_create_on_stream(int device, void* cuda_stream, oibvh_ctx** out);
int oibvh_ctx_destroy(oibvh_ctx* ctx);
int oibvh_ctx_synchronize(oibvh_ctx* ctx);
int oibvh_ctx_get_stream(oibvh_ctx* ctx, void** cuda_stream);
/* number of kernel launches this context has enqueued so far (graph replays count their kernel nodes) */
int oibvh_ctx_launch_count(oibvh_ctx* ctx, uint64_t* launches);
/* enable per-stage CUDA-event timing (adds event records, no host syncs); read the last values with stage_ms.
 * Replaces the reference's kernelLaunch() stopwatch (include/cuda/utils.cuh:36-59), without its per-launch sync. */
int oibvh_ctx_enable_timing(oibvh_ctx* ctx, int enable);
int oibvh_ctx_stage_ms(oibvh_ctx* ctx, float ms[OIBVH_STAGE_COUNT]);

/* ---- whole-frame CUDA graphs: everything enqueued on this context between begin and end (transform, build, refit,
 *      detect_async) is recorded instead of executed and can then be replayed with one launch. The reference has
 *      no counterpart (it synchronises after every kernel). Handles used inside must have run once before (buffers
 *      allocated); a graph becomes stale -- launch returns OIBVH_ERR_INVALID -- if a work queue is regrown later. */
int oibvh_ctx_capture_begin(oibvh_ctx* ctx);
int oibvh_ctx_capture_end(oibvh_ctx* ctx, oibvh_graph** out);
/* abandon a capture after a failed call (the stream would otherwise stay in capture mode); no graph is produced */
int oibvh_ctx_capture_abort(oibvh_ctx* ctx);
int oibvh_graph_launch(oibvh_graph* graph);
int oibvh_graph_destroy(oibvh_graph* graph);

/* ---- tree: OibvhTree (include/cuda/oibvhTree.cuh:44-93) ------------------------------------------------- */
/* OibvhTree(mesh) + setup() (src/cuda/oibvhTree.cu:9-15, 157-174): takes packed xyz positions [V*3], triangle
 * indices [T*3] and the mesh AABB computed at Mesh construction (src/utils/mesh.cpp:91-98; min xyz, max xyz),
 * uploads them once. T >= 2 (the reference underflows for T = 1, oibvhTree.cu:317-322). */
int oibvh_tree_create(oibvh_ctx* ctx, const float* host_positions, uint32_t V, const uint32_t* host_indices,
                      uint32_t T, const float mesh_aabb[6], oibvh_tree** out);
/* same with inputs already resident in device memory (copied device-to-device, the caller keeps its buffers) */
int oibvh_tree_create_from_device(oibvh_ctx* ctx, const float* dev_positions, uint32_t V,
                                  const uint32_t* dev_indices, uint32_t T, const float mesh_aabb[6],
                                  oibvh_tree** out);
/* OibvhTree(other, mesh) copy constructor (src/cuda/oibvhTree.cu:17-33): same sorted faces / permutation /
 * node array / positions as `other`, sharing nothing on the device. */
int oibvh_tree_clone(const oibvh_tree* other, oibvh_tree** out);
/* A replica of `src` on ANOTHER context / device of this process (the multi-GPU detection keeps every tree on every
 * GPU; it is also what lets the facade honour Scene::detectCollision(DeviceType::GPUk) for any k,
 * include/cuda/scene.cuh:12-24): same sorted faces, permutation, positions and node array, copied device to device
 * (peer copy over NVLink when the devices differ). oibvh_tree_sync_replica refreshes positions and nodes (and the
 * face order, if `src` has been rebuilt) after `src` changed: cheaper than a refit from host data and bit-identical
 * by construction. Both calls wait for `src`'s pending work. */
int oibvh_tree_replicate(const oibvh_tree* src, oibvh_ctx* dst_ctx, oibvh_tree** out);
int oibvh_tree_sync_replica(oibvh_tree* replica, const oibvh_tree* src);
int oibvh_tree_destroy(oibvh_tree* tree);
/* the "copy Mesh positions" head of OibvhTree::refit (src/cuda/oibvhTree.cu:196-199, 208) */
int oibvh_tree_set_positions(oibvh_tree* tree, const float* host_positions);
int oibvh_tree_set_positions_from_device(oibvh_tree* tree, const float* dev_positions);
/* Mesh::transform / transform_vec4_kernel (src/utils/mesh.cpp:187-213, src/cuda/transform.cu:18-40) applied to the
 * device-resident positions: p = M * (p, 1), column-major M, glm operation order, no FMA contraction. */
int oibvh_tree_transform(oibvh_tree* tree, const float M[16]);
/* OibvhTree::build (src/cuda/oibvhTree.cu:237-388): Morton keys, stable sort, implicit-layout AABB reduction */
int oibvh_tree_build(oibvh_tree* tree);
/* Build several trees of the same context together (extension; the reference builds one tree per call, one
 * OibvhTree::build per object of a many-body scene). Results are identical to separate builds.
 *  - trees of up to 4096 triangles: ALL of them in one launch, one thread block per tree;
 *  - larger trees: the keys of 2..4 trees are sorted by ONE cooperative launch (the grid barriers of a single sort);
 *    otherwise consecutive oibvh_tree_build calls.
 * The device table naming the trees is kept while the same list is passed again (no upload on later calls; only such
 * repeat calls may be captured into a graph). */
int oibvh_tree_build_many(oibvh_tree* const* trees, uint32_t n);
/* OibvhTree::refit (src/cuda/oibvhTree.cu:193-235) on the positions currently on the device */
int oibvh_tree_refit(oibvh_tree* tree);
/* the per-object refit loop of a many-body frame: small trees in one launch, larger ones one launch each */
int oibvh_tree_refit_many(oibvh_tree* const* trees, uint32_t n);
/* Rigid transform + refit of several trees in one call: host_mats = n x 16 floats (column-major), apply[i] != 0 selects
 * the trees that are transformed (NULL = all); every tree is then refitted. The transform of one tree runs under the
 * refit of another (two streams) instead of in front of both; capturable with host matrices (they travel as kernel
 * arguments). Same arithmetic and bits as oibvh_tree_transform followed by oibvh_tree_refit_many. */
int oibvh_tree_transform_refit_many(oibvh_tree* const* trees, uint32_t n, const float* host_mats,
                                    const unsigned char* apply);
/* oibvh_tree_transform for n trees in one launch; mats = n column-major 4x4 matrices (host / device memory) */
int oibvh_tree_transform_many(oibvh_tree* const* trees, uint32_t n, const float* host_mats);
int oibvh_tree_transform_many_from_device(oibvh_tree* const* trees, uint32_t n, const float* dev_mats);
/* getPrimCount / vertex count / oibvh_get_size(T) / getDepth() = ilog2(N)  (oibvhTree.cu:45-53) */
int oibvh_tree_get_info(const oibvh_tree* tree, uint32_t* T, uint32_t* V, uint32_t* N, uint32_t* depth);
int oibvh_tree_is_built(const oibvh_tree* tree, int* built);
/* results the reference copies back after every build/refit (oibvhTree.cu:232, 375-377): nodes [N] in real-index
 * order, sorted faces [T*3]; plus perm [T] (sorted position -> input face id), which the reference discards.
 * Any pointer may be NULL. Synchronises. */
int oibvh_tree_download(oibvh_tree* tree, oibvh_aabb* host_nodes, uint32_t* host_sorted_faces, uint32_t* host_perm);
int oibvh_tree_download_positions(oibvh_tree* tree, float* host_positions);
/* sorted 30-bit Morton keys [T] of the last build (debug / parity) */
int oibvh_tree_download_keys(oibvh_tree* tree, uint32_t* host_sorted_keys);
/* device views for zero-copy consumers (valid until the tree is destroyed). Positions are stored on the device as
 * 16-byte records (x, y, z, 1) -- V x 4 floats -- so that a gathered vertex is one 128-bit load. */
int oibvh_tree_device_views(oibvh_tree* tree, const oibvh_aabb** dev_nodes, const uint32_t** dev_sorted_faces,
                            const float** dev_positions);

/* ---- scene: Scene (include/cuda/scene.cuh:26-67) -------------------------------------------------------- */
int oibvh_scene_create(oibvh_ctx* ctx, oibvh_scene** out);
int oibvh_scene_destroy(oibvh_scene* scene);
/* Scene::addOibvhTree (src/cuda/scene.cu:95-129); the tree must be built; bvh index = insertion order.
 * Trees are referenced, not copied: refits between detections are honoured like in the reference. */
int oibvh_scene_add_tree(oibvh_scene* scene, oibvh_tree* tree);
/* restrict this scene to shard `rank` of `world` (seed BVTT nodes are dealt round-robin); default 0 of 1 */
int oibvh_scene_set_shard(oibvh_scene* scene, uint32_t rank, uint32_t world);
/* ---- opt-in extensions beyond the reference's behaviour ----------------------------------
 * Self-collision: the reference only tests object pairs i < j (src/cuda/scene.cu:192-223). With it enabled every
 * object is also tested against itself: records (i, i, triA, triB) with triA < triB for the intersecting triangle
 * pairs of object i that share NO vertex index (triangles with a common vertex always touch there: neighbours, not
 * collisions). Same overlap and SAT rules as between objects.
 * Temporal coherence: the reference restarts the traversal from entryLevel on every detectCollision
 * (src/cuda/scene.cu:236). A coherent scene records, during one detection, a complete cut of the BVTT `cut_depth`
 * levels above the leaves (0 = default 6) and starts the following detections from it, as long as the trees are only
 * refitted / transformed (any build, added tree or changed shard records a new cut). The pair set is exactly that of
 * a detection from the roots, for any motion; the cut buffer is sized by `candidate_records` of oibvh_scene_reserve
 * (like the candidate list).
 * Scenes of more than 4096 object pairs ignore the setting. */
int oibvh_scene_set_self_collision(oibvh_scene* scene, int enable);
int oibvh_scene_set_coherence(oibvh_scene* scene, int enable, uint32_t cut_depth);

/* size the work queues up front (records): BVTT front (x2), candidate list, pair list. Queues otherwise start at
 * 2^20 / 2^20 / 2^19 records and are regrown after an overflow; a multi-GPU scene cannot regrow (its pair list is
 * mapped by the other ranks), so reserve before oibvh_mgpu_export. Synchronises; invalidates captured graphs. */
int oibvh_scene_reserve(oibvh_scene* scene, uint32_t front_records, uint32_t candidate_records, uint32_t pair_records);

/* ---- multi-GPU detection: the reference's DeviceType::GPU0..GPU8 + cudaSetDevice stub (include/cuda/scene.cuh:12-24,
 *      src/cuda/scene.cu:229), redesigned: one scene per GPU over the SAME trees (every rank builds /
 *      refits its own replica), oibvh_scene_set_shard(rank, world) deals the BVTT front of round 0 to the ranks, and
 *      every rank's narrow phase appends its hits DIRECTLY to the gathering rank's pair list through a peer mapping
 *      (NVLink): no collective and no host round trip per frame. Ranks may be threads of one process or one process
 *      per GPU (CUDA IPC); the 160-byte handle is plain bytes the caller moves between them (MPI, a socket,
 *      torch.distributed ...).
 *        root:    oibvh_scene_reserve(..); oibvh_scene_set_shard(s, 0, W); oibvh_mgpu_export(s, &h);  -> send h
 *        others:  oibvh_scene_set_shard(s, r, W); oibvh_mgpu_attach(s, &h);
 *        frame:   every rank calls oibvh_scene_detect_async (or replays its graph) the SAME number of times. The
 *                 root's detection completes only after every rank's hits of that frame have landed, so its
 *                 get_counts / get_pairs / device_pairs return the gathered list; on the other ranks they report
 *                 that rank's own counts and get_pairs fails (the list lives on the root).
 *      A rank that never launches its frame makes the root's wait time out: OIBVH_ERR_INTERNAL from get_counts. */
typedef struct oibvh_mgpu_handle
{
    unsigned char bytes[160];
} oibvh_mgpu_handle;
int oibvh_mgpu_export(oibvh_scene* root_scene, oibvh_mgpu_handle* out);
int oibvh_mgpu_attach(oibvh_scene* scene, const oibvh_mgpu_handle* root);
/* leave multi-GPU mode (unmaps the root's buffers); every rank, after the last frame has completed everywhere */
int oibvh_mgpu_detach(oibvh_scene* scene);
/* root only, optional: open the next frame (zero the counter block, let the other ranks append) ahead of the root's
 * own detection -- lets a single device play several ranks one after the other (tests) */
int oibvh_mgpu_open_frame(oibvh_scene* root_scene);

/* Scene::detectCollision(GPUx, entryLevel, expandLevels) (src/cuda/scene.cu:157-185, 225-446): broad + narrow phase
 * over all object pairs i<j. entry_level / expand_levels keep the reference meaning (seed level; levels descended
 * per round); expand_levels = 0 lets the library choose the schedule (3 levels per round, round 0 absorbs the
 * remainder, entry_level ignored). The resulting pair SET does not depend on either value.
 * n_candidates may be NULL. Synchronises. */
int oibvh_scene_detect(oibvh_scene* scene, uint32_t entry_level, uint32_t expand_levels, uint32_t* n_pairs,
                       uint32_t* n_candidates);
/* enqueue only; read the result with oibvh_scene_get_counts */
int oibvh_scene_detect_async(oibvh_scene* scene, uint32_t entry_level, uint32_t expand_levels);
int oibvh_scene_get_counts(oibvh_scene* scene, uint32_t* n_pairs, uint32_t* n_candidates);
/* Scene::m_intTriPairs (scene.cu:414-416): n_pairs records, order unspecified (as in the reference) */
int oibvh_scene_get_pairs(oibvh_scene* scene, oibvh_int_tri_pair* host_pairs);
/* device view of the pair list of the last detection (for a collective gather by the caller) */
int oibvh_scene_device_pairs(oibvh_scene* scene, const oibvh_int_tri_pair** dev_pairs, uint32_t* n_pairs);
/* SM-clock cycles CTA 0 of the detection kernel spent in each phase of the last detection: [0] seeding, [1] the
 * queue-driven traversal (with the narrow warps testing candidates beside it), [2] the rest of the narrow phase,
 * [3] leaving the queue empty.
 * Returns the number of phases written. Replaces the reference's per-launch stopwatch (scene.cu:299-302, 419). */
int oibvh_scene_get_phase_cycles(oibvh_scene* scene, uint32_t* cycles, uint32_t max_phases, uint32_t* n_phases);
/* device view of the counter block of the last detection: word 0 = candidates, word 1 = pairs (lets a caller chain
 * a collective on the stream without a host round trip). Does not synchronise. After the first detection the block
 * (512 bytes = 32 records) is the head of the pair-list allocation: the pair list returned by
 * oibvh_scene_device_pairs starts exactly 512 bytes behind it, so ONE collective over
 * [counters, counters + 512 + n * 16) moves the counts and the first n pair records together. Both pointers change
 * when the queues are re-allocated after an overflow (query them again after oibvh_scene_get_counts). */
int oibvh_scene_device_counters(oibvh_scene* scene, const uint32_t** dev_counters);
/* current capacity (records) of the device pair list returned by oibvh_scene_device_pairs */
int oibvh_scene_pair_capacity(oibvh_scene* scene, uint32_t* capacity);
/* BVTT statistics of the last detection: tested[l] = BVTT nodes taken from the work queue whose side-A node is at
 * tree level l (the traversal has no rounds: it is queue-driven). Returns the number of levels written (<= max_rounds). */
int oibvh_scene_get_round_stats(oibvh_scene* scene, uint32_t* tested, uint32_t max_rounds, uint32_t* n_rounds);

/* Scene::convertToVertexArray (src/cuda/scene.cu:68-93) as a device-side gather: pair i of the last detection
 * contributes six packed float3 (A.v0, A.v1, A.v2, B.v0, B.v1, B.v2) = 18 floats, in the order of the pair list.
 * _device: enqueue only, into a caller-provided DEVICE buffer of capacity_pairs * 18 floats; the pair count is read
 * on the device (min(n_pairs, capacity_pairs) records are written), so the call can follow oibvh_scene_detect_async
 * on the stream or sit inside a captured frame graph -- a renderer consumes the result without any D2H of pair lists.
 * The host variant synchronises and writes n_pairs * 18 floats. */
int oibvh_scene_pair_vertices_device(oibvh_scene* scene, float* dev_vertices, uint32_t capacity_pairs);
int oibvh_scene_pair_vertices(oibvh_scene* scene, float* host_vertices);
/* OibvhTree::convertToVertexArray (src/cuda/oibvhTree.cu:69-124, makeCube src/utils/utils.cpp:15-70): wireframe
 * boxes of the first min(internal nodes, max_nodes) nodes (the reference draws 256): 8 corners (24 floats) and 12
 * edges (24 indices, offset by 8 per box) each, corner arithmetic as in the reference. Host buffers; *n_boxes
 * receives the number of boxes written. Synchronises. */
int oibvh_tree_box_wireframe(oibvh_tree* tree, uint32_t max_nodes, float* host_vertices, uint32_t* host_indices,
                             uint32_t* n_boxes);

#ifdef __cplusplus
}
#endif
#endif /* OIBVH_B200_H */
