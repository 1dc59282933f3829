// Host-visible launch interface of the oibvh_b200 kernels (implemented in tree_kernels.cu / collide_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oibvh
{

struct MeshAabb
{
    float v[6]; // min xyz, max xyz (Mesh::m_aabb at construction, src/utils/mesh.cpp:91-98)
};
struct Mat4
{
    float m[16]; // column-major like glm::mat4
};

// ---- streaming radix sort (trees beyond the single-wave capacity): 30-bit keys, four 8-bit onesweep passes ----
constexpr int kRadixBits = 8;
constexpr int kRadixPasses = 4;
constexpr int kSortItemsPerThread = 16;
constexpr int kSortTile = (1 << kRadixBits) * kSortItemsPerThread;

// ---- build / refit ----
// boundary layout conversion: packed xyz <-> float4 (w = 1), packed index triples -> uint4 (w = 0)
cudaError_t launch_pack_positions(const float* xyz, float4* pos4, uint32_t V, cudaStream_t s);
cudaError_t launch_unpack_positions(const float4* pos4, float* xyz, uint32_t V, cudaStream_t s);
cudaError_t launch_pack_faces(const uint32_t* f3, uint4* faces4, uint32_t T, cudaStream_t s);
cudaError_t launch_morton_hist(const uint4* faces4, const float4* pos4, uint32_t T, const MeshAabb& mesh,
                               uint32_t* keys, uint32_t* hist, cudaStream_t s);
uint32_t onesweep_tiles(uint32_t T);
cudaError_t launch_onesweep_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                                 uint32_t* vals_out, uint32_t T, int pass, const uint32_t* hist, uint32_t* status,
                                 uint32_t* ticket, cudaStream_t s);
// Cooperative 3 x 10-bit LSD sort with ballot ranking (sort_lsd.cu): keys[i] sorted in place, permutation in vals[i];
// rec[i] = 2 T[i] scratch records; ctl = kMaxLsdJobs blocks of lsd_sort_ctl_words() words, zeroed once.
constexpr int kMaxLsdJobs = 4;
cudaError_t lsd_sort_configure();
uint32_t lsd_sort_capacity();
uint32_t lsd_sort_capacity_multi();
size_t lsd_sort_ctl_words();
// status: the context's device status word (bit 0 = a grid barrier of the sort timed out: the order is invalid)
cudaError_t launch_lsd_sort_many(uint32_t n, uint32_t* const* keys, uint32_t* const* vals, uint2* const* rec,
                                 const uint32_t* T, uint32_t* ctl, uint32_t* status, cudaStream_t s);
cudaError_t tree_emit_configure();
// arrival counters of the hierarchical top-of-tree completion (zeroed once; the kernel re-arms them)
size_t emit_counter_words(uint32_t T);
// status: the context's device status word (bit 1 = the finisher gave up waiting for a chunk)
cudaError_t launch_tree_emit(bool build, const uint4* faces_in4, const uint32_t* perm, uint32_t* faces_sorted,
                             const float4* pos4, float* nodes, uint32_t T, uint32_t* done_counter, uint32_t* status,
                             cudaStream_t s);
cudaError_t launch_transform(float4* pos4, uint32_t V, const Mat4& M, cudaStream_t s);

// ---- many small trees per launch (many-body scenes) ----
constexpr uint32_t kSmallTreeMax = 4096; // one CTA sorts and reduces the whole tree in shared memory / L2
struct SmallTreeDesc
{
    const uint4* faces_in; // T x (i0, i1, i2, 0), input order
    const float4* pos;
    uint32_t* faces; // T x 3, Morton order (written by a build, read by a refit)
    float* nodes;
    uint32_t* keys; // sorted keys / permutation, written by a build
    uint32_t* perm;
    MeshAabb mesh;
    uint32_t T, L;
};
// builds sort with a bitonic network up to kSmallBitonicSplit triangles and with a shared-memory radix sort above:
// descs[0, n_bitonic) must be the trees of the first kind, descs[n_bitonic, n) the others (refits do not care)
constexpr uint32_t kSmallBitonicSplit = 512;
cudaError_t small_trees_configure();
cudaError_t launch_small_trees(bool build, const SmallTreeDesc* descs, uint32_t n_bitonic, uint32_t n, cudaStream_t s);
struct XformDesc
{
    float4* pos;
    uint32_t V;
    uint32_t block0; // first 256-vertex block of this tree in the launch
};
// descs: n descriptors followed by total_blocks uint32 (the tree of every 256-vertex block)
cudaError_t launch_transform_many(const XformDesc* descs, uint32_t n, uint32_t total_blocks, const float* mats,
                                  cudaStream_t s);

// ---- collision ----
// one entry of the device object table (Scene::m_aabbOffsets/m_primOffsets/m_vertexOffsets/m_primCounts of the
// reference, src/cuda/scene.cu:95-129, become direct views of each tree's device buffers)
struct ObjDesc
{
    const float* nodes;    // N x 6 floats, real-index order
    const uint32_t* faces; // T x 3, Morton order
    const float4* pos;     // V x (x, y, z, 1)
    uint32_t T;
    uint32_t L; // leaf level = ceil(log2 T)
};

// BVTT node: (objA, objB, nodeA, nodeB) with node = level << kNodeLevelShift | position-in-level
constexpr int kNodeLevelShift = 26;
constexpr uint32_t kNodePosMask = (1u << kNodeLevelShift) - 1;

// a BVTT round descends at most this many levels per side (4^5 = 1024 combinations per node pair); a larger first-round
// value selects the dense level-k0 seeding of scenes with very few objects (collide_kernels.cu: dense_seed_phase)
constexpr uint32_t kMaxExpandLevels = 5;
constexpr uint32_t kMaxDenseSeedLevel = 11;

// counters block in device memory (uint32 each)
enum
{
    CTR_CANDIDATES = 0,
    CTR_PAIRS = 1,
    CTR_OVERFLOW = 2, // bit 0: work queue, bit 1: cut buffer, bit 2: pairs, bit 3: a wait timed out, bit 4: multi-GPU wait timed out,
                      // bit 5: candidate list
    CTR_BARRIER = 3,  // (unused since the seeding barrier was removed; kept so that the layout is stable)
    CTR_CUT = 4,      // records written by a recording detection (temporal coherence)
    CTR_CAND_TAIL = 5, // candidates (leaf pairs with overlapping boxes) appended to the candidate list
    CTR_FRONT0 = 8,   // CTR_FRONT0 + l = BVTT nodes processed whose side-A node is at tree level l (statistics)
    CTR_MAX_ROUNDS = 32,
    CTR_TIME0 = 64,   // CTR_TIME0 + i = SM cycle counter (low 32 bits) of CTA 0 at phase boundary i
    CTR_WORDS_TIME = 64,
    // work-queue state, each hot word on a 128-byte line of its own
    CTR_Q_STATE = 128, // 64-bit: finished items (low half) | pushed records (high half)
    CTR_Q_TAIL = 129,  // = the high half: pushed records, the next free slot
    CTR_Q_HEAD = 160,  // claimed slots
    CTR_Q_STOP = 192,  // traversal finished (or aborted)
    CTR_WORDS = 256
};

// ---- multi-GPU detection: replicated BVH, round 0 dealt to `world` ranks, and every rank's narrow phase
// appends its hits DIRECTLY to the gathering rank's pair list through a peer mapping (NVLink), so a frame needs no
// collective. Protocol words live in a small persistent block per scene (never zeroed per frame):
enum
{
    MG_FRAME = 0,  // detections completed on this scene since attach (written by the last CTA to leave)
    MG_OPEN = 1,   // root only: last frame whose counter block has been zeroed (remote ranks poll it before appending)
    MG_DONE = 2,   // root only: remote completions so far (monotonic: every remote rank adds 1 per frame)
    MG_EXIT = 3,   // CTAs that have left the current launch (the last one re-arms it)
    MG_FAIL = 4,   // sticky: a wait timed out
    MG_CUT = 8,    // temporal coherence: number of records of the recorded BVTT cut
    MG_WORDS = 64
};
struct MgpuArgs
{
    uint32_t mode;          // 0 = single GPU, 1 = gathering rank (root), 2 = remote rank
    uint32_t world;
    uint32_t* state;        // this scene's protocol block
    uint32_t* root_state;   // remote: the root's protocol block (peer mapping)
    uint32_t* root_counters; // remote: the root's counter block
    uint4* root_pairs;      // remote: the root's pair list
    uint32_t root_pair_cap;
};
// options of one detection
struct DetectOpts
{
    uint32_t mode;       // 0 = from the roots, 1 = from the roots + record the BVTT cut, 2 = from the recorded cut
    uint32_t self;       // also every object against itself (pairs of non-adjacent triangles of one mesh)
    uint4* cut;          // cut records: (objA, objB, level << 26 | pos, level << 26 | pos)
    uint32_t cut_cap;
    uint32_t cut_depth;  // the cut lies this many levels above the leaves
    uint32_t* cut_state; // persistent: [0] = number of records of the recorded cut
    uint4* cand;         // candidate list: (objA, objB, triA, triB), tested by the narrow phase after the traversal
    uint32_t cand_cap;
};
// root: zero the counter block of the coming frame, then publish MG_OPEN = frame (system scope)
cudaError_t launch_mgpu_open(uint32_t* counters, uint32_t* state, cudaStream_t s);

// Whole detection in one cooperative launch: seeds (one root pair per object pair i<j, or every node pair of level
// levels0 when levels0 > kMaxExpandLevels) -> queue-driven traversal (root pairs descend levels0 levels, the others
// `levels`; rank/world shard the children of the root pairs) with the narrow phase beside it (opt.cand = candidate list). `queue` must hold the
// empty marker (all bits set) in every slot -- the kernel leaves it so --, counters must be zeroed before the launch.
// grid_blocks comes from collide_configure().
cudaError_t collide_configure(int* grid_blocks);
cudaError_t launch_collide(int grid_blocks, const ObjDesc* objs, uint32_t n_obj, uint4* queue, uint32_t queue_cap,
                           uint4* pairs, uint32_t pair_cap, uint32_t* counters, uint32_t levels0, uint32_t levels,
                           uint32_t rank, uint32_t world, const MgpuArgs& mg, const DetectOpts& opt, cudaStream_t s);

// collided-triangle vertex stream (Scene::convertToVertexArray) and node-box wireframes
// (OibvhTree::convertToVertexArray) as device-side gathers
cudaError_t launch_pair_vertices(const ObjDesc* objs, const uint4* pairs, uint32_t pair_cap, const uint32_t* counters,
                                 float* out, uint32_t out_cap_pairs, cudaStream_t s);
cudaError_t launch_box_wireframe(const float* nodes, uint32_t n, float* verts, uint32_t* idx, cudaStream_t s);

} // namespace oibvh
