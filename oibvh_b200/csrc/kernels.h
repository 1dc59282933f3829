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

// ---- radix sort configuration: 30-bit keys ----
constexpr int kRadixBits = 8;
constexpr int kRadixPasses = 4;
constexpr int kSortItemsPerThread = 16;
constexpr int kSortTile = (1 << kRadixBits) * kSortItemsPerThread;

// ---- build / refit ----
cudaError_t launch_morton_hist(const uint32_t* faces, const float* pos, uint32_t T, const MeshAabb& mesh,
                               uint32_t* keys, uint32_t* hist, cudaStream_t s);
uint32_t onesweep_tiles(uint32_t T);
cudaError_t launch_onesweep_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                                 uint32_t* vals_out, uint32_t T, int pass, const uint32_t* hist, uint32_t* status,
                                 uint32_t* ticket, cudaStream_t s);
cudaError_t tree_emit_configure();
cudaError_t launch_tree_emit(bool build, const uint32_t* faces_in, const uint32_t* perm, uint32_t* faces_sorted,
                             const float* pos, float* nodes, uint32_t T, uint32_t* done_counter, cudaStream_t s);
cudaError_t launch_transform(float* pos, uint32_t V, const Mat4& M, cudaStream_t s);

// ---- collision ----
// one entry of the device object table (Scene::m_aabbOffsets/m_primOffsets/m_vertexOffsets/m_primCounts of the
// reference, src/cuda/scene.cu:95-129, become direct views of each tree's device buffers)
struct ObjDesc
{
    const float* nodes;    // N x 6 floats, real-index order
    const uint32_t* faces; // T x 3, Morton order
    const float* pos;      // V x 3
    uint32_t T;
    uint32_t L; // leaf level = ceil(log2 T)
};

// BVTT node: (objA, objB, nodeA, nodeB) with node = level << kNodeLevelShift | position-in-level
constexpr int kNodeLevelShift = 26;
constexpr uint32_t kNodePosMask = (1u << kNodeLevelShift) - 1;

// counters block in device memory (uint32 each)
enum
{
    CTR_CANDIDATES = 0,
    CTR_PAIRS = 1,
    CTR_OVERFLOW = 2, // bit 0: front, bit 1: candidates, bit 2: pairs
    CTR_FRONT0 = 8,   // CTR_FRONT0 + r = size of the front consumed by round r
    CTR_MAX_ROUNDS = 48,
    CTR_WORDS = CTR_FRONT0 + CTR_MAX_ROUNDS + 8
};

// front[p] = (i, j, root, root) for the p-th object pair i < j; counters[CTR_FRONT0] = number of pairs
cudaError_t launch_seed(uint32_t n_obj, uint4* front, uint32_t front_cap, uint32_t* counters, cudaStream_t s);
// round `round`: consumes counters[CTR_FRONT0 + round] nodes of `in`, descends `levels` levels per side.
// rank/world shard the children of round 0. grid_hint = expected front size (sizes the persistent grid).
cudaError_t launch_expand(const ObjDesc* objs, const uint4* in, uint4* out, uint32_t front_cap, uint4* cand,
                          uint32_t cand_cap, uint32_t* counters, uint32_t round, uint32_t levels, uint32_t rank,
                          uint32_t world, uint32_t grid_hint, cudaStream_t s);
cudaError_t launch_narrow(const ObjDesc* objs, const uint4* cand, uint32_t cand_cap, uint4* pairs,
                          uint32_t pair_cap, uint32_t* counters, uint32_t grid_hint, cudaStream_t s);

} // namespace oibvh
