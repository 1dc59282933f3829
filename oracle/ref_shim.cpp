// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Headless link shim + C wrapper around the UNMODIFIED reference CPU path
// (SimpleBVH + SimpleCollide + Mesh + triangleIntersect), compiled where the sources lie
// under /root/reference by oracle/Makefile (target `ref`) into oracle/_ref/liboibvh_ref.so.
// Nothing from the reference is copied into this repo: this file only *includes* reference
// headers at build time and provides
//   (1) no-op definitions for the 14 OpenGL entry points / 2 Shader setters the reference's
//       CPU classes reference (they own GL buffers in their constructors), and a host
//       implementation of Transform::transformVec4 (reference: src/cuda/transform.cu:18-40
//       runs `M * v` per vertex on the GPU; here the same glm expression on the host), and
//   (2) an extern "C" surface so tests can drive the reference classes through ctypes.
//
// Used for: pinning oracle/oibvh_oracle.c (the CPU restatement) and generating tests/golden/.

#include <cstdint>
#include <cstring>
#include <memory>
#include <queue>
#include <sstream>
#include <vector>

#include <glad/glad.h>

#define private public
#define protected public
#include "cpu/simpleBVH.h"
#include "cpu/simpleCollide.h"
#undef private
#undef protected
#include "utils/mesh.h"
#include "utils/utils.h"
#include "cuda/oibvh.cuh"

// ---------------------------------------------------------------------------------------------
// (1) link shim: GL entry points as no-ops
// ---------------------------------------------------------------------------------------------
static void APIENTRY nop_gen(GLsizei n, GLuint* ids)
{
    for (GLsizei i = 0; i < n; i++) ids[i] = 0;
}
static void APIENTRY nop_del(GLsizei, const GLuint*) {}
static void APIENTRY nop_bind1(GLuint) {}
static void APIENTRY nop_bind2(GLenum, GLuint) {}
static void APIENTRY nop_enum(GLenum) {}
static void APIENTRY nop_enum2(GLenum, GLenum) {}
static void APIENTRY nop_bufdata(GLenum, GLsizeiptr, const void*, GLenum) {}
static void APIENTRY nop_drawarrays(GLenum, GLint, GLsizei) {}
static void APIENTRY nop_drawelements(GLenum, GLsizei, GLenum, const void*) {}
static void APIENTRY nop_vap(GLuint, GLint, GLenum, GLboolean, GLsizei, const void*) {}

PFNGLGENVERTEXARRAYSPROC glad_glGenVertexArrays = nop_gen;
PFNGLGENBUFFERSPROC glad_glGenBuffers = nop_gen;
PFNGLDELETEVERTEXARRAYSPROC glad_glDeleteVertexArrays = nop_del;
PFNGLDELETEBUFFERSPROC glad_glDeleteBuffers = nop_del;
PFNGLBINDVERTEXARRAYPROC glad_glBindVertexArray = nop_bind1;
PFNGLENABLEVERTEXATTRIBARRAYPROC glad_glEnableVertexAttribArray = nop_bind1;
PFNGLBINDBUFFERPROC glad_glBindBuffer = nop_bind2;
PFNGLBINDTEXTUREPROC glad_glBindTexture = nop_bind2;
PFNGLACTIVETEXTUREPROC glad_glActiveTexture = nop_enum;
PFNGLPOLYGONMODEPROC glad_glPolygonMode = nop_enum2;
PFNGLBUFFERDATAPROC glad_glBufferData = nop_bufdata;
PFNGLDRAWARRAYSPROC glad_glDrawArrays = nop_drawarrays;
PFNGLDRAWELEMENTSPROC glad_glDrawElements = nop_drawelements;
PFNGLVERTEXATTRIBPOINTERPROC glad_glVertexAttribPointer = nop_vap;

void Shader::setInt(const std::string&, int) const {}
void Shader::setBool(const std::string&, bool) const {}

#ifndef REF_SHIM_DEVICE_TRANSFORM // the GPU build (oracle/ref_gpu_main.cu) links the reference's own transform.cu
Transform::Transform() : m_deviceVec4s(nullptr) {}
Transform::~Transform() {}
void Transform::transformVec4(std::vector<glm::vec4>& vec4s, const glm::mat4 transformMat)
{
    for (auto& v : vec4s) v = transformMat * v;
}
#endif

// ---------------------------------------------------------------------------------------------
// (2) C wrapper
// ---------------------------------------------------------------------------------------------
namespace
{
struct MuteCout
{
    std::ostringstream sink; // declared first: must be alive before its buffer is installed
    std::streambuf* old;
    MuteCout() : sink(), old(std::cout.rdbuf(sink.rdbuf())) {}
    ~MuteCout() { std::cout.rdbuf(old); }
};

struct RefMesh
{
    std::shared_ptr<Mesh> mesh;
};
struct RefBvh
{
    std::shared_ptr<SimpleBVH> bvh;
};
struct RefCollide
{
    std::unique_ptr<SimpleCollide> collide;
};
} // namespace

extern "C"
{

    // ---- Mesh (include/utils/mesh.h:71-187) ---------------------------------------------------
    void* ref_mesh_create(const float* pos, uint32_t V, const uint32_t* idx, uint32_t T)
    {
        std::vector<Vertex> verts(V);
        std::memset(verts.data(), 0, sizeof(Vertex) * V);
        for (uint32_t i = 0; i < V; i++) verts[i].m_position = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        std::vector<unsigned int> indices(idx, idx + 3 * (size_t)T);
        auto* h = new RefMesh;
        h->mesh = std::make_shared<Mesh>(verts, indices);
        return h;
    }
    void* ref_mesh_clone(void* m)
    {
        auto* h = new RefMesh;
        h->mesh = std::make_shared<Mesh>(*static_cast<RefMesh*>(m)->mesh);
        return h;
    }
    void ref_mesh_destroy(void* m) { delete static_cast<RefMesh*>(m); }
    void ref_mesh_translate(void* m, float x, float y, float z)
    {
        static_cast<RefMesh*>(m)->mesh->translate(glm::vec3(x, y, z));
    }
    void ref_mesh_rotate(void* m, float ax, float ay, float az, float angle_deg)
    {
        static_cast<RefMesh*>(m)->mesh->rotate(glm::vec3(ax, ay, az), angle_deg);
    }
    void ref_mesh_transform(void* m, const float* col_major16)
    {
        glm::mat4 M;
        std::memcpy(&M[0][0], col_major16, 16 * sizeof(float));
        static_cast<RefMesh*>(m)->mesh->transform(M);
    }
    void ref_mesh_set_positions(void* m, const float* pos)
    {
        auto& mesh = *static_cast<RefMesh*>(m)->mesh;
        for (uint32_t i = 0; i < mesh.m_verticesCount; i++)
            mesh.m_vertices[i].m_position = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    }
    void ref_mesh_get_positions(void* m, float* out)
    {
        auto& mesh = *static_cast<RefMesh*>(m)->mesh;
        for (uint32_t i = 0; i < mesh.m_verticesCount; i++)
        {
            out[3 * i] = mesh.m_vertices[i].m_position.x;
            out[3 * i + 1] = mesh.m_vertices[i].m_position.y;
            out[3 * i + 2] = mesh.m_vertices[i].m_position.z;
        }
    }
    void ref_mesh_get_aabb(void* m, float* out6)
    {
        auto& mesh = *static_cast<RefMesh*>(m)->mesh;
        std::memcpy(out6, &mesh.m_aabb.m_minimum, 12);
        std::memcpy(out6 + 3, &mesh.m_aabb.m_maximum, 12);
    }
    // last translation/rotation matrix is not kept by the reference; expose the glm builders instead
    void ref_glm_translate(float x, float y, float z, float* out16)
    {
        glm::mat4 M = glm::translate(glm::identity<glm::mat4>(), glm::vec3(x, y, z));
        std::memcpy(out16, &M[0][0], 64);
    }
    void ref_glm_rotate_about(const float* center3, float ax, float ay, float az, float angle_deg, float* out16)
    {
        // same sequence as Mesh::rotate (src/utils/mesh.cpp:171-178)
        glm::vec3 c(center3[0], center3[1], center3[2]);
        glm::mat4 M = glm::identity<glm::mat4>();
        M = glm::translate(M, c);
        M = glm::rotate(M, glm::radians(angle_deg), glm::vec3(ax, ay, az));
        M = glm::translate(M, -c);
        std::memcpy(out16, &M[0][0], 64);
    }
    void ref_mesh_get_center(void* m, float* out3)
    {
        auto& mesh = *static_cast<RefMesh*>(m)->mesh;
        std::memcpy(out3, &mesh.m_center, 12);
    }

    // ---- SimpleBVH (include/cpu/simpleBVH.h:20-54) ---------------------------------------------
    void* ref_bvh_create(void* m)
    {
        MuteCout mute;
        auto* h = new RefBvh;
        h->bvh = std::make_shared<SimpleBVH>(static_cast<RefMesh*>(m)->mesh);
        return h;
    }
    void ref_bvh_destroy(void* b) { delete static_cast<RefBvh*>(b); }
    void ref_bvh_build(void* b)
    {
        MuteCout mute;
        static_cast<RefBvh*>(b)->bvh->build();
    }
    void ref_bvh_refit(void* b)
    {
        auto& bvh = *static_cast<RefBvh*>(b)->bvh;
        bvh.unRefit();
        bvh.refit();
    }
    uint32_t ref_bvh_node_count(void* b) { return static_cast<RefBvh*>(b)->bvh->m_nodeCount; }
    uint32_t ref_bvh_depth(void* b) { return static_cast<RefBvh*>(b)->bvh->m_depth; }
    // BFS order == the order SimpleBVH::log writes (src/cpu/simpleBVH.cpp:56-79); out: N x 6 floats, tri: N ints
    uint32_t ref_bvh_dump_bfs(void* b, float* out_aabbs, int32_t* out_tri)
    {
        auto& bvh = *static_cast<RefBvh*>(b)->bvh;
        std::queue<std::shared_ptr<simple_bvh_node_t>> q;
        q.push(bvh.m_root);
        uint32_t n = 0;
        while (!q.empty())
        {
            auto node = q.front();
            q.pop();
            if (out_aabbs)
            {
                std::memcpy(out_aabbs + 6 * (size_t)n, &node->m_aabb.m_minimum, 12);
                std::memcpy(out_aabbs + 6 * (size_t)n + 3, &node->m_aabb.m_maximum, 12);
            }
            if (out_tri) out_tri[n] = node->m_triId;
            n++;
            if (node->m_left) q.push(node->m_left);
            if (node->m_right) q.push(node->m_right);
        }
        return n;
    }

    // ---- SimpleCollide (include/cpu/simpleCollide.h:8-52) --------------------------------------
    void* ref_collide_create()
    {
        auto* h = new RefCollide;
        h->collide.reset(new SimpleCollide());
        return h;
    }
    void ref_collide_destroy(void* c) { delete static_cast<RefCollide*>(c); }
    void ref_collide_add(void* c, void* b)
    {
        static_cast<RefCollide*>(c)->collide->addSimpleBVH(static_cast<RefBvh*>(b)->bvh);
    }
    uint32_t ref_collide_detect(void* c)
    {
        auto& col = *static_cast<RefCollide*>(c)->collide;
        col.detect(false);
        return col.getIntTriPairCount();
    }
    void ref_collide_get_pairs(void* c, uint32_t* out4)
    {
        auto& col = *static_cast<RefCollide*>(c)->collide;
        static_assert(sizeof(int_tri_pair_node_t) == 16, "pair record is 16 bytes");
        if (!col.m_intTriPairs.empty())
            std::memcpy(out4, col.m_intTriPairs.data(), col.m_intTriPairs.size() * sizeof(int_tri_pair_node_t));
    }

    // SimpleCollide::convertToVertexArray (src/cpu/simpleCollide.cpp:191-216): six vec3 per intersecting pair
    uint32_t ref_collide_vertex_array(void* c, float* out /* pairs * 18 */)
    {
        auto& col = *static_cast<RefCollide*>(c)->collide;
        col.convertToVertexArray();
        static_assert(sizeof(glm::vec3) == 12, "vec3 is three packed floats");
        if (out && !col.m_vertices.empty())
            std::memcpy(out, col.m_vertices.data(), col.m_vertices.size() * sizeof(glm::vec3));
        return (uint32_t)col.m_vertices.size();
    }
    // Node-box wireframes: the loop body of OibvhTree::convertToVertexArray (src/cuda/oibvhTree.cu:80-113; that file
    // is CUDA host code and is not compiled here) around the UNMODIFIED makeCube (src/utils/utils.cpp:15-70).
    void ref_box_wireframe(const float* nodes6, uint32_t n, float* verts /* n*24 */, uint32_t* idx /* n*24 */)
    {
        for (uint32_t i = 0; i < n; i++)
        {
            aabb_box_t aabb;
            aabb.m_minimum = glm::vec3(nodes6[6 * i], nodes6[6 * i + 1], nodes6[6 * i + 2]);
            aabb.m_maximum = glm::vec3(nodes6[6 * i + 3], nodes6[6 * i + 4], nodes6[6 * i + 5]);
            std::vector<glm::vec3> cubeVertices;
            std::vector<unsigned int> cubeIndices;
            makeCube(0.5f * (aabb.m_maximum.x - aabb.m_minimum.x), 0.5f * (aabb.m_maximum.y - aabb.m_minimum.y),
                     0.5 * (aabb.m_maximum.z - aabb.m_minimum.z), cubeVertices, cubeIndices);
            const glm::vec3 diff = aabb.m_minimum - cubeVertices[4];
            for (size_t k = 0; k < cubeVertices.size(); k++)
            {
                const glm::vec3 pos = cubeVertices[k] + diff;
                verts[24 * i + 3 * k] = pos.x;
                verts[24 * i + 3 * k + 1] = pos.y;
                verts[24 * i + 3 * k + 2] = pos.z;
            }
            for (size_t j = 0; j < cubeIndices.size(); j++) idx[24 * i + j] = cubeIndices[j] + (unsigned)cubeVertices.size() * i;
        }
    }

    // ---- free functions ------------------------------------------------------------------------
    // src/utils/utils.cpp:97-169
    int ref_triangle_intersect(const float* p /*3x3*/, const float* q /*3x3*/)
    {
        return triangleIntersect(glm::vec3(p[0], p[1], p[2]),
                                 glm::vec3(p[3], p[4], p[5]),
                                 glm::vec3(p[6], p[7], p[8]),
                                 glm::vec3(q[0], q[1], q[2]),
                                 glm::vec3(q[3], q[4], q[5]),
                                 glm::vec3(q[6], q[7], q[8]))
            ? 1
            : 0;
    }
    int ref_aabb_overlap(const float* a6, const float* b6)
    {
        aabb_box_t a, b;
        a.init(glm::vec3(a6[0], a6[1], a6[2]), glm::vec3(a6[3], a6[4], a6[5]));
        b.init(glm::vec3(b6[0], b6[1], b6[2]), glm::vec3(b6[3], b6[4], b6[5]));
        return a.overlap(b) ? 1 : 0;
    }
    // include/cuda/oibvh.cuh:56-182 host versions of the layout math
    uint32_t ref_oibvh_get_size(uint32_t t) { return oibvh_get_size(t); }
    uint32_t ref_oibvh_implicit_to_real(uint32_t i, uint32_t leafLev, uint32_t vl)
    {
        return oibvh_implicit_to_real(i, leafLev, vl);
    }
    uint32_t ref_oibvh_real_to_implicit(uint32_t r, uint32_t leafLev, uint32_t vl)
    {
        return oibvh_real_to_implicit(r, leafLev, vl);
    }
    int ref_oibvh_have_rchild(uint32_t i, uint32_t leafLev, uint32_t vl) { return oibvh_have_rchild(i, leafLev, vl); }
    uint32_t ref_oibvh_most_right_valid(uint32_t level, uint32_t leafLev, uint32_t vl)
    {
        return oibvh_get_most_right_valid_implicitIdx(level, leafLev, vl);
    }
    uint32_t ref_oibvh_level_real_count(uint32_t level, uint32_t leafLev, uint32_t vl)
    {
        return oibvh_level_real_node_count(level, leafLev, vl);
    }

} // extern "C"
