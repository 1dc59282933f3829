// C++ facade of oibvh_b200: the reference's host operator surface on top of the C ABI (include/oibvh_b200.h).
//
// Same class / method / member names as the reference so that its driver code (src/main.cpp:127-151 set-up,
// :208-319 per-frame loop, minus the GL/ImGui calls) compiles against this header:
//
//   aabb_box_t, tri_pair_node_t, int_tri_pair_node_t      include/utils/utils.h:13-61
//   Vertex, Mesh{translate, rotate*, transform, m_vertices, m_indices, m_aabb, ...}   include/utils/mesh.h:20-187
//   OibvhTree{OibvhTree(mesh), OibvhTree(other, mesh), build, refit, getDepth, getPrimCount}   include/cuda/oibvhTree.cuh:44-93
//   DeviceType, Scene{addOibvhTree, detectCollision, getIntTriPairCount}              include/cuda/scene.cuh:12-67
//
// Differences, all deliberate:
//   * no OpenGL: draw()/VAO/VBO members are gone (rendering is out of scope);
//   * trees live on the device; m_aabbTree / m_faces / m_positions are refreshed from the device by syncHost()
//     (the reference copies them back after every build/refit, oibvhTree.cu:232, 375-377) -- call it when you read them;
//   * failures throw std::runtime_error carrying oibvh_last_error() (the reference checks nothing);
//   * vector/matrix types: define OIBVH_FACADE_USE_GLM before including to use glm::vec3/mat4 (drop-in inside the
//     reference tree); otherwise the minimal oibvh::vec types below, which follow glm's arithmetic order exactly.
// Header-only; link with -loibvh_b200.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../oibvh_b200.h"

#ifdef OIBVH_FACADE_USE_GLM
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
namespace oibvh_math = glm;
#else
namespace oibvh_math
{
struct vec2
{
    float x = 0, y = 0;
};
struct vec3
{
    float x = 0, y = 0, z = 0;
    vec3() = default;
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct uvec3
{
    unsigned int x = 0, y = 0, z = 0;
};
struct vec4
{
    float x = 0, y = 0, z = 0, w = 0;
};
// column-major 4x4 like glm::mat4: c[j] is column j
struct mat4
{
    vec4 c[4];
    vec4& operator[](int j) { return c[j]; }
    const vec4& operator[](int j) const { return c[j]; }
};
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator*(const vec4& a, float s) { return vec4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec3 min(const vec3& x, const vec3& y) // glm::min(x, y) = (y < x) ? y : x per component
{
    return vec3(y.x < x.x ? y.x : x.x, y.y < x.y ? y.y : x.y, y.z < x.z ? y.z : x.z);
}
inline vec3 max(const vec3& x, const vec3& y) // glm::max(x, y) = (x < y) ? y : x
{
    return vec3(x.x < y.x ? y.x : x.x, x.y < y.y ? y.y : x.y, x.z < y.z ? y.z : x.z);
}
inline mat4 identity()
{
    mat4 m;
    m[0].x = m[1].y = m[2].z = m[3].w = 1.0f;
    return m;
}
inline float radians(float deg) { return deg * static_cast<float>(0.01745329251994329576923690768489); }
// glm::translate (third/glm/ext/matrix_transform.inl:10-15)
inline mat4 translate(const mat4& m, const vec3& v)
{
    mat4 r = m;
    r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return r;
}
// glm::rotate (third/glm/ext/matrix_transform.inl:18-46)
inline mat4 rotate(const mat4& m, float angle, const vec3& v)
{
    const float c = std::cos(angle), s = std::sin(angle);
    const float inv = 1.0f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    const vec3 axis(v.x * inv, v.y * inv, v.z * inv);
    const vec3 temp((1.0f - c) * axis.x, (1.0f - c) * axis.y, (1.0f - c) * axis.z);
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    mat4 r;
    r[0] = m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2];
    r[1] = m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2];
    r[2] = m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2];
    r[3] = m[3];
    return r;
}
// glm mat4 * vec4 (third/glm/detail/type_mat4x4.inl:561-572): (m0*x + m1*y) + (m2*z + m3*w)
inline vec4 operator*(const mat4& m, const vec4& v) { return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w); }
} // namespace oibvh_math
#endif

// ---------------------------------------------------------------------------------------------------------
// plain records (include/utils/utils.h:13-61)
// ---------------------------------------------------------------------------------------------------------
typedef struct aabb_box
{
    oibvh_math::vec3 m_minimum{std::numeric_limits<float>::max()};
    oibvh_math::vec3 m_maximum{-std::numeric_limits<float>::max()};
    void merge(const aabb_box& o)
    {
        m_minimum = oibvh_math::min(m_minimum, o.m_minimum);
        m_maximum = oibvh_math::max(m_maximum, o.m_maximum);
    }
    bool overlap(const aabb_box& o) const
    {
        return (m_minimum.x <= o.m_maximum.x && m_maximum.x >= o.m_minimum.x) &&
               (m_minimum.y <= o.m_maximum.y && m_maximum.y >= o.m_minimum.y) &&
               (m_minimum.z <= o.m_maximum.z && m_maximum.z >= o.m_minimum.z);
    }
} aabb_box_t;
static_assert(sizeof(aabb_box_t) == sizeof(oibvh_aabb), "aabb_box_t must stay 24 bytes");

typedef struct tri_pair_node
{
    unsigned int m_triIndex[2];
} tri_pair_node_t;

typedef struct int_tri_pair_node
{
    unsigned int m_bvhIndex[2];
    unsigned int m_triIndex[2];
} int_tri_pair_node_t;
static_assert(sizeof(int_tri_pair_node_t) == sizeof(oibvh_int_tri_pair), "int_tri_pair_node_t must stay 16 bytes");

namespace oibvh_detail
{
// glm has no non-template identity(): glm::mat4(1.0f) there, the mini type's identity() otherwise
inline oibvh_math::mat4 mat4_identity()
{
#ifdef OIBVH_FACADE_USE_GLM
    return oibvh_math::mat4(1.0f);
#else
    return oibvh_math::identity();
#endif
}
inline void check(int rc)
{
    if (rc != OIBVH_OK) throw std::runtime_error(std::string("oibvh_b200: ") + oibvh_last_error());
}
// one context per device, created on first use (the reference uses the current device + default stream). The table
// is guarded by a mutex; a context itself is NOT thread-safe (include/oibvh_b200.h): threads that drive the same
// device through this facade must serialise their calls, threads on different devices need not.
inline oibvh_ctx* context(int device = 0)
{
    static std::mutex guard;
    static oibvh_ctx* ctx[16] = {nullptr};
    if (device < 0 || device >= 16) throw std::runtime_error("oibvh_b200: bad device index");
    std::lock_guard<std::mutex> lock(guard);
    if (!ctx[device]) check(oibvh_ctx_create(device, &ctx[device]));
    return ctx[device];
}
} // namespace oibvh_detail

// ---------------------------------------------------------------------------------------------------------
// Mesh (include/utils/mesh.h) -- geometry and transforms only
// ---------------------------------------------------------------------------------------------------------
struct Vertex
{
    oibvh_math::vec3 m_position;
    oibvh_math::vec3 m_normal;
    oibvh_math::vec2 m_texCoords;
};

class Mesh
{
public:
    Mesh() = delete;
    Mesh(const std::vector<Vertex>& vertices, const std::vector<unsigned int>& indices)
        : m_verticesCount((unsigned)vertices.size())
        , m_facesCount((unsigned)(indices.size() / 3))
        , m_vertices(vertices)
        , m_indices(indices)
    {
        // Mesh::setupAABB (src/utils/mesh.cpp:91-98): note the argument order
        for (const auto& v : m_vertices)
        {
            m_aabb.m_maximum = oibvh_math::max(v.m_position, m_aabb.m_maximum);
            m_aabb.m_minimum = oibvh_math::min(v.m_position, m_aabb.m_minimum);
        }
        // Mesh::setupCenter (:145-153)
        m_center = oibvh_math::vec3(0.0f);
        for (const auto& v : m_vertices)
        {
            m_center.x += v.m_position.x;
            m_center.y += v.m_position.y;
            m_center.z += v.m_position.z;
        }
        const float n = (float)m_verticesCount;
        m_center.x /= n;
        m_center.y /= n;
        m_center.z /= n;
    }
    Mesh(const Mesh&) = default;

    void rotateX(const float angle = 1.0f) { rotate(oibvh_math::vec3(1.0f, 0.0f, 0.0f), angle); }
    void rotateY(const float angle = 1.0f) { rotate(oibvh_math::vec3(0.0f, 1.0f, 0.0f), angle); }
    void rotateZ(const float angle = 1.0f) { rotate(oibvh_math::vec3(0.0f, 0.0f, 1.0f), angle); }
    // Mesh::rotate (mesh.cpp:171-178): about the mesh centre, angle in degrees
    void rotate(const oibvh_math::vec3 axis, const float angle)
    {
        oibvh_math::mat4 m = oibvh_detail::mat4_identity();
        m = oibvh_math::translate(m, m_center);
        m = oibvh_math::rotate(m, oibvh_math::radians(angle), axis);
        m = oibvh_math::translate(m, -m_center);
        transform(m);
    }
    void translate(const oibvh_math::vec3 t) { transform(oibvh_math::translate(oibvh_detail::mat4_identity(), t)); }
    // Mesh::transform (mesh.cpp:187-213). The reference round-trips every vertex through a GPU kernel; the values
    // are the same glm expression evaluated here on the host.
    void transform(const oibvh_math::mat4 M)
    {
        const oibvh_math::vec4 c = M * oibvh_math::vec4{m_center.x, m_center.y, m_center.z, 1.0f};
        m_center = oibvh_math::vec3(c.x, c.y, c.z);
        for (auto& v : m_vertices)
        {
            const oibvh_math::vec4 p = M * oibvh_math::vec4{v.m_position.x, v.m_position.y, v.m_position.z, 1.0f};
            v.m_position = oibvh_math::vec3(p.x, p.y, p.z);
        }
    }

    unsigned int m_verticesCount;
    unsigned int m_facesCount;
    std::vector<Vertex> m_vertices;
    std::vector<unsigned int> m_indices;
    aabb_box_t m_aabb;
    oibvh_math::vec3 m_center;
};

// ---------------------------------------------------------------------------------------------------------
// OibvhTree (include/cuda/oibvhTree.cuh:44-93)
// ---------------------------------------------------------------------------------------------------------
class Scene;
class OibvhTree
{
public:
    OibvhTree() = delete;
    OibvhTree(const OibvhTree&) = delete;
    OibvhTree& operator=(const OibvhTree&) = delete;

    explicit OibvhTree(const std::shared_ptr<Mesh> mesh) : m_buildDone(false), m_mesh(mesh)
    {
        std::vector<float> pos = packedPositions();
        const float aabb[6] = {mesh->m_aabb.m_minimum.x, mesh->m_aabb.m_minimum.y, mesh->m_aabb.m_minimum.z,
                               mesh->m_aabb.m_maximum.x, mesh->m_aabb.m_maximum.y, mesh->m_aabb.m_maximum.z};
        oibvh_detail::check(oibvh_tree_create(oibvh_detail::context(), pos.data(), mesh->m_verticesCount,
                                              mesh->m_indices.data(), mesh->m_facesCount, aabb, &m_handle));
    }
    // copy constructor of the reference (src/cuda/oibvhTree.cu:17-33): same Morton order / tree as `other`
    OibvhTree(const std::shared_ptr<OibvhTree> other, const std::shared_ptr<Mesh> mesh)
        : m_buildDone(other->m_buildDone), m_mesh(mesh)
    {
        oibvh_detail::check(oibvh_tree_clone(other->m_handle, &m_handle));
    }
    ~OibvhTree() { oibvh_tree_destroy(m_handle); }

    void build()
    {
        if (m_static && m_buildDone && !m_dirty) return; // a static tree is built once
        oibvh_detail::check(oibvh_tree_build(m_handle));
        m_buildDone = true;
        m_dirty = false;
        m_version++;
    }
    // re-reads the mesh positions like the reference (oibvhTree.cu:196-199)
    void refit()
    {
        if (m_static && !m_dirty) return; // a static tree is neither re-uploaded nor refitted until markDirty()
        std::vector<float> pos = packedPositions();
        oibvh_detail::check(oibvh_tree_set_positions(m_handle, pos.data()));
        oibvh_detail::check(oibvh_tree_refit(m_handle));
        oibvh_detail::check(oibvh_ctx_synchronize(oibvh_detail::context())); // `pos` is about to go away
        m_dirty = false;
        m_version++;
    }
    // Extension: the reference re-uploads and refits every tree every frame whether or not its mesh moved
    // (main.cpp:248-262). A tree marked static keeps what it has: build() and refit() return at once until markDirty()
    // says the mesh changed (the Mesh members are public, so the facade cannot see a change by itself).
    void setStatic(bool isStatic)
    {
        m_static = isStatic;
        m_dirty = true; // the next build() / refit() still runs once
    }
    void markDirty() { m_dirty = true; }
    // Extension: Mesh::transform of the vertices ON THE DEVICE followed by the refit, without the host round trip
    // (same arithmetic, bit-identical positions: oibvh_tree_transform). The host Mesh is transformed too, so both stay
    // equal like in the reference (src/utils/mesh.cpp:187-213).
    void transformAndRefitOnDevice(const oibvh_math::mat4& M)
    {
        static_assert(sizeof(oibvh_math::mat4) == 64, "mat4 is 16 packed floats, column-major");
        m_mesh->transform(M);
        oibvh_detail::check(oibvh_tree_transform(m_handle, reinterpret_cast<const float*>(&M)));
        oibvh_detail::check(oibvh_tree_refit(m_handle));
        m_version++;
    }
    unsigned int getDepth() const
    {
        uint32_t d = 0;
        oibvh_detail::check(oibvh_tree_get_info(m_handle, nullptr, nullptr, nullptr, &d));
        return d;
    }
    unsigned int getPrimCount() const
    {
        uint32_t t = 0;
        oibvh_detail::check(oibvh_tree_get_info(m_handle, &t, nullptr, nullptr, nullptr));
        return t;
    }
    // refresh m_aabbTree / m_faces / m_positions (and the sorted->input face permutation) from the device
    void syncHost()
    {
        uint32_t T = 0, V = 0, N = 0;
        oibvh_detail::check(oibvh_tree_get_info(m_handle, &T, &V, &N, nullptr));
        m_aabbTree.resize(N);
        m_faces.resize(T);
        m_perm.resize(T);
        static_assert(sizeof(oibvh_math::uvec3) == 12, "uvec3 is three packed uint32");
        oibvh_detail::check(oibvh_tree_download(m_handle, reinterpret_cast<oibvh_aabb*>(m_aabbTree.data()),
                                                reinterpret_cast<uint32_t*>(m_faces.data()), m_perm.data()));
        std::vector<float> p(3 * (size_t)V);
        oibvh_detail::check(oibvh_tree_download_positions(m_handle, p.data()));
        m_positions.resize(V);
        for (uint32_t i = 0; i < V; i++) m_positions[i] = oibvh_math::vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    }
    oibvh_tree* handle() const { return m_handle; }
    const std::shared_ptr<Mesh>& mesh() const { return m_mesh; } // extension: the mesh this tree was created on

    // Many-body extensions (the reference loops over its objects, one build / refit call each): every tree of up to
    // 4096 triangles is processed by one thread block of ONE launch. Results equal the per-tree calls.
    static void buildMany(const std::vector<std::shared_ptr<OibvhTree>>& trees)
    {
        std::vector<oibvh_tree*> h = handles(trees);
        oibvh_detail::check(oibvh_tree_build_many(h.data(), (uint32_t)h.size()));
        for (auto& t : trees)
        {
            t->m_buildDone = true;
            t->m_version++;
        }
    }
    // refit on the positions currently on the device (after transformMany / oibvh_tree_transform)
    static void refitManyOnDevice(const std::vector<std::shared_ptr<OibvhTree>>& trees)
    {
        std::vector<oibvh_tree*> h = handles(trees);
        oibvh_detail::check(oibvh_tree_refit_many(h.data(), (uint32_t)h.size()));
        for (auto& t : trees) t->m_version++;
    }
    // Mesh::transform of every object on its device-resident vertices, one matrix per tree
    static void transformMany(const std::vector<std::shared_ptr<OibvhTree>>& trees,
                              const std::vector<oibvh_math::mat4>& mats)
    {
        if (mats.size() != trees.size()) throw std::runtime_error("transformMany: one matrix per tree");
        static_assert(sizeof(oibvh_math::mat4) == 64, "mat4 is 16 packed floats, column-major");
        std::vector<oibvh_tree*> h = handles(trees);
        oibvh_detail::check(oibvh_tree_transform_many(h.data(), (uint32_t)h.size(),
                                                      reinterpret_cast<const float*>(mats.data())));
        for (auto& t : trees) t->m_version++;
    }

    // OibvhTree::convertToVertexArray (src/cuda/oibvhTree.cu:69-124) without the GL upload: wireframe boxes of the
    // first min(internal nodes, 256) nodes into m_vertices (8 corners per box) / m_indices (12 edges per box)
    void convertToVertexArray()
    {
        uint32_t T = 0, N = 0, n = 0;
        oibvh_detail::check(oibvh_tree_get_info(m_handle, &T, nullptr, &N, nullptr));
        const uint32_t boxes = std::min<uint32_t>(N - T, 256u);
        m_vertices.resize((size_t)boxes * 8);
        m_indices.resize((size_t)boxes * 24);
        oibvh_detail::check(oibvh_tree_box_wireframe(m_handle, 256u, reinterpret_cast<float*>(m_vertices.data()),
                                                     m_indices.data(), &n));
    }
    std::vector<oibvh_math::vec3> m_vertices;    // (after convertToVertexArray)
    std::vector<unsigned int> m_indices;

    std::vector<aabb_box_t> m_aabbTree;          // N nodes, real-index order (after syncHost)
    std::vector<oibvh_math::uvec3> m_faces;      // Morton-sorted faces (after syncHost)
    std::vector<oibvh_math::vec3> m_positions;   // (after syncHost)
    std::vector<uint32_t> m_perm;                // sorted position -> input face id (extension)
    bool m_buildDone;

private:
    static std::vector<oibvh_tree*> handles(const std::vector<std::shared_ptr<OibvhTree>>& trees)
    {
        std::vector<oibvh_tree*> h;
        h.reserve(trees.size());
        for (auto& t : trees) h.push_back(t->m_handle);
        return h;
    }
    std::vector<float> packedPositions() const
    {
        std::vector<float> pos(3 * (size_t)m_mesh->m_verticesCount);
        for (unsigned i = 0; i < m_mesh->m_verticesCount; i++)
        {
            pos[3 * i] = m_mesh->m_vertices[i].m_position.x;
            pos[3 * i + 1] = m_mesh->m_vertices[i].m_position.y;
            pos[3 * i + 2] = m_mesh->m_vertices[i].m_position.z;
        }
        return pos;
    }
    std::shared_ptr<Mesh> m_mesh;
    oibvh_tree* m_handle = nullptr;
    bool m_static = false, m_dirty = true;
    unsigned long long m_version = 0; // bumped by everything that changes the device tree (replicas follow it)
    friend class Scene;
};

// ---------------------------------------------------------------------------------------------------------
// Scene (include/cuda/scene.cuh:12-67)
// ---------------------------------------------------------------------------------------------------------
enum class DeviceType
{
    CPU = -1,
    GPU0,
    GPU1,
    GPU2,
    GPU3,
    GPU4,
    GPU5,
    GPU6,
    GPU7,
    GPU8
};

class Scene
{
public:
    Scene() { oibvh_detail::check(oibvh_scene_create(oibvh_detail::context(), &m_handle)); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    ~Scene()
    {
        for (auto& kv : m_replicas)
        {
            oibvh_scene_destroy(kv.second.scene);
            for (auto* t : kv.second.trees) oibvh_tree_destroy(t);
        }
        oibvh_scene_destroy(m_handle);
    }

    void addOibvhTree(std::shared_ptr<OibvhTree> oibvhTree)
    {
        oibvh_detail::check(oibvh_scene_add_tree(m_handle, oibvhTree->m_handle));
        m_oibvhTrees.push_back(oibvhTree);
    }
    // Scene::detectCollision (src/cuda/scene.cu:157-185). DeviceType::CPU is an empty TODO in the reference
    // (scene.cu:187-190); there is no CPU path here, so it throws.
    void detectCollision(const DeviceType deviceType = DeviceType::GPU0, const unsigned int entryLevel = 0,
                         const unsigned int expandLevels = 1)
    {
        const int dev = static_cast<int>(deviceType);
        if (dev < 0) throw std::runtime_error("oibvh_b200: DeviceType::CPU is an empty TODO in the reference and there is no CPU path here");
        // the reference asserts deviceType < m_deviceCount (scene.cu:170)
        if (dev >= oibvh_device_count()) throw std::runtime_error("oibvh_b200: no such CUDA device");
        oibvh_scene* scene = m_handle;
        if (dev != 0)
        {
            // GPUk, k > 0: the reference only calls cudaSetDevice(k) and then uses buffers it allocated on another
            // device (scene.cu:229, 15-21). Here the detection really runs on device k, on replicas of the trees
            // that are created on first use and refreshed (device-to-device) whenever a tree has changed since.
            Replica& r = m_replicas[dev];
            oibvh_ctx* ctx = oibvh_detail::context(dev);
            if (!r.scene) oibvh_detail::check(oibvh_scene_create(ctx, &r.scene));
            for (size_t i = 0; i < m_oibvhTrees.size(); i++)
            {
                OibvhTree& t = *m_oibvhTrees[i];
                if (i >= r.trees.size())
                {
                    oibvh_tree* rep = nullptr;
                    oibvh_detail::check(oibvh_tree_replicate(t.m_handle, ctx, &rep));
                    oibvh_detail::check(oibvh_scene_add_tree(r.scene, rep));
                    r.trees.push_back(rep);
                    r.versions.push_back(t.m_version);
                }
                else if (r.versions[i] != t.m_version)
                {
                    oibvh_detail::check(oibvh_tree_sync_replica(r.trees[i], t.m_handle));
                    r.versions[i] = t.m_version;
                }
            }
            scene = r.scene;
        }
        uint32_t n = 0, c = 0;
        oibvh_detail::check(oibvh_scene_detect(scene, entryLevel, expandLevels, &n, &c));
        m_intTriPairCount = n;
        m_candidateCount = c;
        m_intTriPairs.resize(n);
        oibvh_detail::check(oibvh_scene_get_pairs(scene, reinterpret_cast<oibvh_int_tri_pair*>(m_intTriPairs.data())));
        m_lastScene = scene;
    }
    // Opt-in extensions (see include/oibvh_b200.h): also test every object against itself, and start
    // detections from a recorded BVTT cut while the trees are only refitted. Both apply to the GPU0 scene.
    void setSelfCollision(bool enable) { oibvh_detail::check(oibvh_scene_set_self_collision(m_handle, enable ? 1 : 0)); }
    void setTemporalCoherence(bool enable, unsigned int cutDepth = 0)
    {
        oibvh_detail::check(oibvh_scene_set_coherence(m_handle, enable ? 1 : 0, cutDepth));
    }
    unsigned int getIntTriPairCount() const { return m_intTriPairCount; }
    unsigned int getCandidateCount() const { return m_candidateCount; } // extension
    // Scene::convertToVertexArray (src/cuda/scene.cu:68-93) without the GL upload: m_vertices = six vec3 per pair
    // (triangle A, then triangle B), gathered on the device from the trees' current positions
    void convertToVertexArray()
    {
        m_vertices.resize((size_t)m_intTriPairCount * 6);
        static_assert(sizeof(oibvh_math::vec3) == 12, "vec3 is three packed floats");
        oibvh_detail::check(oibvh_scene_pair_vertices(m_lastScene ? m_lastScene : m_handle,
                                                      reinterpret_cast<float*>(m_vertices.data())));
    }
    std::vector<oibvh_math::vec3> m_vertices;

    std::vector<int_tri_pair_node_t> m_intTriPairs; // {bvhA < bvhB, triA, triB}; tri = Morton-sorted position

private:
    struct Replica // the scene and its trees on another device (detectCollision(DeviceType::GPUk))
    {
        oibvh_scene* scene = nullptr;
        std::vector<oibvh_tree*> trees;
        std::vector<unsigned long long> versions;
    };
    std::map<int, Replica> m_replicas;
    oibvh_scene* m_lastScene = nullptr;
    std::vector<std::shared_ptr<OibvhTree>> m_oibvhTrees;
    oibvh_scene* m_handle = nullptr;
    unsigned int m_intTriPairCount = 0;
    unsigned int m_candidateCount = 0;
};
