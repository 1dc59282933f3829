// oibvh_b200 -- Model: mesh input for the headless driver.
//
// Replaces the reference's ASSIMP-backed loader (include/utils/model.h:17-123, src/utils/model.cpp:40-186) for the
// one thing the collision path needs from it: `Model(path)` -> `m_meshes[0]` (positions + triangle indices),
// `Model(const Model&)` deep copy (main.cpp:132: bunny2 = copy of bunny1), `m_verticesCount`, `m_facesCount`,
// `m_aabb`. Textures, materials, normals and draw() are rendering and are not mirrored.
//
//   * OBJ reader: `v x y z` and `f a b c ...` records only (`a/b/c`, `a//c` and negative indices accepted, polygons
//     fan-triangulated like aiProcess_Triangulate). ASSIMP's JoinIdenticalVertices (model.cpp:48-51) re-orders
//     vertices; here vertex and face ids are the file's own order (with `joinIdenticalVertices()` as an opt-in
//     that merges bit-identical positions, first occurrence wins).
//   * Loop subdivision (BASELINE.json configs[1]: "two Loop-subdivided bunnies at 1M triangles each"): one step
//     splits every triangle into four and smooths with Loop's masks (Warren's beta; boundary edges use the
//     cubic-spline masks). The reference has no subdivision code; this is the standard scheme.
//   * Generators for the synthetic scenes the configs need when no asset is available (bunny.obj is not shipped
//     with the reference): icosphere, UV sphere (the survey's known-answer mesh), noisy blob, height-field terrain,
//     cube.
#ifndef OIBVH_MODEL_HPP
#define OIBVH_MODEL_HPP

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

#include "oibvh.hpp"

namespace oibvh_mesh
{
using oibvh_math::vec3;

struct RawMesh
{
    std::vector<vec3> positions;
    std::vector<unsigned int> indices; // 3 per triangle
    size_t faces() const { return indices.size() / 3; }
};

inline std::shared_ptr<Mesh> toMesh(const RawMesh& r)
{
    std::vector<Vertex> verts(r.positions.size());
    for (size_t i = 0; i < verts.size(); i++) verts[i].m_position = r.positions[i];
    return std::make_shared<Mesh>(verts, r.indices);
}

// ---- OBJ ---------------------------------------------------------------------------------------------------
inline RawMesh readObj(std::istream& in)
{
    RawMesh m;
    std::string line;
    while (std::getline(in, line))
    {
        if (line.size() < 2) continue;
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t'))
        {
            std::istringstream ls(line.substr(2));
            vec3 p(0.0f);
            ls >> p.x >> p.y >> p.z;
            if (!ls) throw std::runtime_error("OBJ: malformed vertex record: " + line);
            m.positions.push_back(p);
        }
        else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t'))
        {
            std::istringstream ls(line.substr(2));
            std::string tok;
            std::vector<unsigned int> poly;
            while (ls >> tok)
            {
                const long v = std::strtol(tok.c_str(), nullptr, 10); // "a", "a/b", "a/b/c", "a//c": the leading number
                if (v == 0) throw std::runtime_error("OBJ: malformed face record: " + line);
                const long idx = v > 0 ? v - 1 : (long)m.positions.size() + v;
                if (idx < 0 || idx >= (long)m.positions.size())
                    throw std::runtime_error("OBJ: vertex index out of range: " + line);
                poly.push_back((unsigned int)idx);
            }
            if (poly.size() < 3) throw std::runtime_error("OBJ: face with fewer than 3 vertices: " + line);
            for (size_t k = 1; k + 1 < poly.size(); k++) // fan triangulation
            {
                m.indices.push_back(poly[0]);
                m.indices.push_back(poly[k]);
                m.indices.push_back(poly[k + 1]);
            }
        }
    }
    if (m.positions.empty() || m.indices.empty()) throw std::runtime_error("OBJ: no geometry");
    return m;
}

inline RawMesh readObjFile(const std::string& path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    return readObj(f);
}

inline void writeObj(std::ostream& out, const RawMesh& m)
{
    char buf[128];
    for (const auto& p : m.positions)
    {
        snprintf(buf, sizeof buf, "v %.9g %.9g %.9g\n", p.x, p.y, p.z); // %.9g round-trips fp32
        out << buf;
    }
    for (size_t t = 0; t < m.faces(); t++)
        out << "f " << m.indices[3 * t] + 1 << ' ' << m.indices[3 * t + 1] + 1 << ' ' << m.indices[3 * t + 2] + 1 << '\n';
}

// merge bit-identical positions (first occurrence keeps its id, later ids are compacted in order)
inline RawMesh joinIdenticalVertices(const RawMesh& in)
{
    struct Key
    {
        uint32_t w[3];
        bool operator<(const Key& o) const { return std::memcmp(w, o.w, sizeof w) < 0; }
    };
    std::map<Key, unsigned int> seen;
    std::vector<unsigned int> remap(in.positions.size());
    RawMesh out;
    for (size_t i = 0; i < in.positions.size(); i++)
    {
        Key k;
        std::memcpy(k.w, &in.positions[i], sizeof k.w);
        auto it = seen.find(k);
        if (it == seen.end())
        {
            it = seen.emplace(k, (unsigned int)out.positions.size()).first;
            out.positions.push_back(in.positions[i]);
        }
        remap[i] = it->second;
    }
    out.indices.resize(in.indices.size());
    for (size_t i = 0; i < in.indices.size(); i++) out.indices[i] = remap[in.indices[i]];
    return out;
}

// ---- Loop subdivision ------------------------------------------------------------------------------------------
// One step: vertices keep their ids (0..V-1, re-positioned by the vertex mask), one new vertex per edge follows
// (ids V.. in order of first appearance while walking the faces), and face t becomes faces 4t..4t+3:
// (v0, e01, e20), (v1, e12, e01), (v2, e20, e12), (e01, e12, e20) -- orientation preserved.
inline RawMesh loopSubdivide(const RawMesh& in)
{
    const size_t V = in.positions.size(), T = in.faces();
    struct Edge
    {
        unsigned int id;      // new vertex id
        unsigned int opp[2];  // opposite vertices of the (up to two) incident faces
        unsigned int nfaces;
    };
    std::unordered_map<uint64_t, Edge> edges;
    edges.reserve(T * 2);
    auto key = [](unsigned int a, unsigned int b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; };
    unsigned int next = (unsigned int)V;
    std::vector<unsigned int> eid(3 * T);
    for (size_t t = 0; t < T; t++)
        for (int k = 0; k < 3; k++)
        {
            const unsigned int a = in.indices[3 * t + k], b = in.indices[3 * t + (k + 1) % 3], c = in.indices[3 * t + (k + 2) % 3];
            auto it = edges.find(key(a, b));
            if (it == edges.end()) it = edges.emplace(key(a, b), Edge{next++, {c, c}, 0}).first;
            Edge& e = it->second;
            if (e.nfaces < 2) e.opp[e.nfaces] = c;
            e.nfaces++;
            eid[3 * t + k] = e.id;
        }
    RawMesh out;
    out.positions.assign(next, vec3(0.0f));
    // vertex neighbourhoods: interior vertex -> all edge neighbours; boundary vertex -> its two boundary neighbours
    std::vector<double> sx(V, 0.0), sy(V, 0.0), sz(V, 0.0), bx(V, 0.0), by(V, 0.0), bz(V, 0.0);
    std::vector<unsigned int> valence(V, 0), bcount(V, 0);
    for (const auto& kv : edges)
    {
        const unsigned int a = (unsigned int)(kv.first >> 32), b = (unsigned int)(kv.first & 0xffffffffu);
        const Edge& e = kv.second;
        const vec3 pa = in.positions[a], pb = in.positions[b];
        valence[a]++; valence[b]++;
        sx[a] += pb.x; sy[a] += pb.y; sz[a] += pb.z;
        sx[b] += pa.x; sy[b] += pa.y; sz[b] += pa.z;
        if (e.nfaces == 1)
        {
            bcount[a]++; bcount[b]++;
            bx[a] += pb.x; by[a] += pb.y; bz[a] += pb.z;
            bx[b] += pa.x; by[b] += pa.y; bz[b] += pa.z;
            out.positions[e.id] = vec3((float)(0.5 * ((double)pa.x + pb.x)), (float)(0.5 * ((double)pa.y + pb.y)),
                                       (float)(0.5 * ((double)pa.z + pb.z)));
        }
        else
        {
            const vec3 pc = in.positions[e.opp[0]], pd = in.positions[e.opp[1]];
            out.positions[e.id] = vec3((float)(0.375 * ((double)pa.x + pb.x) + 0.125 * ((double)pc.x + pd.x)),
                                       (float)(0.375 * ((double)pa.y + pb.y) + 0.125 * ((double)pc.y + pd.y)),
                                       (float)(0.375 * ((double)pa.z + pb.z) + 0.125 * ((double)pc.z + pd.z)));
        }
    }
    for (size_t v = 0; v < V; v++)
    {
        const vec3 p = in.positions[v];
        if (valence[v] == 0)
            out.positions[v] = p; // unreferenced vertex
        else if (bcount[v] >= 2)
            out.positions[v] = vec3((float)(0.75 * p.x + 0.125 * bx[v]), (float)(0.75 * p.y + 0.125 * by[v]),
                                    (float)(0.75 * p.z + 0.125 * bz[v]));
        else
        {
            const double n = (double)valence[v];
            const double beta = n == 3.0 ? 3.0 / 16.0 : 3.0 / (8.0 * n); // Warren's weights
            out.positions[v] = vec3((float)((1.0 - n * beta) * p.x + beta * sx[v]), (float)((1.0 - n * beta) * p.y + beta * sy[v]),
                                    (float)((1.0 - n * beta) * p.z + beta * sz[v]));
        }
    }
    out.indices.reserve(12 * T);
    for (size_t t = 0; t < T; t++)
    {
        const unsigned int v0 = in.indices[3 * t], v1 = in.indices[3 * t + 1], v2 = in.indices[3 * t + 2];
        const unsigned int e01 = eid[3 * t], e12 = eid[3 * t + 1], e20 = eid[3 * t + 2];
        const unsigned int f[12] = {v0, e01, e20, v1, e12, e01, v2, e20, e12, e01, e12, e20};
        out.indices.insert(out.indices.end(), f, f + 12);
    }
    return out;
}

inline RawMesh loopSubdivide(RawMesh m, unsigned int steps)
{
    for (unsigned int s = 0; s < steps; s++) m = loopSubdivide(m);
    return m;
}

// ---- generators --------------------------------------------------------------------------------------------
inline RawMesh makeCubeMesh(float half = 0.5f)
{
    RawMesh m;
    for (int i = 0; i < 8; i++) m.positions.push_back(vec3((i & 1) ? half : -half, (i & 2) ? half : -half, (i & 4) ? half : -half));
    const unsigned int f[36] = {0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4, 2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6, 1, 3, 5, 3, 7, 5};
    m.indices.assign(f, f + 36);
    return m;
}

inline RawMesh makeIcosphere(unsigned int subdiv, float radius = 1.0f)
{
    RawMesh m;
    const float t = (1.0f + std::sqrt(5.0f)) * 0.5f;
    const float v[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                            {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    for (auto& p : v) m.positions.push_back(vec3(p[0], p[1], p[2]));
    const unsigned int f[60] = {0, 11, 5, 0, 5, 1, 0, 1, 7, 0, 7, 10, 0, 10, 11, 1, 5, 9, 5, 11, 4, 11, 10, 2, 10, 7, 6, 7, 1, 8,
                                3, 9, 4, 3, 4, 2, 3, 2, 6, 3, 6, 8, 3, 8, 9, 4, 9, 5, 2, 4, 11, 6, 2, 10, 8, 6, 7, 9, 8, 1};
    m.indices.assign(f, f + 60);
    for (unsigned int s = 0; s < subdiv; s++)
    {
        std::unordered_map<uint64_t, unsigned int> mid;
        RawMesh n;
        n.positions = m.positions;
        auto midpoint = [&](unsigned int a, unsigned int b)
        {
            const uint64_t k = a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a;
            auto it = mid.find(k);
            if (it != mid.end()) return it->second;
            const vec3 pa = n.positions[a], pb = n.positions[b];
            n.positions.push_back(vec3(0.5f * (pa.x + pb.x), 0.5f * (pa.y + pb.y), 0.5f * (pa.z + pb.z)));
            return mid[k] = (unsigned int)n.positions.size() - 1;
        };
        for (size_t i = 0; i < m.faces(); i++)
        {
            const unsigned int a = m.indices[3 * i], b = m.indices[3 * i + 1], c = m.indices[3 * i + 2];
            const unsigned int ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            const unsigned int g[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
            n.indices.insert(n.indices.end(), g, g + 12);
        }
        m = n;
    }
    for (auto& p : m.positions)
    {
        const float l = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
        p = vec3(p.x / l * radius, p.y / l * radius, p.z / l * radius);
    }
    return m;
}

// the survey's known-answer UV sphere: (n+1) x n vertices, 2 n^2 faces incl. pole slivers
inline RawMesh makeUvSphere(unsigned int n)
{
    RawMesh m;
    const float pi = 3.14159265358979323846f;
    for (unsigned int j = 0; j <= n; j++)
        for (unsigned int i = 0; i < n; i++)
        {
            const float th = pi * (float)j / (float)n, ph = 2.0f * pi * (float)i / (float)n;
            m.positions.push_back(vec3(sinf(th) * cosf(ph), cosf(th), sinf(th) * sinf(ph)));
        }
    for (unsigned int j = 0; j < n; j++)
        for (unsigned int i = 0; i < n; i++)
        {
            const unsigned int a = j * n + i, b = j * n + (i + 1) % n, c = (j + 1) * n + i, d = (j + 1) * n + (i + 1) % n;
            const unsigned int f[6] = {a, b, c, b, d, c};
            m.indices.insert(m.indices.end(), f, f + 6);
        }
    return m;
}

// small deterministic hash noise (no libm-dependent generator state): value in [-1, 1]
inline float hashNoise(uint32_t x, uint32_t y, uint32_t seed)
{
    uint32_t h = x * 0x9E3779B1u ^ (y + 0x7F4A7C15u) * 0x85EBCA77u ^ seed * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (float)(h & 0xffffffu) / 8388607.5f - 1.0f;
}

// bunny stand-in: closed genus-0 blob = UV sphere with smooth low-frequency radial displacement; nu x nv quads
inline RawMesh makeBlob(unsigned int nu, unsigned int nv, uint32_t seed = 1234, float amplitude = 0.15f)
{
    RawMesh m;
    const float pi = 3.14159265358979323846f;
    float amp[6], phase[6];
    for (int k = 0; k < 6; k++)
    {
        amp[k] = amplitude * (0.5f + 0.5f * hashNoise(k, 17, seed)) / (float)(1 + k / 2);
        phase[k] = pi * hashNoise(k, 91, seed);
    }
    m.positions.push_back(vec3(0, 1, 0)); // north pole (displaced below)
    for (unsigned int j = 1; j < nv; j++)
        for (unsigned int i = 0; i < nu; i++)
        {
            const float th = pi * (float)j / (float)nv, ph = 2.0f * pi * (float)i / (float)nu;
            m.positions.push_back(vec3(sinf(th) * cosf(ph), cosf(th), sinf(th) * sinf(ph)));
        }
    m.positions.push_back(vec3(0, -1, 0));
    for (auto& p : m.positions)
    {
        float r = 1.0f;
        for (int k = 0; k < 6; k++)
            r += amp[k] * sinf((float)(k + 1) * (p.x * 1.3f + p.y * 0.7f) + phase[k]) * cosf((float)(k / 2 + 1) * p.z * 1.9f + phase[5 - k]);
        p = vec3(p.x * r, p.y * r, p.z * r);
    }
    const unsigned int south = (unsigned int)m.positions.size() - 1;
    auto ring = [&](unsigned int j, unsigned int i) { return 1 + (j - 1) * nu + i % nu; };
    for (unsigned int i = 0; i < nu; i++)
    {
        const unsigned int f[3] = {0, ring(1, i + 1), ring(1, i)};
        m.indices.insert(m.indices.end(), f, f + 3);
    }
    for (unsigned int j = 1; j + 1 < nv; j++)
        for (unsigned int i = 0; i < nu; i++)
        {
            const unsigned int a = ring(j, i), b = ring(j, i + 1), c = ring(j + 1, i), d = ring(j + 1, i + 1);
            const unsigned int f[6] = {a, b, c, b, d, c};
            m.indices.insert(m.indices.end(), f, f + 6);
        }
    for (unsigned int i = 0; i < nu; i++)
    {
        const unsigned int f[3] = {south, ring(nv - 1, i), ring(nv - 1, i + 1)};
        m.indices.insert(m.indices.end(), f, f + 3);
    }
    return m;
}

// height-field terrain over [-sx, sx] x [-sz, sz]: n x n quads, value noise summed over `octaves`
inline RawMesh makeTerrain(unsigned int n, float half_extent = 4.0f, float height = 0.6f, uint32_t seed = 1234, int octaves = 5)
{
    RawMesh m;
    auto smooth = [&](float x, float y, uint32_t s)
    {
        const float fx = std::floor(x), fy = std::floor(y);
        const uint32_t ix = (uint32_t)(int32_t)fx, iy = (uint32_t)(int32_t)fy;
        float tx = x - fx, ty = y - fy;
        tx = tx * tx * (3.0f - 2.0f * tx);
        ty = ty * ty * (3.0f - 2.0f * ty);
        const float a = hashNoise(ix, iy, s), b = hashNoise(ix + 1, iy, s), c = hashNoise(ix, iy + 1, s), d = hashNoise(ix + 1, iy + 1, s);
        return (a + (b - a) * tx) + ((c + (d - c) * tx) - (a + (b - a) * tx)) * ty;
    };
    for (unsigned int j = 0; j <= n; j++)
        for (unsigned int i = 0; i <= n; i++)
        {
            const float u = (float)i / (float)n, v = (float)j / (float)n;
            float h = 0.0f, a = 1.0f, f = 4.0f;
            for (int o = 0; o < octaves; o++)
            {
                h += a * smooth(u * f, v * f, seed + (uint32_t)o);
                a *= 0.5f;
                f *= 2.0f;
            }
            m.positions.push_back(vec3((2.0f * u - 1.0f) * half_extent, height * h, (2.0f * v - 1.0f) * half_extent));
        }
    for (unsigned int j = 0; j < n; j++)
        for (unsigned int i = 0; i < n; i++)
        {
            const unsigned int a = j * (n + 1) + i, b = a + 1, c = a + n + 1, d = c + 1;
            const unsigned int f[6] = {a, c, b, b, c, d};
            m.indices.insert(m.indices.end(), f, f + 6);
        }
    return m;
}
} // namespace oibvh_mesh

// ---------------------------------------------------------------------------------------------------------
// Model (include/utils/model.h:17-123): the loader surface main.cpp uses
// ---------------------------------------------------------------------------------------------------------
class Model
{
public:
    Model() = delete;
    // .obj file (model.cpp:40-61 without ASSIMP); `gamma` is a texture option of the reference and is ignored
    explicit Model(const std::string& path, const bool gamma = false) { (void)gamma; init(oibvh_mesh::readObjFile(path)); }
    explicit Model(const oibvh_mesh::RawMesh& raw) { init(raw); }
    // deep copy (model.cpp:22-38): every mesh is duplicated so that transforms of the copy leave the source alone
    Model(const Model& other) : m_verticesCount(other.m_verticesCount), m_facesCount(other.m_facesCount), m_aabb(other.m_aabb)
    {
        for (const auto& m : other.m_meshes) m_meshes.push_back(std::make_shared<Mesh>(*m));
    }

    unsigned int m_verticesCount = 0;
    unsigned int m_facesCount = 0;
    aabb_box_t m_aabb;
    std::vector<std::shared_ptr<Mesh>> m_meshes;

private:
    void init(const oibvh_mesh::RawMesh& raw)
    {
        m_meshes.push_back(oibvh_mesh::toMesh(raw));
        m_verticesCount = m_meshes[0]->m_verticesCount;
        m_facesCount = m_meshes[0]->m_facesCount;
        m_aabb = m_meshes[0]->m_aabb;
    }
};

#endif // OIBVH_MODEL_HPP
