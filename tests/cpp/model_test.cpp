// Host-only checks of include/oibvh/model.hpp (OBJ reader / writer, Loop subdivision, generators, Model copy).
// Prints one "ok <name>" line per check; exits non-zero on the first failure. No GPU work.
#include <cstdio>
#include <set>
#include <sstream>

#include "oibvh/model.hpp"

using namespace oibvh_mesh;

static int g_fail = 0;
#define CHECK(cond, name)                                                                 \
    do                                                                                    \
    {                                                                                     \
        if (cond) printf("ok %s\n", name);                                                \
        else { printf("FAIL %s (%s:%d)\n", name, __FILE__, __LINE__); g_fail = 1; }       \
    } while (0)

static size_t edgeCount(const RawMesh& m, size_t* boundary = nullptr)
{
    std::map<std::pair<unsigned, unsigned>, int> e;
    for (size_t t = 0; t < m.faces(); t++)
        for (int k = 0; k < 3; k++)
        {
            unsigned a = m.indices[3 * t + k], b = m.indices[3 * t + (k + 1) % 3];
            if (a > b) std::swap(a, b);
            e[{a, b}]++;
        }
    if (boundary)
    {
        *boundary = 0;
        for (auto& kv : e) *boundary += kv.second == 1;
    }
    return e.size();
}

int main()
{
    // ---- OBJ: polygons, slashes, negative indices, comments ----
    {
        std::istringstream in("# quad + tri\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\nvn 0 0 1\nvt 0 0\n"
                              "f 1/1/1 2/1/1 3/1/1 4/1/1\nf -1 1//1 2//1\n");
        RawMesh m = readObj(in);
        CHECK(m.positions.size() == 5 && m.faces() == 3, "obj_counts");
        const unsigned want[9] = {0, 1, 2, 0, 2, 3, 4, 0, 1};
        CHECK(std::equal(want, want + 9, m.indices.begin()), "obj_fan_and_negative_indices");
        std::ostringstream out;
        writeObj(out, m);
        std::istringstream back(out.str());
        RawMesh r = readObj(back);
        CHECK(r.indices == m.indices && memcmp(r.positions.data(), m.positions.data(), 12 * m.positions.size()) == 0,
              "obj_round_trip_bit_exact");
        bool threw = false;
        try { std::istringstream bad("v 0 0 0\nf 1 2 3\n"); readObj(bad); } catch (const std::exception&) { threw = true; }
        CHECK(threw, "obj_out_of_range_index_rejected");
    }
    // ---- Loop subdivision on a closed mesh ----
    {
        RawMesh ico = makeIcosphere(0);
        const size_t E = edgeCount(ico);
        RawMesh s = loopSubdivide(ico);
        size_t b = 1;
        const size_t E2 = edgeCount(s, &b);
        CHECK(s.faces() == 4 * ico.faces() && s.positions.size() == ico.positions.size() + E, "loop_counts");
        CHECK(b == 0 && (long)s.positions.size() - (long)E2 + (long)s.faces() == 2, "loop_closed_genus0");
        // an icosahedron on the unit sphere stays inside it and symmetric: all old vertices share one radius
        float r0 = -1.f, dev = 0.f;
        for (size_t i = 0; i < ico.positions.size(); i++)
        {
            const auto& p = s.positions[i];
            const float r = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
            if (r0 < 0) r0 = r;
            dev = std::max(dev, std::fabs(r - r0));
        }
        const auto& q = ico.positions[0];
        const float rin = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
        CHECK(dev < 1e-5f && r0 < rin && r0 > 0.5f * rin, "loop_vertex_mask_contracts_symmetrically");
        RawMesh s3 = loopSubdivide(ico, 3);
        CHECK(s3.faces() == 20 * 64, "loop_three_steps");
        // orientation preserved: signed volume keeps its sign
        auto vol = [](const RawMesh& m)
        {
            double v = 0;
            for (size_t t = 0; t < m.faces(); t++)
            {
                const auto &a = m.positions[m.indices[3 * t]], &b = m.positions[m.indices[3 * t + 1]], &c = m.positions[m.indices[3 * t + 2]];
                v += a.x * (b.y * c.z - b.z * c.y) - a.y * (b.x * c.z - b.z * c.x) + a.z * (b.x * c.y - b.y * c.x);
            }
            return v;
        };
        CHECK(vol(ico) * vol(s3) > 0 && std::fabs(vol(s3)) < std::fabs(vol(ico)), "loop_orientation_and_shrinkage");
    }
    // ---- Loop subdivision with boundaries: a flat grid stays flat, its outline stays put ----
    {
        RawMesh g = makeTerrain(4, 1.0f, 0.0f);
        RawMesh s = loopSubdivide(g);
        size_t b0 = 0, b1 = 0;
        edgeCount(g, &b0);
        edgeCount(s, &b1);
        bool flat = true, inside = true;
        for (auto& p : s.positions)
        {
            flat = flat && p.y == 0.0f;
            inside = inside && std::fabs(p.x) <= 1.0f && std::fabs(p.z) <= 1.0f;
        }
        CHECK(s.faces() == 4 * g.faces() && b1 == 2 * b0, "loop_boundary_counts");
        CHECK(flat && inside, "loop_boundary_flat_grid");
        // a straight boundary is reproduced exactly by the cubic-spline masks: corner-adjacent edge midpoints stay on the line
        bool onEdge = false;
        for (auto& p : s.positions) onEdge = onEdge || (p.z == -1.0f && p.x > -1.0f && p.x < -0.5f);
        CHECK(onEdge, "loop_boundary_midpoints_on_outline");
    }
    // ---- generators ----
    {
        RawMesh blob = makeBlob(136, 129);
        size_t b = 1;
        const size_t E = edgeCount(blob, &b);
        CHECK(blob.faces() == 34816 && b == 0 && (long)blob.positions.size() - (long)E + (long)blob.faces() == 2, "blob_closed_34816");
        CHECK(makeBlob(1024, 513).faces() == (1u << 20), "blob_2pow20");
        CHECK(makeUvSphere(64).faces() == 8192 && makeUvSphere(64).positions.size() == 65 * 64, "uv_sphere_survey_shape");
        CHECK(makeCubeMesh().faces() == 12 && makeTerrain(8).faces() == 128, "cube_terrain_counts");
        RawMesh dup = makeCubeMesh();
        dup.positions.push_back(dup.positions[3]);
        dup.indices[0] = 8;
        RawMesh j = joinIdenticalVertices(dup);
        CHECK(j.positions.size() == 8 && j.indices[0] == 3, "join_identical_vertices");
    }
    // ---- Model: deep copy like Model(const Model&) (model.cpp:22-38) ----
    {
        Model a(makeIcosphere(1));
        Model b(a);
        b.m_meshes[0]->translate(oibvh_math::vec3(1.0f, 0.0f, 0.0f));
        CHECK(a.m_meshes[0]->m_vertices[0].m_position.x != b.m_meshes[0]->m_vertices[0].m_position.x, "model_copy_is_deep");
        CHECK(a.m_facesCount == 80 && a.m_verticesCount == 42 && a.m_aabb.m_maximum.x > 0.9f, "model_counts_and_aabb");
    }
    return g_fail;
}
