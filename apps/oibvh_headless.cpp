// oibvh_headless -- the reference's driver without the window.
//
// src/main.cpp:127-151 (set-up) and :208-319 (per-frame loop) re-enacted against the facade, with the GLFW/ImGui
// parts dropped: load or generate the meshes, build, copy-construct, transform, refit, detectCollision(GPU0, 4, 3),
// report "check result" style counts per frame. One preset per BASELINE.json config; `--scale` shrinks the triangle
// counts so the same scenarios run as tests. The colliding pair set of the last frame can be dumped in canonical
// form (bvhA, bvhB, inputFaceA, inputFaceB; sorted) for comparison against the CPU oracle (tests/test_gpu_headless.py)
// -- the reference's own check (SimpleCollide::check, main.cpp:306-319) compares only the counts.
//
//   oibvh_headless --config N [--scale S] [--frames F] [--obj file.obj] [--subdivide K] [--entry E] [--expand X]
//                  [--dump pairs.bin] [--dump-mesh out.obj] [--bodies B]
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "oibvh/model.hpp"

using oibvh_math::vec3;
using oibvh_mesh::RawMesh;
using Clock = std::chrono::steady_clock;

namespace
{
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

struct Options
{
    int config = 1;
    double scale = 1.0;
    int frames = 3;
    std::string obj, dump, dump_mesh;
    int subdivide = 0;
    unsigned entry = 4, expand = 3; // main.cpp:284
    int bodies = 0;
};

unsigned scaled(unsigned n, double scale, unsigned lo)
{
    const double v = std::max<double>(lo, std::floor((double)n * std::sqrt(scale) + 0.5));
    return (unsigned)v;
}

RawMesh bodyMesh(const Options& o, unsigned nu, unsigned nv)
{
    RawMesh m = o.obj.empty() ? oibvh_mesh::makeBlob(scaled(nu, o.scale, 8), scaled(nv, o.scale, 6)) : oibvh_mesh::readObjFile(o.obj);
    return oibvh_mesh::loopSubdivide(m, (unsigned)o.subdivide);
}

vec3 extent(const Mesh& m) { return vec3(m.m_aabb.m_maximum.x - m.m_aabb.m_minimum.x, m.m_aabb.m_maximum.y - m.m_aabb.m_minimum.y, m.m_aabb.m_maximum.z - m.m_aabb.m_minimum.z); }

void dumpCanonical(const Options& o, const Scene& scene, const std::vector<std::shared_ptr<OibvhTree>>& trees)
{
    if (o.dump.empty()) return;
    for (auto& t : trees) t->syncHost();
    std::vector<std::array<uint32_t, 4>> canon;
    for (const auto& p : scene.m_intTriPairs)
        canon.push_back({p.m_bvhIndex[0], p.m_bvhIndex[1], trees[p.m_bvhIndex[0]]->m_perm[p.m_triIndex[0]],
                         trees[p.m_bvhIndex[1]]->m_perm[p.m_triIndex[1]]});
    std::sort(canon.begin(), canon.end());
    FILE* f = fopen(o.dump.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + o.dump);
    fwrite(canon.data(), 16, canon.size(), f);
    fclose(f);
    // positions of every body as the detection saw them, so the checker rebuilds the same scene
    f = fopen((o.dump + ".scene").c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + o.dump + ".scene");
    const uint32_t n = (uint32_t)trees.size();
    fwrite(&n, 4, 1, f);
    for (auto& t : trees)
    {
        const uint32_t V = (uint32_t)t->m_positions.size(), T = t->getPrimCount();
        fwrite(&V, 4, 1, f);
        fwrite(&T, 4, 1, f);
        fwrite(t->m_positions.data(), 12, V, f);
        fwrite(t->mesh()->m_indices.data(), 4, 3 * (size_t)T, f);
        fwrite(&t->mesh()->m_aabb, 24, 1, f);
    }
    fclose(f);
}

void report(int frame, const Scene& s, double ms)
{
    printf("frame %d pairs %u candidates %u ms %.3f\n", frame, s.getIntTriPairCount(), s.getCandidateCount(), ms);
}

// configs[0] / configs[1]: body vs transformed copy (main.cpp:127-151, :240-284)
int twoBodies(const Options& o, bool rebuildEveryFrame, unsigned nu, unsigned nv)
{
    Model body1(bodyMesh(o, nu, nv));
    if (!o.dump_mesh.empty())
    {
        std::ofstream f(o.dump_mesh);
        RawMesh r;
        for (auto& v : body1.m_meshes[0]->m_vertices) r.positions.push_back(v.m_position);
        r.indices = body1.m_meshes[0]->m_indices;
        oibvh_mesh::writeObj(f, r);
    }
    auto t0 = Clock::now();
    auto tree1 = std::make_shared<OibvhTree>(body1.m_meshes[0]);
    tree1->build();
    Model body2(body1);
    auto tree2 = std::make_shared<OibvhTree>(tree1, body2.m_meshes[0]);
    const vec3 ext = extent(*body1.m_meshes[0]);
    body2.m_meshes[0]->translate(vec3(0.5f * ext.x, 0.02f * ext.y, 0.01f * ext.z));
    tree2->refit();
    Scene scene;
    scene.addOibvhTree(tree1);
    scene.addOibvhTree(tree2);
    scene.detectCollision(DeviceType::GPU0, o.entry, o.expand);
    printf("setup faces %u x2 depth %u ms %.3f\n", tree1->getPrimCount(), tree1->getDepth(), ms_since(t0));
    report(-1, scene, 0.0);
    for (int f = 0; f < o.frames; f++)
    {
        t0 = Clock::now();
        body2.m_meshes[0]->rotateZ(1.0f); // rigid 1 deg/frame about the body's centre (mesh.cpp:155-178)
        if (rebuildEveryFrame)
        {
            // configs[1]: "full pipeline per frame" = build + refit of both bodies, broad, narrow
            tree1->build();
            tree1->refit();
            tree2->refit();  // uploads the rotated positions ...
            tree2->build();  // ... which the rebuild sorts; m_aabb stays the construction-time box (mesh.cpp:91-98)
            tree2->refit();
        }
        else
            tree2->refit();
        scene.detectCollision(DeviceType::GPU0, o.entry, o.expand);
        report(f, scene, ms_since(t0));
    }
    dumpCanonical(o, scene, {tree1, tree2});
    return 0;
}

// configs[2]: deforming mesh (refit only, never rebuilt) vs static obstacle
int deforming(const Options& o)
{
    Options oo = o;
    Model cloth(bodyMesh(oo, 2048, 1025));
    Model obstacle(oibvh_mesh::makeIcosphere(o.scale >= 1.0 ? 7 : (o.scale >= 0.05 ? 5 : 3), 0.6f));
    auto rest = cloth.m_meshes[0]->m_vertices;
    obstacle.m_meshes[0]->translate(vec3(1.15f, 0.1f, 0.0f));
    auto tc = std::make_shared<OibvhTree>(cloth.m_meshes[0]);
    auto to = std::make_shared<OibvhTree>(obstacle.m_meshes[0]);
    tc->build();
    to->build();
    Scene scene;
    scene.addOibvhTree(tc);
    scene.addOibvhTree(to);
    printf("setup faces %u + %u\n", tc->getPrimCount(), to->getPrimCount());
    for (int f = 0; f < o.frames; f++)
    {
        auto t0 = Clock::now();
        const float ph = 0.35f * (float)(f + 1);
        auto& verts = cloth.m_meshes[0]->m_vertices;
        for (size_t i = 0; i < verts.size(); i++)
        {
            const vec3 p = rest[i].m_position;
            const float s = 1.0f + 0.08f * sinf(3.0f * p.y + ph) * cosf(2.0f * p.z - 0.5f * ph);
            verts[i].m_position = vec3(p.x * s, p.y * s, p.z * s);
        }
        tc->refit();
        scene.detectCollision(DeviceType::GPU0, o.entry, o.expand);
        report(f, scene, ms_since(t0));
    }
    dumpCanonical(o, scene, {tc, to});
    return 0;
}

// configs[3]: many instanced bodies in a box; one launch per step for all of them
int manyBodies(const Options& o)
{
    const int n = o.bodies > 0 ? o.bodies : std::max(8, (int)std::floor(4096.0 * o.scale + 0.5));
    const RawMesh proto = oibvh_mesh::makeBlob(32, 33); // 2048 triangles
    const RawMesh cube = oibvh_mesh::makeCubeMesh(0.8f);
    const int side = (int)std::ceil(std::cbrt((double)n));
    std::vector<std::shared_ptr<Mesh>> meshes;
    std::vector<std::shared_ptr<OibvhTree>> trees;
    std::vector<oibvh_math::mat4> mats;
    for (int i = 0; i < n; i++)
    {
        auto m = oibvh_mesh::toMesh(i % 5 == 4 ? cube : proto);
        const float jx = 0.45f * oibvh_mesh::hashNoise(i, 1, 99), jy = 0.45f * oibvh_mesh::hashNoise(i, 2, 99), jz = 0.45f * oibvh_mesh::hashNoise(i, 3, 99);
        m->translate(vec3(2.4f * (float)(i % side) + jx, 2.4f * (float)((i / side) % side) + jy, 2.4f * (float)(i / (side * side)) + jz));
        meshes.push_back(m);
        trees.push_back(std::make_shared<OibvhTree>(m)); // m_aabb = the translated construction-time box
        mats.push_back(oibvh_math::translate(oibvh_detail::mat4_identity(), vec3(0.04f * oibvh_mesh::hashNoise(i, 4, 99), 0.04f * oibvh_mesh::hashNoise(i, 5, 99), 0.0f)));
    }
    auto t0 = Clock::now();
    OibvhTree::buildMany(trees);
    Scene scene;
    for (auto& t : trees) scene.addOibvhTree(t);
    scene.detectCollision(DeviceType::GPU0, o.entry, o.expand);
    printf("setup bodies %d ms %.3f\n", n, ms_since(t0));
    report(-1, scene, 0.0);
    for (int f = 0; f < o.frames; f++)
    {
        t0 = Clock::now();
        OibvhTree::transformMany(trees, mats); // device-resident Mesh::transform, one launch
        OibvhTree::refitManyOnDevice(trees);
        scene.detectCollision(DeviceType::GPU0, o.entry, o.expand);
        report(f, scene, ms_since(t0));
    }
    dumpCanonical(o, scene, trees);
    return 0;
}

// configs[4]: large terrain vs one body pressed into it
int terrain(const Options& o)
{
    Model ground(oibvh_mesh::makeTerrain(scaled(2897, o.scale, 16)));
    Model body(bodyMesh(o, 1024, 513));
    body.m_meshes[0]->translate(vec3(0.3f, 0.75f, -0.2f));
    auto tg = std::make_shared<OibvhTree>(ground.m_meshes[0]);
    auto tb = std::make_shared<OibvhTree>(body.m_meshes[0]);
    auto t0 = Clock::now();
    tg->build();
    tb->build();
    Scene scene;
    scene.addOibvhTree(tg);
    scene.addOibvhTree(tb);
    printf("setup faces %u + %u ms %.3f\n", tg->getPrimCount(), tb->getPrimCount(), ms_since(t0));
    for (int f = 0; f < o.frames; f++)
    {
        t0 = Clock::now();
        body.m_meshes[0]->translate(vec3(0.05f, -0.02f, 0.03f));
        tb->refit();
        scene.detectCollision(DeviceType::GPU0, o.entry, o.expand);
        report(f, scene, ms_since(t0));
    }
    dumpCanonical(o, scene, {tg, tb});
    return 0;
}
} // namespace

int main(int argc, char** argv)
{
    Options o;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        auto next = [&]() -> const char*
        {
            if (i + 1 >= argc) { fprintf(stderr, "missing value after %s\n", a.c_str()); exit(64); }
            return argv[++i];
        };
        if (a == "--config") o.config = atoi(next());
        else if (a == "--scale") o.scale = atof(next());
        else if (a == "--frames") o.frames = atoi(next());
        else if (a == "--obj") o.obj = next();
        else if (a == "--subdivide") o.subdivide = atoi(next());
        else if (a == "--entry") o.entry = (unsigned)atoi(next());
        else if (a == "--expand") o.expand = (unsigned)atoi(next());
        else if (a == "--dump") o.dump = next();
        else if (a == "--dump-mesh") o.dump_mesh = next();
        else if (a == "--bodies") o.bodies = atoi(next());
        else
        {
            fprintf(stderr, "usage: %s --config 1..5 [--scale S] [--frames F] [--obj file] [--subdivide K] [--entry E] "
                            "[--expand X] [--dump pairs.bin] [--dump-mesh out.obj] [--bodies B]\n", argv[0]);
            return 64;
        }
    }
    try
    {
        switch (o.config)
        {
        case 1: return twoBodies(o, false, 136, 129);  // ~35 K triangles per body, refit + detect per frame
        case 2: return twoBodies(o, true, 1024, 513);  // 2^20 triangles per body, full pipeline per frame
        case 3: return deforming(o);
        case 4: return manyBodies(o);
        case 5: return terrain(o);
        default: fprintf(stderr, "--config must be 1..5\n"); return 64;
        }
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
