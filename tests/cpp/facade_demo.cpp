// Headless re-enactment of the reference driver (src/main.cpp:127-151 set-up, :240-319 per-frame loop) against
// the facade: mesh -> OibvhTree::build -> copy-constructed second tree -> translate -> refit -> Scene ->
// detectCollision(GPU0, 4, 3) -> "check result" style report. Input: the survey's known-answer UV sphere
// (SURVEY.md Appendix A), so the expected counts are known: n = 64 -> 2093 candidates, 456 pairs.
// Writes the canonical pair list (bvhA, bvhB, inputFaceA, inputFaceB) to argv[2] for the pytest harness.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "oibvh/oibvh.hpp"

static std::shared_ptr<Mesh> uvSphere(unsigned n)
{
    const float pi = 3.14159265358979323846f;
    std::vector<Vertex> verts;
    std::vector<unsigned int> idx;
    for (unsigned j = 0; j <= n; j++)
        for (unsigned i = 0; i < n; i++)
        {
            const float th = pi * (float)j / (float)n, ph = 2.0f * pi * (float)i / (float)n;
            Vertex v;
            v.m_position = oibvh_math::vec3(sinf(th) * cosf(ph), cosf(th), sinf(th) * sinf(ph));
            verts.push_back(v);
        }
    for (unsigned j = 0; j < n; j++)
        for (unsigned i = 0; i < n; i++)
        {
            const unsigned a = j * n + i, b = j * n + (i + 1) % n, c = (j + 1) * n + i, d = (j + 1) * n + (i + 1) % n;
            const unsigned f[6] = {a, b, c, b, d, c};
            idx.insert(idx.end(), f, f + 6);
        }
    return std::make_shared<Mesh>(verts, idx);
}

int main(int argc, char** argv)
{
    const unsigned n = argc > 1 ? (unsigned)atoi(argv[1]) : 64;
    try
    {
        auto mesh1 = uvSphere(n);
        auto tree1 = std::make_shared<OibvhTree>(mesh1);
        tree1->build();
        auto mesh2 = std::make_shared<Mesh>(*mesh1);
        auto tree2 = std::make_shared<OibvhTree>(tree1, mesh2);
        mesh2->translate(oibvh_math::vec3(1.0f, 0.1f, 0.05f));
        tree2->refit();
        Scene scene;
        scene.addOibvhTree(tree1);
        scene.addOibvhTree(tree2);
        scene.detectCollision(DeviceType::GPU0, 4, 3);
        printf("faces %u depth %u candidates %u pairs %u\n", tree1->getPrimCount(), tree1->getDepth(),
               scene.getCandidateCount(), scene.getIntTriPairCount());

        tree1->syncHost();
        tree2->syncHost();
        std::vector<std::array<uint32_t, 4>> canon;
        for (const auto& p : scene.m_intTriPairs)
            canon.push_back({p.m_bvhIndex[0], p.m_bvhIndex[1], tree1->m_perm[p.m_triIndex[0]], tree2->m_perm[p.m_triIndex[1]]});
        std::sort(canon.begin(), canon.end());
        if (argc > 2)
        {
            FILE* f = fopen(argv[2], "wb");
            if (!f) return 3;
            fwrite(canon.data(), 16, canon.size(), f);
            fclose(f);
        }
        // a few frames of the per-frame loop: rotate, refit, detect (main.cpp:240-284)
        for (int frame = 0; frame < 3; frame++)
        {
            mesh2->rotateX();
            tree2->refit();
            scene.detectCollision(DeviceType::GPU0, 4, 3);
            printf("frame %d pairs %u\n", frame, scene.getIntTriPairCount());
        }
        // many-body extension: 27 small spheres on a grid, built / moved / refitted with one launch per step,
        // against the same scene handled one object at a time like the reference's loops
        {
            std::vector<std::shared_ptr<Mesh>> meshesA, meshesB;
            std::vector<std::shared_ptr<OibvhTree>> many, single;
            std::vector<oibvh_math::mat4> mats;
            for (int i = 0; i < 27; i++)
            {
                const oibvh_math::vec3 c(1.6f * (float)(i % 3), 1.6f * (float)((i / 3) % 3), 1.6f * (float)(i / 9));
                for (int copy = 0; copy < 2; copy++)
                {
                    auto m = uvSphere(8 + (unsigned)(i % 4));
                    m->translate(c);
                    (copy ? meshesB : meshesA).push_back(m);
                    // trees are created on the translated meshes (m_aabb is fixed at Mesh construction, as in the reference)
                    (copy ? single : many).push_back(std::make_shared<OibvhTree>(m));
                }
                mats.push_back(oibvh_math::translate(oibvh_detail::mat4_identity(),
                                                     oibvh_math::vec3(0.05f * (float)(i % 5), -0.03f * (float)(i % 3), 0.0f)));
            }
            OibvhTree::buildMany(many);
            for (auto& t : single) t->build();
            Scene sm, ss;
            for (auto& t : many) sm.addOibvhTree(t);
            for (auto& t : single) ss.addOibvhTree(t);
            sm.detectCollision(DeviceType::GPU0, 4, 3);
            ss.detectCollision(DeviceType::GPU0, 4, 3);
            const unsigned p0m = sm.getIntTriPairCount(), p0s = ss.getIntTriPairCount();
            OibvhTree::transformMany(many, mats);
            OibvhTree::refitManyOnDevice(many);
            for (size_t i = 0; i < single.size(); i++)
            {
                meshesB[i]->transform(mats[i]);
                single[i]->refit();
            }
            sm.detectCollision(DeviceType::GPU0, 4, 3);
            ss.detectCollision(DeviceType::GPU0, 4, 3);
            bool same_nodes = true;
            for (size_t i = 0; i < many.size(); i++)
            {
                many[i]->syncHost();
                single[i]->syncHost();
                same_nodes = same_nodes && many[i]->m_aabbTree.size() == single[i]->m_aabbTree.size() &&
                             memcmp(many[i]->m_aabbTree.data(), single[i]->m_aabbTree.data(),
                                    sizeof(aabb_box_t) * many[i]->m_aabbTree.size()) == 0;
            }
            printf("manybody bodies %zu pairs %u %u moved %u %u same_nodes %d\n", many.size(), p0m, p0s,
                   sm.getIntTriPairCount(), ss.getIntTriPairCount(), same_nodes ? 1 : 0);
        }
        // root box sanity: tree root == union of leaf boxes == mesh bounds
        const aabb_box_t& root = tree1->m_aabbTree[0];
        printf("root [%g %g %g] [%g %g %g]\n", root.m_minimum.x, root.m_minimum.y, root.m_minimum.z, root.m_maximum.x,
               root.m_maximum.y, root.m_maximum.z);
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
