// Drop-in check: the reference driver's set-up and frame lines (src/main.cpp:127-151, 240-284) written against the
// facade in GLM mode -- glm::vec3 / glm::mat4 are the tree's own types, as inside the reference tree. Compiled by
// tests/test_model_io.py against the reference's vendored glm (third/glm) where /root/reference exists; the golden
// counts of the same flow are checked on the GPU by tests/test_gpu_facade.py (mini-type mode).
// (OIBVH_FACADE_USE_GLM is given on the command line: the same file is also built with the facade's mini types, and
// tests/test_gpu_facade.py checks that both binaries print the same bits)
#include <cstdio>
#include <cstring>
#include <oibvh/model.hpp>
#ifndef OIBVH_FACADE_USE_GLM
namespace glm = oibvh_math;
#endif
int main()
{
    Model bunny1(oibvh_mesh::makeIcosphere(2));
    std::shared_ptr<OibvhTree> treeBunny1 = std::make_shared<OibvhTree>(bunny1.m_meshes[0]);
    treeBunny1->build();
    Model bunny2(bunny1);
    std::shared_ptr<OibvhTree> treeBunny2 = std::make_shared<OibvhTree>(treeBunny1, bunny2.m_meshes[0]);
    bunny2.m_meshes[0]->translate(glm::vec3(1.0f, 0.0f, 0.0f));
    treeBunny2->refit();
    Scene scene;
    scene.addOibvhTree(treeBunny1);
    scene.addOibvhTree(treeBunny2);
    bunny2.m_meshes[0]->rotateX();
    treeBunny2->refit();
    scene.detectCollision(DeviceType::GPU0, 4, 3);
    scene.convertToVertexArray();
    treeBunny1->convertToVertexArray();
    treeBunny2->syncHost();
    uint32_t h = 2166136261u; // FNV-1a over every float the frame produced: vertex stream, wireframes, moved positions
    auto mix = [&](const void* p, size_t n)
    {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 16777619u;
    };
    // the pair list (hence the vertex stream) comes in a different order every run: combine per-pair hashes commutatively
    uint32_t stream = 0;
    for (size_t i = 0; i + 6 <= scene.m_vertices.size(); i += 6)
    {
        h = 2166136261u;
        mix(&scene.m_vertices[i], 6 * sizeof(glm::vec3));
        stream += h;
    }
    h = 2166136261u ^ stream;
    mix(treeBunny1->m_vertices.data(), treeBunny1->m_vertices.size() * sizeof(glm::vec3));
    mix(treeBunny2->m_positions.data(), treeBunny2->m_positions.size() * sizeof(glm::vec3));
    printf("pairs %u candidates %u depth %u prims %u vertices %zu boxes %zu hash %08x\n", scene.getIntTriPairCount(),
           scene.getCandidateCount(), treeBunny1->getDepth(), treeBunny1->getPrimCount(), scene.m_vertices.size(),
           treeBunny1->m_indices.size() / 24, h);
    return 0;
}
