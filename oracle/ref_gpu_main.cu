// TEST / BASELINE INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Headless driver for the UNMODIFIED reference GPU path (src/cuda/{oibvh,oibvhTree,collide,scene,transform}.cu +
// src/utils/{utils,mesh}.cpp), compiled for sm_100a where the sources lie under /root/reference by oracle/Makefile
// (target `refgpu`) into oracle/_ref/ref_gpu_bench. It is "the recompiled kernel to beat" on the same B200: bench.py
// reports its frame time beside ours ("The reference GPU path (unmodified .cu, sm_100a) is timed too
// where its caps allow"), and tests/test_gpu_refgpu.py checks that its pair set equals ours.
//
//   ref_gpu_bench <mesh.bin> <frames> <out.bin>
// mesh.bin: uint32 V, T; float pos[V*3]; uint32 idx[T*3]; float offsetB[3]; float rot_axis[3]; float rot_deg
// Frame = what bench.py's frame does through the reference's own classes (main.cpp:127-151, 240-284):
//   treeA->build(); treeB->build(); meshB->rotate(axis, deg); treeA->refit(); treeB->refit();
//   scene.detectCollision(GPU0, 4, 3)
// stdout: one line per frame "frame i build_ms refit_ms detect_ms total_ms pairs"; out.bin: the last frame's pairs
// as rows (bvhA, bvhB, vA0, vA1, vA2, vB0, vB1, vB2) -- triangles named by their vertex ids, because the reference
// discards the Morton permutation.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#define private public
#define protected public
#include "cuda/oibvhTree.cuh"
#include "cuda/scene.cuh"
#undef private
#undef protected

#define REF_SHIM_DEVICE_TRANSFORM 1
#include "ref_shim.cpp" // GL no-op link shim (and the CPU wrappers, unused here)

using Clock = std::chrono::steady_clock;
static double ms(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }

struct NullBuf : std::streambuf
{
    int overflow(int c) override { return c; }
};

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s mesh.bin frames out.bin\n", argv[0]); return 64; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    uint32_t V = 0, T = 0;
    if (fread(&V, 4, 1, f) != 1 || fread(&T, 4, 1, f) != 1) return 2;
    std::vector<float> pos(3 * (size_t)V);
    std::vector<unsigned int> idx(3 * (size_t)T);
    float xf[7];
    if (fread(pos.data(), 4, pos.size(), f) != pos.size() || fread(idx.data(), 4, idx.size(), f) != idx.size() ||
        fread(xf, 4, 7, f) != 7)
        return 2;
    fclose(f);
    const int frames = atoi(argv[2]);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }

    NullBuf nb;
    std::streambuf* old = std::cout.rdbuf(); // the reference prints per-kernel timings from build()
    if (!getenv("REF_GPU_COUT")) std::cout.rdbuf(&nb);
    std::vector<Vertex> verts(V);
    std::memset(verts.data(), 0, sizeof(Vertex) * V);
    for (uint32_t i = 0; i < V; i++) verts[i].m_position = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    auto meshA = std::make_shared<Mesh>(verts, idx);
    auto meshB = std::make_shared<Mesh>(*meshA);
    const bool verbose = getenv("REF_GPU_VERBOSE") != nullptr;
    auto note = [&](const char* what)
    {
        if (!verbose) return;
        cudaDeviceSynchronize();
        fprintf(stderr, "[ref_gpu] %s (cuda: %s)\n", what, cudaGetErrorString(cudaPeekAtLastError()));
    };
    auto treeA = std::make_shared<OibvhTree>(meshA);
    note("tree A constructed");
    treeA->build();
    note("tree A built");
    auto treeB = std::make_shared<OibvhTree>(treeA, meshB);
    meshB->translate(glm::vec3(xf[0], xf[1], xf[2]));
    treeB->build();
    note("tree B built");
    Scene scene;
    scene.addOibvhTree(treeA);
    scene.addOibvhTree(treeB);
    scene.detectCollision(DeviceType::GPU0, 4, 3);
    note("first detection done");
    cudaDeviceSynchronize();
    for (int i = 0; i < frames; i++)
    {
        const auto t0 = Clock::now();
        treeA->build();
        treeB->build();
        cudaDeviceSynchronize();
        const auto t1 = Clock::now();
        meshB->rotate(glm::vec3(xf[3], xf[4], xf[5]), xf[6]);
        treeA->refit();
        treeB->refit();
        cudaDeviceSynchronize();
        const auto t2 = Clock::now();
        scene.detectCollision(DeviceType::GPU0, 4, 3);
        cudaDeviceSynchronize();
        const auto t3 = Clock::now();
        printf("frame %d %.4f %.4f %.4f %.4f %u\n", i, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t0, t3), scene.getIntTriPairCount());
        fflush(stdout);
    }
    std::cout.rdbuf(old);
    const cudaError_t e = cudaGetLastError(); // the reference checks no CUDA call; surface a sticky error here
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error after the frames: %s\n", cudaGetErrorString(e)); return 4; }
    f = fopen(argv[3], "wb");
    if (!f) return 2;
    for (unsigned i = 0; i < scene.getIntTriPairCount(); i++)
    {
        const auto p = scene.m_intTriPairs[i];
        const glm::uvec3 a = treeA->m_faces[p.m_triIndex[0]], b = treeB->m_faces[p.m_triIndex[1]];
        const uint32_t row[8] = {p.m_bvhIndex[0], p.m_bvhIndex[1], a.x, a.y, a.z, b.x, b.y, b.z};
        fwrite(row, 4, 8, f);
    }
    fclose(f);
    // final positions of body B (the reference's Mesh::transform result) so the checker can feed the same floats
    f = fopen((std::string(argv[3]) + ".posB").c_str(), "wb");
    for (uint32_t i = 0; i < V; i++) fwrite(&meshB->m_vertices[i].m_position, 12, 1, f);
    fclose(f);
    return 0;
}
