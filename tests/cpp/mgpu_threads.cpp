// The multi-GPU C ABI driven from C++ host threads, one thread per rank (the reference's only
// multi-device artefact is the DeviceType enum + cudaSetDevice, include/cuda/scene.cuh:12-24, src/cuda/scene.cu:229).
//   rank r: its own context on device r % device_count, replicas of both trees (rank 0 builds, the others get them
//   with oibvh_tree_replicate), its own scene with oibvh_scene_set_shard(r, W); rank 0 exports the handle of its pair
//   list, the others attach; every rank then runs F frames (rotate body B on its device, refit, detect) WITHOUT any
//   synchronisation between the threads -- the ranks meet on the device. After every frame rank 0 compares the
//   gathered pair set with a single-GPU detection of the same trees.
//   With more ranks than devices (a single-GPU test box) ranks share a device, where their machine-filling detection
//   kernels cannot run side by side: the frames are then ORDERED on the host -- rank 0 opens the frame
//   (oibvh_mgpu_open_frame), the other ranks detect and finish, rank 0 detects last -- which is what that entry point
//   is for. One rank per device needs none of this.
// Also exercises the facade: Scene::detectCollision(DeviceType::GPU1) where a second device exists.
// Prints "ok ranks W frames F pairs P" or a failure and a non-zero exit code.
#include <algorithm>
#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>

#include "oibvh/model.hpp"
#include "oibvh/oibvh.hpp"

#define CK(call)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        if ((call) != OIBVH_OK)                                                                    \
        {                                                                                          \
            fprintf(stderr, "%s failed: %s\n", #call, oibvh_last_error());                         \
            failed = true;                                                                         \
            return;                                                                                \
        }                                                                                          \
    } while (0)

static std::atomic<bool> failed{false};

struct Shared
{
    std::mutex m;
    std::condition_variable cv;
    oibvh_tree* root_trees[2] = {nullptr, nullptr};
    bool trees_ready = false;
    oibvh_mgpu_handle handle;
    bool handle_ready = false;
    int attached = 0, finished = 0;
    int opened = 0, remote_done = 0; // ordered mode (ranks sharing a device)
};

static std::vector<std::array<uint32_t, 4>> sorted_pairs(oibvh_scene* sc)
{
    uint32_t n = 0;
    if (oibvh_scene_get_counts(sc, &n, nullptr) != OIBVH_OK) return {};
    std::vector<std::array<uint32_t, 4>> p(n);
    if (n && oibvh_scene_get_pairs(sc, reinterpret_cast<oibvh_int_tri_pair*>(p.data())) != OIBVH_OK) return {};
    std::sort(p.begin(), p.end());
    return p;
}

static void rank_main(int rank, int world, int frames, Shared* sh, const oibvh_mesh::RawMesh* raw)
{
    const int dev = rank % std::max(1, oibvh_device_count());
    oibvh_ctx* ctx = nullptr;
    CK(oibvh_ctx_create(dev, &ctx));
    oibvh_tree* tree[2] = {nullptr, nullptr};
    if (rank == 0)
    {
        std::vector<float> pos;
        for (const auto& v : raw->positions)
        {
            pos.push_back(v.x);
            pos.push_back(v.y);
            pos.push_back(v.z);
        }
        float aabb[6] = {1e30f, 1e30f, 1e30f, -1e30f, -1e30f, -1e30f};
        for (size_t i = 0; i < pos.size(); i++)
        {
            aabb[i % 3] = std::min(aabb[i % 3], pos[i]);
            aabb[3 + i % 3] = std::max(aabb[3 + i % 3], pos[i]);
        }
        const uint32_t V = (uint32_t)raw->positions.size(), T = (uint32_t)(raw->indices.size() / 3);
        CK(oibvh_tree_create(ctx, pos.data(), V, raw->indices.data(), T, aabb, &tree[0]));
        CK(oibvh_tree_build(tree[0]));
        CK(oibvh_tree_clone(tree[0], &tree[1]));
        const float shift[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.9f, 0.1f, 0.05f, 1};
        CK(oibvh_tree_transform(tree[1], shift));
        CK(oibvh_tree_refit(tree[1]));
        CK(oibvh_ctx_synchronize(ctx));
        std::lock_guard<std::mutex> lock(sh->m);
        sh->root_trees[0] = tree[0];
        sh->root_trees[1] = tree[1];
        sh->trees_ready = true;
        sh->cv.notify_all();
    }
    else
    {
        std::unique_lock<std::mutex> lock(sh->m);
        sh->cv.wait(lock, [&] { return sh->trees_ready || failed.load(); });
        if (failed) return;
        // rank 0 is idle (it waits for the attachments below), so its trees can be read
        CK(oibvh_tree_replicate(sh->root_trees[0], ctx, &tree[0]));
        CK(oibvh_tree_replicate(sh->root_trees[1], ctx, &tree[1]));
    }
    oibvh_scene* sc = nullptr;
    CK(oibvh_scene_create(ctx, &sc));
    CK(oibvh_scene_add_tree(sc, tree[0]));
    CK(oibvh_scene_add_tree(sc, tree[1]));
    CK(oibvh_scene_set_shard(sc, (uint32_t)rank, (uint32_t)world));
    if (rank == 0)
    {
        CK(oibvh_scene_reserve(sc, 1u << 20, 1u << 20, 1u << 18));
        std::unique_lock<std::mutex> lock(sh->m);
        CK(oibvh_mgpu_export(sc, &sh->handle));
        sh->handle_ready = true;
        sh->cv.notify_all();
        sh->cv.wait(lock, [&] { return sh->attached == world - 1 || failed.load(); });
    }
    else
    {
        std::unique_lock<std::mutex> lock(sh->m);
        sh->cv.wait(lock, [&] { return sh->handle_ready || failed.load(); });
        if (failed) return;
        CK(oibvh_mgpu_attach(sc, &sh->handle));
        sh->attached++;
        sh->cv.notify_all();
    }
    if (failed) return;
    // a second, plain scene on rank 0 for the single-GPU answer
    oibvh_scene* plain = nullptr;
    if (rank == 0)
    {
        CK(oibvh_scene_create(ctx, &plain));
        CK(oibvh_scene_add_tree(plain, tree[0]));
        CK(oibvh_scene_add_tree(plain, tree[1]));
    }
    // rotation of body B about z by one degree (any fixed matrix will do: every rank applies the same one)
    const float c = 0.99984770f, s = 0.01745241f;
    const float rot[16] = {c, s, 0, 0, -s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    size_t last_pairs = 0;
    const bool ordered = world > oibvh_device_count();
    for (int f = 0; f < frames && !failed; f++)
    {
        CK(oibvh_tree_transform(tree[1], rot));
        CK(oibvh_tree_refit(tree[1]));
        if (ordered)
        {
            std::unique_lock<std::mutex> lock(sh->m);
            if (rank == 0)
            {
                CK(oibvh_mgpu_open_frame(sc));
                CK(oibvh_ctx_synchronize(ctx));
                sh->opened = f + 1;
                sh->cv.notify_all();
                sh->cv.wait(lock, [&] { return sh->remote_done == (world - 1) * (f + 1) || failed.load(); });
            }
            else
            {
                sh->cv.wait(lock, [&] { return sh->opened == f + 1 || failed.load(); });
                if (failed) break;
                lock.unlock();
                uint32_t n = 0;
                CK(oibvh_scene_detect_async(sc, 4, 0));
                CK(oibvh_scene_get_counts(sc, &n, nullptr)); // this rank's frame has completed
                lock.lock();
                sh->remote_done++;
                sh->cv.notify_all();
                continue;
            }
        }
        if (failed) break;
        CK(oibvh_scene_detect_async(sc, 4, 0));
        if (rank == 0)
        {
            const auto got = sorted_pairs(sc); // waits for every rank's hits of this frame
            uint32_t n = 0;
            CK(oibvh_scene_detect(plain, 4, 0, &n, nullptr));
            const auto want = sorted_pairs(plain);
            if (got != want || want.empty())
            {
                fprintf(stderr, "frame %d: gathered %zu pairs, single GPU %zu\n", f, got.size(), want.size());
                failed = true;
            }
            last_pairs = want.size();
        }
        else
        {
            uint32_t n = 0;
            CK(oibvh_scene_get_counts(sc, &n, nullptr)); // this rank's own hits
        }
    }
    {
        // nobody unmaps before every rank has finished its last frame
        std::unique_lock<std::mutex> lock(sh->m);
        sh->finished++;
        sh->cv.notify_all();
        sh->cv.wait(lock, [&] { return sh->finished == world || failed.load(); });
    }
    oibvh_mgpu_detach(sc);
    if (rank == 0 && !failed) printf("ok ranks %d frames %d pairs %zu\n", world, frames, last_pairs);
    oibvh_scene_destroy(plain);
    oibvh_scene_destroy(sc);
    if (rank != 0) // rank 0's trees may still be read by late replications: it outlives everyone (joined last)
    {
        oibvh_tree_destroy(tree[0]);
        oibvh_tree_destroy(tree[1]);
        oibvh_ctx_destroy(ctx);
    }
}

int main(int argc, char** argv)
{
    const int world = argc > 1 ? atoi(argv[1]) : 2, frames = argc > 2 ? atoi(argv[2]) : 5;
    if (oibvh_device_count() < 1)
    {
        fprintf(stderr, "no CUDA device\n");
        return 2;
    }
    const oibvh_mesh::RawMesh raw = oibvh_mesh::makeBlob(160, 128, 7);
    Shared sh;
    std::vector<std::thread> threads;
    for (int r = 0; r < world; r++) threads.emplace_back(rank_main, r, world, frames, &sh, &raw);
    for (auto& t : threads) t.join();
    if (failed) return 1;

    // facade: Scene::detectCollision on another device gives the same pairs as on GPU0
    if (oibvh_device_count() >= 2)
    {
        try
        {
            auto mesh1 = oibvh_mesh::toMesh(raw);
            auto tree1 = std::make_shared<OibvhTree>(mesh1);
            tree1->build();
            auto mesh2 = std::make_shared<Mesh>(*mesh1);
            auto tree2 = std::make_shared<OibvhTree>(tree1, mesh2);
            mesh2->translate(oibvh_math::vec3(0.9f, 0.1f, 0.05f));
            tree2->refit();
            Scene scene;
            scene.addOibvhTree(tree1);
            scene.addOibvhTree(tree2);
            scene.detectCollision(DeviceType::GPU0, 4, 3);
            auto a = scene.m_intTriPairs;
            scene.detectCollision(DeviceType::GPU1, 4, 3);
            auto b = scene.m_intTriPairs;
            mesh2->rotateZ(2.0f);
            tree2->refit(); // the replica on GPU1 follows
            scene.detectCollision(DeviceType::GPU1, 4, 3);
            auto c1 = scene.m_intTriPairs;
            scene.detectCollision(DeviceType::GPU0, 4, 3);
            auto c0 = scene.m_intTriPairs;
            auto key = [](const int_tri_pair_node_t& p) {
                return std::array<uint32_t, 4>{p.m_bvhIndex[0], p.m_bvhIndex[1], p.m_triIndex[0], p.m_triIndex[1]};
            };
            auto canon = [&](std::vector<int_tri_pair_node_t>& v) {
                std::vector<std::array<uint32_t, 4>> k;
                for (auto& p : v) k.push_back(key(p));
                std::sort(k.begin(), k.end());
                return k;
            };
            if (canon(a) != canon(b) || canon(c0) != canon(c1) || a.empty())
            {
                fprintf(stderr, "facade: GPU1 pairs differ from GPU0 (%zu vs %zu, %zu vs %zu)\n", a.size(), b.size(),
                        c0.size(), c1.size());
                return 1;
            }
            printf("facade GPU1 ok pairs %zu then %zu\n", a.size(), c0.size());
        }
        catch (const std::exception& e)
        {
            fprintf(stderr, "facade: %s\n", e.what());
            return 1;
        }
    }
    return 0;
}
