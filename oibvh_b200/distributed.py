"""Multi-GPU plumbing: one process per GPU, the BVH replicated, the BVTT front of round 0 dealt to the ranks
(oibvh_scene_set_shard), and only the pair list gathered.

The data path needs NO collective per frame: every rank's narrow phase appends its hits directly to rank 0's pair list
through a CUDA-IPC peer mapping over NVLink (C ABI: oibvh_mgpu_export / oibvh_mgpu_attach; protocol in
csrc/collide_kernels.cu). `torch.distributed` is only the transport of the 160-byte handle at set-up (`attach`) and
the barrier around timed regions. `gather_pairs` is the collective alternative (all-gather of counts, then of the
padded 16-byte records; NCCL over NVLink on GPUs, gloo in the CPU tests), kept as the cross-check of the peer path.
"""
import torch
import torch.distributed as dist


def gather_pairs(local_pairs, n_local, group=None):
    """local_pairs: int32 tensor [cap>=n_local, 4] on this rank's device (rows beyond n_local are ignored).
    Returns the concatenation over ranks ([sum n, 4], rank order) on every rank."""
    world = dist.get_world_size(group)
    dev = local_pairs.device
    cnt = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    send = torch.zeros((width, 4), dtype=torch.int32, device=dev)
    if n_local:
        send[:n_local] = local_pairs[:n_local]
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([recv[r][:counts[r]] for r in range(world)], dim=0)


def attach(scene, rank, world, group=None, src=0):
    """Put `scene` (one per rank, same trees on every rank) into multi-GPU mode: rank `src` exports the handle of its
    pair list, the handle travels by broadcast, the others map it. Collective: every rank of the group calls it."""
    scene.set_shard(rank, world)
    box = [scene.mgpu_export() if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    if rank != src:
        scene.mgpu_attach(box[0])
    dist.barrier(group=group)


def detach(scene, group=None):
    """leave multi-GPU mode after the last frame has completed on every rank"""
    scene.ctx.synchronize()
    dist.barrier(group=group)
    scene.mgpu_detach()
    dist.barrier(group=group)


def pairs_tensor_from_device_ptr(ptr, n, device):
    """zero-copy int32 [n,4] view of the scene's device pair list (oibvh_scene_device_pairs)"""
    if n == 0:
        return torch.empty((0, 4), dtype=torch.int32, device=device)

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n), 4), "typestr": "<i4", "data": (int(ptr), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(h, device=device)
