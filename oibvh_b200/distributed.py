"""Multi-GPU plumbing: one process per GPU, the BVH replicated, the BVTT seed front sharded
(oibvh_scene_set_shard), and only the pair list exchanged.

`torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries the single exchange step of the
path: an all-gather of the per-rank pair counts followed by an all-gather of the 16-byte pair records padded
to the largest shard. Messages are KB-MB, i.e. latency-bound; nothing else crosses GPUs.
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


# ---- fixed-size exchange: ONE collective per frame -----------------------------------------------------------------
# A scene's 512-byte counter block is the head of its pair-list allocation (include/oibvh_b200.h,
# oibvh_scene_device_counters), so [counter block | first `cap` pair records] is one contiguous device range.
HEAD_RECORDS = 32  # 512 bytes / 16-byte records; row 0 = (candidates, pairs, overflow flags, barrier word)


def block_view(counters_ptr, cap, device):
    """zero-copy int32 [HEAD_RECORDS + cap, 4] view of a scene's [counter block | pair list] device range"""
    return pairs_tensor_from_device_ptr(counters_ptr, HEAD_RECORDS + int(cap), device)


def gather_blocks(out, block, group=None):
    """all-gather every rank's block into out ([world * (HEAD_RECORDS + cap), 4]); enqueue-only on NCCL"""
    dist.all_gather_into_tensor(out, block, group=group)


def unpack_blocks(gathered, world, cap):
    """-> (per-rank pair counts, list of per-rank [min(count, cap), 4] views, truncated?) from a gathered buffer"""
    v = gathered.view(world, HEAD_RECORDS + int(cap), 4)
    counts = [int(c) for c in v[:, 0, 1].tolist()]
    parts = [v[r, HEAD_RECORDS:HEAD_RECORDS + min(counts[r], int(cap))] for r in range(world)]
    return counts, parts, max(counts) > int(cap)


def agree_capacity(n_local, floor=4096, ceiling=None, device=None, group=None):
    """Fixed per-rank record count of the block exchange, IDENTICAL on every rank (a collective with different sizes
    per rank never completes): 4x the LARGEST per-rank pair count (all-reduce MAX), rounded up to a power of two,
    at least `floor`, at most `ceiling` (the smallest pair-list capacity, also agreed by all-reduce MIN)."""
    t = torch.tensor([int(n_local), -int(ceiling) if ceiling is not None else -(1 << 62)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    n_max, ceil_min = int(t[0].item()), -int(t[1].item())
    cap = int(floor)
    while cap < 4 * n_max:
        cap *= 2
    return min(cap, ceil_min)
