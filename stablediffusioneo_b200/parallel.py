"""Multi-GPU plumbing for the denoising path: replicas only.

Every image/prompt is independent for all DDIM steps — GroupNorm, LayerNorm and attention are per-sample — so the
path shards by image: one process per GPU, rank r takes images[r::world], each rank holds a full weight copy and its
own step graph, and nothing is exchanged inside the loop. The single collective is an all-gather of the final latents
(or uint8 images) after the loop (NCCL over NVLink on GPUs, gloo in the CPU tests); it is a few KB per image, so it is
one launch at the very end and is not fused with any kernel."""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Indices of the images this rank denoises: a strided partition, balanced to within one item."""
    return list(range(rank, n_items, world))


def owner_of(index, world):
    return index % world


def gather_by_image(local, n_items, rank=None, world=None, group=None):
    """local: [len(shard_indices(n_items, rank, world)), ...] results of this rank, in shard order.
    Returns the full [n_items, ...] tensor in image order on every rank (all-gather + un-stride)."""
    if not dist.is_available() or not dist.is_initialized():
        assert local.shape[0] == n_items
        return local
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    per_rank = (n_items + world - 1) // world
    pad = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        if idx:
            out[idx] = bufs[r][: len(idx)]
    return out
