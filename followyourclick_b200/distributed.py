"""Multi-GPU plumbing: clips shard one per GPU (the reference's own strategy: DistributedSampler over prompts,
scripts/inference.py:44-51,259-269), weights replicated, and exactly ONE collective - an all_gather of the decoded
frames at the end of the clip.  Nothing inside a clip crosses GPUs (cross-frame GroupNorm, temporal and spatial attention
all reduce within the clip - SURVEY 8e: "replicas only" within a clip)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """One process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)."""
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group(backend, init_method="env://", **kw)
    return dist


def shard_clips(n_clips, rank, world_size):
    """Indices of the clips this rank renders: r, r + W, r + 2W, ... (DistributedSampler order without shuffling;
    unlike the sampler no clip is duplicated to pad the last round)."""
    return list(range(rank, n_clips, world_size))


def gather_frames(video, group=None):
    """video (b, 3, F, H, W) on every rank -> (world * b, 3, F, H, W) on every rank, rank-major order: the single
    collective of the path (NCCL all_gather over NVLink/NVSwitch on GPU, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    video = video.contiguous()
    out = torch.empty((world * video.shape[0],) + tuple(video.shape[1:]), dtype=video.dtype, device=video.device)
    dist.all_gather_into_tensor(out, video, group=group)
    return out


def gather_clip_results(videos_by_index, n_clips, group=None):
    """Ragged case (n_clips not a multiple of world size): every rank contributes its {clip index: video} dict and gets
    the full list back in clip order.  Pads the short ranks with zeros so a single all_gather suffices."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per_rank = (n_clips + world - 1) // world
    proto = next(iter(videos_by_index.values())) if videos_by_index else None
    shape = [None]
    dist.all_gather_object(shape_list := [None] * world, None if proto is None else (tuple(proto.shape), str(proto.dtype)), group=group)
    meta = next(m for m in shape_list if m is not None)
    dtype = getattr(torch, meta[1].split(".")[-1])
    dev = proto.device if proto is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    local = torch.zeros((per_rank,) + meta[0], dtype=dtype, device=dev)
    for j, idx in enumerate(shard_clips(n_clips, rank, world)):
        local[j] = videos_by_index[idx]
    allv = gather_frames(local.flatten(0, 1), group=group).view((world, per_rank) + meta[0])
    return [allv[i % world, i // world] for i in range(n_clips)]
