"""Frame sharding across the GPUs of one box (one process per GPU).

The hot path partitions by frame (and, inside a frame, by tile): frames are
independent units, there is NO exchange step during decode.  Frame i
of a batch goes to rank i % world ("round robin", keeps per-rank work equal when
frames differ in size).  The only collective is the optional gather of the
decoded uint16 frames (NCCL all_gather over NVLink on the GPU box, gloo in the
CPU tests) -- plumbing provided by torch.distributed."""
from typing import List, Sequence


def frames_of_rank(nframes: int, rank: int, world: int) -> List[int]:
    """Indices of the frames rank `rank` decodes."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, nframes, world))


def owner_of_frame(i: int, world: int) -> int:
    return i % world


def max_frames_per_rank(nframes: int, world: int) -> int:
    return (nframes + world - 1) // world


def gather_frames(local_frames, nframes: int, dist=None, group=None):
    """All ranks end up with all `nframes` decoded frames, in frame order.

    local_frames: tensor [n_local, ...] holding this rank's frames in the order of
    frames_of_rank().  Ranks with fewer frames are padded for the collective.
    Works with any torch.distributed backend (nccl on GPUs, gloo on CPU)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_frames
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = max_frames_per_rank(nframes, world)
    shape = (per,) + tuple(local_frames.shape[1:])
    padded = torch.zeros(shape, dtype=local_frames.dtype, device=local_frames.device)
    padded[:local_frames.shape[0]] = local_frames
    out = torch.empty((world,) + shape, dtype=local_frames.dtype, device=local_frames.device)
    # the collective moves raw bytes (uint8): every backend supports it
    dist.all_gather_into_tensor(out.view(-1).view(torch.uint8), padded.view(-1).view(torch.uint8),
                                group=group)
    # out[r, k] is frame r + k*world
    full = out.permute(1, 0, *range(2, out.dim())).reshape((per * world,) + shape[1:])
    assert frames_of_rank(nframes, rank, world) == list(range(rank, nframes, world))
    return full[:nframes]
