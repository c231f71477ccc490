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


def gather_frames(local_frames, nframes: int, dist=None, group=None, out=None, reorder=True):
    """All ranks end up with all `nframes` decoded frames.

    local_frames: tensor [n_local, ...] holding this rank's frames in the order of
    frames_of_rank().  Ranks with fewer frames are padded for the collective.
    out: optional preallocated result buffer of shape [world, per, ...] (per =
    max_frames_per_rank); without it one is allocated per call.
    reorder=True returns the frames in frame order (one extra device copy of the whole
    batch); reorder=False returns the collective's own layout, a [world, per, ...] view
    in which frame r + k*world sits at [r, k] -- no copy besides the collective.
    Works with any torch.distributed backend (nccl on GPUs, gloo on CPU)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_frames
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = max_frames_per_rank(nframes, world)
    shape = (per,) + tuple(local_frames.shape[1:])
    if local_frames.shape[0] == per and local_frames.is_contiguous():
        padded = local_frames  # (the usual case: every rank holds `per` frames)
    else:
        padded = torch.zeros(shape, dtype=local_frames.dtype, device=local_frames.device)
        padded[:local_frames.shape[0]] = local_frames
    if out is None:
        out = torch.empty((world,) + shape, dtype=local_frames.dtype, device=local_frames.device)
    elif tuple(out.shape) != (world,) + shape or out.dtype != local_frames.dtype:
        raise ValueError("gather_frames: `out` must have shape [world, per, ...] and the frames' dtype")
    # the collective moves raw bytes (uint8): every backend supports it
    dist.all_gather_into_tensor(out.view(-1).view(torch.uint8), padded.view(-1).view(torch.uint8),
                                group=group)
    assert frames_of_rank(nframes, rank, world) == list(range(rank, nframes, world))
    if not reorder:
        return out
    # out[r, k] is frame r + k*world
    full = out.permute(1, 0, *range(2, out.dim())).reshape((per * world,) + shape[1:])
    return full[:nframes]
