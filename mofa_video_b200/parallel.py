"""Multi-GPU plumbing for the one place the path shards: independent clips on the batch axis
(SURVEY.md §8e).  One process per GPU (torchrun), weights replicated, clip k -> rank k mod W, no collective
during a clip; the only exchange is the final gather of uint8 frames to rank 0."""
import torch
import torch.distributed as dist


def clips_for_rank(n_clips, world, rank):
    """Round-robin assignment of clip indices (clip k -> rank k mod world)."""
    return list(range(rank, n_clips, world))


def gather_frames(frames_u8, dst=0):
    """frames_u8: uint8 [T, H, W, 3] of this rank's clip.  Returns the list of all ranks' clips on `dst`
    (None elsewhere).  NCCL on device tensors, gloo on CPU tensors; a single collective, no reduction."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [frames_u8]
    world, rank = dist.get_world_size(), dist.get_rank()
    bufs = [torch.empty_like(frames_u8) for _ in range(world)] if rank == dst else None
    dist.gather(frames_u8.contiguous(), bufs, dst=dst)
    return bufs
