"""Multi-GPU split of the path-tracing hot path — new work, the reference is single-GPU.

Frames are independent given the scene (per-pixel RNG seed = tea(pixel, frame), pathtrace.rgen:47), so the
scene + BVH are replicated, rank r of R renders frames f = first + r (mod R), accumulates a per-rank SUM of
frame colours (SOLB_ACCUM_SUM), and ONE collective — a float32 sum-reduce of the W x H x float4 buffers to
rank 0 over NCCL/NVLink — is the path's only exchange; rank 0 then resolves sum / count with the reference's
display transform.  torch.distributed is the plumbing (nccl on GPUs, gloo in the CPU tests)."""


def frames_for_rank(rank, world, n_frames, first=0):
    """Frames of [first, first + n_frames) owned by `rank`: f = first + rank (mod world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(first + rank, first + n_frames, world))


def reduce_accum(tensor, dst=0, group=None):
    """Sum the per-rank accumulation buffers (xyz = sum of frame colours, w = frame count) onto `dst`."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allgather_rows(tensor, world, group=None):
    """Tile split for single-sample interactive frames: each rank renders H/world contiguous rows of
    one frame; gather them into the full image on every rank."""
    import torch
    import torch.distributed as dist

    parts = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(parts, tensor, group=group)
    return torch.cat(parts, dim=0)
