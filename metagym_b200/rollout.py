"""Multi-GPU plumbing: shard independent envs over ranks and collate rollout chunks for the learner.

Envs never interact, so the data path has NO collective: rank r steps the contiguous global env block
[r*n_local, (r+1)*n_local) with `env_index_base = r*n_local` (per-env random streams are keyed by the global index, so
results do not depend on the sharding).  The one collective is the all-gather of a rollout chunk
{obs, act, rew, done}[T, n_local, ...] -> [T, world*n_local, ...] on every rank (NCCL over NVLink/NVSwitch on GPUs; the
same code runs on gloo for the CPU tests).
"""
import os


def shard_range(num_envs_global, rank, world_size):
    """Contiguous block of global env indices owned by `rank` -> (base, count).  Remainder goes to the low ranks."""
    q, r = divmod(int(num_envs_global), int(world_size))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def all_gather_rollout(chunk, group=None):
    """chunk: dict name -> tensor [T, n_local, ...] (same n_local on every rank).  Returns dict name -> tensor
    [T, world*n_local, ...] ordered by global env index.  One all_gather_into_tensor per field."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = {}
    for k, v in chunk.items():
        if v is None:
            continue
        if world == 1:
            out[k] = v
            continue
        src = v
        if src.dtype == torch.bool:
            src = src.to(torch.uint8)
        src = src.contiguous()
        # gather along a new leading rank axis, then fold it into the env axis: [W, T, n, ...] -> [T, W*n, ...]
        T, n = src.shape[0], src.shape[1]
        buf = torch.empty((world * T,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        dist.all_gather_into_tensor(buf, src, group=group)        # rank-major concatenation along dim 0
        g = buf.view((world, T) + tuple(src.shape[1:])).movedim(0, 1).reshape((T, world * n) + tuple(src.shape[2:]))
        out[k] = g.to(torch.bool) if v.dtype == torch.bool else g
    return out


def rollout_bytes(chunk):
    return sum(int(v.numel()) * v.element_size() for v in chunk.values() if v is not None)
