"""Clip sharding across the GPUs of one box -- host-side plumbing only.

Clips are independent plugin-instance chains, so rank r of N renders a contiguous clip range with
its own engine and nothing is exchanged while rendering.  The one collective is the gather of the
per-clip records (16 floats per clip) after the render: NCCL on the device records of
jb_metrics_device (bench.py), gloo on host tensors in the CPU tests.
"""
import numpy as np

REC = 16  # floats per record (jb_metrics)


def clip_pitch(n_clips):
    """Engine-side pitch of the structure-of-arrays record block: clips rounded up to a warp."""
    return (int(n_clips) + 31) // 32 * 32


def shard_range(n_clips, rank, world):
    """[lo, hi) of the clips rank `rank` of `world` renders: contiguous, balanced to within one clip."""
    n_clips, rank, world = int(n_clips), int(rank), int(world)
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_records_soa(records, pitch):
    """[n][16] per-clip records -> the engine's device layout [16][pitch] (flat), zero padded."""
    records = np.asarray(records, dtype=np.float32)
    out = np.zeros((REC, pitch), dtype=np.float32)
    out[:, :records.shape[0]] = records.T
    return out.reshape(-1)


def unpack_gathered(flat, counts, pitch):
    """What all_gather_into_tensor returns (world x [16][pitch], flat) -> [sum(counts)][16] in clip order."""
    flat = np.asarray(flat, dtype=np.float32).reshape(len(counts), REC, pitch)
    return np.concatenate([flat[r, :, :counts[r]].T for r in range(len(counts))], axis=0)


def gather_records(local_soa, world, dist=None):
    """all_gather of every rank's [16][pitch] record block (torch tensor, device or host).
    Every rank must use the same pitch (bench.py: same clips per GPU; tests: pitch of the largest shard)."""
    import torch
    if world == 1:
        return local_soa.clone()
    if dist is None:
        import torch.distributed as dist
    out = torch.empty(world * local_soa.numel(), dtype=local_soa.dtype, device=local_soa.device)
    dist.all_gather_into_tensor(out, local_soa)
    return out
