"""Frame sharding across the GPUs of one box: frames are independent, so rank g of G owns a
contiguous block of the frame index space and no data-path collective is needed.  The only exchange is
the optional gather of per-frame peak records (16-32 B/frame) onto every rank, which uses
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def frames_per_rank(total_frames: int, world: int) -> int:
    """Equal block size: ceil(F / G).  The last ranks may hold padding frames (all-zero input), so that
    all-gather counts are equal; padding is dropped again by `gather_peaks`."""
    if total_frames < 0 or world <= 0:
        raise ValueError("total_frames >= 0 and world > 0 required")
    return -(-total_frames // world)


def shard_range(total_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[start, end) of the real frames owned by `rank` (end - start may be < frames_per_rank, even 0)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} out of range for world {world}")
    per = frames_per_rank(total_frames, world)
    start = min(total_frames, rank * per)
    return start, min(total_frames, start + per)


def stft_sample_span(start: int, end: int, hop: int, frame_len: int) -> tuple[int, int]:
    """Sample range [lo, hi) a rank needs for STFT frames [start, end): its span plus the (frame_len - hop) halo."""
    if end <= start:
        return 0, 0
    return start * hop, (end - 1) * hop + frame_len


def gather_peaks(local_peaks, total_frames: int, group=None):
    """All-gather the per-rank peak records (a torch uint8/structured-byte tensor of shape
    (frames_per_rank, record_bytes), padded) into one (total_frames, record_bytes) tensor on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = frames_per_rank(total_frames, world)
    if local_peaks.shape[0] != per:
        raise ValueError(f"every rank must pass frames_per_rank={per} records (pad with zeros), got {local_peaks.shape[0]}")
    out = torch.empty((world * per,) + tuple(local_peaks.shape[1:]), dtype=local_peaks.dtype, device=local_peaks.device)
    dist.all_gather_into_tensor(out, local_peaks.contiguous(), group=group)
    return out[:total_frames]


def peaks_from_bytes(buf: np.ndarray, precision: str):
    """View gathered record bytes as the structured peak dtype of the plan precision."""
    from ._lib import PEAK_F32, PEAK_F64
    return np.ascontiguousarray(buf).view(PEAK_F64 if precision == "f64" else PEAK_F32).reshape(-1)
