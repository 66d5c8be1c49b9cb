"""Clip sharding across ranks/devices.  Frames share no state, so the only
multi-GPU logic is a static partition of the clip list: contiguous ranges,
balanced on cumulative frame count.  There is no collective on the data path;
ranks only agree on a max-over-ranks time."""
from __future__ import annotations

import numpy as np


def shard_range(n_clips: int, world: int, rank: int) -> tuple[int, int]:
    """Equal-length clips: contiguous [c0, c1) of rank `rank`."""
    return n_clips * rank // world, n_clips * (rank + 1) // world


def shard_by_frames(frames_per_clip, world: int) -> list[tuple[int, int]]:
    """Ragged clips: contiguous ranges balanced on cumulative frame count (the
    same rule mb_extract_multi applies across devices)."""
    f = np.asarray(frames_per_clip, dtype=np.int64)
    prefix = np.concatenate([[0], np.cumsum(f)])
    total = int(prefix[-1])
    cuts = [0]
    for d in range(1, world):
        c = int(np.searchsorted(prefix, total * d // world, side="left"))
        cuts.append(min(max(c, cuts[-1]), len(f)))
    cuts.append(len(f))
    return [(cuts[d], cuts[d + 1]) for d in range(world)]


def aggregate_throughput(frames_all_ranks: int, steps: int, ms_max_over_ranks: float) -> float:
    """Whole-job frames/s: all ranks' frames over the slowest rank's time."""
    return frames_all_ranks * steps / (ms_max_over_ranks * 1e-3)
