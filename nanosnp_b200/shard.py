"""Region sharding of the s1+s2 path across GPUs.

Every site depends only on reads overlapping [c-16, c+16], so contigs are cut into fixed regions [s, e);
each region is computed on [s-16, e+16) and emits the sites with s <= c < e.  Regions are independent:
there is no data-path collective.  Per-rank call lists are gathered to rank 0 (torch.distributed
gather_object: NCCL/gloo plumbing only) and concatenated in (contig, position) order BEFORE the record
logic runs in 1000-site batches per contig, because predict.py's output depends on batch composition
(SURVEY 8a, row P13).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

FLANK = 16


@dataclass(frozen=True)
class Region:
    contig: str
    contig_index: int
    contig_len: int
    emit_start: int      # sites with emit_start <= c < emit_end belong to this region
    emit_end: int

    @property
    def start(self) -> int:          # computed span = emitted span + halo
        return max(0, self.emit_start - FLANK)

    @property
    def end(self) -> int:
        return min(self.contig_len, self.emit_end + FLANK)

    @property
    def length(self) -> int:
        return self.end - self.start


def plan_regions(contigs: Sequence[Tuple[str, int]], region_len: int) -> List[Region]:
    """Cuts each (name, length) contig into regions of at most region_len emitted positions."""
    out = []
    for ci, (name, L) in enumerate(contigs):
        n = max(1, -(-L // region_len))
        step = -(-L // n)
        for k in range(n):
            s, e = k * step, min(L, (k + 1) * step)
            if s < e:
                out.append(Region(name, ci, L, s, e))
    return out


def assign_lpt(regions: Sequence[Region], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of regions to ranks (balances 25 unequal contigs).
    Deterministic; returns per-rank lists of region indices, each in (contig, position) order."""
    order = sorted(range(len(regions)), key=lambda i: (-regions[i].length, i))
    load = [0] * world_size
    mine: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += regions[i].length
        mine[r].append(i)
    for m in mine:
        m.sort()
    return mine


def read_range_for_region(read_pos: np.ndarray, max_span: int, region: Region) -> Tuple[int, int]:
    """Conservative [lo, hi) slice of a position-sorted read array that can overlap the region's computed span.
    Reads spanning a boundary are handed to both neighbours (at most one read length of duplication)."""
    lo = int(np.searchsorted(read_pos, region.start - max_span, side="left"))
    hi = int(np.searchsorted(read_pos, region.end, side="left"))
    return lo, hi


def merge_site_lists(parts: Sequence[Dict[str, np.ndarray]]) -> Dict[str, np.ndarray]:
    """Concatenates per-region results {contig_index, pos, ...} and orders them by (contig_index, pos)."""
    parts = [p for p in parts if p is not None and len(p["pos"])]
    if not parts:
        return {}
    keys = parts[0].keys()
    cat = {k: np.concatenate([p[k] for p in parts]) for k in keys}
    order = np.lexsort((cat["pos"], cat["contig_index"]))
    return {k: v[order] for k, v in cat.items()}


def gather_to_rank0(local_parts: List[Dict[str, np.ndarray]]):
    """torch.distributed plumbing: returns the merged site list on rank 0, None elsewhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_site_lists(local_parts)
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local_parts, gathered, dst=0)
    if dist.get_rank() != 0:
        return None
    flat = [p for parts in gathered for p in parts]
    return merge_site_lists(flat)
