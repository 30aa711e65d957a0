"""Sharding of one barcode-sorted read stream over ranks.

Clouds are independent once the global k-mer table exists, so reads shard by contiguous
ranges of CLOUDS.  A shard always ends right after a read that carries PG_READ_CHANGE - the
point where the reference flushes a cloud (count_kmer.cpp:251-270) - so every cloud,
including the pair its off-by-one steals from the next barcode, lies wholly inside one
shard, and concatenating the ranks' rows in rank order reproduces the single-process row
order (file order, count_kmer.cpp:283-292).  The only exchange step of the path is the sum
of the per-rank dense count tables (one all-reduce).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PG_READ_CHANGE = 1


@dataclass(frozen=True)
class Shard:
    rank: int
    read_lo: int   # reads [read_lo, read_hi)
    read_hi: int
    group_lo: int  # clouds [group_lo, group_hi) of the global numbering
    group_hi: int
    byte_lo: int
    byte_hi: int

    @property
    def n_reads(self):
        return self.read_hi - self.read_lo

    @property
    def n_groups(self):
        return self.group_hi - self.group_lo


def plan_shards(read_off: np.ndarray, read_flag: np.ndarray, n_ranks: int) -> list[Shard]:
    """Cut the stream into n_ranks contiguous shards of ~equal bytes at cloud boundaries.

    The last cloud of the stream (the reads after the last PG_READ_CHANGE) belongs to the
    last shard.  Ranks can end up empty when there are fewer clouds than ranks."""
    read_off = np.asarray(read_off, dtype=np.int64)
    read_flag = np.asarray(read_flag, dtype=np.uint8)
    n_reads = len(read_flag)
    n_groups = 1 + int((read_flag & PG_READ_CHANGE).sum())
    total = int(read_off[-1]) if n_reads else 0
    # candidate cut points: just after every flagged read
    cut_reads = np.flatnonzero(read_flag & PG_READ_CHANGE) + 1           # read index where the next cloud starts
    cut_bytes = read_off[cut_reads] if len(cut_reads) else np.zeros(0, np.int64)
    shards, read_lo, group_lo = [], 0, 0
    for r in range(n_ranks):
        if r == n_ranks - 1 or len(cut_reads) == 0:
            read_hi, group_hi = (n_reads, n_groups) if r == n_ranks - 1 else (read_lo, group_lo)
        else:
            target = total * (r + 1) // n_ranks
            i = int(np.searchsorted(cut_bytes, target, side="left"))     # first cut at or after the target
            i = min(i, len(cut_reads) - 1)
            read_hi = max(int(cut_reads[i]), read_lo)
            group_hi = max(i + 1, group_lo)                               # clouds 0..i end at or before this cut
        shards.append(Shard(r, read_lo, read_hi, group_lo, group_hi, int(read_off[read_lo]) if n_reads else 0,
                            int(read_off[read_hi]) if n_reads else 0))
        read_lo, group_lo = read_hi, group_hi
    return shards


def slice_shard(shard: Shard, seq: np.ndarray, read_off: np.ndarray, read_flag: np.ndarray, group_keep: np.ndarray, qual=None):
    """Views of one shard's arrays with offsets rebased to 0 (the layout pg_reads expects)."""
    off = np.ascontiguousarray(read_off[shard.read_lo:shard.read_hi + 1] - read_off[shard.read_lo])
    flag = read_flag[shard.read_lo:shard.read_hi]
    # Locally the shard has 1 + (its own PG_READ_CHANGE flags) clouds - that is what pg_featurize checks.  A shard that is
    # not the last one ends on a flush, which locally opens one more, empty (dropped) cloud; the last shard's slice of
    # group_keep already holds the trailing cloud of the stream; an empty rank owns one empty cloud.
    n_local = 1 + int((np.asarray(flag) & PG_READ_CHANGE).sum())
    keep = np.zeros(n_local, dtype=np.uint8)
    own = np.asarray(group_keep[shard.group_lo:shard.group_hi], dtype=np.uint8)[:n_local]
    keep[:len(own)] = own
    s = seq[shard.byte_lo:shard.byte_hi]
    q = qual[shard.byte_lo:shard.byte_hi] if qual is not None else None
    return s, off, flag, keep, q
