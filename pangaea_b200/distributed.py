"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

    rank r:  upload shard r -> pack -> count into the local dense table
             all_reduce(SUM) of the dense count tables            <- the ONE exchange step
             fused abundance + TNF over the rank's own clouds -> normalise
    rows stay sharded by cloud range (rank order == file order) or are gathered to rank 0.

Why an all-reduce and not the owner-partitioned all-to-all the north star sketches: for
k <= 16 the table is a dense array (k = 15: 2 GiB); summing it across 8 B200s over NVSwitch
moves 2 x 7/8 x 2 GiB per rank (~6 ms at the measured 725 GB/s bus bandwidth) and makes every
later look-up local, whereas routing 4 B per k-mer occurrence to its owner moves ~86 GB per
rank at C5 scale, twice.  The hash-table mode (k > 16) has no dense view and
is single-GPU in this round.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .shard import plan_shards, slice_shard


def extract_features_sharded(ctx: "_lib.Context", seq, read_off, read_flag, group_keep, qual=None, rank=None, world=None, group=None):
    """Whole path for this rank's shard of a host-resident stream.  Returns
    (Features for the shard's clouds, Shard).  Collective: every rank must call it."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    shard = plan_shards(read_off, read_flag, world)[rank]
    s, off, flag, keep, q = slice_shard(shard, seq, read_off, read_flag, np.asarray(group_keep, dtype=np.uint8), qual)
    reads = _lib.make_reads(np.ascontiguousarray(s), off, np.ascontiguousarray(flag), qual=None if q is None else np.ascontiguousarray(q))
    ctx.table_clear()
    batch = ctx.upload(reads)
    ctx.count(batch)
    if world > 1:
        table = ctx.table_as_torch()
        ctx.all_reduce_table(table, group=group)  # NCCL over NVLink / NVSwitch, overlapped with grouping + TNF below
    feats = ctx.featurize(batch, keep)
    feats.normalize()
    batch.free()
    return feats, shard


def extract_features_from_file(ctx: "_lib.Context", path, rank=None, world=None, group=None, want_qual=False, batch_seq_bytes=None):
    """Whole path for this rank's BYTE RANGE of a plain-text interleaved FASTQ: the file is parsed once in aggregate.

    Every rank counts the newlines of its 1/world of the file (pg_fastq_count_lines); the counts are all-gathered - the only
    host-side exchange - so each rank knows the line number at its cut and can snap to a record boundary and then to the
    next cloud flush (csrc/fastq.cpp: align_start; the rank before ends at the same point).  Each rank then streams its
    range through the GPU (pangaea_b200/stream.py); between the count pass and the featurize pass the dense tables are
    summed with one all-reduce.  Rows stay sharded: rank order == file order.
    Returns (names list[str], Features).  Collective: every rank must call it."""
    import os

    import torch
    import torch.distributed as dist

    from . import stream as stream_mod

    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    size = os.path.getsize(path)
    lo, hi = size * rank // world, size * (rank + 1) // world
    before = 0
    if world > 1:
        mine = torch.tensor([_lib.count_lines(path, lo, hi)], dtype=torch.int64, device=f"cuda:{ctx.params.device}" if dist.get_backend(group) == "nccl" else "cpu")
        counts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(counts, mine, group=group)
        before = int(sum(int(c) for c in counts[:rank]))
    table = ctx.table_as_torch() if world > 1 else None

    def reduce_table():
        if world > 1:
            ctx.all_reduce_table(table, group=group)

    def open_stream():
        return _lib.FastqStream(path, None, want_qual=want_qual, pinned=True, target_seq_bytes=batch_seq_bytes or stream_mod.DEFAULT_BATCH_SEQ_BYTES,
                                byte_lo=lo, byte_hi=hi if rank + 1 < world else -1, lines_before_lo=before)

    return stream_mod.extract_features_streaming(ctx, open_stream, reduce_table=reduce_table)


def _all_to_all(send, in_splits, out_splits, group):
    """all_to_all_single on device tensors; backends without a CUDA all-to-all (gloo, used by the one-GPU tests) go through the host"""
    import torch
    import torch.distributed as dist

    recv = torch.empty(int(sum(out_splits)), dtype=send.dtype, device=send.device)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, list(out_splits), list(in_splits), group=group)
        # the collective is ordered on torch's stream; what consumes `recv` runs on the ctx stream, which is not ordered after it
        torch.cuda.current_stream(send.device).synchronize()
        return recv
    h_send, h_recv = send.cpu(), torch.empty(int(sum(out_splits)), dtype=send.dtype)
    dist.all_to_all_single(h_recv, h_send, list(out_splits), list(in_splits), group=group)
    recv.copy_(h_recv)
    return recv


def extract_features_owner_partitioned(ctx: "_lib.Context", reads, group_keep, n_groups=None, group=None, seg_words=1 << 21):
    """The north star's multi-GPU table: owner = hash(canonical k-mer) mod n_ranks, one all-to-all of keys for count insertion,
    one all-to-all of keys + one of counts for the abundance queries (csrc/exchange.cuh).  This is the multi-rank form of the
    HASH table (k > 16) - no dense view exists to all-reduce - and works for the dense table as well.  `reads`: this rank's
    batch (pg_reads over host arrays, clouds whole - e.g. a shard of shard.plan_shards or a rank's byte range of a file).
    Every rank's table ends up holding only the k-mers it owns.  Collective: every rank must call it with the same seg_words.
    Returns Features for the rank's clouds."""
    import torch
    import torch.distributed as dist

    L = _lib.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = f"cuda:{ctx.params.device}"
    keep = np.ascontiguousarray(group_keep, dtype=np.uint8) if isinstance(group_keep, np.ndarray) else group_keep
    ctx.table_clear()
    batch = ctx.upload(reads)
    n_words = int(L.pg_batch_n_words(batch.h))
    # every rank runs the same number of exchange rounds (ranks with less data send empty segments)
    rounds = torch.tensor([(n_words + seg_words - 1) // seg_words], dtype=torch.int64, device=dev if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(rounds, op=dist.ReduceOp.MAX, group=group)
    rounds = int(rounds)

    def keys_by_owner(seg, feature):
        w0 = min(n_words, seg * seg_words)
        w1 = min(n_words, w0 + seg_words)
        n = 32 * (w1 - w0)
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        ordered = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        dest = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        counts = np.zeros(world, dtype=np.int64)
        ctx._ck(L.pg_batch_window_keys(ctx.h, batch.h, w0, w1, int(feature), keys.data_ptr()))
        ctx._ck(L.pg_keys_partition(ctx.h, keys.data_ptr(), n, world, ordered.data_ptr(), dest.data_ptr(), counts.ctypes.data))
        send_n = counts.tolist()
        c = torch.tensor(send_n, dtype=torch.int64, device=dev)
        recv_n = _all_to_all(c, [1] * world, [1] * world, group).tolist()  # how many keys every rank sends me
        return w0, w1, ordered[: sum(send_n)], dest, send_n, recv_n

    # ---- count: keys -> owners ----
    for seg in range(rounds):
        w0, w1, ordered, dest, send_n, recv_n = keys_by_owner(seg, feature=False)
        mine = _all_to_all(ordered, send_n, recv_n, group)
        ctx._ck(L.pg_table_add_keys(ctx.h, mine.data_ptr(), mine.numel()))
        ctx.synchronize()
    dist.barrier(group=group)  # every insert everywhere is done before the first query
    # ---- featurize: grouping + TNF locally, counts from the owners ----
    feats = ctx.featurize(batch, keep, n_groups, no_abundance=True)
    for seg in range(rounds):
        w0, w1, ordered, dest, send_n, recv_n = keys_by_owner(seg, feature=True)
        queries = _all_to_all(ordered, send_n, recv_n, group)
        answers = torch.empty(max(queries.numel(), 1), dtype=torch.int32, device=dev)
        ctx._ck(L.pg_table_lookup_keys(ctx.h, queries.data_ptr(), queries.numel(), answers.data_ptr()))
        back = _all_to_all(answers[: queries.numel()], recv_n, send_n, group)   # counts in the order the keys were sent
        ctx._ck(L.pg_features_add_counts(ctx.h, feats.h, batch.h, w0, w1, dest.data_ptr(), back.data_ptr() if back.numel() else dest.data_ptr()))
        ctx.synchronize()
    feats.normalize()
    batch.free()
    return feats


def gather_rows_named(names, feats: "_lib.Features", group=None):
    """Rank 0 gets (names, abundance int32, tnf int32) of all ranks in rank order (= file order); other ranks get None."""
    import torch.distributed as dist

    abd, tnf = feats.raw()
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object((list(names), abd, tnf), out, dst=0, group=group)
    if out is None:
        return None
    return (np.array(sum((p[0] for p in out), []), dtype=object), np.concatenate([p[1] for p in out]), np.concatenate([p[2] for p in out]))


def gather_rows(feats: "_lib.Features", shard, labels_of_group, group=None):
    """Rank 0 gets (names, abundance int32, tnf int32) for ALL clouds in file order; other
    ranks get None.  labels_of_group: callable global cloud index -> label."""
    import torch.distributed as dist

    abd, tnf = feats.raw()
    groups = feats.row_groups() + shard.group_lo
    payload = (groups, abd, tnf)
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(payload, out, dst=0, group=group)
    if out is None:
        return None
    groups = np.concatenate([p[0] for p in out])
    names = np.array([labels_of_group(int(g)) for g in groups], dtype=object)
    return names, np.concatenate([p[1] for p in out]), np.concatenate([p[2] for p in out])
