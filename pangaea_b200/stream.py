"""Streaming featurization: a FASTQ of any size through the GPU in batches.

The reference streams its input one cloud at a time (count_kmer.cpp:236-282) and never holds more than a
few clouds in memory.  Here the unit is a batch (csrc/fastq.cpp: pg_fastq_stream_*), and because the
abundance histogram needs the GLOBAL k-mer counts, the flow is two passes:

    pass 1  for every batch: parse (host threads) -> pinned buffers -> H2D -> 2-bit pack -> count into the table.
            The next batch is parsed while the GPU works on the current one.  The packed batch (0.5 B per base)
            stays in HBM when it fits (pg_batch_compact); otherwise the file is parsed again in pass 2.
    pass 2  for every batch: cloud grouping, TNF, abundance look-ups -> rows.  Rows of consecutive batches are the
            reference's row order.

With one batch this is exactly pg_extract_features (shared partition, one upload).
"""
from __future__ import annotations

import queue
import threading

import numpy as np

from . import _lib

DEFAULT_BATCH_SEQ_BYTES = 6 << 30


def keep_partitions(ctx, seq_bytes_total) -> bool:
    """May the batches of a stream keep the partition of their k-mer windows (what pg_count2's keep_partition asks for) until the
    featurize pass?  Yes when all of them plus the packed batches fit in about half of the device memory - the featurize pass then
    skips its own partition (a fifth of its time); the rest is for entries in flight and the matrices.  A kept partition is
    allocated at region capacity: 4 B per base position x 1.5 slack x 1.08 padding = 6.5 B per sequence byte."""
    _, total = ctx.mem_info()
    return seq_bytes_total * (6.5 + 0.6) < 0.5 * total


class _Prefetch:
    """iterate a FastqStream one batch ahead on a feeder thread (pg_fastq_stream_next releases the GIL)"""

    def __init__(self, stream, depth=1):
        self.q = queue.Queue(maxsize=depth)
        self.t = threading.Thread(target=self._run, args=(stream,), daemon=True)
        self.t.start()

    def _run(self, stream):
        try:
            while True:
                fq = stream.next()
                self.q.put(fq)
                if fq is None:
                    return
        except BaseException as e:  # handed to the consumer
            self.q.put(e)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item


class DeviceIngest:
    """Batches of a plain-text INTERLEAVED FASTQ parsed ON THE DEVICE (csrc/ingest.cuh, pg_ingest_text): the host only
    moves raw file bytes into a pinned staging buffer (all cores) and from there over PCIe; the line index, getBarcode,
    the cloud flags and the 2-bit pack happen in HBM.  While the GPU works on window w the host stages window w + 1; the
    part of w behind its last cloud flush (or an incomplete record) is put in front of it.

    Iterating yields (Batch, keep uint8[], labels list[str], is_last_batch)."""

    def __init__(self, ctx, path, window_bytes=1 << 30, slack_bytes=64 << 20):
        self.ctx, self.path, self.window, self.slack = ctx, path, int(window_bytes), int(slack_bytes)

    def __iter__(self):
        import torch

        ctx = self.ctx
        n = _size(self.path)
        if n == 0:
            batch, labels, keep, _, _ = ctx.ingest_text(b"", b"", 0, final=True)
            yield batch, keep, labels, True
            return
        L = _lib.lib()
        window, slack = self.window, self.slack
        state = {}

        def alloc():
            state["bufs"] = [torch.empty(slack + window, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            state["views"] = [b.numpy() for b in state["bufs"]]

        fs_path = _lib.os.fsencode(self.path)

        debug = _lib.os.environ.get("PG_INGEST_DEBUG") == "1"

        def stage(k, lo, hi):  # file bytes [lo, hi) -> views[k][slack : slack + hi - lo]: pread on all cores, no mapping
            import time

            t0 = time.perf_counter()
            if hi > lo and L.pg_parallel_pread(fs_path, lo, hi - lo, state["views"][k][slack:].ctypes.data) != 0:
                state["error"] = _lib.PgError(-4, f"cannot read {self.path!r}")  # (may run on the staging thread)
            if debug:
                dt = time.perf_counter() - t0
                print(f"[ingest] staged {hi - lo} bytes in {1e3 * dt:.1f} ms ({(hi - lo) / 1e9 / max(dt, 1e-9):.1f} GB/s)", file=_lib.sys.stderr, flush=True)

        alloc()
        last, rt = b"", 0
        k, head = 0, 0                      # head: bytes in front of views[k][slack] that open the chunk (tail of the previous one)
        lo, hi = 0, min(n, window)          # file range staged at views[k][slack:]
        stage(k, lo, hi)
        while True:
            final = hi >= n
            chunk = state["views"][k][slack - head: slack + (hi - lo)]
            nxt, nlo, nhi = None, hi, min(n, hi + window)
            if not final:                   # stage the next window while the GPU parses this one
                nxt = threading.Thread(target=stage, args=(k ^ 1, nlo, nhi))
                nxt.start()
            batch, labels, keep, consumed, rt = ctx.ingest_text(chunk, last, rt, final=final, n_bytes=len(chunk))
            if nxt:
                nxt.join()
            if "error" in state:
                raise state["error"]
            restart = None
            if batch is None:               # no cloud flush inside the chunk (a cloud larger than the window): take a larger one
                if final:
                    raise _lib.PgError(-5, "device ingest made no progress on the last chunk")
                window *= 2
                restart = lo - head
            else:
                last = labels[-1].encode("utf-8", "surrogateescape")
                yield batch, keep, labels, final
                if final:
                    return
                tail = len(chunk) - consumed    # bytes of this chunk that the next batch starts with
                if tail > slack:
                    slack = 2 * tail
                    restart = lo - head + consumed
                else:
                    if tail:
                        state["views"][k ^ 1][slack - tail: slack] = chunk[consumed:]
                    k, head, lo, hi = k ^ 1, tail, nlo, nhi
            if restart is not None:         # rare: new buffers, staged synchronously from `restart`
                del chunk
                alloc()
                k, head, lo, hi = 0, 0, restart, min(n, restart + window)
                stage(k, lo, hi)


def _size(path):
    import os

    return os.path.getsize(path)


def extract_features_streaming(ctx: "_lib.Context", open_stream, clear_table=True, reduce_table=None, resident_fraction=0.45, seq_bytes_hint=None):
    """Whole path over a stream of batches.

    open_stream: callable -> a fresh _lib.FastqStream (called again for pass 2 when the packed batches do not fit).
    reduce_table: optional callable run between the passes (the multi-GPU all-reduce of the count tables).
    Returns (names list[str], Features) - Features holds the rows of all batches on the device."""
    if clear_table:
        ctx.table_clear()
    _, total = ctx.mem_info()
    budget = int(total * resident_fraction)
    stream = open_stream()
    first = stream.next()
    if first is None:  # empty input
        stream.close()
        empty = _lib.make_reads(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.uint8))
        b = ctx.upload_count(empty, keep_partition=False)
        if reduce_table:
            reduce_table()
        f = ctx.featurize(b, np.zeros(1, np.uint8))
        b.free()
        return [], f
    second = stream.next()
    if second is None:
        # one batch: keep the partition for the featurize pass (the fast path of pg_extract_features)
        stream.close()
        b = ctx.upload_count(first.reads, keep_partition=True)
        if reduce_table:
            reduce_table()
        f = ctx.featurize(b, first.group_keep, first.n_groups)
        labels = first.labels()
        names = [labels[g] for g in f.row_groups().tolist()]
        b.free()
        first.close()
        return names, f

    # ---- pass 1: count every batch ----
    held, resident, keep_resident = [], 0, True
    keep_part = seq_bytes_hint is not None and keep_partitions(ctx, seq_bytes_hint)

    def count_one(fq):
        nonlocal resident, keep_resident
        b = ctx.upload_count(fq.reads, keep_partition=keep_part)
        ctx.synchronize()  # the copies read fq's host buffers
        packed = fq.reads.n_bytes // 2 + 9 * fq.reads.n_reads
        if keep_resident and resident + packed > budget:
            keep_resident = False
            for hb in held:
                hb[0].free()
                hb[0] = None
        if keep_resident:
            b.compact()
            resident += packed
        else:
            b.free()
            b = None
        keep = np.ctypeslib.as_array(_lib.C.cast(fq.group_keep, _lib.C.POINTER(_lib.C.c_uint8)), shape=(fq.n_groups,)).copy()
        held.append([b, keep, fq.labels()])
        fq.close()

    count_one(first)
    count_one(second)
    for fq in _Prefetch(stream):
        count_one(fq)
    stream.close()
    if reduce_table:
        reduce_table()

    # ---- pass 2: featurize every batch ----
    names, parts = [], []
    if keep_resident:
        for b, keep, labels in held:
            f = ctx.featurize(b, keep)
            names += [labels[g] for g in f.row_groups().tolist()]
            b.free()
            parts.append(f)
    else:
        stream = open_stream()
        for i, fq in enumerate(_Prefetch(stream)):
            _, keep, labels = held[i]
            if fq.n_groups != len(keep):
                raise _lib.PgError(-5, "the input changed between the two passes")
            b = ctx.upload(fq.reads)
            f = ctx.featurize(b, keep)
            ctx.synchronize()
            names += [labels[g] for g in f.row_groups().tolist()]
            b.free()
            fq.close()
            parts.append(f)
        stream.close()
    feats = ctx.concat_features(parts)
    for f in parts:
        f.free()
    return names, feats


def extract_features_device_ingest(ctx: "_lib.Context", path, window_bytes=1 << 30, clear_table=True, reduce_table=None, resident_fraction=0.45,
                                   byte_range=None):
    """The same two-pass flow with the DEVICE parser (DeviceIngest) as the source of batches - plain-text interleaved FASTQ
    only; gzip, paired or hostile input goes through extract_features_streaming.  Returns (names list[str], Features)."""
    if clear_table:
        ctx.table_clear()
    _, total = ctx.mem_info()
    budget = int(total * resident_fraction)
    held, resident, keep_resident = [], 0, True
    keep_part = keep_partitions(ctx, 0.45 * _size(path))  # (sequence lines are ~40 % of a FASTQ file's bytes)
    for batch, keep, labels, is_last in DeviceIngest(ctx, path, window_bytes):
        single = is_last and not held
        ctx.count(batch, keep_partition=single or keep_part)
        n_reads, n_bytes = batch.shape()
        packed = n_bytes // 2 + 9 * n_reads
        if keep_resident and not single and resident + packed > budget:
            keep_resident = False
            for hb in held:
                hb[0].free()
                hb[0] = None
        if keep_resident:
            if not single:
                batch.compact()
            resident += packed
        else:
            batch.free()
            batch = None
        held.append([batch, keep, labels])
    if reduce_table:
        reduce_table()
    names, parts = [], []
    if keep_resident:
        for b, keep, labels in held:
            f = ctx.featurize(b, keep)
            names += [labels[g] for g in f.row_groups().tolist()]
            b.free()
            parts.append(f)
    else:
        for i, (b, keep, labels, _) in enumerate(DeviceIngest(ctx, path, window_bytes)):
            if len(keep) != len(held[i][1]):
                raise _lib.PgError(-5, "the input changed between the two passes")
            f = ctx.featurize(b, keep)
            names += [labels[g] for g in f.row_groups().tolist()]
            b.free()
            parts.append(f)
    if len(parts) == 1:
        return names, parts[0]
    feats = ctx.concat_features(parts)
    for f in parts:
        f.free()
    return names, feats
