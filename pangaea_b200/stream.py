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
    moves raw file bytes; the line index, getBarcode, the cloud flags and the 2-bit pack happen in HBM.

    Three stages run concurrently on consecutive windows of the file: (1) `pread` by all host cores into one of two pinned
    buffers (~38 GB/s on the 16-core box), (2) the H2D copy of the window on a copy stream (PCIe), (3) the device parse and
    whatever the consumer does with the batch (the count pass).  The part of a window behind its last cloud flush - or an
    incomplete record - is moved, device to device, in front of the next window's bytes.

    Iterating yields (Batch, keep uint8[], labels list[str], is_last_batch)."""

    def __init__(self, ctx, path, window_bytes=1 << 30, slack_bytes=64 << 20):
        self.ctx, self.path, self.window, self.slack = ctx, path, int(window_bytes), int(slack_bytes)

    def __iter__(self):
        import time

        import torch

        ctx = self.ctx
        n = _size(self.path)
        if n == 0:
            batch, labels, keep, _, _ = ctx.ingest_text(b"", b"", 0, final=True)
            yield batch, keep, labels, True
            return
        L = _lib.lib()
        W, slack = max(1, min(self.window, n)), self.slack  # (a small file: one window of its own size, not a 1 GiB pinned buffer)
        n_win = (n + W - 1) // W
        dev_name = f"cuda:{ctx.params.device}"
        host = [torch.empty(W, dtype=torch.uint8, pin_memory=True) for _ in range(min(2, n_win))]
        dev = [torch.empty(slack + W, dtype=torch.uint8, device=dev_name) for _ in range(min(2, n_win))]
        copy_stream = torch.cuda.Stream(device=dev_name)
        fs_path = _lib.os.fsencode(self.path)
        debug = _lib.os.environ.get("PG_INGEST_DEBUG") == "1"
        issued = [threading.Event() for _ in range(n_win)]   # the H2D of window w is queued (h2d[w] is its CUDA event)
        parsed = [threading.Event() for _ in range(n_win)]   # the parse of window w has returned: its device buffer may be refilled
        h2d = [None] * n_win
        failure = []

        def producer():
            try:
                for w in range(n_win):
                    k = w % 2
                    if w >= 2:
                        h2d[w - 2].synchronize()   # the pinned buffer is read no more
                        parsed[w - 2].wait()       # the device buffer is parsed (and its tail moved on)
                        if failure:
                            return
                    lo, hi = w * W, min(n, (w + 1) * W)
                    t0 = time.perf_counter()
                    if L.pg_parallel_pread(fs_path, lo, hi - lo, host[k].data_ptr()) != 0:
                        raise _lib.PgError(-4, f"cannot read {self.path!r}")
                    if debug:
                        dt = time.perf_counter() - t0
                        print(f"[ingest] staged {hi - lo} bytes in {1e3 * dt:.1f} ms ({(hi - lo) / 1e9 / max(dt, 1e-9):.1f} GB/s)", file=_lib.sys.stderr, flush=True)
                    with torch.cuda.stream(copy_stream):
                        dev[k][slack: slack + hi - lo].copy_(host[k][: hi - lo], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    h2d[w] = ev
                    issued[w].set()
            except BaseException as e:  # handed to the consumer
                failure.append(e)
                for ev in issued:
                    ev.set()

        feeder = threading.Thread(target=producer, daemon=True)
        feeder.start()
        last, rt, tail = b"", 0, 0
        try:
            for w in range(n_win):
                k = w % 2
                issued[w].wait()
                if failure:
                    raise failure[0]
                h2d[w].synchronize()
                final = w == n_win - 1
                n_w = min(n, (w + 1) * W) - w * W
                chunk_len = tail + n_w
                base = dev[k].data_ptr() + slack - tail
                batch, labels, keep, consumed, rt = ctx.ingest_text(base, last, rt, final=final, device_text=True, n_bytes=chunk_len)
                if batch is None:  # no cloud flush inside the window: everything is carried into the next one
                    if final:
                        raise _lib.PgError(-5, "device ingest made no progress on the last window")
                    consumed = 0
                new_tail = chunk_len - consumed
                if not final:
                    if new_tail > slack:
                        raise _WindowTooSmall(new_tail)
                    if new_tail:  # device to device, in front of the next window's bytes (a different region than its H2D writes)
                        src = dev[k][slack - tail + consumed: slack - tail + chunk_len]
                        dev[k ^ 1][slack - new_tail: slack].copy_(src)
                        torch.cuda.current_stream(dev_name).synchronize()
                tail = new_tail
                parsed[w].set()
                if batch is not None:
                    last = labels[-1].encode("utf-8", "surrogateescape")
                    yield batch, keep, labels, final
        finally:
            if not failure:
                failure.append(GeneratorExit())  # tells the producer to stop at its next wait
            for ev in parsed:
                ev.set()
            feeder.join(timeout=60)


class _WindowTooSmall(Exception):
    """a cloud (or the run behind a window's last flush) does not fit the slack in front of a window"""


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
                                   slack_bytes=64 << 20):
    """The same two-pass flow with the DEVICE parser (DeviceIngest) as the source of batches - plain-text interleaved FASTQ
    only; gzip, paired or hostile input goes through extract_features_streaming.  Returns (names list[str], Features)."""
    while True:
        try:
            return _device_ingest_two_pass(ctx, path, window_bytes, slack_bytes, clear_table, reduce_table, resident_fraction)
        except _WindowTooSmall as e:  # rare: start over with room for the longest carried tail seen
            slack_bytes = max(2 * slack_bytes, 2 * int(e.args[0]))
            window_bytes = max(window_bytes, slack_bytes)
            clear_table = True


def _device_ingest_two_pass(ctx, path, window_bytes, slack_bytes, clear_table, reduce_table, resident_fraction):
    if clear_table:
        ctx.table_clear()
    _, total = ctx.mem_info()
    budget = int(total * resident_fraction)
    import time

    debug = _lib.os.environ.get("PG_INGEST_DEBUG") == "1"
    t_start = time.perf_counter()
    held, resident, keep_resident = [], 0, True
    keep_part = keep_partitions(ctx, 0.45 * _size(path))  # (sequence lines are ~40 % of a FASTQ file's bytes)
    for batch, keep, labels, is_last in DeviceIngest(ctx, path, window_bytes, slack_bytes):
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
    if debug:
        ctx.synchronize()
        print(f"[ingest] pass 1 (stage, copy, parse, count; {len(held)} batches): {1e3 * (time.perf_counter() - t_start):.1f} ms", file=_lib.sys.stderr, flush=True)
        t_start = time.perf_counter()
    names, parts = [], []
    if keep_resident:
        for b, keep, labels in held:
            f = ctx.featurize(b, keep)
            names += [labels[g] for g in f.row_groups().tolist()]
            b.free()
            parts.append(f)
        if debug:
            ctx.synchronize()
            print(f"[ingest] pass 2 (featurize, row labels): {1e3 * (time.perf_counter() - t_start):.1f} ms", file=_lib.sys.stderr, flush=True)
    else:
        for i, (b, keep, labels, _) in enumerate(DeviceIngest(ctx, path, window_bytes, slack_bytes)):
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
