"""ctypes binding of include/pangaea_b200.h (the C-ABI of libpangaea_b200.so).

No fallback of any kind lives here: if the shared library is missing the import
raises, and if there is no CUDA device ``Context()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PG_LIB_PATH") or os.path.join(HERE, "libpangaea_b200.so")  # PG_LIB_PATH: A/B builds (tools/)

PG_READ_CHANGE, PG_READ_NOFEAT = 1, 2
PG_TABLE_AUTO, PG_TABLE_DENSE, PG_TABLE_HASH, PG_TABLE_NONE = 0, 1, 2, 3
T_PACK, T_COUNT, T_GROUP, T_FEAT, T_NORM, T_ALL, T_COUNT_SCATTER, T_FEAT_SCATTER, T_TNF, T_COUNT_SPLIT, T_COLLECT = range(11)
ABD_RAW, TNF_RAW, ABD, TNF, WEIGHTS = range(5)
PG_FQ_QUAL, PG_FQ_PINNED = 1, 2


class PgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pangaea_b200 error {code}: {msg}")
        self.code = code


class pg_params(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("k", C.c_int32), ("tnf_k", C.c_int32), ("window_size", C.c_int32),
        ("vector_size", C.c_int32), ("min_length", C.c_int64), ("min_qual_char", C.c_int32),
        ("table_mode", C.c_int32), ("table_capacity", C.c_uint64),
    ]


class pg_reads(C.Structure):
    _fields_ = [
        ("seq", C.c_void_p), ("qual", C.c_void_p), ("read_off", C.c_void_p), ("read_flag", C.c_void_p),
        ("n_reads", C.c_int64), ("n_bytes", C.c_int64),
    ]


# every symbol include/pangaea_b200.h declares: (restype, argtypes)
_vp, _i64, _i32, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_int
_P = C.POINTER
SIGNATURES = {
    "pg_default_params": (None, [_P(pg_params)]),
    "pg_create": (_int, [_P(pg_params), _P(_vp)]),
    "pg_destroy": (None, [_vp]),
    "pg_last_error": (C.c_char_p, [_vp]),
    "pg_device_count": (_int, []),
    "pg_tnf_dim": (_int, [_int]),
    "pg_synchronize": (_int, [_vp]),
    "pg_stream": (_vp, [_vp]),
    "pg_batch_upload": (_int, [_vp, _P(pg_reads), _P(_vp)]),
    "pg_batch_adopt": (_int, [_vp, _P(pg_reads), _P(_vp)]),
    "pg_batch_free": (None, [_vp, _vp]),
    "pg_batch_n_groups": (_i64, [_vp]),
    "pg_batch_shape": (None, [_vp, _P(_i64), _P(_i64)]),
    "pg_batch_download": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "pg_parallel_memcpy": (_int, [_vp, _vp, _i64]),
    "pg_count": (_int, [_vp, _vp]),
    "pg_count2": (_int, [_vp, _vp, _int]),
    "pg_batch_upload_count": (_int, [_vp, _P(pg_reads), _int, _P(_vp)]),
    "pg_batch_compact": (_int, [_vp, _vp]),
    "pg_features_concat": (_int, [_vp, _P(_vp), _i32, _P(_vp)]),
    "pg_table_clear": (_int, [_vp]),
    "pg_table_set": (_int, [_vp, _vp, _vp, _i64]),
    "pg_table_get": (_int, [_vp, _vp, _vp, _i64]),
    "pg_table_size": (_int, [_vp, _P(_i64)]),
    "pg_table_export": (_int, [_vp, _vp, _vp, _i64, _P(_i64)]),
    "pg_table_dense_view": (_int, [_vp, _P(_vp), _P(_i64)]),
    "pg_table_wait_event": (_int, [_vp, _vp]),
    "pg_table_clamp": (_int, [_vp, C.c_uint32]),
    "pg_featurize": (_int, [_vp, _vp, _vp, _i64, _P(_vp)]),
    "pg_featurize2": (_int, [_vp, _vp, _vp, _i64, _int, _P(_vp)]),
    "pg_batch_n_words": (_i64, [_vp]),
    "pg_batch_window_keys": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "pg_keys_partition": (_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pg_table_add_keys": (_int, [_vp, _vp, _i64]),
    "pg_table_lookup_keys": (_int, [_vp, _vp, _i64, _vp]),
    "pg_features_add_counts": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "pg_features_free": (None, [_vp, _vp]),
    "pg_features_rows": (_i64, [_vp]),
    "pg_features_abd_dim": (_i32, [_vp]),
    "pg_features_tnf_dim": (_i32, [_vp]),
    "pg_features_row_groups": (_int, [_vp, _vp, _vp]),
    "pg_features_copy_raw": (_int, [_vp, _vp, _vp, _vp]),
    "pg_normalize": (_int, [_vp, _vp]),
    "pg_features_copy_normalized": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "pg_features_from_raw": (_int, [_vp, _vp, _vp, _i64, _i32, _i32, _P(_vp)]),
    "pg_features_dlpack": (_vp, [_vp, _vp, _int]),
    "pg_features_device_ptr": (_vp, [_vp, _int]),
    "pg_extract_features": (_int, [_vp, _P(pg_reads), _vp, _i64, _P(_vp)]),
    "pg_fastq_parse": (_int, [C.c_char_p, C.c_char_p, _int, _P(_vp)]),
    "pg_fastq_count_lines": (_int, [C.c_char_p, _i64, _i64, _P(_i64)]),
    "pg_fastq_stream_open": (_int, [C.c_char_p, C.c_char_p, _int, _i64, _i64, _i64, _P(_vp)]),
    "pg_fastq_stream_next": (_int, [_vp, _i64, _P(_vp)]),
    "pg_fastq_stream_close": (None, [_vp]),
    "pg_fastq_free": (None, [_vp]),
    "pg_fastq_reads": (None, [_vp, _P(pg_reads)]),
    "pg_fastq_n_groups": (_i64, [_vp]),
    "pg_fastq_group_keep": (_vp, [_vp]),
    "pg_fastq_group_label": (C.c_char_p, [_vp, _i64]),
    "pg_fastq_group_labels": (_i64, [_vp, _vp, _i64, _vp]),
    "pg_mem_info": (_int, [_vp, _P(_i64), _P(_i64)]),
    "pg_trim": (_int, [_vp]),
    "pg_ingest_text": (_int, [_vp, _vp, _i64, _int, C.c_char_p, _i64, _P(_i32), _P(_i64), _P(_vp), _P(_vp)]),
    "pg_ingest_n_groups": (_i64, [_vp]),
    "pg_ingest_group_keep": (_vp, [_vp]),
    "pg_ingest_group_labels": (_i64, [_vp, _vp, _i64, _vp]),
    "pg_ingest_free": (None, [_vp]),
    "pg_fastq_sort_by_barcode": (_int, [_vp, _vp, _i64, _vp, _i64, _P(_i64)]),
    "pg_parallel_pread": (_int, [C.c_char_p, _i64, _i64, _vp]),
    "pg_preprocess_stlfr": (_int, [_vp, _vp, _i64, _vp, _i64, _int, _vp, _i64, _P(_i64), _vp, _i64, _P(_i64)]),
    "pg_preprocess_tellseq": (_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _P(_i64), _vp, _i64, _P(_i64), _vp, _i64, _P(_i64)]),
    "pg_extract_open": (_int, [_vp, _vp, _i64, _P(_vp)]),
    "pg_extract_n_runs": (_i64, [_vp]),
    "pg_extract_run_labels": (_i64, [_vp, _vp, _i64, _vp]),
    "pg_extract_route": (_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "pg_extract_copy": (_int, [_vp, _vp, _vp, _vp]),
    "pg_extract_close": (None, [_vp, _vp]),
    "pg_sampler_create": (_int, [_vp, _vp, _i64, C.c_double, _P(_vp)]),
    "pg_sampler_free": (None, [_vp, _vp]),
    "pg_sampler_draw": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "pg_sampler_draw_unique_round": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, _P(_i64)]),
    "pg_features_gather": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "pg_synth_generate": (_int, [_vp, _i64, _i32, _i64, _vp, _vp, _i64, _i32, _i32, C.c_double, C.c_double, C.c_uint64, _vp, _vp, _vp]),
    "pg_synth_generate2": (_int, [_vp, _i64, _i32, _i64, _vp, _vp, _i64, _i32, _i32, C.c_double, C.c_double, C.c_uint64, _i64, _i64, _vp, _vp, _vp]),
    "pg_timing_reset": (_int, [_vp]),
    "pg_timing_get": (_int, [_vp, _int, _P(C.c_double), _P(_i64)]),
}

_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it was not built (python -m pangaea_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing - build it with `python -m pangaea_b200.build` (there is no CPU path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (host pinned or device)
        return a.data_ptr()
    return int(a)


def make_reads(seq, read_off, read_flag, qual=None, n_reads=None, n_bytes=None) -> pg_reads:
    """pg_reads over numpy arrays / torch tensors / raw addresses (caller keeps them alive)."""
    r = pg_reads()
    r.seq, r.qual, r.read_off, r.read_flag = _ptr(seq), _ptr(qual), _ptr(read_off), _ptr(read_flag)
    r.n_reads = int(n_reads if n_reads is not None else len(read_flag))
    r.n_bytes = int(n_bytes if n_bytes is not None else len(seq))
    r._keepalive = (seq, qual, read_off, read_flag)  # the struct only holds raw addresses
    return r


class Fastq:
    """One batch of the host FASTQ reader (csrc/fastq.cpp): replaces the getline loops of count_kmer.cpp:181-282.
    ``Fastq(path)`` parses the whole input as one batch; ``FastqStream`` yields batches."""

    def __init__(self, path1=None, path2=None, want_qual=False, pinned=False, handle=None):
        if handle is None:
            h = _vp()
            flags = (PG_FQ_QUAL if want_qual else 0) | (PG_FQ_PINNED if pinned else 0)
            rc = lib().pg_fastq_parse(os.fsencode(path1), os.fsencode(path2) if path2 else None, flags, C.byref(h))
            if rc != 0:
                raise PgError(rc, f"cannot read {path1!r}" + (f" / {path2!r}" if path2 else ""))
            handle = h
        self.h = handle
        self.reads = pg_reads()
        lib().pg_fastq_reads(self.h, C.byref(self.reads))
        self.n_groups = int(lib().pg_fastq_n_groups(self.h))
        self.group_keep = lib().pg_fastq_group_keep(self.h)

    def label(self, g: int) -> str:
        return lib().pg_fastq_group_label(self.h, g).decode("utf-8", "surrogateescape")

    def labels(self):
        """all cloud labels as a list of str (one bulk call instead of n_groups ctypes calls)"""
        need = int(lib().pg_fastq_group_labels(self.h, None, 0, None))
        buf = np.empty(max(need, 1), dtype=np.uint8)
        off = np.empty(self.n_groups + 1, dtype=np.int64)
        lib().pg_fastq_group_labels(self.h, buf.ctypes.data, need, off.ctypes.data)
        blob = buf[:need].tobytes()
        o = off.tolist()
        return [blob[o[g]:o[g + 1]].decode("utf-8", "surrogateescape") for g in range(self.n_groups)]

    def arrays(self):
        """numpy views (seq, read_off, read_flag, keep) - valid while self lives."""
        r = self.reads
        as_np = lambda p, n, t: np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(n,)) if n else np.zeros(0, dtype=t)
        return (as_np(r.seq, r.n_bytes, C.c_uint8), as_np(r.read_off, r.n_reads + 1, C.c_int64),
                as_np(r.read_flag, r.n_reads, C.c_uint8), as_np(self.group_keep, self.n_groups, C.c_uint8))

    def close(self):
        if getattr(self, "h", None):
            lib().pg_fastq_free(self.h)
            self.h = None

    __del__ = close


def count_lines(path, byte_lo=0, byte_hi=-1) -> int:
    n = _i64()
    rc = lib().pg_fastq_count_lines(os.fsencode(path), int(byte_lo), int(byte_hi), C.byref(n))
    if rc != 0:
        raise PgError(rc, f"cannot read {path!r}")
    return int(n.value)


class FastqStream:
    """Batches of about ``target_seq_bytes`` sequence bytes, each ending at a cloud flush (pg_fastq_stream_*).
    ``byte_lo`` / ``byte_hi`` / ``lines_before_lo``: this rank's byte range of a plain-text interleaved file."""

    def __init__(self, path1, path2=None, want_qual=False, pinned=False, target_seq_bytes=0, byte_lo=0, byte_hi=-1, lines_before_lo=0):
        h = _vp()
        flags = (PG_FQ_QUAL if want_qual else 0) | (PG_FQ_PINNED if pinned else 0)
        rc = lib().pg_fastq_stream_open(os.fsencode(path1), os.fsencode(path2) if path2 else None, flags, int(byte_lo), int(byte_hi),
                                        int(lines_before_lo), C.byref(h))
        if rc != 0:
            raise PgError(rc, f"cannot read {path1!r}" + (f" / {path2!r}" if path2 else ""))
        self.h, self.target, self.path1 = h, int(target_seq_bytes), path1

    def next(self):
        """-> Fastq, or None at the end of the stream (ctypes releases the GIL: call it from a feeder thread)."""
        out = _vp()
        rc = lib().pg_fastq_stream_next(self.h, self.target, C.byref(out))
        if rc != 0:
            raise PgError(rc, f"error while reading {self.path1!r} (I/O error, truncated gzip stream or out of memory)")
        return Fastq(handle=out) if out.value else None

    def __iter__(self):
        while True:
            fq = self.next()
            if fq is None:
                return
            yield fq

    def close(self):
        if getattr(self, "h", None):
            lib().pg_fastq_stream_close(self.h)
            self.h = None

    __del__ = close


class Features:
    """Per-cloud matrices resident in HBM (pg_features)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle
        L = lib()
        self.rows = int(L.pg_features_rows(handle))
        self.abd_dim = int(L.pg_features_abd_dim(handle))
        self.tnf_dim = int(L.pg_features_tnf_dim(handle))

    def row_groups(self) -> np.ndarray:
        out = np.empty(self.rows, dtype=np.int64)
        self.ctx._ck(lib().pg_features_row_groups(self.ctx.h, self.h, out.ctypes.data))
        return out

    def raw(self):
        """(abd int32 [rows, v], tnf int32 [rows, 136]) copied to host."""
        abd = np.empty((self.rows, self.abd_dim), dtype=np.int32)
        tnf = np.empty((self.rows, self.tnf_dim), dtype=np.int32)
        self.ctx._ck(lib().pg_features_copy_raw(self.ctx.h, self.h, abd.ctypes.data, tnf.ctypes.data))
        return abd, tnf

    def normalize(self):
        self.ctx._ck(lib().pg_normalize(self.ctx.h, self.h))
        return self

    def normalized(self, abd=None, tnf=None, weights=None):
        """(abd f32, tnf f32, weights f64) copied to host (optionally into given pinned arrays)."""
        self.normalize()
        abd = np.empty((self.rows, self.abd_dim), dtype=np.float32) if abd is None else abd
        tnf = np.empty((self.rows, self.tnf_dim), dtype=np.float32) if tnf is None else tnf
        weights = np.empty(self.rows, dtype=np.float64) if weights is None else weights
        self.ctx._ck(lib().pg_features_copy_normalized(self.ctx.h, self.h, _ptr(abd), _ptr(tnf), _ptr(weights)))
        return abd, tnf, weights

    def dlpack(self, which: int):
        """PyCapsule "dltensor" over the device buffer (zero copy; torch.from_dlpack consumes it)."""
        if which >= ABD:
            self.normalize()
        p = lib().pg_features_dlpack(self.ctx.h, self.h, which)
        if not p:
            raise PgError(-1, self.ctx.last_error())
        new = C.pythonapi.PyCapsule_New
        new.restype, new.argtypes = C.py_object, [C.c_void_p, C.c_char_p, C.c_void_p]
        return new(p, b"dltensor", None)

    def torch(self, which: int):
        import torch

        return torch.from_dlpack(self.dlpack(which))

    def free(self):
        if getattr(self, "h", None):
            lib().pg_features_free(self.ctx.h, self.h)
            self.h = None

    __del__ = free


class Batch:
    def __init__(self, ctx: "Context", handle, keepalive=None):
        self.ctx, self.h, self._keep = ctx, handle, keepalive

    def shape(self):
        nr, nb = _i64(), _i64()
        lib().pg_batch_shape(self.h, C.byref(nr), C.byref(nb))
        return int(nr.value), int(nb.value)

    def download(self, want_seq=True):
        """(seq uint8[n_bytes] | None, read_off int64[n_reads + 1], read_flag uint8[n_reads]) copied back to the host"""
        nr, nb = self.shape()
        seq = np.empty(nb, dtype=np.uint8) if want_seq else None
        off, flag = np.empty(nr + 1, dtype=np.int64), np.empty(nr, dtype=np.uint8)
        self.ctx._ck(lib().pg_batch_download(self.ctx.h, self.h, seq.ctypes.data if want_seq and nb else None, off.ctypes.data, flag.ctypes.data if nr else None))
        return seq, off, flag

    def compact(self):
        """drop the ASCII bases (and a kept partition); the packed stream stays for pg_featurize"""
        self.ctx._ck(lib().pg_batch_compact(self.ctx.h, self.h))
        self._keep = None
        return self

    def free(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().pg_batch_free(self.ctx.h, self.h)
            self.h = None

    __del__ = free


class Context:
    """One CUDA device + stream + k-mer table (pg_ctx)."""

    def __init__(self, device=0, k=15, tnf_k=4, window_size=10, vector_size=400, min_length=2000,
                 min_qual_char=0, table_mode=PG_TABLE_AUTO, table_capacity=0):
        L = lib()
        p = pg_params()
        L.pg_default_params(C.byref(p))
        p.device, p.k, p.tnf_k, p.window_size, p.vector_size = device, k, tnf_k, window_size, vector_size
        p.min_length, p.min_qual_char, p.table_mode, p.table_capacity = min_length, min_qual_char, table_mode, table_capacity
        self.params = p
        self.h = None
        h = _vp()
        rc = L.pg_create(C.byref(p), C.byref(h))
        if rc != 0:
            raise PgError(rc, L.pg_last_error(None).decode())
        self.h = h
        self.tnf_dim = int(L.pg_tnf_dim(tnf_k))

    def last_error(self) -> str:
        return lib().pg_last_error(self.h).decode()

    def _ck(self, rc):
        if rc != 0:
            raise PgError(rc, self.last_error())

    def close(self):
        if getattr(self, "h", None):
            lib().pg_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def stream(self) -> int:
        return int(lib().pg_stream(self.h) or 0)

    def synchronize(self):
        self._ck(lib().pg_synchronize(self.h))

    def mem_info(self):
        """(bytes this ctx can still get, total device bytes)"""
        f, t = _i64(), _i64()
        self._ck(lib().pg_mem_info(self.h, C.byref(f), C.byref(t)))
        return int(f.value), int(t.value)

    def trim(self):
        """release the idle device memory the ctx keeps for reuse (live batches / features / the table stay)"""
        self._ck(lib().pg_trim(self.h))

    # ---- batches -------------------------------------------------------------
    def upload(self, reads: pg_reads, keepalive=None) -> Batch:
        h = _vp()
        self._ck(lib().pg_batch_upload(self.h, C.byref(reads), C.byref(h)))
        return Batch(self, h, keepalive)

    def adopt(self, reads: pg_reads, keepalive=None) -> Batch:
        h = _vp()
        self._ck(lib().pg_batch_adopt(self.h, C.byref(reads), C.byref(h)))
        return Batch(self, h, keepalive)

    # ---- table ---------------------------------------------------------------
    def count(self, batch: Batch, keep_partition=True):
        self._ck(lib().pg_count2(self.h, batch.h, int(keep_partition)))

    def upload_count(self, reads: pg_reads, keep_partition=False, keepalive=None) -> Batch:
        """one batch of a stream: pipelined upload + count; the table is not cleared"""
        h = _vp()
        self._ck(lib().pg_batch_upload_count(self.h, C.byref(reads), int(keep_partition), C.byref(h)))
        return Batch(self, h, keepalive)

    def concat_features(self, parts) -> "Features":
        arr = (_vp * len(parts))(*[p.h for p in parts])
        h = _vp()
        self._ck(lib().pg_features_concat(self.h, arr, len(parts), C.byref(h)))
        return Features(self, h)

    def table_clear(self):
        self._ck(lib().pg_table_clear(self.h))

    def table_set(self, keys, counts):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        self._ck(lib().pg_table_set(self.h, keys.ctypes.data, counts.ctypes.data, len(keys)))

    def table_get(self, keys) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(len(keys), dtype=np.uint32)
        self._ck(lib().pg_table_get(self.h, keys.ctypes.data, out.ctypes.data, len(keys)))
        return out

    def table_size(self) -> int:
        n = _i64()
        self._ck(lib().pg_table_size(self.h, C.byref(n)))
        return int(n.value)

    def table_export(self):
        """(keys uint64 ascending - reference canonical form, counts uint32)."""
        n = self.table_size()
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        got = _i64()
        self._ck(lib().pg_table_export(self.h, keys.ctypes.data, counts.ctypes.data, n, C.byref(got)))
        return keys[: got.value], counts[: got.value]

    def table_dense_view(self):
        p, n = _vp(), _i64()
        self._ck(lib().pg_table_dense_view(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def table_clamp(self, max_count: int):
        """counter = min(counter, max_count) over the dense table (before an int32 sum across ranks)."""
        self._ck(lib().pg_table_clamp(self.h, int(max_count)))

    def table_wait_event(self, cuda_event: int):
        """Later table readers of this ctx wait for the CUDA event (raw cudaEvent_t, e.g. torch.cuda.Event.cuda_event)."""
        self._ck(lib().pg_table_wait_event(self.h, C.c_void_p(int(cuda_event))))

    def all_reduce_table(self, table_t, group=None):
        """Sum the dense count tables across ranks (NCCL) WITHOUT stalling this ctx: the collective runs on a side stream
        ordered after the count kernels, and only the kernels that read the table wait for it - cloud grouping and the TNF
        kernel of the next pg_featurize overlap the all-reduce.  Keep the returned objects alive until that call returns."""
        import torch
        import torch.distributed as dist

        dev = f"cuda:{self.params.device}"
        world = dist.get_world_size(group)
        if world > 1:  # counters saturate at 2^31 - 1: keep the int32 sum of `world` tables below bit 31 (pg_table_clamp)
            if self.params.window_size * self.params.vector_size > (2 ** 31 - 1) // world:
                raise PgError(-1, f"window_size * vector_size must be <= (2^31 - 1) / {world} ranks")
            self.table_clamp((2 ** 31 - 1) // world)
        counted = torch.cuda.Event()
        with torch.cuda.stream(torch.cuda.ExternalStream(self.stream, device=dev)):
            counted.record()
        side = getattr(self, "_side_stream", None) or torch.cuda.Stream(device=dev)
        self._side_stream = side
        side.wait_event(counted)
        with torch.cuda.stream(side):
            work = dist.all_reduce(table_t, group=group, async_op=True)
            work.wait()  # orders `side` after NCCL's stream; does not block the host
            done = torch.cuda.Event()
            done.record(side)
        self.table_wait_event(done.cuda_event)
        self._pending_reduce = (work, done, counted)
        return self._pending_reduce

    def table_as_torch(self):
        """The dense counter array as an int32 CUDA tensor sharing memory (for all_reduce)."""
        import torch

        self.synchronize()  # the ctx stream is non-blocking: torch's stream is not ordered after it
        ptr, n = self.table_dense_view()

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Arr(), device=f"cuda:{self.params.device}")

    # ---- features ------------------------------------------------------------
    def featurize(self, batch: Batch, group_keep, n_groups=None, no_abundance=False) -> Features:
        if isinstance(group_keep, np.ndarray):
            group_keep = np.ascontiguousarray(group_keep, dtype=np.uint8)
            n_groups = len(group_keep) if n_groups is None else n_groups
        h = _vp()
        self._ck(lib().pg_featurize2(self.h, batch.h, _ptr(group_keep), int(n_groups), 1 if no_abundance else 0, C.byref(h)))
        return Features(self, h)

    def extract_features(self, reads: pg_reads, group_keep, n_groups=None) -> Features:
        """Whole path from HOST buffers: upload, count, featurize, normalize."""
        if isinstance(group_keep, np.ndarray):
            group_keep = np.ascontiguousarray(group_keep, dtype=np.uint8)
            n_groups = len(group_keep) if n_groups is None else n_groups
        h = _vp()
        self._ck(lib().pg_extract_features(self.h, C.byref(reads), _ptr(group_keep), int(n_groups), C.byref(h)))
        return Features(self, h)

    def features_from_raw(self, abd, tnf) -> Features:
        abd = np.ascontiguousarray(abd, dtype=np.uint32)
        tnf = np.ascontiguousarray(tnf, dtype=np.uint32)
        h = _vp()
        self._ck(lib().pg_features_from_raw(self.h, abd.ctypes.data, tnf.ctypes.data, abd.shape[0], abd.shape[1], tnf.shape[1], C.byref(h)))
        return Features(self, h)

    # ---- FASTQ text on the device ---------------------------------------------
    def ingest_text(self, text, last_barcode=b"", read_type=0, final=True, device_text=False, n_bytes=None):
        """Device parse of a chunk of interleaved FASTQ text -> (Batch | None, labels list[str], keep uint8[], consumed, read_type)."""
        n = int(n_bytes if n_bytes is not None else len(text))
        rt, consumed, hb, hi = _i32(int(read_type)), _i64(0), _vp(), _vp()
        flags = (1 if final else 0) | (2 if device_text else 0)
        if isinstance(text, (bytes, bytearray)):
            holder = np.frombuffer(text, dtype=np.uint8)
            addr = holder.ctypes.data if n else None
        else:
            holder, addr = text, _ptr(text)
        lb = bytes(last_barcode)
        self._ck(lib().pg_ingest_text(self.h, addr, n, flags, lb, len(lb), C.byref(rt), C.byref(consumed), C.byref(hb), C.byref(hi)))
        if not hb.value:
            return None, [], np.zeros(0, np.uint8), 0, int(rt.value)
        L = lib()
        ng = int(L.pg_ingest_n_groups(hi))
        keep = np.ctypeslib.as_array(C.cast(L.pg_ingest_group_keep(hi), C.POINTER(C.c_uint8)), shape=(ng,)).copy()
        need = int(L.pg_ingest_group_labels(hi, None, 0, None))
        buf = np.empty(max(need, 1), dtype=np.uint8)
        off = np.empty(ng + 1, dtype=np.int64)
        L.pg_ingest_group_labels(hi, buf.ctypes.data, need, off.ctypes.data)
        blob, o = buf[:need].tobytes(), off.tolist()
        labels = [blob[o[g]:o[g + 1]].decode("utf-8", "surrogateescape") for g in range(ng)]
        L.pg_ingest_free(hi)
        return Batch(self, hb), labels, keep, int(consumed.value), int(rt.value)

    def sort_fastq_by_barcode(self, text) -> bytes:
        """run_pangaea:237-252 on the device: interleaved FASTQ text -> text sorted by BX:Z: tag."""
        src = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, dtype=np.uint8)
        n = len(src)
        out = np.empty(n + n // 64 + 16, dtype=np.uint8)  # tabs become newlines 1:1; a missing last newline adds one byte
        got = _i64(0)
        self._ck(lib().pg_fastq_sort_by_barcode(self.h, src.ctypes.data if n else None, n, out.ctypes.data, len(out), C.byref(got)))
        return out[: got.value].tobytes()

    # ---- instrumentation -------------------------------------------------------
    def timing_reset(self):
        self._ck(lib().pg_timing_reset(self.h))

    def timing(self, which=T_ALL):
        ms, n = C.c_double(), _i64()
        self._ck(lib().pg_timing_get(self.h, which, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)
