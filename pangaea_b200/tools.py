"""The text tools either side of the featurization path, on the device.  Each function takes the
arguments of the reference command line it replaces and leaves the same files:

    sort_by_barcode      awk | LANG=C sort -k1,1 | cut -f2- | tr "\\t" "\\n"           src/run_pangaea:237-252
    preprocess_stlfr     bin/preprocess_stlfr -1 R1 -2 R2 -o PREFIX -n [-l]           src/cpptools/preprocess_stlfr.cpp
    preprocess_tellseq   bin/preprocess_tellseq -1 R1 -2 R2 -l I1 -o PREFIX           src/cpptools/preprocess_tellseq.cpp
    extract_reads        bin/extract_reads -i INTERLEAVED -c CLUSTERS -o PREFIX       src/cpptools/extract_reads.cpp

The bytes go through HBM (csrc/ingest.cuh, csrc/transform.cuh); the host reads and writes files.  Inputs ending in ``.gz``
are decompressed on the host, as the reference does with gzstream.
"""
from __future__ import annotations

import ctypes as C
import gzip

import numpy as np

from . import _lib


def _read(path) -> np.ndarray:
    if str(path).endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return np.frombuffer(f.read(), dtype=np.uint8)
    return np.fromfile(path, dtype=np.uint8)


def _ctx(ctx, device=0):
    return ctx or _lib.Context(device=device, table_mode=_lib.PG_TABLE_NONE)


def sort_by_barcode(input_fastq, output_sorted, ctx=None):
    """run_pangaea:237-252: interleaved FASTQ -> the same records sorted by BX:Z: tag (untagged last)."""
    ctx = _ctx(ctx)
    text = _read(input_fastq)
    out = ctx.sort_fastq_by_barcode(text)
    with open(output_sorted, "wb") as f:
        f.write(out)
    return output_sorted


def preprocess_stlfr(reads1, reads2, output, number=True, library=False, ctx=None):
    """preprocess_stlfr -n [-l]: writes <output>_1.fq and <output>_2.fq."""
    if not number:
        raise NotImplementedError("the whitelist mode of preprocess_stlfr (without -n) is not part of this path; run_pangaea passes -n")
    ctx = _ctx(ctx)
    a, b = _read(reads1), _read(reads2)
    L = _lib.lib()
    cap1, cap2 = len(a) + len(a) // 4 + 64, len(b) + len(a) // 2 + 64
    for _ in range(2):
        o1, o2 = np.empty(cap1, np.uint8), np.empty(cap2, np.uint8)
        n1, n2 = C.c_int64(0), C.c_int64(0)
        rc = L.pg_preprocess_stlfr(ctx.h, a.ctypes.data if len(a) else None, len(a), b.ctypes.data if len(b) else None, len(b), int(library),
                                   o1.ctypes.data, cap1, C.byref(n1), o2.ctypes.data, cap2, C.byref(n2))
        if rc != 0 and (n1.value > cap1 or n2.value > cap2):
            cap1, cap2 = n1.value + 64, n2.value + 64
            continue
        ctx._ck(rc)
        break
    o1[: n1.value].tofile(output + "_1.fq")
    o2[: n2.value].tofile(output + "_2.fq")
    return output + "_1.fq", output + "_2.fq"


def preprocess_tellseq(reads1, reads2, l1, output, ctx=None):
    """preprocess_tellseq: writes <output>_1.fq, <output>_2.fq and <output>.wl."""
    ctx = _ctx(ctx)
    a, b, i = _read(reads1), _read(reads2), _read(l1)
    L = _lib.lib()
    cap1, cap2, capw = len(a) + len(a) // 2 + 64, len(b) + len(a) // 2 + 64, len(i) + 64
    o1, o2, ow = np.empty(cap1, np.uint8), np.empty(cap2, np.uint8), np.empty(capw, np.uint8)
    n1, n2, nw = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    ptr = lambda x: x.ctypes.data if len(x) else None
    ctx._ck(L.pg_preprocess_tellseq(ctx.h, ptr(a), len(a), ptr(b), len(b), ptr(i), len(i), o1.ctypes.data, cap1, C.byref(n1), o2.ctypes.data, cap2,
                                    C.byref(n2), ow.ctypes.data, capw, C.byref(nw)))
    o1[: n1.value].tofile(output + "_1.fq")
    o2[: n2.value].tofile(output + "_2.fq")
    ow[: nw.value].tofile(output + ".wl")
    return output + "_1.fq", output + "_2.fq", output + ".wl"


def parse_clusters(path):
    """extract_reads.cpp:58-84 restated: -> (cluster names in file order without "-1" lines, {barcode: cluster index}).
    A later line overwrites the cluster of a barcode it repeats."""
    names, barcode2cluster = [], {}
    with open(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    for line in lines:
        pos = line.find(b"\t")
        name = line if pos < 0 else line[:pos]
        if name == b"-1":
            continue
        cid = len(names)
        names.append(name)
        # the reference walks the rest of the line, one comma-separated barcode at a time (a line without a tab wraps
        # std::string::npos to 0 and reads the whole line as one barcode)
        rest = line if pos < 0 else line[pos + 1:]
        for bc in rest.split(b","):  # ("name\t" and "" map the empty barcode, exactly as the reference's loop does)
            barcode2cluster[bc] = cid
    return names, barcode2cluster


def extract_reads(interleaved, clusters, output, ctx=None):
    """extract_reads -i: writes <output>_bin<cluster>.fq and <output>_bin<cluster>.barcode for every cluster of the tsv."""
    ctx = _ctx(ctx)
    names, b2c = parse_clusters(clusters)
    text = _read(interleaved)
    L = _lib.lib()
    h = _lib._vp()
    ctx._ck(L.pg_extract_open(ctx.h, text.ctypes.data if len(text) else None, len(text), C.byref(h)))
    try:
        n_runs = int(L.pg_extract_n_runs(h))
        need = int(L.pg_extract_run_labels(h, None, 0, None))
        buf, off = np.empty(max(need, 1), np.uint8), np.empty(n_runs + 1, np.int64)
        L.pg_extract_run_labels(h, buf.ctypes.data, need, off.ctypes.data)
        blob, o = buf[:need].tobytes(), off.tolist()
        cl = np.array([b2c.get(blob[o[r]:o[r + 1]], -1) for r in range(n_runs)], dtype=np.int32)
        nc = len(names)
        fq_start, bc_start = np.zeros(nc + 1, np.int64), np.zeros(nc + 1, np.int64)
        ctx._ck(L.pg_extract_route(ctx.h, h, cl.ctypes.data, nc, fq_start.ctypes.data, bc_start.ctypes.data))
        fq, bc = np.empty(max(int(fq_start[-1]), 1), np.uint8), np.empty(max(int(bc_start[-1]), 1), np.uint8)
        ctx._ck(L.pg_extract_copy(ctx.h, h, fq.ctypes.data, bc.ctypes.data))
    finally:
        L.pg_extract_close(ctx.h, h)
    written = []
    for c, name in enumerate(names):  # a name that repeats reopens (truncates) its files, as the reference's fstream does
        stem = output + "_bin" + name.decode("utf-8", "surrogateescape")
        fq[int(fq_start[c]): int(fq_start[c + 1])].tofile(stem + ".fq")
        bc[int(bc_start[c]): int(bc_start[c + 1])].tofile(stem + ".barcode")
        written.append(stem)
    return written
