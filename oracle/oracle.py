"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker.  The
product package ``pangaea_b200`` never does.

Two checkers live here:

* ``libpg_oracle.so`` - the C restatement in ``oracle/pg_oracle.c`` (each function
  cites the reference file:line it follows);
* ``oracle/_ref/count_kmer`` and ``oracle/_ref/count_tnf`` - the UNMODIFIED reference
  tools compiled from ``/root/reference/src/cpptools`` by ``oracle/Makefile``.

Parity status: grouping / abundance / TNF are pinned against ``oracle/_ref``; the
jellyfish stage (global counts) is **parity unpinned** - see ``pg_oracle.c``.
All citations are relative to ``/root/reference/``.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpg_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_COUNT_KMER = os.path.join(REF_DIR, "count_kmer")
REF_COUNT_TNF = os.path.join(REF_DIR, "count_tnf")


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when the reference tree is mounted, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.access(REF_COUNT_KMER, os.X_OK) and os.access(REF_COUNT_TNF, os.X_OK)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        build(ref=False)
    L = C.CDLL(LIB_PATH)
    u64, i64, vp, cp = C.c_uint64, C.c_int64, C.c_void_p, C.c_char_p
    L.pgo_revcomp.restype = u64
    L.pgo_revcomp.argtypes = [u64, C.c_int]
    L.pgo_canonical.restype = u64
    L.pgo_canonical.argtypes = [u64, C.c_int]
    L.pgo_table_new.restype = vp
    L.pgo_table_free.argtypes = [vp]
    L.pgo_table_size.restype = u64
    L.pgo_table_size.argtypes = [vp]
    L.pgo_table_set.argtypes = [vp, u64, u64]
    L.pgo_table_add.argtypes = [vp, u64, u64]
    L.pgo_table_get.restype = C.c_int
    L.pgo_table_get.argtypes = [vp, u64, C.POINTER(u64)]
    L.pgo_table_items.argtypes = [vp, vp, vp]
    L.pgo_count_read.argtypes = [vp, cp, i64, cp, C.c_int, C.c_int]
    L.pgo_count_fastq.restype = C.c_int
    L.pgo_count_fastq.argtypes = [vp, cp, C.c_int, C.c_int]
    L.pgo_dump_write.restype = C.c_int
    L.pgo_dump_write.argtypes = [vp, cp, C.c_int]
    L.pgo_dump_load.restype = C.c_int
    L.pgo_dump_load.argtypes = [vp, cp, C.c_int]
    L.pgo_parse_header.argtypes = [C.POINTER(C.c_int), cp, cp, cp, C.c_int]
    L.pgo_tnf_lut.restype = C.c_int
    L.pgo_tnf_lut.argtypes = [C.c_int, vp]
    L.pgo_abundance.restype = vp
    L.pgo_abundance.argtypes = [cp, cp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pgo_tnf.restype = vp
    L.pgo_tnf.argtypes = [cp, cp, C.c_int, C.c_int]
    L.pgo_rows_free.argtypes = [vp]
    L.pgo_rows_n.restype = i64
    L.pgo_rows_n.argtypes = [vp]
    L.pgo_rows_dim.restype = i64
    L.pgo_rows_dim.argtypes = [vp]
    L.pgo_rows_label.restype = cp
    L.pgo_rows_label.argtypes = [vp, i64]
    L.pgo_rows_vals.restype = C.POINTER(C.c_double)
    L.pgo_rows_vals.argtypes = [vp]
    _lib = L
    return L


# ----------------------------------------------------------------------------
# k-mer arithmetic (count_kmer.cpp:11-21, :73-86)
# ----------------------------------------------------------------------------
CODE = {"A": 0, "C": 1, "T": 2, "G": 3}
LETTER = "ACTG"


def encode(kmer: str) -> int:
    v = 0
    for ch in kmer:
        v = (v << 2) | CODE[ch]
    return v


def decode(v: int, k: int) -> str:
    return "".join(LETTER[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def revcomp(v: int, k: int) -> int:
    return int(lib().pgo_revcomp(v, k))


def canonical(v: int, k: int) -> int:
    return int(lib().pgo_canonical(v, k))


def tnf_lut(k: int = 4):
    """(lut[4**k] -> column, n_columns); count_tnf.cpp:54-76,138-164."""
    lut = np.empty(4 ** k, dtype=np.int32)
    n = lib().pgo_tnf_lut(k, lut.ctypes.data)
    return lut, int(n)


def parse_header(line: str, read_type: int = 0):
    """getBarcode (count_kmer.cpp:25-53) -> (name, barcode, read_type_after)."""
    rt = C.c_int(read_type)
    name = C.create_string_buffer(4096)
    bc = C.create_string_buffer(4096)
    lib().pgo_parse_header(C.byref(rt), line.encode(), name, bc, 4096)
    return name.value.decode(), bc.value.decode(), rt.value


class Table:
    """k-mer -> count map (stands for the reference's unordered_map, count_kmer.cpp:139)."""

    def __init__(self):
        self.h = lib().pgo_table_new()

    def __del__(self):
        if getattr(self, "h", None):
            lib().pgo_table_free(self.h)
            self.h = None

    def __len__(self):
        return int(lib().pgo_table_size(self.h))

    def set(self, key: int, count: int):
        lib().pgo_table_set(self.h, key, count)

    def get(self, key: int):
        v = C.c_uint64()
        return int(v.value) if lib().pgo_table_get(self.h, key, C.byref(v)) else None

    def items(self):
        n = len(self)
        keys = np.empty(n, dtype=np.uint64)
        vals = np.empty(n, dtype=np.uint64)
        lib().pgo_table_items(self.h, keys.ctypes.data, vals.ctypes.data)
        order = np.argsort(keys)
        return keys[order], vals[order]

    def count_read(self, seq: bytes, k: int, qual: bytes | None = None, min_qual: int = 0):
        lib().pgo_count_read(self.h, seq, len(seq), qual, k, min_qual)

    def count_fastq(self, path: str, k: int, min_qual: int = 0):
        if lib().pgo_count_fastq(self.h, path.encode(), k, min_qual) != 0:
            raise FileNotFoundError(path)

    def write_dump(self, path: str, k: int):
        if lib().pgo_dump_write(self.h, path.encode(), k) != 0:
            raise OSError(path)

    def load_dump(self, path: str, k: int):
        if lib().pgo_dump_load(self.h, path.encode(), k) != 0:
            raise FileNotFoundError(path)


def count_fastq(paths, k: int = 15, min_qual: int = 0) -> Table:
    """jellyfish count -C stand-in over one or more FASTQ files (feature.py:76-94)."""
    t = Table()
    for p in [paths] if isinstance(paths, str) else paths:
        t.count_fastq(p, k, min_qual)
    return t


def _rows(h):
    L = lib()
    n, d = int(L.pgo_rows_n(h)), int(L.pgo_rows_dim(h))
    labels = np.array([L.pgo_rows_label(h, i).decode() for i in range(n)], dtype=object)
    vals = np.ctypeslib.as_array(L.pgo_rows_vals(h), shape=(n, d)).copy() if n else np.zeros((0, d))
    L.pgo_rows_free(h)
    return labels, vals.astype(np.int64)


def abundance(path1, path2, table: Table, k=15, mlen=2000, vs=400, ws=10):
    """count_kmer main() + countKmer (count_kmer.cpp:55-108,181-292) -> (labels, int64[G, vs])."""
    return _rows(lib().pgo_abundance(path1.encode(), (path2 or "").encode(), table.h, k, mlen, vs, ws))


def tnf(path1, path2=None, k=4, mlen=2000):
    """count_tnf main() + countKmer (count_tnf.cpp:78-113,166-302) -> (labels, int64[G, 136])."""
    return _rows(lib().pgo_tnf(path1.encode(), (path2 or "").encode(), k, mlen))


def featurize(path1, path2=None, k=15, tnf_k=4, mlen=2000, vs=400, ws=10, min_qual=0, table: Table | None = None):
    """Whole step 1 as Feature.extract_features does it (feature.py:28-39)."""
    if table is None:
        table = count_fastq([path1] + ([path2] if path2 else []), k, min_qual)
    n1, abd = abundance(path1, path2, table, k, mlen, vs, ws)
    n2, t = tnf(path1, path2, tnf_k, mlen)
    assert (n1 == n2).all()  # feature.py:35
    return n1, abd, t


# ----------------------------------------------------------------------------
# a16: Data.__init__ (src/data.py:9-22)
# ----------------------------------------------------------------------------
def normalize_l1(x: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize(x, "l1") for a dense array: float64 copy,
    norms = sum |x| per row, zero norms replaced by 1, divide (data.py:16,21)."""
    x = np.asarray(x).astype(np.float64)
    norms = np.abs(x).sum(axis=1)
    norms[norms == 0.0] = 1.0
    return x / norms[:, None]


def data_init(abd: np.ndarray, tnf_: np.ndarray):
    """-> (abd f32 [G,vs], tnf f32 [G,136], weights f64 [G]); data.py:16-21."""
    nabd = normalize_l1(abd)
    weights = (nabd.max(axis=1) ** 2).astype(np.float64) if nabd.shape[0] else np.zeros(0)
    return nabd.astype(np.float32), normalize_l1(tnf_).astype(np.float32), weights


def text_round(m: np.ndarray) -> np.ndarray:
    """What reaches pandas after the tools' CSV round trip: `ostream << double` prints 6 significant digits
    (count_kmer.cpp:211, count_tnf.cpp:204), so a tally >= 10^6 comes back as float("%.6g" % tally) - KAT-5.
    Python's % formatting and glibc's printf round the exact decimal value the same way (ties to even)."""
    m = np.asarray(m)
    big = m >= 1_000_000
    if not big.any():
        return m
    out = m.astype(np.float64)
    out[big] = [float("%.6g" % int(v)) for v in m[big]]
    return out


# ----------------------------------------------------------------------------
# the compiled reference tools
# ----------------------------------------------------------------------------
def _read_csv(path):
    """What feature.py:115,139 does: pd.read_csv(header=None); col 0 = labels."""
    import pandas as pd

    if os.path.getsize(path) == 0 or not gzip.open(path, "rb").read(1):
        return np.array([], dtype=object), None
    df = pd.read_csv(path, header=None, dtype={0: str})
    return df[0].to_numpy(), df.drop(columns=0).to_numpy()


def _read_csv_raw(path):
    """The tool's output as it wrote it: rows end with "\\n" only, the label is everything before the first comma.  (pandas
    also breaks rows at a bare "\\r": a label that carries one - a CRLF header whose label runs to the end of the line -
    garbles the reference's own read-back, feature.py:115; the differential tests on hostile text compare what the binary
    wrote.)"""
    text = gzip.open(path, "rb").read()
    if not text:
        return np.array([], dtype=object), None
    rows = [r for r in text.split(b"\n") if r]
    labels = np.array([r.split(b",", 1)[0].decode("utf-8", "surrogateescape") for r in rows], dtype=object)
    vals = np.array([[float(x) for x in r.split(b",")[1:]] for r in rows])
    return labels, vals


def ref_count_tnf(out_gz, interleaved=None, reads1=None, reads2=None, k=4, mlen=2000, threads=4, raw=False):
    cmd = [REF_COUNT_TNF, "-k", str(k), "-t", str(threads), "-l", str(mlen), "-o", out_gz]
    cmd += ["-i", interleaved] if interleaved else ["-1", reads1, "-2", reads2]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return _read_csv_raw(out_gz) if raw else _read_csv(out_gz)


def ref_count_kmer(out_gz, dump, interleaved=None, reads1=None, reads2=None, k=15, mlen=2000, vs=400, ws=10, threads=4, raw=False):
    cmd = [REF_COUNT_KMER, "-t", str(threads), "-g", dump, "-k", str(k), "-l", str(mlen), "-w", str(ws), "-v", str(vs), "-o", out_gz]
    cmd += ["-i", interleaved] if interleaved else ["-1", reads1, "-2", reads2]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return _read_csv_raw(out_gz) if raw else _read_csv(out_gz)
