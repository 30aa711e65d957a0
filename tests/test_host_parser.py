"""CPU: the product's host FASTQ reader (csrc/fastq.cpp) + the batch contract of
include/pangaea_b200.h, checked against the golden outputs of the reference binaries and
against the oracle.  No GPU compute is called."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import contract_features
from pangaea_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """Every function declared in include/pangaea_b200.h is exported and bound."""
    hdr = open(os.path.join(ROOT, "include", "pangaea_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_device_fails_loudly():
    L = _lib.lib()
    if L.pg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.PgError, match="no CPU path"):
        _lib.Context()


def test_param_validation_without_device():
    L = _lib.lib()
    p = _lib.pg_params()
    L.pg_default_params(C.byref(p))
    assert (p.k, p.tnf_k, p.window_size, p.vector_size, p.min_length) == (15, 4, 10, 400, 2000)
    h = C.c_void_p()
    for field, bad in (("k", 0), ("k", 32), ("tnf_k", 7), ("window_size", 0), ("vector_size", 0), ("table_capacity", 3)):
        q = _lib.pg_params()
        L.pg_default_params(C.byref(q))
        setattr(q, field, bad)
        assert L.pg_create(C.byref(q), C.byref(h)) == -1, field
    assert [L.pg_tnf_dim(k) for k in (1, 2, 3, 4, 5, 6)] == [2, 10, 32, 136, 512, 2080]


def test_parser_contract_reproduces_reference_outputs(golden, oracle):
    p = golden.params
    fq = _lib.Fastq(golden.path1, golden.reads2)
    seq, off, flag, keep = fq.arrays()
    labels = [fq.label(g) for g in range(fq.n_groups)]
    assert off[0] == 0 and off[-1] == len(seq) and len(off) == len(flag) + 1
    table = oracle.Table()
    table.load_dump(golden.dump, p["k"])
    names, abd, tnf = contract_features(seq, off, flag, keep, labels, table, p["k"], p["tnf_k"], p["min_length"],
                                        p["vector_size"], p["window_size"])
    assert names == list(golden.abd_labels) == list(golden.tnf_labels)
    assert np.array_equal(abd, golden.abd)
    assert np.array_equal(tnf, golden.tnf)


def test_parser_reads_gzip_and_missing_file(tmp_path, oracle):
    import gzip
    import shutil

    src = os.path.join(ROOT, "tests", "golden", "kat1_interleaved_10x", "reads.fq")
    gz = tmp_path / "reads.fq.gz"
    with open(src, "rb") as a, gzip.open(gz, "wb") as b:
        shutil.copyfileobj(a, b)
    plain, zipped = _lib.Fastq(src), _lib.Fastq(str(gz))
    for x, y in zip(plain.arrays(), zipped.arrays()):
        assert np.array_equal(x, y)
    with pytest.raises(_lib.PgError):
        _lib.Fastq(str(tmp_path / "nope.fq"))


def test_parser_qualities_follow_the_sequence_layout(tmp_path):
    (tmp_path / "i.fq").write_bytes(b"@a BX:Z:AA-1\nACGT\n+\nIII?\n@a BX:Z:AA-1\nGG\n+\n>>\n")
    fq = _lib.Fastq(str(tmp_path / "i.fq"), want_qual=True)
    r = fq.reads
    q = np.ctypeslib.as_array(C.cast(r.qual, C.POINTER(C.c_uint8)), shape=(r.n_bytes,))
    assert bytes(q) == b"III?\xff>>\xff"
    assert bytes(fq.arrays()[0]) == b"ACGT\nGG\n"


def test_long_lines_and_chunk_boundaries(tmp_path, oracle):
    """Lines longer than / straddling the reader's 16 MiB chunk."""
    rng = np.random.default_rng(3)
    big = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(1 << 24) + 12345))
    recs = b"".join(b"@r%d BX:Z:%s-1\n%s\n+\n%s\n" % (i, bc, s, b"I" * len(s))
                    for i, (bc, s) in enumerate([(b"AA", big), (b"AA", b"ACGT"), (b"CC", big[:70000]), (b"CC", b"TTGA")]))
    (tmp_path / "i.fq").write_bytes(recs)
    fq = _lib.Fastq(str(tmp_path / "i.fq"))
    seq, off, flag, keep = fq.arrays()
    assert np.diff(off).tolist() == [len(big) + 1, 5, 70001, 5]
    assert flag.tolist() == [0, 1, 0, 1] and keep.tolist() == [0, 1, 1]
    assert bytes(seq[off[2]:off[3]]) == big[:70000] + b"\n"


def _parse_all(path, want_qual=False):
    fq = _lib.Fastq(path, want_qual=want_qual)
    seq, off, flag, keep = (a.copy() for a in fq.arrays())
    r = fq.reads
    qual = bytes(np.ctypeslib.as_array(C.cast(r.qual, C.POINTER(C.c_uint8)), shape=(r.n_bytes,))) if want_qual and r.n_bytes else b""
    labels = [fq.label(g) for g in range(fq.n_groups)]
    fq.close()
    return seq, off, flag, keep, labels, qual


@pytest.mark.parametrize("threads", ["2", "3", "7", "16"])
def test_parallel_reader_equals_sequential_reader(tmp_path, monkeypatch, threads):
    """The multi-threaded reader of plain-text interleaved files (fastq.cpp: parse_interleaved_parallel) takes the same
    decisions as the sequential loop: golden files of the reference tools (ragged / hostile text included), a stLFR file
    whose read_type latches late, a truncated last record, barcodes that change exactly at the thread cuts."""
    from pangaea_b200 import synth

    files = [os.path.join(ROOT, "tests", "golden", d, "reads.fq") for d in ("kat1_interleaved_10x", "edge_ragged", "synth_10x_l2000")]
    data = synth.generate(n_barcodes=40, mean_pairs=3, read_len=37, n_genomes=2, genome_len=5000, frag_len=1000, seed=5, unbarcoded_pairs=7)
    files.append(synth.write_interleaved(str(tmp_path / "stlfr.fq"), data, style="stlfr"))
    # headers without any barcode first, then stLFR ones: the type latches in the middle of the file
    late = b"".join(b"@p%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(10)) + open(files[-1], "rb").read()
    (tmp_path / "late.fq").write_bytes(late)
    files.append(str(tmp_path / "late.fq"))
    (tmp_path / "trunc.fq").write_bytes(open(files[0], "rb").read()[:-37])  # ends inside a record, no final newline
    files.append(str(tmp_path / "trunc.fq"))
    (tmp_path / "one.fq").write_bytes(b"@a BX:Z:AA-1\nACGT\n+\nIIII\n@a BX:Z:AA-1\nGG\n+\n>>")
    files.append(str(tmp_path / "one.fq"))
    for path in files:
        monkeypatch.setenv("PG_FASTQ_THREADS", "1")
        want = _parse_all(path, want_qual=True)
        monkeypatch.setenv("PG_FASTQ_THREADS", threads)
        monkeypatch.setenv("PG_FASTQ_PARALLEL_MIN", "0")
        got = _parse_all(path, want_qual=True)
        monkeypatch.delenv("PG_FASTQ_PARALLEL_MIN")
        for a, b in zip(want[:4], got[:4]):
            assert np.array_equal(a, b), path
        assert want[4] == got[4] and want[5] == got[5], path


def test_reference_csv_writer_matches_the_reference_tools(tmp_path, golden):
    """Feature(write_csv=True) leaves the same text the reference binaries wrote (golden abundance.csv / tnf.csv), including
    the precision-6 scientific notation of tallies >= 10^6 (KAT-5)."""
    import gzip

    from pangaea_b200.feature import write_reference_csv

    for labels, m, name in ((golden.abd_labels, golden.abd, "abundance.csv"), (golden.tnf_labels, golden.tnf, "tnf.csv")):
        out = tmp_path / (name + ".gz")
        write_reference_csv(str(out), labels, m)
        want = open(os.path.join(golden.dir, name)).read()
        assert gzip.open(out, "rt").read() == want
    big = np.array([[1114930, 3, 999999, 1000000, 12345678]])
    write_reference_csv(str(tmp_path / "big.gz"), ["AC"], big)
    assert gzip.open(tmp_path / "big.gz", "rt").read() == "AC,1.11493e+06,3,999999,1e+06,1.23457e+07\n"


def test_parallel_reader_property_random_text(tmp_path, monkeypatch):
    """Property test (hypothesis): on arbitrary hostile text - blank lines, '\\r', headers with or without BX / '#', records cut
    anywhere, no final newline - the parallel reader returns exactly what the sequential one returns."""
    from hypothesis import given, settings, strategies as st

    header = st.one_of(
        st.builds(lambda n, bc: b"@r%d BX:Z:%s-1" % (n, bc), st.integers(0, 99), st.sampled_from([b"AAAA", b"AAAC", b"CC", b"", b"GG-TT"])),
        st.builds(lambda n, bc, m: b"@r%d#%s/%d" % (n, bc, m), st.integers(0, 99), st.sampled_from([b"1_1_1", b"0_0_0", b"2_2_2", b""]), st.integers(1, 2)),
        st.sampled_from([b"@plain", b"", b"@x\tBX:Z:AAAA", b"@y BX:Z:", b"# /", b"@z\r"]))
    seq = st.one_of(st.text(alphabet="ACGTNacgt", min_size=0, max_size=40).map(str.encode), st.sampled_from([b"", b"ACGT\r", b"@ACGT", b"+"]))
    line = st.one_of(header, seq, st.sampled_from([b"+", b"IIII", b"", b"????"]))
    record = st.tuples(header, seq, st.just(b"+"), seq).map(lambda t: list(t))
    text = st.one_of(st.lists(record, max_size=24).map(lambda rs: [l for r in rs for l in r]), st.lists(line, max_size=60))

    counter = [0]

    @settings(max_examples=400, deadline=None)
    @given(lines=text, final_newline=st.booleans(), threads=st.sampled_from(["2", "3", "5"]))
    def check(lines, final_newline, threads):
        counter[0] += 1
        path = str(tmp_path / f"h{counter[0] % 4}.fq")
        data = b"\n".join(lines) + (b"\n" if final_newline and lines else b"")
        open(path, "wb").write(data)
        monkeypatch.setenv("PG_FASTQ_THREADS", "1")
        want = _parse_all(path, want_qual=True)
        monkeypatch.setenv("PG_FASTQ_THREADS", threads)
        monkeypatch.setenv("PG_FASTQ_PARALLEL_MIN", "0")
        got = _parse_all(path, want_qual=True)
        monkeypatch.delenv("PG_FASTQ_PARALLEL_MIN")
        for a, b in zip(want[:4], got[:4]):
            assert np.array_equal(a, b), data
        assert want[4] == got[4] and want[5] == got[5], data

    check()
