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
        monkeypatch.setenv("PG_FASTQ_SEQUENTIAL", "1")   # the getline-style reader (also used for gzip / paired input)
        want = _parse_all(path, want_qual=True)
        monkeypatch.delenv("PG_FASTQ_SEQUENTIAL")
        monkeypatch.setenv("PG_FASTQ_THREADS", threads)
        got = _parse_all(path, want_qual=True)
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
        monkeypatch.setenv("PG_FASTQ_SEQUENTIAL", "1")   # the getline-style reader (also used for gzip / paired input)
        want = _parse_all(path, want_qual=True)
        monkeypatch.delenv("PG_FASTQ_SEQUENTIAL")
        monkeypatch.setenv("PG_FASTQ_THREADS", threads)
        got = _parse_all(path, want_qual=True)
        for a, b in zip(want[:4], got[:4]):
            assert np.array_equal(a, b), data
        assert want[4] == got[4] and want[5] == got[5], data

    check()


# ------------------------------------------------------------------------------------
# the reader as a stream of batches, and byte ranges for ranks
# ------------------------------------------------------------------------------------
def _chunk_tuple(fq, want_qual):
    seq, off, flag, keep = (a.copy() for a in fq.arrays())
    r = fq.reads
    qual = bytes(np.ctypeslib.as_array(C.cast(r.qual, C.POINTER(C.c_uint8)), shape=(r.n_bytes,))) if want_qual and r.n_bytes else b""
    labels = [fq.label(g) for g in range(fq.n_groups)]
    return seq, off, flag, keep, labels, qual


def _stream_all(path, target, want_qual=True, path2=None, **kw):
    s = _lib.FastqStream(path, path2, want_qual=want_qual, target_seq_bytes=target, **kw)
    chunks = []
    for fq in s:
        chunks.append(_chunk_tuple(fq, want_qual))
        fq.close()
    s.close()
    return chunks


def _concat(chunks):
    """what the batches of a stream add up to, in the layout of one batch"""
    if not chunks:
        return np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.uint8), [""], b""
    seq = np.concatenate([c[0] for c in chunks])
    base, offs = 0, [np.zeros(1, np.int64)]
    for c in chunks:
        offs.append(c[1][1:] + base)
        base += int(c[1][-1])
    flag = np.concatenate([c[2] for c in chunks])
    labels = list(chunks[0][4])
    for prev, c in zip(chunks, chunks[1:]):
        assert c[4][0] == prev[4][-1], "label 0 of a batch = label of the cloud that was open when the previous one ended"
        labels += c[4][1:]
    return seq, np.concatenate(offs), flag, labels, b"".join(c[5] for c in chunks)


def _check_cut_points(chunks):
    for c in chunks[:-1]:
        seq, off, flag, keep, labels, _ = c
        assert len(flag), "no empty batch in the middle of a stream"
        assert (flag[-1] & 1) or labels[-1] == "", "a batch ends at a cloud flush or inside a cloud that is dropped whole"
        assert len(keep) == len(labels) == 1 + int((flag & 1).sum())
        assert [bool(k) for k in keep] == [l != "" for l in labels]


def _stream_files(tmp_path):
    from pangaea_b200 import synth

    files = [os.path.join(ROOT, "tests", "golden", d, "reads.fq") for d in ("kat1_interleaved_10x", "edge_ragged", "synth_10x_l2000")]
    data = synth.generate(n_barcodes=60, mean_pairs=3, read_len=37, n_genomes=2, genome_len=5000, frag_len=1000, seed=5, unbarcoded_pairs=25)
    files.append(synth.write_interleaved(str(tmp_path / "stlfr.fq"), data, style="stlfr"))
    files.append(synth.write_interleaved(str(tmp_path / "tenx.fq"), data))
    (tmp_path / "trunc.fq").write_bytes(open(files[-1], "rb").read()[:-37])
    files.append(str(tmp_path / "trunc.fq"))
    # single-pair barcodes back to back (every pair flushes), then a long one
    recs = b"".join(b"@r%d BX:Z:%s-1\nACGTACGTAC\n+\nIIIIIIIIII\n" % (i // 2, b"ACGT"[(i // 2) % 4:(i // 2) % 4 + 1] * 3) for i in range(40))
    recs += b"".join(b"@q%d BX:Z:TTT-1\nGGGTACGTAC\n+\nIIIIIIIIII\n" % (i // 2) for i in range(60))
    (tmp_path / "singles.fq").write_bytes(recs)
    files.append(str(tmp_path / "singles.fq"))
    return files


@pytest.mark.parametrize("sequential", [False, True])
def test_stream_batches_add_up_to_the_whole_file(tmp_path, monkeypatch, sequential):
    if sequential:
        monkeypatch.setenv("PG_FASTQ_SEQUENTIAL", "1")
    monkeypatch.setenv("PG_FASTQ_THREADS", "3")
    for path in _stream_files(tmp_path):
        whole = _parse_all(path, want_qual=True)
        for target in (1, 150, 1000, 20_000, 10 ** 9):
            chunks = _stream_all(path, target)
            _check_cut_points(chunks)
            seq, off, flag, labels, qual = _concat(chunks)
            assert np.array_equal(seq, whole[0]) and np.array_equal(off, whole[1]) and np.array_equal(flag, whole[2]), (path, target)
            assert labels == whole[4] and qual == whole[5], (path, target)
            if target <= 150 and len(whole[2]) > 20:
                assert len(chunks) > 2


def test_stream_of_paired_and_gzip_input(tmp_path):
    import gzip
    import shutil

    g = os.path.join(ROOT, "tests", "golden", "synth_paired_minqual")
    whole = _lib.Fastq(os.path.join(g, "r1.fq"), os.path.join(g, "r2.fq"), want_qual=True)
    wt = _chunk_tuple(whole, True)
    for target in (300, 5000):
        chunks = _stream_all(os.path.join(g, "r1.fq"), target, path2=os.path.join(g, "r2.fq"))
        assert len(chunks) > 1
        _check_cut_points(chunks)
        seq, off, flag, labels, qual = _concat(chunks)
        assert np.array_equal(seq, wt[0]) and np.array_equal(off, wt[1]) and np.array_equal(flag, wt[2]) and labels == wt[4] and qual == wt[5]
    src = os.path.join(ROOT, "tests", "golden", "synth_10x_l2000", "reads.fq")
    gz = tmp_path / "reads.fq.gz"
    with open(src, "rb") as a, gzip.open(gz, "wb") as b:
        shutil.copyfileobj(a, b)
    wt = _parse_all(src, want_qual=True)
    chunks = _stream_all(str(gz), 4000)
    assert len(chunks) > 2
    seq, off, flag, labels, qual = _concat(chunks)
    assert np.array_equal(seq, wt[0]) and np.array_equal(flag, wt[2]) and labels == wt[4] and qual == wt[5]
    # a truncated gzip stream is an error, not a shorter file
    raw = gz.read_bytes()
    (tmp_path / "cut.fq.gz").write_bytes(raw[: len(raw) * 2 // 3])
    with pytest.raises(_lib.PgError):
        _lib.Fastq(str(tmp_path / "cut.fq.gz"))
    with pytest.raises(_lib.PgError):  # byte ranges need plain text
        _lib.FastqStream(str(gz), byte_lo=10, byte_hi=100, lines_before_lo=0)


def _ranges(path, world):
    size = os.path.getsize(path)
    cuts = [size * r // world for r in range(world)] + [size]
    lines = [_lib.count_lines(path, cuts[r], cuts[r + 1]) for r in range(world)]
    before = [sum(lines[:r]) for r in range(world)]
    return [(cuts[r], cuts[r + 1] if r + 1 < world else -1, before[r]) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_rank_byte_ranges_partition_the_file(tmp_path, monkeypatch, world):
    """Every rank opens its own byte range (pg_fastq_stream_open byte_lo / byte_hi); exchanged: newline counts only.  The
    ranks' batches, in rank order, add up to the whole file - and each rank starts at a cloud flush, so its clouds are whole."""
    monkeypatch.setenv("PG_FASTQ_THREADS", "2")
    for path in _stream_files(tmp_path):
        whole = _parse_all(path, want_qual=True)
        for target in (0, 700):
            chunks = []
            for lo, hi, before in _ranges(path, world):
                mine = _stream_all(path, target, byte_lo=lo, byte_hi=hi, lines_before_lo=before)
                chunks += mine
            _check_cut_points(chunks)
            seq, off, flag, labels, qual = _concat(chunks)
            assert np.array_equal(seq, whole[0]) and np.array_equal(off, whole[1]) and np.array_equal(flag, whole[2]), (path, world, target)
            assert labels == whole[4] and qual == whole[5], (path, world, target)


def test_stream_property_random_text(tmp_path, monkeypatch):
    """hypothesis: hostile text, random batch sizes and rank counts - batches and ranges always add up to the one-batch parse."""
    from hypothesis import given, settings, strategies as st

    header = st.one_of(
        st.builds(lambda n, bc: b"@r%d BX:Z:%s-1" % (n, bc), st.integers(0, 99), st.sampled_from([b"AAAA", b"AAAC", b"CC", b"", b"GG-TT"])),
        st.builds(lambda n, bc, m: b"@r%d#%s/%d" % (n, bc, m), st.integers(0, 99), st.sampled_from([b"1_1_1", b"0_0_0", b"2_2_2", b""]), st.integers(1, 2)),
        st.sampled_from([b"@plain", b"", b"@x\tBX:Z:AAAA", b"@y BX:Z:", b"# /", b"@z\r"]))
    seq = st.one_of(st.text(alphabet="ACGTNacgt", min_size=0, max_size=40).map(str.encode), st.sampled_from([b"", b"ACGT\r", b"@ACGT", b"+"]))
    line = st.one_of(header, seq, st.sampled_from([b"+", b"IIII", b"", b"????"]))
    record = st.tuples(header, seq, st.just(b"+"), seq).map(lambda t: list(t))
    text = st.one_of(st.lists(record, max_size=40).map(lambda rs: [l for r in rs for l in r]), st.lists(line, max_size=80))
    counter = [0]

    @settings(max_examples=300, deadline=None)
    @given(lines=text, final_newline=st.booleans(), threads=st.sampled_from(["1", "2", "3"]), target=st.sampled_from([1, 30, 200, 0]),
           world=st.integers(1, 4), sequential=st.booleans())
    def check(lines, final_newline, threads, target, world, sequential):
        counter[0] += 1
        path = str(tmp_path / f"s{counter[0] % 4}.fq")
        data = b"\n".join(lines) + (b"\n" if final_newline and lines else b"")
        open(path, "wb").write(data)
        monkeypatch.setenv("PG_FASTQ_THREADS", threads)
        whole = _parse_all(path, want_qual=True)
        if sequential:
            monkeypatch.setenv("PG_FASTQ_SEQUENTIAL", "1")
            chunks = _stream_all(path, target)
            monkeypatch.delenv("PG_FASTQ_SEQUENTIAL")
        else:
            chunks = []
            for lo, hi, before in _ranges(path, world):
                chunks += _stream_all(path, target, byte_lo=lo, byte_hi=hi, lines_before_lo=before)
        _check_cut_points(chunks)
        seq_, off, flag, labels, qual = _concat(chunks)
        assert np.array_equal(seq_, whole[0]) and np.array_equal(off, whole[1]) and np.array_equal(flag, whole[2]), data
        assert labels == whole[4] and qual == whole[5], data

    check()


@pytest.mark.parametrize("target", [400, 3000])
def test_streamed_batches_give_the_reference_rows(tmp_path, oracle, golden, target):
    """Contract-level check of the streaming flow: count over all batches, featurize batch by batch, concatenate the rows
    == the reference tools' outputs for the whole file (golden vectors)."""
    if golden.reads2 and not golden.interleaved:
        path1, path2 = golden.reads1, golden.reads2
    else:
        path1, path2 = golden.path1, None
    p = golden.params
    table = oracle.Table()
    table.load_dump(golden.dump, p["k"])
    names, abd, tnf = [], [], []
    n_chunks = 0
    s = _lib.FastqStream(path1, path2, target_seq_bytes=target)
    for fq in s:
        n_chunks += 1
        seq, off, flag, keep = fq.arrays()
        labels = [fq.label(g) for g in range(fq.n_groups)]
        n, a, t = contract_features(seq, off, flag, keep, labels, table, p["k"], p["tnf_k"], p["min_length"], p["vector_size"], p["window_size"])
        names += n; abd.append(a); tnf.append(t)
        fq.close()
    assert names == list(golden.abd_labels) == list(golden.tnf_labels)
    if names:
        assert np.array_equal(np.concatenate(abd), golden.abd) and np.array_equal(np.concatenate(tnf), golden.tnf)


@pytest.mark.parametrize("seed", range(100, 108))
def test_host_reader_on_hostile_headers_matches_the_oracle(tmp_path, oracle, seed):
    """The hostile interleaved texts of tests/test_oracle_vs_ref.py (where the oracle is checked against the reference binaries):
    host reader -> batch contract evaluated in Python == oracle.  Labels with a trailing '\\r' included."""
    from test_oracle_vs_ref import _hostile_text

    path = str(tmp_path / "h.fq")
    open(path, "wb").write(_hostile_text(seed))
    k, tnf_k, mlen, vs, ws = 9, 3, 30, 11, 2
    table = oracle.count_fastq([path], k)
    want_names, want_abd, want_tnf = oracle.featurize(path, None, k=k, tnf_k=tnf_k, mlen=mlen, vs=vs, ws=ws, table=table)
    fq = _lib.Fastq(path)
    seq, off, flag, keep = fq.arrays()
    labels = [fq.label(g) for g in range(fq.n_groups)]
    names, abd, tnf = contract_features(seq, off, flag, keep, labels, table, k, tnf_k, mlen, vs, ws)
    assert names == list(want_names)
    assert len(names) == 0 or (np.array_equal(abd, want_abd) and np.array_equal(tnf, want_tnf))


@pytest.mark.parametrize("seed", range(200, 206))
def test_host_reader_on_hostile_paired_files_matches_the_oracle(tmp_path, oracle, seed):
    """Paired files whose mates sometimes disagree in name or barcode (PG_READ_NOFEAT: counted, in no cloud)."""
    from test_oracle_vs_ref import _hostile_pair_files

    p1, p2 = _hostile_pair_files(seed, tmp_path)
    k, tnf_k, mlen, vs, ws = 9, 3, 30, 11, 2
    table = oracle.count_fastq([p1, p2], k)
    want_names, want_abd, want_tnf = oracle.featurize(p1, p2, k=k, tnf_k=tnf_k, mlen=mlen, vs=vs, ws=ws, table=table)
    fq = _lib.Fastq(p1, p2)
    seq, off, flag, keep = fq.arrays()
    labels = [fq.label(g) for g in range(fq.n_groups)]
    names, abd, tnf = contract_features(seq, off, flag, keep, labels, table, k, tnf_k, mlen, vs, ws)
    assert names == list(want_names)
    assert len(names) == 0 or (np.array_equal(abd, want_abd) and np.array_equal(tnf, want_tnf))
