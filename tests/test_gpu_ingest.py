"""GPU: FASTQ text on the device (csrc/ingest.cuh) - SURVEY.md §8f.1.
* pg_ingest_text (device line index + getBarcode + cloud flags) against the host reader, which the CPU suite pins to the
  reference tools, and against the oracle's features;
* pg_fastq_sort_by_barcode against the reference's own `LANG=C sort -k1,1 | cut | tr` pipeline (oracle/ingest_oracle.py)."""
import os

import numpy as np
import pytest

from pangaea_b200 import _lib, stream, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files(tmp_path):
    files = [os.path.join(ROOT, "tests", "golden", d, "reads.fq") for d in ("kat1_interleaved_10x", "edge_ragged", "synth_10x_l2000")]
    data = synth.generate(n_barcodes=60, mean_pairs=3, read_len=37, n_genomes=2, genome_len=5000, frag_len=1000, seed=5, unbarcoded_pairs=25)
    files.append(synth.write_interleaved(str(tmp_path / "stlfr.fq"), data, style="stlfr"))
    files.append(synth.write_interleaved(str(tmp_path / "tenx.fq"), data))
    late = b"".join(b"@p%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(10)) + open(files[-2], "rb").read()
    (tmp_path / "late.fq").write_bytes(late)   # the read type latches in the middle of the file
    files.append(str(tmp_path / "late.fq"))
    (tmp_path / "trunc.fq").write_bytes(open(files[-2], "rb").read()[:-37])  # ends inside a record, no final newline
    files.append(str(tmp_path / "trunc.fq"))
    (tmp_path / "odd.fq").write_bytes(b"@a BX:Z:\nAC\r\n+\n\n@a#1_2_3\n\n+\nII\n@b BX:Z:GG-TT-1\nACGT\n+\nIIII\n@b\tBX:Z:GG\nTT\n+\nII")
    files.append(str(tmp_path / "odd.fq"))
    (tmp_path / "empty.fq").write_bytes(b"")
    files.append(str(tmp_path / "empty.fq"))
    return files


def _host(path):
    fq = _lib.Fastq(path)
    seq, off, flag, keep = (a.copy() for a in fq.arrays())
    return seq, off, flag, keep, fq.labels()


def _device_chunks(ctx, text, window):
    """drive pg_ingest_text the way DeviceIngest does, with tiny windows"""
    out, pos, last, rt = [], 0, b"", 0
    n = len(text)
    w = window
    while True:
        final = pos + w >= n
        batch, labels, keep, consumed, rt = ctx.ingest_text(text[pos:pos + w], last, rt, final=final)
        if batch is None:
            assert not final
            w *= 2
            continue
        seq, off, flag = batch.download()
        batch.free()
        out.append((seq, off, flag, keep, labels))
        last = labels[-1].encode("utf-8", "surrogateescape")
        if final:
            return out
        assert consumed > 0
        pos += consumed
        w = window


def _concat(chunks):
    seq = np.concatenate([c[0] for c in chunks])
    base, offs = 0, [np.zeros(1, np.int64)]
    for c in chunks:
        offs.append(c[1][1:] + base)
        base += int(c[1][-1])
    labels = list(chunks[0][4])
    for prev, c in zip(chunks, chunks[1:]):
        assert c[4][0] == prev[4][-1]
        labels += c[4][1:]
    return seq, np.concatenate(offs), np.concatenate([c[2] for c in chunks]), labels


@pytest.mark.parametrize("window", [10 ** 9, 3000, 400])
def test_device_parse_equals_host_parse(tmp_path, window):
    ctx = _lib.Context(k=11)
    for path in _files(tmp_path):
        want = _host(path)
        text = open(path, "rb").read()
        chunks = _device_chunks(ctx, text, window)
        seq, off, flag, labels = _concat(chunks)
        # the device batch keeps two read slots per record; a record cut off by the end of the file leaves empty slots
        live = np.diff(off) > 0
        assert np.array_equal(seq, want[0]), path
        assert np.array_equal(np.diff(off)[live], np.diff(want[1])), path
        assert np.array_equal(flag[live], want[2]) and not flag[~live].any(), path
        assert labels == want[4], path
        for c in chunks:
            assert [bool(k) for k in c[3]] == [l != "" for l in c[4]]
        if window < 1000 and len(want[2]) > 40:
            assert len(chunks) > 2


def test_device_ingest_features_equal_oracle(tmp_path, oracle):
    data = synth.generate(n_barcodes=150, mean_pairs=16, read_len=100, n_genomes=3, genome_len=60_000, frag_len=9_000, seed=78,
                          unbarcoded_pairs=40, n_rate=0.002)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    ctx = _lib.Context()
    for window, resident in ((1 << 30, 0.45), (300_000, 0.45), (300_000, 0.0)):
        g_names, feats = stream.extract_features_device_ingest(ctx, path, window_bytes=window, resident_fraction=resident)
        g_abd, g_tnf = feats.raw()
        assert g_names == list(names) and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf), (window, resident)
        feats.free()
    # the slack in front of a window is too small for the tail behind the last flush: the driver starts over with a larger one
    g_names, feats = stream.extract_features_device_ingest(ctx, path, window_bytes=200_000, slack_bytes=64)
    g_abd, g_tnf = feats.raw()
    assert g_names == list(names) and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    assert len(list(stream.DeviceIngest(ctx, path, window_bytes=200_000))) > 3


def _unsorted_fastq(rng, n_pairs=3000, long_names=False):
    """pairs in random order: space-separated BX tags of different lengths (one a prefix of another), untagged pairs,
    duplicated read names (ties beyond the radix key), a tab inside a header"""
    tags = [b"ACGTACGTACGTACGT-1", b"ACGTACGTACGTACGT", b"ACGTACGTACGTACGA-1", b"TTTT-1", b"1_2_3", b"~zz", b"A"]
    recs = []
    for i in range(n_pairs):
        t = tags[int(rng.integers(0, len(tags)))] if rng.random() < 0.85 else None
        name = (b"@" + b"N" * 120 + b":%d" % (i % 50)) if long_names else b"@r%d" % int(rng.integers(0, 400))
        hdr = name + (b" BX:Z:" + t if t else b"")
        if i % 97 == 0:
            hdr = name + b"\tBX:Z:" + (t or b"Q")  # the script's tr turns this tab into a newline
        s1 = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(5, 40))))
        s2 = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(5, 40))))
        recs.append(hdr + b"\n" + s1 + b"\n+\n" + b"I" * len(s1) + b"\n" + hdr + b"\n" + s2 + b"\n+\n" + b"I" * len(s2) + b"\n")
    return b"".join(recs)


@pytest.mark.parametrize("long_names", [False, True])
def test_device_sort_equals_gnu_sort(long_names):
    from oracle import ingest_oracle as I

    ctx = _lib.Context(table_mode=_lib.PG_TABLE_NONE)
    rng = np.random.default_rng(12)
    text = _unsorted_fastq(rng, long_names=long_names)
    want = I.barcode_sort(text)
    got = ctx.sort_fastq_by_barcode(text)
    assert got == want
    assert ctx.sort_fastq_by_barcode(text[:-1]) == want  # a missing final newline is added (awk prints one)
    assert ctx.sort_fastq_by_barcode(b"") == b""
    with pytest.raises(_lib.PgError):
        ctx.sort_fastq_by_barcode(b"x\n" + text)         # not whole '@' records: refused, never silently mis-sorted


def test_sorted_file_then_features_equal_oracle(tmp_path, oracle):
    """the two device steps chained as run_pangaea chains them: sort by barcode, then featurize the sorted file"""
    from oracle import ingest_oracle as I

    data = synth.generate(n_barcodes=80, mean_pairs=14, read_len=100, n_genomes=3, genome_len=50_000, frag_len=8_000, seed=5, unbarcoded_pairs=30)
    path = synth.write_interleaved(str(tmp_path / "sorted_in.fq"), data)
    text = open(path, "rb").read().replace(b"\tBX:Z:", b" BX:Z:")  # seqtk mergepe leaves a space (run_pangaea:221)
    recs = [b"".join(r) for r in zip(*[iter(text.splitlines(keepends=True))] * 8)]
    rng = np.random.default_rng(3)
    shuffled = b"".join(recs[i] for i in rng.permutation(len(recs)))
    ctx = _lib.Context()
    got = ctx.sort_fastq_by_barcode(shuffled)
    assert got == I.barcode_sort(shuffled)
    (tmp_path / "sorted.fq").write_bytes(got)
    names, abd, tnf = oracle.featurize(str(tmp_path / "sorted.fq"), None)
    g_names, feats = stream.extract_features_device_ingest(ctx, str(tmp_path / "sorted.fq"), window_bytes=500_000)
    g_abd, g_tnf = feats.raw()
    assert g_names == list(names) and len(names) > 50 and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)


def test_device_parse_property_random_text(tmp_path):
    """hypothesis: on arbitrary hostile text - blank lines, '\\r', headers with or without BX / '#', records cut anywhere, no final
    newline - the device parser (windows of random size) returns exactly what the host reader returns."""
    from hypothesis import given, settings, strategies as st

    header = st.one_of(
        st.builds(lambda n, bc: b"@r%d BX:Z:%s-1" % (n, bc), st.integers(0, 99), st.sampled_from([b"AAAA", b"AAAC", b"CC", b"", b"GG-TT"])),
        st.builds(lambda n, bc, m: b"@r%d#%s/%d" % (n, bc, m), st.integers(0, 99), st.sampled_from([b"1_1_1", b"0_0_0", b"2_2_2", b""]), st.integers(1, 2)),
        st.sampled_from([b"@plain", b"", b"@x\tBX:Z:AAAA", b"@y BX:Z:", b"# /", b"@z\r", b"BX:Z", b"@q BX:Z:A#B/1"]))
    seq = st.one_of(st.text(alphabet="ACGTNacgt", min_size=0, max_size=40).map(str.encode), st.sampled_from([b"", b"ACGT\r", b"@ACGT", b"+"]))
    line = st.one_of(header, seq, st.sampled_from([b"+", b"IIII", b"", b"????"]))
    record = st.tuples(header, seq, st.just(b"+"), seq).map(lambda t: list(t))
    text = st.one_of(st.lists(record, max_size=40).map(lambda rs: [l for r in rs for l in r]), st.lists(line, max_size=80))
    ctx = _lib.Context(table_mode=_lib.PG_TABLE_NONE)
    counter = [0]

    @settings(max_examples=150, deadline=None)
    @given(lines=text, final_newline=st.booleans(), window=st.sampled_from([40, 150, 600, 10 ** 6]))
    def check(lines, final_newline, window):
        counter[0] += 1
        data = b"\n".join(lines) + (b"\n" if final_newline and lines else b"")
        path = str(tmp_path / f"p{counter[0] % 4}.fq")
        open(path, "wb").write(data)
        want = _host(path)
        chunks = _device_chunks(ctx, data, window)
        seq_, off, flag, labels = _concat(chunks)
        live = np.diff(off) > 0
        assert np.array_equal(seq_, want[0]), data
        assert np.array_equal(np.diff(off)[live], np.diff(want[1])) and np.array_equal(flag[live], want[2]) and not flag[~live].any(), data
        assert labels == want[4], data

    check()
