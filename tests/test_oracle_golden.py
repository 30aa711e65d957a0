"""CPU: the C restatement (oracle/pg_oracle.c) against the golden vectors that the
unmodified reference binaries produced (tests/golden/make_golden.py)."""
import numpy as np


def test_oracle_matches_reference_outputs(golden, oracle):
    p = golden.params
    table = oracle.Table()
    table.load_dump(golden.dump, p["k"])  # count_kmer.cpp:139-170
    labels, abd = oracle.abundance(golden.path1, golden.reads2, table, p["k"], p["min_length"], p["vector_size"], p["window_size"])
    assert list(labels) == list(golden.abd_labels)
    assert np.array_equal(abd, golden.abd)
    labels, tnf = oracle.tnf(golden.path1, golden.reads2, p["tnf_k"], p["min_length"])
    assert list(labels) == list(golden.tnf_labels)
    assert np.array_equal(tnf, golden.tnf)


def test_oracle_counter_reproduces_golden_dump(golden, oracle):
    """The dumps of the synthetic cases were written by the jellyfish stand-in; the
    hand-made KAT-4 dump is skipped (it is not a count of its reads)."""
    if golden.name == "kat4_bins":
        return
    p = golden.params
    files = [f for f in (golden.interleaved, golden.reads1, golden.reads2) if f]
    t = oracle.count_fastq(files, p["k"], p.get("min_qual", 0))
    want = oracle.Table()
    want.load_dump(golden.dump, p["k"])
    k1, v1 = t.items()
    k2, v2 = want.items()
    assert np.array_equal(k1, k2) and np.array_equal(v1, v2)


def test_kat1_boundary_quirk(oracle):
    """each cloud loses its first pair to the previous label."""
    from conftest import GoldenCase

    g = GoldenCase("kat1_interleaved_10x")
    assert list(g.tnf_labels) == ["AAAA", "CCCC", "GGGG"]
    assert g.tnf.shape == (3, 136)
    assert g.tnf.sum(axis=1).tolist() == [162, 108, 108]


def test_kat2_paired_rules(oracle):
    from conftest import GoldenCase

    g = GoldenCase("kat2_paired_stlfr")
    assert list(g.tnf_labels) == ["1_1_1", "2_2_2", "3_3_3"]
    assert g.tnf.sum(axis=1).tolist() == [222, 148, 74]


def test_kat3_tnf_column_order(oracle):
    lut, n = oracle.tnf_lut(4)
    assert n == 136
    cols = {}
    for code in range(256):
        if oracle.canonical(code, 4) == code:
            cols[int(lut[code])] = oracle.decode(code, 4)
    names = [cols[i] for i in range(136)]
    assert names[:10] == ["AAAA", "AAAC", "AAAT", "AAAG", "AACA", "AACC", "AACT", "AACG", "AATA", "AATC"]
    assert names[-3:] == ["GTAC", "GTCC", "GGCC"]
    assert [oracle.tnf_lut(k)[1] for k in (1, 2, 3, 5)] == [2, 10, 32, 512]


def test_revcomp_and_encoding(oracle):
    assert oracle.encode("ACTG") == 0b00011011
    assert oracle.revcomp(oracle.encode("AACG"), 4) == oracle.encode("CGTT")
    rng = np.random.default_rng(0)
    for k in (1, 4, 15, 16, 21, 31):
        for _ in range(50):
            v = int(rng.integers(0, 4 ** k, dtype=np.uint64)) if k < 32 else 0
            s = oracle.decode(v, k)
            rc = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
            assert oracle.revcomp(v, k) == oracle.encode(rc)
            assert oracle.revcomp(oracle.revcomp(v, k), k) == v


def test_header_rules(oracle):
    """getBarcode, count_kmer.cpp:25-53."""
    ph = oracle.parse_header
    assert ph("@r1 BX:Z:ACGT-1") == ("@r1", "ACGT", 1)
    assert ph("@r1\tBX:Z:ACGT-1\tXX") == ("@r1", "ACGT", 1)
    assert ph("@r1\tBX:Z:ACGT") == ("@r1", "ACGT", 1)
    assert ph("@r1#12_3_4/1") == ("@r1", "12_3_4", 2)
    assert ph("@r1#0_0_0/1") == ("@r1", "", 2)
    assert ph("@plain") == ("@plain", "", 0)
    # the type latches on the first decisive header and is not re-inferred
    assert ph("@r1#1_2_3/1 BX:Z:AAA-1", 0)[2] == 1
    assert ph("@r1#1_2_3/1", 1) == ("@r1#1_2_3/1", "", 1)
    # stLFR latch, header without '#': size_t wrap-around makes the whole line the barcode
    assert ph("@r1 BX:Z:AAA-1", 2) == ("@r1 BX:Z:AAA-1", "@r1 BX:Z:AAA-1", 2)


def test_data_init_matches_sklearn(oracle):
    """Data.__init__, src/data.py:16-21 (sklearn is installed here and on the box)."""
    from sklearn.preprocessing import normalize

    rng = np.random.default_rng(5)
    abd = rng.integers(0, 50, size=(64, 400)).astype(np.int64)
    abd[3] = 0
    tnf = rng.integers(0, 2000, size=(64, 136)).astype(np.int64)
    a, t, w = oracle.data_init(abd, tnf)
    na = normalize(abd, "l1")
    assert np.array_equal(a, na.astype(np.float32))
    assert np.array_equal(t, normalize(tnf, "l1").astype(np.float32))
    assert np.array_equal(w, np.array([na[i].max() ** 2 for i in range(64)], dtype=np.float64))
    assert a.dtype == np.float32 and w.dtype == np.float64


def test_text_round_matches_libc_printf_and_the_integer_rule(oracle):
    """oracle.text_round (python "%.6g") == glibc printf("%.6g") == the integer rule csrc/normalize.cuh:text_round6 uses
    (round to 6 significant digits, ties to even), on random tallies and on exact ties."""
    import ctypes

    libc = ctypes.CDLL("libc.so.6")
    libc.snprintf.restype = ctypes.c_int
    buf = ctypes.create_string_buffer(64)

    def c_round(v):
        libc.snprintf(buf, ctypes.c_size_t(64), b"%.6g", ctypes.c_double(float(v)))
        return float(buf.value)

    def int_rule(v):  # mirrors text_round6
        if v < 1_000_000:
            return v
        q = 10 if v < 10 ** 7 else 100 if v < 10 ** 8 else 1000 if v < 10 ** 9 else 10000
        r, base = v % q, v - v % q
        up = 2 * r > q or (2 * r == q and (base // q) & 1)
        return base + q if up else base

    rng = np.random.default_rng(0)
    vals = list(rng.integers(0, 2 ** 32, size=4000)) + list(rng.integers(999_990, 1_000_100, size=200))
    vals += [1_000_005, 1_000_015, 1_000_025, 12_345_650, 12_345_750, 999_999, 1_000_000, 9_999_995, 9_999_985, 99_999_950, 99_999_850,
             4_294_967_295, 4_294_965_000, 2_500_005_000, 1_114_930, 1_114_934, 1_114_935, 1_114_936]
    for v in map(int, vals):
        want = c_round(v)
        assert float(int_rule(v)) == want, v
        assert float(oracle.text_round(np.array([v], dtype=np.int64))[0]) == want, v
