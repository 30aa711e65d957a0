"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the
C-ABI (pangaea_b200._lib -> libpangaea_b200.so); the oracle and the golden vectors of the
reference binaries are the checkers.  Integer results are compared bit-exactly."""
import argparse
import os

import numpy as np
import pytest

from helpers import contract_features, load_dump_arrays
from pangaea_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _ctx(**kw):
    return _lib.Context(**kw)


def _names(fq, feats):
    return [fq.label(int(g)) for g in feats.row_groups()]


def _oracle_table_arrays(t):
    k, v = t.items()
    return k.astype(np.uint64), v.astype(np.uint64)


# ------------------------------------------------------------------------------------
# golden vectors produced by the unmodified reference binaries
# ------------------------------------------------------------------------------------
def test_golden_reference_outputs(golden):
    """count_kmer / count_tnf outputs, table loaded from the same dump (`-g`)."""
    p = golden.params
    fq = _lib.Fastq(golden.path1, golden.reads2)
    ctx = _ctx(k=p["k"], tnf_k=p["tnf_k"], window_size=p["window_size"], vector_size=p["vector_size"], min_length=p["min_length"])
    keys, counts = load_dump_arrays(golden.dump, p["k"])
    ctx.table_set(keys, counts)
    batch = ctx.upload(fq.reads)
    feats = ctx.featurize(batch, fq.group_keep, fq.n_groups)
    abd, tnf = feats.raw()
    assert _names(fq, feats) == list(golden.abd_labels) == list(golden.tnf_labels)
    assert np.array_equal(abd, golden.abd)
    assert np.array_equal(tnf, golden.tnf)


def test_golden_counts_match_oracle_counter(golden, oracle):
    """k-mer table after pg_count == the oracle's jellyfish stand-in (bit-exact counts)."""
    if golden.name == "kat4_bins":
        pytest.skip("hand-made dump")
    p = golden.params
    mq = p.get("min_qual", 0)
    fq = _lib.Fastq(golden.path1, golden.reads2, want_qual=bool(mq))
    ctx = _ctx(k=p["k"], min_qual_char=mq)
    batch = ctx.upload(fq.reads)
    ctx.count(batch)
    keys, counts = ctx.table_export()
    want = oracle.Table()
    want.load_dump(golden.dump, p["k"])
    wk, wv = _oracle_table_arrays(want)
    assert np.array_equal(keys, wk)
    assert np.array_equal(counts.astype(np.uint64), wv)
    assert ctx.table_size() == len(wk)
    # spot look-ups, in either orientation
    probe = wk[:: max(1, len(wk) // 50)]
    assert np.array_equal(ctx.table_get(probe).astype(np.uint64), wv[:: max(1, len(wk) // 50)])
    rc = np.array([oracle.revcomp(int(x), p["k"]) for x in probe], dtype=np.uint64)
    assert np.array_equal(ctx.table_get(rc).astype(np.uint64), wv[:: max(1, len(wk) // 50)])


def test_golden_whole_path_from_host_buffers(golden, oracle):
    """pg_extract_features (count + featurize + normalize) vs oracle over the same files."""
    if golden.name == "kat4_bins":
        pytest.skip("hand-made dump")
    p = golden.params
    mq = p.get("min_qual", 0)
    fq = _lib.Fastq(golden.path1, golden.reads2, want_qual=bool(mq))
    ctx = _ctx(k=p["k"], tnf_k=p["tnf_k"], window_size=p["window_size"], vector_size=p["vector_size"],
               min_length=p["min_length"], min_qual_char=mq)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    abd, tnf = feats.raw()
    assert _names(fq, feats) == list(golden.abd_labels)
    assert np.array_equal(abd, golden.abd) and np.array_equal(tnf, golden.tnf)
    a, t, w = feats.normalized()
    oa, ot, ow = oracle.data_init(golden.abd, golden.tnf)
    assert np.array_equal(a, oa) and np.array_equal(t, ot) and np.array_equal(w, ow)


# ------------------------------------------------------------------------------------
# seeded random inputs vs the oracle
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,mode", [(15, _lib.PG_TABLE_AUTO), (11, _lib.PG_TABLE_AUTO), (12, _lib.PG_TABLE_AUTO), (14, _lib.PG_TABLE_AUTO),
                                    (16, _lib.PG_TABLE_AUTO), (15, _lib.PG_TABLE_HASH), (21, _lib.PG_TABLE_AUTO), (31, _lib.PG_TABLE_AUTO),
                                    (1, _lib.PG_TABLE_AUTO), (7, _lib.PG_TABLE_HASH)])
def test_random_vs_oracle(tmp_path, oracle, k, mode):
    data = synth.generate(n_barcodes=60, mean_pairs=12, read_len=100, n_genomes=3, genome_len=40_000, frag_len=6_000,
                          seed=100 + k, unbarcoded_pairs=9, n_rate=0.004, lower_rate=0.003)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    names, abd, tnf = oracle.featurize(path, None, k=k, tnf_k=4, mlen=1500, vs=50, ws=2)
    fq = _lib.Fastq(path)
    ctx = _ctx(k=k, window_size=2, vector_size=50, min_length=1500, table_mode=mode, table_capacity=1 << 18 if (mode == _lib.PG_TABLE_HASH or k > 16) else 0)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    g_abd, g_tnf = feats.raw()
    assert _names(fq, feats) == list(names)
    assert np.array_equal(g_tnf, tnf)
    assert np.array_equal(g_abd, abd)
    keys, counts = ctx.table_export()
    wk, wv = _oracle_table_arrays(oracle.count_fastq(path, k))
    assert np.array_equal(keys, wk) and np.array_equal(counts.astype(np.uint64), wv)


@pytest.mark.parametrize("tnf_k", [1, 2, 3, 5, 6])
def test_tnf_k_variants(tmp_path, oracle, tnf_k):
    data = synth.generate(n_barcodes=12, mean_pairs=10, read_len=80, seed=7 + tnf_k, n_rate=0.01)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    names, tnf = oracle.tnf(path, None, tnf_k, 0)
    fq = _lib.Fastq(path)
    ctx = _ctx(k=9, tnf_k=tnf_k, min_length=0)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    assert _names(fq, feats) == list(names)
    assert np.array_equal(feats.raw()[1], tnf)


def test_tiny_clouds_overflow_the_slots(tmp_path, oracle):
    """One pair per barcode (the hybrid 'per-read' shape): dozens of clouds per tile, so the
    block-private slots overflow into direct global reductions."""
    data = synth.generate(n_barcodes=700, mean_pairs=1, read_len=150, n_genomes=2, genome_len=30_000, frag_len=3_000, seed=5)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    names, abd, tnf = oracle.featurize(path, None, k=15, mlen=0)
    fq = _lib.Fastq(path)
    ctx = _ctx(min_length=0)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    assert _names(fq, feats) == list(names)
    g_abd, g_tnf = feats.raw()
    assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)


@pytest.mark.parametrize("env", [{}, {"PG_SEG_WORDS": "512"}, {"PG_FORCE_DIRECT": "1"}, {"PG_REGION_SLACK": "0.3"},
                                 {"PG_REGION_SLACK": "0.02", "PG_SEG_WORDS": "2048"}, {"PG_COUNT_L2": "1"},
                                 {"PG_COUNT_L2": "1", "PG_REGION_SLACK": "0.3"}, {"PG_COUNT_L2": "1", "PG_SEG_WORDS": "1024"},
                                 {"PG_NO_SHARED": "1"}, {"PG_NO_SHARED": "1", "PG_SEG_WORDS": "1024"}, {"PG_NO_SHARED": "1", "PG_REGION_SLACK": "0.3"},
                                 {"PG_SEG_WORDS": "1024", "PG_STASH_SEGMENTS": "3"}, {"PG_SEG_WORDS": "2048", "PG_STASH_SEGMENTS": "1", "PG_COUNT_L2": "1"}])
def test_sliced_and_direct_table_paths_agree(tmp_path, oracle, monkeypatch, env):
    """k = 15 uses the L2-sliced path (bucket.cuh) by default; PG_SEG_WORDS forces many
    segments on a small input, PG_FORCE_DIRECT the one-kernel path, PG_REGION_SLACK < 1 makes
    slice regions (and sub-slice regions of the two-level count) overflow so runs are applied by the
    scatter / split kernels, PG_COUNT_L2 applies the count entries with L2 atomics instead of the
    shared-memory sub-slices (count2.cuh), PG_NO_SHARED partitions separately for the count and the
    featurize pass instead of once for both (the default when clouds are >= 64 bytes; an overflow there
    makes pg_featurize fall back to its own partition), PG_STASH_SEGMENTS keeps only the first segments of the shared
    partition (what a batch too large for HBM gets).  All must equal the oracle."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    data = synth.generate(n_barcodes=150, mean_pairs=14, read_len=100, n_genomes=3, genome_len=60_000, frag_len=8_000,
                          seed=77, unbarcoded_pairs=40, n_rate=0.003, lower_rate=0.002)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    fq = _lib.Fastq(path)
    ctx = _ctx()
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    g_abd, g_tnf = feats.raw()
    assert _names(fq, feats) == list(names)
    assert np.array_equal(g_tnf, tnf) and np.array_equal(g_abd, abd)
    keys, counts = ctx.table_export()
    wk, wv = _oracle_table_arrays(oracle.count_fastq(path, 15))
    assert np.array_equal(keys, wk) and np.array_equal(counts.astype(np.uint64), wv)


def test_ragged_lengths_and_long_reads(tmp_path, oracle):
    rng = np.random.default_rng(8)
    recs = []
    for i in range(400):
        bc = b"ACGT"[i // 100:i // 100 + 1] * 8
        for _ in range(2):
            n = int(rng.choice([0, 1, 14, 15, 16, 17, 31, 32, 33, 47, 64, 65, 100, 1000, 5000]))
            s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n, p=[.245, .245, .245, .245, .02]))
            recs.append(b"@r%d BX:Z:%s-1\n%s\n+\n%s\n" % (i, bc, s, b"I" * n))
    path = str(tmp_path / "i.fq")
    open(path, "wb").write(b"".join(recs))
    names, abd, tnf = oracle.featurize(path, None, k=15, mlen=2000)
    fq = _lib.Fastq(path)
    ctx = _ctx()
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    assert _names(fq, feats) == list(names) and len(names) >= 3
    g_abd, g_tnf = feats.raw()
    assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)


@pytest.mark.parametrize("env", [{}, {"PG_COUNT_L2": "1"}, {"PG_NO_SHARED": "1"}])
def test_homopolymers_and_repeats(tmp_path, oracle, monkeypatch, env):
    """Poly-G tails / tandem repeats: many identical consecutive k-mers (they overflow the staging
    rows of the scatter and split kernels - the redo paths - and are folded inside the warp by the L2
    apply) and counts far above window*vector (ignored bins)."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    recs = []
    for i in range(60):
        bc = b"AAAA" if i < 30 else b"CCCC"
        for s in (b"G" * 120, b"ACGTACGTAC" * 12 + b"T" * 30):
            recs.append(b"@r%d BX:Z:%s-1\n%s\n+\n%s\n" % (i, bc, s, b"I" * len(s)))
    path = str(tmp_path / "i.fq")
    open(path, "wb").write(b"".join(recs))
    names, abd, tnf = oracle.featurize(path, None, k=15, mlen=100, vs=400, ws=10)
    fq = _lib.Fastq(path)
    ctx = _ctx(min_length=100)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    g_abd, g_tnf = feats.raw()
    assert _names(fq, feats) == list(names)
    assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    key = oracle.canonical(oracle.encode("G" * 15), 15)
    assert ctx.table_get(np.array([key], dtype=np.uint64))[0] == 60 * 106


def test_empty_and_degenerate_batches():
    ctx = _ctx(min_length=0)
    # nothing at all
    r = _lib.make_reads(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.uint8))
    feats = ctx.extract_features(r, np.zeros(1, np.uint8))
    assert feats.rows == 0 and feats.raw()[0].shape == (0, 400)
    assert feats.normalized()[2].shape == (0,)
    # a single all-N pair carrying a change flag: cloud 0 holds it, cloud 1 is empty
    seq = np.frombuffer(b"NNNNNNNNNNNNNNNNNNNN\nNNNNNNNNNNNNNNNNNNNN\n", dtype=np.uint8)
    r = _lib.make_reads(seq, np.array([0, 21, 42], np.int64), np.array([0, 1], np.uint8))
    feats = ctx.extract_features(r, np.array([1, 1], np.uint8))
    assert feats.rows == 1 and feats.row_groups().tolist() == [0]
    abd, tnf = feats.raw()
    assert abd.sum() == 0 and tnf.sum() == 0
    a, t, w = feats.normalized()
    assert not a.any() and not t.any() and w.tolist() == [0.0]  # zero rows stay zero (sklearn zero-norm rule)
    assert ctx.table_size() == 0
    # wrong n_groups is an error, not a crash
    with pytest.raises(_lib.PgError, match="n_groups"):
        ctx.extract_features(r, np.array([1, 1, 1], np.uint8))
    # featurize before any count
    ctx2 = _ctx()
    b = ctx2.upload(r)
    with pytest.raises(_lib.PgError, match="table is empty"):
        ctx2.featurize(b, np.array([1, 1], np.uint8))


def test_negative_min_length_drops_everything(tmp_path):
    """`reads_seq.size() <= mlen` is unsigned in the reference: -l -1 emits nothing."""
    data = synth.generate(n_barcodes=5, mean_pairs=5, read_len=50, seed=1)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    fq = _lib.Fastq(path)
    ctx = _ctx(min_length=-1)
    assert ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups).rows == 0


def test_hash_table_full_is_reported(tmp_path):
    data = synth.generate(n_barcodes=30, mean_pairs=20, read_len=100, seed=2)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    fq = _lib.Fastq(path)
    ctx = _ctx(k=21, table_capacity=1 << 10)
    with pytest.raises(_lib.PgError, match="full"):
        ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)


def test_table_set_semantics(oracle):
    """kmer2frequency[key] = freq (count_kmer.cpp:166): re-canonicalised, last one wins."""
    ctx = _ctx(k=15)
    a = oracle.encode("ACGTTGCAACGTACG")
    rc = oracle.revcomp(a, 15)
    ctx.table_set(np.array([a, rc, 5], np.uint64), np.array([7, 9, 3], np.uint32))
    assert ctx.table_get(np.array([a, rc, 5, 6], np.uint64)).tolist() == [9, 9, 3, 0]
    keys, counts = ctx.table_export()
    assert keys.tolist() == sorted([oracle.canonical(a, 15), oracle.canonical(5, 15)])
    ctx.table_clear()
    assert ctx.table_size() == 0


# ------------------------------------------------------------------------------------
# Data.__init__ and the zero-copy hand-off
# ------------------------------------------------------------------------------------
def test_normalize_bit_exact_vs_sklearn_semantics(oracle):
    rng = np.random.default_rng(4)
    abd = rng.integers(0, 3000, size=(1000, 400)).astype(np.int64) * (rng.random((1000, 400)) < 0.05)
    abd[17] = 0
    abd[18, 3] = 4_000_000_000  # near the u32 ceiling
    tnf = rng.integers(0, 5000, size=(1000, 136)).astype(np.int64)
    # tallies >= 10^6 reach the reference's Data.__init__ through 6-significant-digit text (KAT-5): exact ties included
    abd[19, :6] = [1_000_005, 1_000_015, 12_345_650, 99_999_950, 1_114_935, 999_999]
    tnf[20, :3] = [2_500_005_000, 1_114_934, 7]
    for ctx in (_ctx(), _ctx(table_mode=_lib.PG_TABLE_NONE)):
        f = ctx.features_from_raw(abd, tnf)
        a, t, w = f.normalized()
        oa, ot, ow = oracle.data_init(oracle.text_round(abd), oracle.text_round(tnf))
        assert a.dtype == np.float32 and w.dtype == np.float64
        assert np.array_equal(a, oa) and np.array_equal(t, ot) and np.array_equal(w, ow)
        ra, rt = f.raw()  # the raw tallies stay exact
        assert np.array_equal(ra.view(np.uint32), abd.astype(np.uint32)) and np.array_equal(rt.view(np.uint32), tnf.astype(np.uint32))


def test_tableless_ctx_only_normalises(tmp_path):
    """Data.__init__ without device features builds a PG_TABLE_NONE ctx: no 2 GiB table, any vector_size, and everything
    that needs the table fails loudly."""
    from pangaea_b200 import Data

    ctx = _ctx(table_mode=_lib.PG_TABLE_NONE, vector_size=20000)
    rng = np.random.default_rng(1)
    abd, tnf = rng.integers(0, 50, size=(7, 20000)), rng.integers(0, 50, size=(7, 136))
    ds = Data(np.arange(7), abd, tnf)
    assert ds.abd.shape == (7, 20000) and np.allclose(ds.abd.sum(axis=1), 1.0, atol=1e-4)
    data = synth.generate(n_barcodes=3, mean_pairs=5, read_len=80, seed=1)
    fq = _lib.Fastq(synth.write_interleaved(str(tmp_path / "i.fq"), data))
    b = ctx.upload(fq.reads)
    with pytest.raises(_lib.PgError, match="PG_TABLE_NONE"):
        ctx.count(b)
    with pytest.raises(_lib.PgError):
        _ctx(window_size=2 ** 20, vector_size=4096)  # w * v > 2^31 - 1: a saturated counter could land in a bin


@pytest.mark.parametrize("env", [{}, {"PG_COUNT_L2": "1"}, {"PG_FORCE_DIRECT": "1"}, {"PG_REGION_SLACK": "0.3"}])
def test_counters_saturate_instead_of_wrapping(tmp_path, oracle, monkeypatch, env):
    """A k-mer seen >= 2^31 times (poly-G at multi-billion-read scale): the dense counter stops at 2^31 - 1, whatever path
    adds to it (shared-memory merge, overflow REDs, direct kernel), and the k-mer is dropped from the histogram exactly as
    the reference drops the true count (count_kmer.cpp:90-92)."""
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    k = 15
    rng = np.random.default_rng(5)
    body = "".join(rng.choice(list("ACGT"), size=3000))
    reads = []
    for i in range(60):  # 60 pairs in one barcode; every read = 40 G (26 poly-G windows) + 60 genome bases
        for _ in range(2):
            o = int(rng.integers(0, len(body) - 60))
            reads.append("G" * 40 + "A" + body[o:o + 59])
    with open(tmp_path / "i.fq", "w") as f:
        for i in range(0, len(reads), 2):
            for r in reads[i:i + 2]:
                f.write(f"@r{i} BX:Z:ACGTACGTACGTACGT-1\n{r}\n+\n{'I' * len(r)}\n")
        for _ in range(2):  # a second barcode closes the first cloud
            f.write(f"@z BX:Z:TTTTACGTACGTACGT-1\n{body[:100]}\n+\n{'I' * 100}\n")
    fq = _lib.Fastq(str(tmp_path / "i.fq"))
    ctx = _ctx(k=k, min_length=100)
    polyg = oracle.encode("G" * k)
    other = oracle.encode(body[100:100 + k])
    near = 2 ** 31 - 1 - 1000                       # the batch adds 120 * 26 = 3120 > 1000 poly-G windows
    ctx.table_set(np.array([polyg, other], np.uint64), np.array([near, 2 ** 31 - 1], np.uint32))
    batch = ctx.upload(fq.reads)
    ctx.count(batch)
    want = oracle.Table()
    for r in reads + [body[:100]] * 2:
        want.count_read(r.encode() + b"\n", k)
    assert want.get(oracle.canonical(polyg, k)) == 120 * 26
    got = ctx.table_get(np.array([polyg, other], np.uint64))
    assert got.tolist() == [2 ** 31 - 1, 2 ** 31 - 1], got
    # every other counter is exact
    keys, counts = ctx.table_export()
    wk, wv = _oracle_table_arrays(want)
    sel = ~np.isin(wk, [oracle.canonical(polyg, k), oracle.canonical(other, k)])
    assert np.array_equal(keys[np.isin(keys, wk[sel])], wk[sel])
    assert np.array_equal(counts[np.isin(keys, wk[sel])].astype(np.uint64), wv[sel])
    # features: the saturated k-mers are beyond the histogram, like their true counts
    feats = ctx.featurize(batch, fq.group_keep, fq.n_groups)
    abd, _ = feats.raw()
    want.set(oracle.canonical(polyg, k), near + 120 * 26)
    want.set(oracle.canonical(other, k), 2 ** 31 - 1 + want.get(oracle.canonical(other, k)))
    names, oabd, _ = oracle.featurize(str(tmp_path / "i.fq"), None, k=k, mlen=100, table=want)
    assert _names(fq, feats) == list(names) and np.array_equal(abd, oabd)
    # a table that holds zero-count markers cannot be counted into
    ctx.table_clear()
    ctx.table_set(np.array([polyg], np.uint64), np.array([0], np.uint32))
    with pytest.raises(_lib.PgError, match="marker"):
        ctx.count(batch)


def test_table_clamp_before_a_sum_across_ranks():
    ctx = _ctx(k=15)
    keys = np.array([5, 6, 7], np.uint64)
    ctx.table_set(keys, np.array([10, 2 ** 31 - 1, 300_000_000], np.uint32))
    ctx.table_clamp((2 ** 31 - 1) // 8)
    assert ctx.table_get(keys).tolist() == [10, (2 ** 31 - 1) // 8, (2 ** 31 - 1) // 8]
    t = ctx.table_as_torch()
    assert int((t.to("cuda").long() * 8).max()) < 2 ** 31   # the int32 sum of 8 such tables cannot reach bit 31


def test_dlpack_zero_copy_to_torch(tmp_path):
    import torch

    data = synth.generate(n_barcodes=20, mean_pairs=15, read_len=100, seed=3)
    path = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    fq = _lib.Fastq(path)
    ctx = _ctx(min_length=0)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    a, t, w = feats.normalized()
    ta, tt, tw, ra = feats.torch(_lib.ABD), feats.torch(_lib.TNF), feats.torch(_lib.WEIGHTS), feats.torch(_lib.ABD_RAW)
    assert ta.is_cuda and ta.dtype == torch.float32 and tuple(ta.shape) == a.shape
    assert tw.dtype == torch.float64 and ra.dtype == torch.int32
    assert ta.data_ptr() == _lib.lib().pg_features_device_ptr(feats.h, _lib.ABD)  # same memory, no copy
    assert np.array_equal(ta.cpu().numpy(), a) and np.array_equal(tt.cpu().numpy(), t) and np.array_equal(tw.cpu().numpy(), w)
    feats.free()  # the tensors keep the buffers alive
    del ctx
    assert np.array_equal(ta.cpu().numpy(), a)
    x = torch.nn.functional.softmax(ta, dim=1)  # usable by an unchanged torch consumer
    assert torch.isfinite(x).all()


# ------------------------------------------------------------------------------------
# the drop-in classes
# ------------------------------------------------------------------------------------
def _args(out, **kw):
    d = dict(tnf_kmer=4, window_size=10, vector_size=400, kmer=15, min_length=2000, threads=4, output=str(out),
             reads1=None, reads2=None, interleaved_reads=None)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.parametrize("ingest", ["device", "host"])
def test_feature_and_data_drop_in(tmp_path, oracle, ingest):
    """The drop-in classes end to end, with the text parsed on the device (the default for a plain interleaved file) and by
    the host reader."""
    from pangaea_b200 import Data, Feature

    data = synth.generate(n_barcodes=40, mean_pairs=20, read_len=100, n_genomes=3, genome_len=50_000, frag_len=8_000, seed=21,
                          unbarcoded_pairs=4)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    ft = Feature(_args(tmp_path, interleaved_reads=path), script_path="unused", ingest=None if ingest == "device" else ingest)
    g_names, g_abd, g_tnf = ft.extract_features(write_csv=True)
    assert ft.ingest_used == ingest  # ("auto" picks the device parser for this input)
    assert g_abd.dtype == np.int64 and g_abd.shape == abd.shape
    assert list(g_names) == list(names) and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    fd = tmp_path / "1.features"
    assert (fd / "feature_finished").read_text() == "feature finished"
    assert (fd / "abundance.k15.v400.w10.m2000.pkl").exists() and (fd / "tnf.m2000.pkl").exists()
    import pandas as pd  # the optional text artefacts parse the way the reference parses its own (feature.py:115,139)
    for csv, want in ((fd / "abundance.k15.v400.w10.m2000.gz", abd), (fd / "tnf.m2000.gz", tnf)):
        df = pd.read_csv(csv, header=None)
        assert list(df[0]) == list(names) and np.array_equal(df.drop(columns=0).to_numpy(), want)
    l_names, l_abd, l_tnf = Feature(_args(tmp_path, interleaved_reads=path), "unused").load_features()
    assert list(l_names) == list(names) and np.array_equal(l_abd, abd) and np.array_equal(l_tnf, tnf)

    oa, ot, ow = oracle.data_init(abd, tnf)
    for ds in (Data(g_names, g_abd, g_tnf), Data(g_names, g_abd, g_tnf, features=ft.features)):
        assert len(ds) == len(names)
        assert np.array_equal(ds.abd, oa) and np.array_equal(ds.tnf, ot) and np.array_equal(ds.weights, ow)
        item = ds[3]
        assert set(item) == {"abd", "tnf", "bc"} and item["bc"] == names[3]
        assert ds.abd_cuda.is_cuda and np.array_equal(ds.abd_cuda.cpu().numpy(), oa)
    with pytest.raises(ValueError):
        Feature(_args(tmp_path / "x"), "unused").extract_features()


def test_feature_paired_mode_uses_min_qual(tmp_path, oracle):
    from conftest import GoldenCase
    from pangaea_b200 import Feature

    g = GoldenCase("synth_paired_minqual")
    p = g.params
    ft = Feature(_args(tmp_path, reads1=g.reads1, reads2=g.reads2, min_length=p["min_length"]), "unused")
    names, abd, tnf = ft.extract_features(write_cache=False)
    assert list(names) == list(g.abd_labels) and np.array_equal(abd, g.abd) and np.array_equal(tnf, g.tnf)


# ------------------------------------------------------------------------------------
# size-independent properties at a larger size (device-generated reads)
# ------------------------------------------------------------------------------------
def test_properties_at_scale():
    import torch

    from bench import make_synthetic_batch  # same generator the benchmark uses

    n_pairs, L = 400_000, 100
    ctx = _ctx()
    s = make_synthetic_batch(ctx, n_pairs=n_pairs, read_len=L, n_barcodes=4000, n_genomes=8, genome_len=300_000, seed=9)
    batch = ctx.adopt(s["reads"], keepalive=s)
    ctx.count(batch)
    # (1) counters sum to the number of valid windows (every window is counted exactly once)
    seq = s["seq"].cpu().numpy().reshape(2 * n_pairs, L + 1)[:, :L]
    valid = np.isin(seq, np.frombuffer(b"ACGT", dtype=np.uint8))
    def windows(k):
        c = np.cumsum(np.concatenate([np.zeros((valid.shape[0], 1), int), valid], axis=1), axis=1)
        return int(((c[:, k:] - c[:, :-k]) == k).sum())
    table = ctx.table_as_torch()
    assert int(table.to(torch.int64).sum()) == windows(15)
    # (2) featurize: TNF row sums = 4-mer windows of the emitted clouds; deterministic; linear in the table
    keep = np.ones(s["n_groups"], np.uint8)
    keep[0] = 0
    f1 = ctx.featurize(batch, keep)
    f2 = ctx.featurize(batch, keep)
    a1, t1 = f1.raw()
    a2, t2 = f2.raw()
    assert np.array_equal(a1, a2) and np.array_equal(t1, t2)
    assert f1.rows > 3000
    groups = f1.row_groups()
    per_read4 = ((np.cumsum(np.concatenate([np.zeros((valid.shape[0], 1), int), valid], axis=1), axis=1)[:, 4:]
                  - np.cumsum(np.concatenate([np.zeros((valid.shape[0], 1), int), valid], axis=1), axis=1)[:, :-4]) == 4).sum(axis=1)
    flags = s["flag"].cpu().numpy()
    gid = np.concatenate([[0], np.cumsum(flags & 1)[:-1]]).astype(np.int64)
    want = np.bincount(gid, weights=per_read4, minlength=s["n_groups"]).astype(np.int64)
    assert np.array_equal(t1.sum(axis=1), want[groups])
    # every look-up finds its k-mer (count >= 1) and almost all land inside the 400 bins
    per_read15 = ((np.cumsum(np.concatenate([np.zeros((valid.shape[0], 1), int), valid], axis=1), axis=1)[:, 15:]
                   - np.cumsum(np.concatenate([np.zeros((valid.shape[0], 1), int), valid], axis=1), axis=1)[:, :-15]) == 15).sum(axis=1)
    want15 = np.bincount(gid, weights=per_read15, minlength=s["n_groups"]).astype(np.int64)
    assert np.array_equal(a1.sum(axis=1), want15[groups])  # no count reaches 4000 at this depth
    # (3) counting the batch a second time doubles every counter: bins shift accordingly
    ctx.count(batch)
    assert int(ctx.table_as_torch().to(torch.int64).sum()) == 2 * windows(15)
    # (4) normalised rows sum to 1, weights in (0, 1]
    a, t, w = f1.normalized()
    assert np.allclose(a.sum(axis=1), 1, atol=1e-5) and np.allclose(t.sum(axis=1), 1, atol=1e-5)
    assert (w > 0).all() and (w <= 1).all()


# ------------------------------------------------------------------------------------
# the other BASELINE.json configurations as parity cases (reduced sizes)
# ------------------------------------------------------------------------------------
def test_config_tellseq_2x150_18bp_barcodes(tmp_path, oracle):
    """configs[2]: TELL-Seq after preprocessing - 2x150 bp, 18-bp barcodes in the BX tag."""
    data = synth.generate(n_barcodes=400, mean_pairs=12, read_len=150, n_genomes=4, genome_len=120_000, frag_len=20_000,
                          barcode_len=18, seed=3, unbarcoded_pairs=30)
    path = synth.write_interleaved(str(tmp_path / "tell.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    fq = _lib.Fastq(path)
    ctx = _ctx()
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    g_abd, g_tnf = feats.raw()
    assert _names(fq, feats) == list(names) and len(names) > 300 and all(len(n) == 18 for n in names)
    assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)


@pytest.mark.parametrize("env", [{}, {"PG_NO_SHARED": "1"}])
def test_config_hybrid_one_cloud_per_pair(tmp_path, oracle, monkeypatch, env):
    """configs[3]: plain short reads, every pair its own (virtual) barcode, min_length 0 - the reference-equivalent of
    per-read features.  Clouds are 302 bytes: a cloud boundary falls into every 10th word and a 512-word
    tile spans ~55 clouds (TNF slots overflow to global reductions)."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    data = synth.generate(n_barcodes=3000, mean_pairs=1, read_len=150, n_genomes=3, genome_len=100_000, frag_len=5_000, seed=4)
    keep_one = np.concatenate([[True], np.array(data["barcode"][1:]) != np.array(data["barcode"][:-1])])  # first pair of every barcode
    data = {"seq1": data["seq1"][keep_one], "seq2": data["seq2"][keep_one], "barcode": [b for b, k in zip(data["barcode"], keep_one) if k],
            "read_len": 150, "n_pairs": int(keep_one.sum())}
    path = synth.write_interleaved(str(tmp_path / "hybrid.fq"), data)
    names, abd, tnf = oracle.featurize(path, None, mlen=0)
    fq = _lib.Fastq(path)
    ctx = _ctx(min_length=0)
    feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
    g_abd, g_tnf = feats.raw()
    assert _names(fq, feats) == list(names) and len(names) == data["n_pairs"] - 1  # the quirk drops the last label
    assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    assert (g_tnf.sum(axis=1) <= 2 * 147).all()  # one pair per row


def test_shared_partition_equals_separate_partitions_at_scale(monkeypatch):
    """The shared partition of pg_count (reused by pg_featurize) and the two separate partitions give identical tables
    and matrices on 1.5 M device-generated 2x150 pairs - including lower-case-free count-only handling and dropped clouds."""
    import torch

    from bench import make_synthetic_batch

    results = []
    for no_shared in ("0", "1"):
        monkeypatch.setenv("PG_NO_SHARED", no_shared)
        ctx = _ctx(min_length=30_200)  # drops about half of the clouds (Poisson(100) pairs x 302 bytes)
        s = make_synthetic_batch(ctx, n_pairs=1_500_000, read_len=150, n_barcodes=15_000, n_genomes=10, genome_len=400_000, seed=11)
        batch = ctx.adopt(s["reads"], keepalive=s)
        ctx.count(batch)
        keep = np.ones(s["n_groups"], np.uint8)
        keep[0] = 0
        keep[5::7] = 0  # some clouds dropped by label as well
        f = ctx.featurize(batch, keep)
        abd, tnf = f.raw()
        results.append((abd, tnf, f.row_groups(), int(ctx.table_as_torch().to(torch.int64).sum()), ctx.table_size()))
        stages = {n: ctx.timing(w)[0] for n, w in (("count_scatter", _lib.T_COUNT_SCATTER), ("feat_scatter", _lib.T_FEAT_SCATTER))}
        assert (stages["feat_scatter"] == 0) == (no_shared == "0")  # the shared path really ran / was really switched off
        f.free(); batch.free(); ctx.close()
    (a0, t0, g0, s0, n0), (a1, t1, g1, s1, n1) = results
    assert 3000 < len(g0) < 14000 and np.array_equal(g0, g1)
    assert s0 == s1 and n0 == n1
    assert np.array_equal(a0, a1) and np.array_equal(t0, t1)


def test_two_batches_one_table(tmp_path, oracle):
    """Multi-batch flow of INTEGRATION.md: count every batch first (the table accumulates), then featurize each.  Both batches
    keep their shared-partition entries alive at the same time; buffers are freed in a different order than they were made, one
    matrix outlives its context through DLPack."""
    import torch

    paths = []
    for seed in (21, 22):
        data = synth.generate(n_barcodes=120, mean_pairs=15, read_len=100, n_genomes=3, genome_len=50_000, frag_len=8_000, seed=seed,
                              unbarcoded_pairs=10, lower_rate=0.001)
        paths.append(synth.write_interleaved(str(tmp_path / f"b{seed}.fq"), data))
    table = oracle.count_fastq(paths, 15)
    want = [oracle.featurize(p, None, table=table) for p in paths]
    ctx = _ctx()
    fqs = [_lib.Fastq(p) for p in paths]
    batches = [ctx.upload(fq.reads) for fq in fqs]
    for b in batches:
        ctx.count(b)
    keys, counts = ctx.table_export()
    wk, wv = _oracle_table_arrays(table)
    assert np.array_equal(keys, wk) and np.array_equal(counts.astype(np.uint64), wv)
    feats = [ctx.featurize(b, fq.group_keep, fq.n_groups) for b, fq in zip(reversed(batches), reversed(fqs))][::-1]
    batches[0].free()
    for (names, abd, tnf), f, fq in zip(want, feats, fqs):
        g_abd, g_tnf = f.raw()
        assert _names(fq, f) == list(names)
        assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    t = feats[1].torch(_lib.ABD_RAW)
    feats[0].free(); feats[1].free(); batches[1].free()
    ctx.close()
    assert np.array_equal(t.cpu().numpy(), want[1][1])
    del t
    torch.cuda.synchronize()


def test_trim_releases_idle_memory_and_the_ctx_keeps_working(tmp_path, oracle):
    """pg_trim: the blocks the ctx caches for reuse (freed matrices, partitions, packed streams) go back to the driver; live
    objects - a batch, its feature set, the k-mer table - are untouched and the next call allocates afresh."""
    data = synth.generate(n_barcodes=200, mean_pairs=20, read_len=100, n_genomes=3, genome_len=60_000, frag_len=8_000, seed=77)
    path = synth.write_interleaved(str(tmp_path / "t.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    ctx = _ctx()
    fq = _lib.Fastq(path)
    held = ctx.upload(fq.reads)
    ctx.count(held)
    first = ctx.featurize(held, fq.group_keep, fq.n_groups)
    for _ in range(3):  # churn: every round returns its blocks to the cache
        b = ctx.upload(fq.reads)
        f = ctx.featurize(b, fq.group_keep, fq.n_groups)
        f.normalize()
        f.free(); b.free()
    free_before, total = ctx.mem_info()
    ctx.trim()
    free_after, _ = ctx.mem_info()
    assert 0 < free_after <= total and free_after >= free_before - (64 << 20)  # (what was idle is now plain free memory)
    g_abd, g_tnf = first.raw()  # made before the trim
    assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    again = ctx.featurize(held, fq.group_keep, fq.n_groups)  # after it
    a_abd, a_tnf = again.raw()
    assert _names(fq, again) == list(names)
    assert np.array_equal(a_abd, abd) and np.array_equal(a_tnf, tnf)
    ctx.trim()  # idempotent, also with nothing cached
    first.free(); again.free(); held.free()
    ctx.close()


def test_pipelined_upload_with_unpaired_reads_and_lower_case(tmp_path, oracle):
    """pg_extract_features on > 1 MB goes through the chunked upload (the count pass runs while later chunks are still
    crossing PCIe).  Paired files with mismatching R1/R2 names (counted, but in no cloud: PG_READ_NOFEAT) and lower-case
    bases (counted, not featurized) exercise the count-only windows on that path."""
    data = synth.generate(n_barcodes=500, mean_pairs=12, read_len=100, n_genomes=4, genome_len=100_000, frag_len=10_000, seed=31,
                          unbarcoded_pairs=50, lower_rate=0.002)
    p1, p2 = synth.write_paired(str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"), data)
    lines = open(p2, "rb").read().split(b"\n")
    for i in range(0, len(lines) - 4, 4 * 37):  # every 37th R2 record gets another name
        lines[i] = b"@other" + lines[i][1:]
    open(p2, "wb").write(b"\n".join(lines))
    names, abd, tnf = oracle.featurize(p1, p2)
    fq = _lib.Fastq(p1, p2)
    assert fq.reads.n_bytes > (1 << 20) and (fq.arrays()[2] & _lib.PG_READ_NOFEAT).any()
    for seg in (None, "8192"):  # one chunk / many chunks
        if seg:
            os.environ["PG_SEG_WORDS"] = seg
        try:
            ctx = _ctx()
            feats = ctx.extract_features(fq.reads, fq.group_keep, fq.n_groups)
            g_abd, g_tnf = feats.raw()
            assert _names(fq, feats) == list(names)
            assert np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
            keys, counts = ctx.table_export()
            wk, wv = _oracle_table_arrays(oracle.count_fastq([p1, p2], 15))
            assert np.array_equal(keys, wk) and np.array_equal(counts.astype(np.uint64), wv)
        finally:
            os.environ.pop("PG_SEG_WORDS", None)


# ------------------------------------------------------------------------------------
# streaming: files of any size in batches (pangaea_b200/stream.py)
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch_bytes,resident", [(60_000, 0.45), (250_000, 0.45), (60_000, 0.0), (10 ** 9, 0.45)])
def test_streamed_batches_equal_one_batch_equal_oracle(tmp_path, oracle, batch_bytes, resident):
    """count all batches, then featurize each (packed batches resident in HBM, or - resident = 0 - the file parsed twice):
    rows == the oracle over the whole file."""
    from pangaea_b200 import stream

    data = synth.generate(n_barcodes=120, mean_pairs=18, read_len=100, n_genomes=3, genome_len=60_000, frag_len=9_000, seed=77,
                          unbarcoded_pairs=40, n_rate=0.002)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    ctx = _ctx()
    n_batches = len(list(_lib.FastqStream(path, target_seq_bytes=batch_bytes)))
    assert n_batches > 1 if batch_bytes < 10 ** 6 else n_batches == 1
    g_names, feats = stream.extract_features_streaming(
        ctx, lambda: _lib.FastqStream(path, pinned=True, target_seq_bytes=batch_bytes), resident_fraction=resident)
    g_abd, g_tnf = feats.raw()
    assert g_names == list(names) and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    a, t, w = feats.normalized()
    oa, ot, ow = oracle.data_init(abd, tnf)
    assert np.array_equal(a, oa) and np.array_equal(t, ot) and np.array_equal(w, ow)


def test_feature_streams_gzip_and_paired_input(tmp_path, oracle):
    import gzip
    import shutil

    from conftest import GoldenCase
    from pangaea_b200 import Feature

    g = GoldenCase("synth_paired_minqual")
    ft = Feature(_args(tmp_path / "p", reads1=g.reads1, reads2=g.reads2, min_length=g.params["min_length"]), "unused", batch_seq_bytes=3000)
    names, abd, tnf = ft.extract_features(write_cache=False)
    assert list(names) == list(g.abd_labels) and np.array_equal(abd, g.abd) and np.array_equal(tnf, g.tnf)
    g = GoldenCase("synth_10x_l2000")
    gz = tmp_path / "reads.fq.gz"
    with open(g.path1, "rb") as a, gzip.open(gz, "wb") as b:
        shutil.copyfileobj(a, b)
    want = oracle.featurize(g.path1, None, mlen=g.params["min_length"])
    ft = Feature(_args(tmp_path / "z", interleaved_reads=str(gz), min_length=g.params["min_length"]), "unused", batch_seq_bytes=20_000)
    names, abd, tnf = ft.extract_features(write_cache=False)
    assert list(names) == list(want[0]) and np.array_equal(abd, want[1]) and np.array_equal(tnf, want[2])
    assert ft.features.rows == len(names)  # all batches' rows as one device-resident feature set
    assert ft.ingest_used == "host"  # gzip: the host reader
    # the same file as plain text: parsed on the device, in several 44 kB windows
    ft = Feature(_args(tmp_path / "d", interleaved_reads=g.path1, min_length=g.params["min_length"]), "unused", batch_seq_bytes=20_000)
    names, abd, tnf = ft.extract_features(write_cache=False)
    assert ft.ingest_used == "device" and os.path.getsize(g.path1) > 3 * 44_445
    assert list(names) == list(want[0]) and np.array_equal(abd, want[1]) and np.array_equal(tnf, want[2])
    with pytest.raises(ValueError):  # the device parser cannot read gzip: asking for it explicitly is an error, not a silent fallback
        Feature(_args(tmp_path / "e", interleaved_reads=str(gz)), "unused", ingest="device").extract_features(write_cache=False)


def test_data_from_device_features_applies_the_csv_rounding(tmp_path, oracle):
    """KAT-5 through the drop-in: clouds with a tally >= 10^6 (bin 0 of 8 000-pair clouds at ~8x coverage).  The reference
    normalises what pandas read from 6-significant-digit text (count_kmer.cpp:211); Data(features=Feature.features) must too."""
    from pangaea_b200 import Data, Feature

    data = synth.generate(n_barcodes=3, mean_pairs=8000, read_len=100, n_genomes=1, genome_len=600_000, frag_len=600_000, seed=9)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    names, abd, tnf = oracle.featurize(path, None)
    assert abd.max() >= 1_000_000 and (abd.max(axis=1) % 10 != 0).any() and ((abd > 0).sum(axis=1) >= 2).all(), "the case must exercise the rounding"
    ft = Feature(_args(tmp_path, interleaved_reads=path), "unused")
    g_names, g_abd, g_tnf = ft.extract_features(write_cache=False)
    assert g_abd.dtype == np.float64 and np.array_equal(g_abd, oracle.text_round(abd)) and np.array_equal(g_tnf, oracle.text_round(tnf))
    oa, ot, ow = oracle.data_init(oracle.text_round(abd), oracle.text_round(tnf))
    for ds in (Data(g_names, g_abd, g_tnf), Data(g_names, g_abd, g_tnf, features=ft.features)):
        assert np.array_equal(ds.abd, oa) and np.array_equal(ds.tnf, ot) and np.array_equal(ds.weights, ow)
    exact = oracle.data_init(abd, tnf)[0]
    assert not np.array_equal(exact, oa), "normalising the unrounded tallies gives different floats"


# ------------------------------------------------------------------------------------
# every BASELINE configuration, production geometry (no PG_* switches): a prefix against the oracle, the full size against
# the direct path
# ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_config_prefix_matches_oracle_with_default_geometry(tmp_path, oracle, name):
    """The first 150 k pairs of each bench configuration (bench.py CONFIGS: read length, pairs per barcode, label length,
    -l), generated on the device exactly as the bench generates them, written out as FASTQ and featurized by the oracle;
    the GPU side runs with the default segment / slice / region geometry, through the host parser and the streaming driver."""
    import bench
    from pangaea_b200 import stream

    cfg = bench.CONFIGS[name]
    n = 150_000
    for var in ("PG_SEG_WORDS", "PG_FEAT_SEG_WORDS", "PG_REGION_SLACK", "PG_FORCE_DIRECT", "PG_NO_SHARED", "PG_COUNT_L2", "PG_STASH_SEGMENTS"):
        assert var not in os.environ
    ctx = _ctx(min_length=cfg["min_length"])
    s = synth.device_batch(ctx, n, cfg["read_len"], n_barcodes=max(1, n // cfg["pairs_per_barcode"]), n_genomes=cfg["n_genomes"], seed=cfg["seed"])
    path = synth.write_batch_fastq(str(tmp_path / "prefix.fq"), s["seq"].cpu().numpy(), cfg["read_len"], s["bc_start"], n,
                                   barcode_len=cfg["barcode_len"])
    names, abd, tnf = oracle.featurize(path, None, mlen=cfg["min_length"])
    # (a) device-resident batch, as the bench runs it
    keep = np.ones(s["n_groups"], np.uint8)
    keep[0] = 0
    b = ctx.adopt(s["reads"], keepalive=s)
    ctx.count(b)
    f = ctx.featurize(b, keep)
    g_abd, g_tnf = f.raw()
    assert f.rows == len(names) and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    f.free(); b.free()
    # (b) the drop-in route: host parser -> streamed batches
    g_names, f = stream.extract_features_streaming(ctx, lambda: _lib.FastqStream(path, pinned=True, target_seq_bytes=40_000_000))
    g_abd, g_tnf = f.raw()
    assert g_names == list(names) and np.array_equal(g_abd, abd) and np.array_equal(g_tnf, tnf)
    a, t, w = f.normalized()
    oa, ot, ow = oracle.data_init(abd, tnf)
    assert np.array_equal(a, oa) and np.array_equal(t, ot) and np.array_equal(w, ow)


def test_full_size_sliced_path_equals_direct_path(monkeypatch):
    """The headline workload at its FULL size (50 M pairs, 5 segments of 2^26 words, 64 slices at real fill, sub-regions at
    real fill): the sliced production path against the direct path (count_kernel / featurize_kernel - no partitioning at
    all, pinned to the oracle by the small tests above).  Tables and matrices must agree bit for bit."""
    import torch

    import bench

    cfg = bench.CONFIGS["c2"]
    n = cfg["pairs"]
    results = {}
    for direct in ("0", "1"):
        monkeypatch.setenv("PG_FORCE_DIRECT", direct)
        ctx = _ctx()
        s = synth.device_batch(ctx, n, cfg["read_len"], n_barcodes=n // cfg["pairs_per_barcode"], n_genomes=cfg["n_genomes"], seed=cfg["seed"])
        keep = np.ones(s["n_groups"], np.uint8)
        keep[0] = 0
        b = ctx.adopt(s["reads"], keepalive=s)
        ctx.count(b)
        f = ctx.featurize(b, keep)
        ctx.synchronize()
        sliced = ctx.timing(_lib.T_COUNT_SCATTER)[0] > 0
        assert sliced == (direct == "0")
        results[direct] = (ctx.table_as_torch().clone(), f.torch(_lib.ABD_RAW).clone(), f.torch(_lib.TNF_RAW).clone())
        f.free(); b.free(); del s
        ctx.close()
        torch.cuda.empty_cache()
    (t0, a0, n0), (t1, a1, n1) = results["0"], results["1"]
    assert a0.shape[0] > 400_000
    assert torch.equal(t0, t1), "k-mer tables differ"
    assert torch.equal(a0, a1) and torch.equal(n0, n1), "feature matrices differ"


def test_compacted_adopted_batch_survives_the_callers_buffers():
    """pg_batch_compact on an ADOPTED batch copies the read offsets / flags: the caller may free (and overwrite) its buffers
    and pg_featurize still sees the batch (the streamed bench configurations rely on it)."""
    import torch

    ctx = _ctx()
    s = synth.device_batch(ctx, 60_000, 100, n_barcodes=600, n_genomes=4, genome_len=200_000, seed=5)
    keep = np.ones(s["n_groups"], np.uint8)
    keep[0] = 0
    b = ctx.adopt(s["reads"])
    ctx.count(b, keep_partition=False)
    want = ctx.featurize(b, keep).raw()
    b.compact()
    ctx.synchronize()
    for t in ("_seq_full", "off", "flag", "seq"):
        s[t].fill_(255 if s[t].dtype == torch.uint8 else -1)   # the caller recycles its memory
    torch.cuda.synchronize()
    got = ctx.featurize(b, keep).raw()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    with pytest.raises(_lib.PgError, match="compacted"):
        b.download()
