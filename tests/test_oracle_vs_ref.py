"""CPU, differential: the C restatement against the compiled reference binaries in
oracle/_ref on seeded random inputs (skipped where oracle/_ref was not built)."""
import os

import numpy as np
import pytest

from pangaea_b200 import synth


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle


@pytest.mark.parametrize("seed,style,mode", [(1, "10x", "i"), (2, "stlfr", "p"), (3, "10x", "p"), (4, "stlfr", "i")])
def test_random_inputs(ref, tmp_path, seed, style, mode):
    O = ref
    rng = np.random.default_rng(seed)
    data = synth.generate(n_barcodes=int(rng.integers(3, 30)), mean_pairs=int(rng.integers(2, 25)),
                          read_len=int(rng.integers(20, 160)), n_genomes=2, genome_len=20_000, frag_len=4_000,
                          seed=seed, unbarcoded_pairs=int(rng.integers(0, 6)), n_rate=0.003, lower_rate=0.002)
    if mode == "i":
        p1, p2 = synth.write_interleaved(str(tmp_path / "i.fq"), data, style=style), None
        kw = dict(interleaved=p1)
    else:
        p1, p2 = synth.write_paired(str(tmp_path / "1.fq"), str(tmp_path / "2.fq"), data, style=style)
        kw = dict(reads1=p1, reads2=p2)
    k, ws, vs = int(rng.integers(5, 20)), int(rng.integers(1, 4)), int(rng.integers(3, 50))
    mlen = int(rng.choice([0, 500, 2000]))
    t = O.count_fastq([p for p in (p1, p2) if p], k)
    dump = str(tmp_path / "d.dump")
    t.write_dump(dump, k)
    labels, abd = O.abundance(p1, p2, t, k, mlen, vs, ws)
    rl, rabd = O.ref_count_kmer(str(tmp_path / "a.gz"), dump, k=k, mlen=mlen, vs=vs, ws=ws, **kw)
    assert list(labels) == list(rl)
    assert len(rl) == 0 or np.array_equal(abd, rabd)
    labels, tnf = O.tnf(p1, p2, 4, mlen)
    rl, rtnf = O.ref_count_tnf(str(tmp_path / "t.gz"), k=4, mlen=mlen, **kw)
    assert list(labels) == list(rl)
    assert len(rl) == 0 or np.array_equal(tnf, rtnf)


def test_kat6_thread_count_does_not_change_output(ref, tmp_path):
    data = synth.generate(n_barcodes=30, mean_pairs=10, read_len=60, seed=9)
    fq = synth.write_interleaved(str(tmp_path / "i.fq"), data)
    a = ref.ref_count_tnf(str(tmp_path / "t1.gz"), interleaved=fq, mlen=0, threads=1)
    b = ref.ref_count_tnf(str(tmp_path / "t8.gz"), interleaved=fq, mlen=0, threads=8)
    assert list(a[0]) == list(b[0]) and np.array_equal(a[1], b[1])


def test_kat7_min_length_counts_separators(ref, tmp_path):
    """Σ(len+1) <= -l drops the row (count_tnf.cpp:81): 3 pairs of 2x30 -> 186."""
    with open(tmp_path / "i.fq", "wb") as f:
        for i, bc in enumerate([b"AA"] + [b"AA"] * 3 + [b"CC"]):
            for _ in range(2):
                f.write(b"@r%d BX:Z:%s-1\n%s\n+\n%s\n" % (i, bc, b"ACGTTGCAAC" * 3, b"I" * 30))
    fq = str(tmp_path / "i.fq")
    # cloud AA holds pairs 1..4 (pair 0 is lost to the quirk, pair 4 = first CC pair) = 4 * 62 = 248
    for mlen, want in ((247, ["AA"]), (248, [])):
        labels, _ = ref.ref_count_tnf(str(tmp_path / f"t{mlen}.gz"), interleaved=fq, mlen=mlen)
        assert list(labels) == want
        assert list(ref.tnf(fq, None, 4, mlen)[0]) == want


def test_kat5_csv_precision_loss_is_text_only(ref, tmp_path):
    """Tallies >= 1e6 print as 1.xxxxxe+06 (ostream precision 6, count_tnf.cpp:204);
    the oracle keeps exact integers, the text rounds them."""
    n_pairs = 2200
    with open(tmp_path / "i.fq", "wb") as f:
        for i in range(n_pairs + 2):
            bc = b"AA" if i <= n_pairs else b"CC"
            for _ in range(2):
                f.write(b"@r%d BX:Z:%s-1\n%s\n+\n%s\n" % (i, bc, b"A" * 250, b"I" * 250))
    fq = str(tmp_path / "i.fq")
    labels, tnf = ref.tnf(fq, None, 4, 0)
    exact = int(tnf[0, 0])
    assert exact == (n_pairs + 1) * 2 * 247 and exact >= 1_000_000
    rl, rt = ref.ref_count_tnf(str(tmp_path / "t.gz"), interleaved=fq, mlen=0)
    assert rt.dtype == np.float64  # pandas sees scientific notation
    assert rt[0, 0] == float("%.6g" % exact) and rt[0, 0] != exact


def _hostile_text(seed):
    """Interleaved FASTQ text built from the header spellings getBarcode (count_kmer.cpp:25-53) has to cope with: BX tags with
    and without the "-1" suffix, behind a tab or a blank, in the middle of other tags, missing; stLFR '#' labels with and
    without the "/1" suffix, the 0_0_0 label, a second '#'; both kinds in one file (the first decisive header latches the type);
    reads shorter than k, empty reads, CRLF, N / lower case / other bytes; sometimes a truncated last record."""
    rng = np.random.default_rng(seed)
    n_clouds = int(rng.integers(2, 12))
    first_kind = "10x" if rng.random() < 0.5 else "stlfr"
    recs = []
    rid = 0
    for c in range(n_clouds):
        bc10 = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(1, 17))))
        bcst = b"%d_%d_%d" % tuple(int(x) for x in rng.integers(0 if rng.random() < 0.15 else 1, 1500, size=3))
        if rng.random() < 0.1:
            bcst = b"0_0_0"
        for _ in range(int(rng.integers(1, 9))):
            kind = first_kind if (c == 0 or rng.random() < 0.85) else ("stlfr" if first_kind == "10x" else "10x")
            r = rng.random()
            if kind == "10x":
                sep = b"\t" if rng.random() < 0.5 else b" "
                if r < 0.55:
                    tag = sep + b"BX:Z:" + bc10 + b"-1"
                elif r < 0.7:
                    tag = sep + b"BX:Z:" + bc10
                elif r < 0.85:
                    tag = sep + b"RX:Z:NNN" + sep + b"BX:Z:" + bc10 + b"-1" + sep + b"QX:Z:III"
                else:
                    tag = b""
                head = b"@r%d" % rid + tag
            else:
                if r < 0.6:
                    head = b"@r%d#" % rid + bcst + b"/%d"
                elif r < 0.8:
                    head = b"@r%d#" % rid + bcst
                elif r < 0.9:
                    head = b"@r#%d#" % rid + bcst + b"/%d"
                else:
                    head = b"@r%d" % rid
            eol = b"\r\n" if rng.random() < 0.05 else b"\n"
            for mate in (1, 2):
                L = int(rng.choice([0, 3, 9, 14, 15, 16, 40, 100, 151]))
                s = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L).tobytes())
                for _ in range(int(rng.integers(0, 3))):
                    if L:
                        s[int(rng.integers(0, L))] = int(rng.choice(np.frombuffer(b"Nacgt.R", np.uint8)))
                h = head % mate if b"%d" in head else head
                recs.append(h + eol + bytes(s) + eol + b"+" + eol + b"I" * L + eol)
            rid += 1
    if rng.random() < 0.3:
        recs.append(b"@tail\tBX:Z:ACGT-1\nACGTACGTACGTACGTACGT\n")  # truncated: R1 sequence only
    return b"".join(recs)


@pytest.mark.parametrize("seed", range(100, 112))
def test_hostile_headers_and_ragged_reads(ref, tmp_path, seed):
    """The restatement against the reference binaries on hostile interleaved text (labels, grouping off-by-one, the latch,
    min-length with separators, short / empty / dirty reads)."""
    O = ref
    rng = np.random.default_rng(seed)
    path = str(tmp_path / "h.fq")
    open(path, "wb").write(_hostile_text(seed))
    k, ws, vs = int(rng.choice([5, 9, 15])), int(rng.integers(1, 4)), int(rng.integers(3, 30))
    mlen = int(rng.choice([0, 30, 300]))
    t = O.count_fastq([path], k)
    dump = str(tmp_path / "d.dump")
    t.write_dump(dump, k)
    labels, abd = O.abundance(path, None, t, k, mlen, vs, ws)
    rl, rabd = O.ref_count_kmer(str(tmp_path / "a.gz"), dump, k=k, mlen=mlen, vs=vs, ws=ws, interleaved=path, raw=True)
    assert list(labels) == list(rl)
    assert len(rl) == 0 or np.array_equal(abd, rabd)
    tk = int(rng.choice([3, 4]))
    labels, tnf = O.tnf(path, None, tk, mlen)
    rl, rtnf = O.ref_count_tnf(str(tmp_path / "t.gz"), k=tk, mlen=mlen, interleaved=path, raw=True)
    assert list(labels) == list(rl)
    assert len(rl) == 0 or np.array_equal(tnf, rtnf)


def _hostile_pair_files(seed, tmp_path):
    rng = np.random.default_rng(seed + 7)
    text = _hostile_text(seed).replace(b"\r\n", b"\n")
    lines = text.split(b"\n")
    recs = [lines[i:i + 4] for i in range(0, len(lines) - 3, 4)]
    recs = recs[: len(recs) // 2 * 2]
    r1, r2 = [], []
    for i in range(0, len(recs), 2):
        a, b = list(recs[i]), list(recs[i + 1])
        u = rng.random()
        if u < 0.12:
            b[0] = b[0].replace(b"@r", b"@x", 1)                      # another read name
        elif u < 0.24:
            b[0] = b[0].replace(b"BX:Z:", b"BX:Z:T").replace(b"#", b"#9", 1)  # another barcode
        elif u < 0.30:
            b[0] = b[0].split(b"\t")[0].split(b" ")[0].split(b"#")[0]  # no barcode at all
        r1.append(b"\n".join(a) + b"\n")
        r2.append(b"\n".join(b) + b"\n")
    p1, p2 = str(tmp_path / "1.fq"), str(tmp_path / "2.fq")
    open(p1, "wb").write(b"".join(r1))
    open(p2, "wb").write(b"".join(r2))
    return p1, p2


@pytest.mark.parametrize("seed", range(200, 208))
def test_hostile_paired_files(ref, tmp_path, seed):
    """Paired mode (count_kmer.cpp:181-235): the hostile records dealt to two files, then some R2 headers changed - another
    read name, another barcode, no barcode - so that pairs disagree (they are k-mer counted but belong to no cloud)."""
    O = ref
    rng = np.random.default_rng(seed)
    p1, p2 = _hostile_pair_files(seed, tmp_path)
    k, ws, vs = int(rng.choice([5, 9, 15])), int(rng.integers(1, 4)), int(rng.integers(3, 30))
    mlen = int(rng.choice([0, 30, 300]))
    t = O.count_fastq([p1, p2], k)
    dump = str(tmp_path / "d.dump")
    t.write_dump(dump, k)
    labels, abd = O.abundance(p1, p2, t, k, mlen, vs, ws)
    rl, rabd = O.ref_count_kmer(str(tmp_path / "a.gz"), dump, k=k, mlen=mlen, vs=vs, ws=ws, reads1=p1, reads2=p2, raw=True)
    assert list(labels) == list(rl)
    assert len(rl) == 0 or np.array_equal(abd, rabd)
    labels, tnf = O.tnf(p1, p2, 4, mlen)
    rl, rtnf = O.ref_count_tnf(str(tmp_path / "t.gz"), k=4, mlen=mlen, reads1=p1, reads2=p2, raw=True)
    assert list(labels) == list(rl)
    assert len(rl) == 0 or np.array_equal(tnf, rtnf)
