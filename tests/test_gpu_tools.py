"""GPU: the text tools of SURVEY.md §8f rows 2 and 4 (pangaea_b200/tools.py -> csrc/transform.cuh) against golden files written
by the unmodified reference binaries preprocess_stlfr, preprocess_tellseq and extract_reads (tests/golden/make_golden_ingest.py)."""
import filecmp
import gzip
import os
import shutil

import pytest

from pangaea_b200 import _lib, tools

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    assert open(a, "rb").read() == open(b, "rb").read(), (a, b)


@pytest.mark.parametrize("tag,library,r2", [("n", False, "r2.fq"), ("nl", True, "r2.fq"), ("short", True, "r2_short.fq")])
def test_preprocess_stlfr_matches_reference_tool(tmp_path, tag, library, r2):
    d = os.path.join(GOLD, "ingest_stlfr")
    o1, o2 = tools.preprocess_stlfr(os.path.join(d, "r1.fq"), os.path.join(d, r2), str(tmp_path / "out"), number=True, library=library)
    _same(o1, os.path.join(d, f"out_{tag}_1.fq"))
    _same(o2, os.path.join(d, f"out_{tag}_2.fq"))


def test_preprocess_stlfr_gzip_input_and_malformed_headers(tmp_path):
    d = os.path.join(GOLD, "ingest_stlfr")
    for n in ("r1.fq", "r2.fq"):
        with open(os.path.join(d, n), "rb") as a, gzip.open(tmp_path / (n + ".gz"), "wb") as b:
            shutil.copyfileobj(a, b)
    o1, o2 = tools.preprocess_stlfr(str(tmp_path / "r1.fq.gz"), str(tmp_path / "r2.fq.gz"), str(tmp_path / "gz"), library=True)
    _same(o1, os.path.join(d, "out_nl_1.fq"))
    (tmp_path / "bad.fq").write_bytes(b"@no_barcode_here\nACGT\n+\nIIII\n")   # the reference tool aborts (std::out_of_range)
    with pytest.raises(_lib.PgError):
        tools.preprocess_stlfr(str(tmp_path / "bad.fq"), str(tmp_path / "bad.fq"), str(tmp_path / "bad"))
    (tmp_path / "empty.fq").write_bytes(b"")
    o1, o2 = tools.preprocess_stlfr(str(tmp_path / "empty.fq"), str(tmp_path / "empty.fq"), str(tmp_path / "e"))
    assert os.path.getsize(o1) == 0 and os.path.getsize(o2) == 0
    with pytest.raises(NotImplementedError):
        tools.preprocess_stlfr(str(tmp_path / "empty.fq"), str(tmp_path / "empty.fq"), str(tmp_path / "e"), number=False)


def test_preprocess_tellseq_matches_reference_tool(tmp_path):
    d = os.path.join(GOLD, "ingest_tellseq")
    o1, o2, wl = tools.preprocess_tellseq(os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq"), os.path.join(d, "i1.fq"), str(tmp_path / "out"))
    _same(o1, os.path.join(d, "out_1.fq"))
    _same(o2, os.path.join(d, "out_2.fq"))
    _same(wl, os.path.join(d, "out.wl"))


@pytest.mark.parametrize("style", ["tenx", "stlfr"])
def test_extract_reads_matches_reference_tool(tmp_path, style):
    d = os.path.join(GOLD, "ingest_extract")
    od = tmp_path / "out"
    od.mkdir()
    tools.extract_reads(os.path.join(d, style + ".fq"), os.path.join(d, f"clusters_{style}.tsv"), str(od / "x"))
    want = os.path.join(d, "out_" + style)
    assert sorted(os.listdir(od)) == sorted(os.listdir(want))
    for f in os.listdir(want):
        _same(str(od / f), os.path.join(want, f))
    assert os.path.getsize(os.path.join(want, "x_bin0.fq")) > 1000 and os.path.getsize(os.path.join(want, "x_bin9.fq")) == 0


def test_sort_by_barcode_file_to_file(tmp_path):
    from oracle import ingest_oracle as I

    src = open(os.path.join(GOLD, "ingest_extract", "tenx.fq"), "rb").read().replace(b"\tBX:Z:", b" BX:Z:")
    recs = [b"".join(r) for r in zip(*[iter(src.splitlines(keepends=True))] * 8)]
    shuffled = b"".join(reversed(recs))
    (tmp_path / "in.fq").write_bytes(shuffled)
    tools.sort_by_barcode(str(tmp_path / "in.fq"), str(tmp_path / "sorted.fq"))
    assert (tmp_path / "sorted.fq").read_bytes() == I.barcode_sort(shuffled)
