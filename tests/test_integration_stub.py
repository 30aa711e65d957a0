"""The ctypes stub printed in INTEGRATION.md §3 is executable documentation: run it verbatim against the built library
and compare with the golden outputs of the reference tools."""
import argparse
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    assert blocks, "no python block in INTEGRATION.md"
    src = blocks[0]
    return src.replace('C.CDLL("libpangaea_b200.so")', f'C.CDLL({os.path.join(ROOT, "pangaea_b200", "libpangaea_b200.so")!r})')


def test_stub_binds_only_declared_symbols():
    from pangaea_b200 import _lib

    used = set(re.findall(r"\bL\.(pg_[a-z0-9_]+)", _stub_source()))
    assert used and used <= set(_lib.SIGNATURES), used - set(_lib.SIGNATURES)


@pytest.mark.gpu
def test_stub_reproduces_golden_outputs(golden):
    if golden.name == "kat4_bins" or golden.params.get("min_qual"):
        pytest.skip("hand-made dump / quality-filtered case: the stub counts with pg_count and takes qualities only in paired mode")
    p = golden.params
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    args = argparse.Namespace(kmer=p["k"], tnf_kmer=p["tnf_k"], window_size=p["window_size"], vector_size=p["vector_size"],
                              min_length=p["min_length"], reads1=None, reads2=None, interleaved_reads=golden.path1)
    if golden.reads2:
        pytest.skip("paired golden cases use --min-qual-char")
    names, abd, tnf = ns["run_b200"](args)
    # the golden abundance was computed from the same reads' own counts (make_golden.py), so the whole path must match
    assert list(names) == list(golden.abd_labels)
    assert np.array_equal(tnf, golden.tnf) and np.array_equal(abd, golden.abd)
