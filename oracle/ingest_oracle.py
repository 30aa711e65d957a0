"""CPU restatement of what sits either side of the featurization path in run_pangaea - TEST
INFRASTRUCTURE ONLY: imported by tests/ and tools/, never by pangaea_b200/.

* barcode_sort: src/run_pangaea:237-252.  The awk program is restated here (the reference needs gawk's three-argument
  match(); this image has mawk only), the rest of the pipeline - `LANG=C sort -k1,1 | cut -f2- | tr "\\t" "\\n"` - is the
  reference's own command line run with the coreutils of this image.  PINNED by construction for the sort order (GNU
  sort decides it); the awk restatement is unpinned and says so.
* preprocess_stlfr / preprocess_tellseq / extract_reads: the compiled reference tools themselves (oracle/_ref, built by
  oracle/Makefile from /root/reference/src/cpptools) are the checker; the golden vectors under tests/golden/ingest_* were
  produced by them (tests/golden/make_golden_ingest.py).
"""
from __future__ import annotations

import os
import re
import subprocess

_TAG = re.compile(rb"BX:Z:[^ \t\n\v\f\r]+")  # [^[:space:]]+ in the C locale


def awk_tag_lines(text: bytes) -> bytes:
    """run_pangaea:237-251 - the awk program: for every line that starts with '@' (as the main loop meets it) the next seven
    lines are pulled in with getline and joined with tabs; the key is the first BX:Z:<non-space> match in the header, or
    "~~~".  A getline at end of input leaves `line` unchanged (awk semantics), so a truncated record repeats its last line."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    out = []
    i = 0
    while i < len(lines):
        hdr = lines[i]
        i += 1
        if not hdr.startswith(b"@"):
            continue
        block = [hdr]
        line = b""  # awk: an unset variable
        for _ in range(7):
            if i < len(lines):
                line = lines[i]
                i += 1
            block.append(line)
        m = _TAG.search(hdr)
        out.append((m.group(0) if m else b"~~~") + b"\t" + b"\t".join(block) + b"\n")
    return b"".join(out)


def barcode_sort(text: bytes) -> bytes:
    """The whole step: awk (restated) | LANG=C sort -k1,1 | cut -f2- | tr "\\t" "\\n" (the reference's own commands)."""
    env = dict(os.environ, LANG="C", LC_ALL="C")
    p = subprocess.run("sort -k1,1 | cut -f2- | tr '\\t' '\\n'", shell=True, input=awk_tag_lines(text), stdout=subprocess.PIPE, env=env, check=True)
    return p.stdout


def weighted_choice(weights, num_samples, replacement):
    """CustomWeightedRandomSampler.__iter__ (src/utils.py:15-23) restated: numpy.random.choice(range(N), size, p, replace) of the
    legacy RandomState, written out (numpy/random/mtrand.pyx: choice).  Consumes numpy's global generator like the reference.
    PINNED: tests/golden/sampler/*.npz were produced by the reference's own class (tests/golden/make_golden_sampler.py)."""
    import numpy as np
    import torch

    w = torch.as_tensor(weights, dtype=torch.double)
    p = w.numpy() / torch.sum(w).numpy()
    if replacement:
        cdf = p.cumsum()
        cdf /= cdf[-1]
        return cdf.searchsorted(np.random.random_sample(num_samples), side="right").astype(np.int64)
    p = p.copy()
    found = np.zeros(num_samples, dtype=np.int64)
    n_uniq = 0
    while n_uniq < num_samples:
        x = np.random.rand(num_samples - n_uniq)
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, unique_indices = np.unique(new, return_index=True)
        unique_indices.sort()
        new = new.take(unique_indices)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found
