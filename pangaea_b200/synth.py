"""Synthetic linked-read generator - host/numpy version for tests
and for the CPU-baseline sample.  bench.py's full-size inputs come from the device
generator in csrc/synth.cuh, which draws from the same model.

Model: a community of ``n_genomes`` uniform-random ACGT genomes with log-normal(0,1)
abundances; every barcode (read cloud) picks one genome (prob. ~ abundance) and one
``frag_len`` fragment of it; pairs are placed uniformly in the fragment, R2 is the
reverse complement of the far end; 0.5 % substitutions, 0.05 % N; upper case only;
headers ``@r<id>\\tBX:Z:<barcode>-1`` (what run_pangaea's preprocessing leaves,
/root/reference/src/run_pangaea:143,156,237-252); records already sorted by barcode
in C-locale byte order; quality line constant ``I``.
"""
from __future__ import annotations

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b
_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def _barcodes(rng, n, length):
    """n distinct ACGT barcodes, sorted bytewise (LANG=C sort order)."""
    seen = set()
    while len(seen) < n:
        for row in _LETTERS[rng.integers(0, 4, size=(n - len(seen), length))]:
            seen.add(row.tobytes())
    return sorted(seen)


def generate(n_barcodes=50, mean_pairs=20, read_len=100, n_genomes=5, genome_len=200_000,
             frag_len=50_000, sub_rate=0.005, n_rate=0.0005, barcode_len=16, seed=0,
             unbarcoded_pairs=0, lower_rate=0.0):
    """-> dict(seq1, seq2 : uint8[P, L]; barcode : list[bytes] per pair ('' = none); ...)"""
    rng = np.random.default_rng(seed)
    genomes = _LETTERS[rng.integers(0, 4, size=(n_genomes, genome_len), dtype=np.uint8)]
    abundance = rng.lognormal(0.0, 1.0, size=n_genomes)
    abundance /= abundance.sum()
    bcs = _barcodes(rng, n_barcodes, barcode_len)
    counts = np.maximum(rng.poisson(mean_pairs, size=n_barcodes), 1)
    P = int(counts.sum()) + unbarcoded_pairs
    frag_len = min(frag_len, genome_len)
    insert = min(max(2 * read_len, 350), frag_len)

    bc_of_pair = np.repeat(np.arange(n_barcodes), counts)
    g_of_bc = rng.choice(n_genomes, size=n_barcodes, p=abundance)
    f_of_bc = rng.integers(0, genome_len - frag_len + 1, size=n_barcodes)
    g = np.concatenate([g_of_bc[bc_of_pair], rng.integers(0, n_genomes, size=unbarcoded_pairs)])
    f0 = np.concatenate([f_of_bc[bc_of_pair], rng.integers(0, genome_len - frag_len + 1, size=unbarcoded_pairs)])
    start = f0 + rng.integers(0, frag_len - insert + 1, size=P)
    idx = np.arange(read_len)
    seq1 = genomes[g[:, None], start[:, None] + idx[None, :]]
    far = start + insert - 1
    seq2 = _COMP[genomes[g[:, None], far[:, None] - idx[None, :]]]
    for s in (seq1, seq2):
        sub = rng.random(s.shape) < sub_rate
        s[sub] = _LETTERS[rng.integers(0, 4, size=int(sub.sum()))]
        s[rng.random(s.shape) < n_rate] = ord("N")
        if lower_rate:
            low = rng.random(s.shape) < lower_rate
            s[low] |= 0x20
    barcode = [bcs[i] for i in bc_of_pair] + [b""] * unbarcoded_pairs
    return {"seq1": seq1, "seq2": seq2, "barcode": barcode, "read_len": read_len, "n_pairs": P}


def _header(i, bc, mate, style):
    if style == "10x":  # after run_pangaea's preprocessing
        return b"@r%d\tBX:Z:%s-1" % (i, bc) if bc else b"@r%d" % i
    if style == "stlfr":  # raw stLFR names, barcode a_b_c, 0_0_0 = none (count_kmer.cpp:36-43)
        return b"@r%d#%s/%d" % (i, bc if bc else b"0_0_0", mate)
    raise ValueError(style)


def write_interleaved(path, data, style="10x", qual=b"I"):
    """8 lines per pair, plain text - the file pangaea.py -i receives."""
    L = data["read_len"]
    q = qual * L
    with open(path, "wb") as f:
        for i in range(data["n_pairs"]):
            bc = data["barcode"][i]
            f.write(_header(i, bc, 1, style) + b"\n" + data["seq1"][i].tobytes() + b"\n+\n" + q + b"\n")
            f.write(_header(i, bc, 2, style) + b"\n" + data["seq2"][i].tobytes() + b"\n+\n" + q + b"\n")
    return path


def write_paired(path1, path2, data, style="10x", qual=b"I"):
    L = data["read_len"]
    q = qual * L
    with open(path1, "wb") as f1, open(path2, "wb") as f2:
        for i in range(data["n_pairs"]):
            bc = data["barcode"][i]
            f1.write(_header(i, bc, 1, style) + b"\n" + data["seq1"][i].tobytes() + b"\n+\n" + q + b"\n")
            f2.write(_header(i, bc, 2 if style == "stlfr" else 1, style) + b"\n" + data["seq2"][i].tobytes() + b"\n+\n" + q + b"\n")
    return path1, path2
