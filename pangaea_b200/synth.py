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


# ----------------------------------------------------------------------------------------------
# device-generated batches (csrc/synth.cuh) as FASTQ text: what bench.py and the tests hand to the
# oracle / the reference tools / the host parser
# ----------------------------------------------------------------------------------------------
def barcode_label(b: int, length=16) -> bytes:
    """barcode index -> ACGT string whose byte order equals the index order (LANG=C sort)."""
    return bytes(b"ACGT"[(b >> (2 * (length - 1 - i))) & 3] for i in range(length))


def write_batch_fastq(path, seq_host, read_len, bc_start, n_pairs, barcode_len=16, bc_base=0, pair_base=0, first_pair=0, append=False):
    """The first n_pairs pairs of a device-generated batch (read bytes + 1 separator per read, R1 then R2) as the
    interleaved, barcode-sorted FASTQ pangaea.py -i gets: headers ``@r<10-digit id>\tBX:Z:<barcode>-1``, quality ``I``.
    Built with numpy (fixed-width records), so tens of millions of pairs take seconds."""
    rl = read_len + 1
    # seq_host holds pairs [first_pair, first_pair + n_pairs) of the batch (bc_start indexes the batch's pairs); bc_base /
    # pair_base are the batch's global offsets (labels and read ids)
    bc_of_pair = (np.searchsorted(bc_start, np.arange(n_pairs) + first_pair, side="right") - 1).astype(np.int64) + bc_base
    bc_of_read = np.repeat(bc_of_pair, 2)
    ids = np.repeat(np.arange(n_pairs, dtype=np.int64) + first_pair + pair_base, 2)
    n = 2 * n_pairs
    hdr = np.zeros((n, 2 + 10 + 6 + barcode_len + 3), dtype=np.uint8)
    hdr[:, 0:2] = np.frombuffer(b"@r", dtype=np.uint8)
    for d in range(10):
        hdr[:, 2 + d] = ord("0") + (ids // 10 ** (9 - d)) % 10
    hdr[:, 12:18] = np.frombuffer(b"\tBX:Z:", dtype=np.uint8)
    for i in range(barcode_len):
        hdr[:, 18 + i] = _LETTERS[(bc_of_read >> (2 * (barcode_len - 1 - i))) & 3]
    hdr[:, 18 + barcode_len:] = np.frombuffer(b"-1\n", dtype=np.uint8)
    rec = np.empty((n, hdr.shape[1] + rl + 2 + rl), dtype=np.uint8)
    rec[:, :hdr.shape[1]] = hdr
    o = hdr.shape[1]
    rec[:, o:o + read_len] = np.asarray(seq_host[: n * rl]).reshape(n, rl)[:, :read_len]
    rec[:, o + read_len] = ord("\n")
    rec[:, o + rl:o + rl + 2] = np.frombuffer(b"+\n", dtype=np.uint8)
    rec[:, o + rl + 2:o + rl + 2 + read_len] = ord("I")
    rec[:, -1] = ord("\n")
    with open(path, "ab" if append else "wb") as f:
        rec.tofile(f)
    return path


def device_batch(ctx, n_pairs, read_len=100, n_barcodes=None, n_genomes=200, genome_len=3_000_000, frag_len=50_000, seed=2,
                 sub_rate=0.005, n_rate=0.0005, bc_base=0, pair_base=0):
    """SURVEY.md §8d model, generated in HBM.  Returns torch device tensors + a pg_reads over them.  Needs a GPU.
    bc_base / pair_base: global index of the batch's first barcode / pair (batches and ranks that share `seed` then
    draw different clouds from the same community)."""
    import torch

    from . import _lib

    dev = f"cuda:{ctx.params.device}"
    n_barcodes = n_barcodes or max(1, n_pairs // 100)
    rng = np.random.default_rng([seed, bc_base])
    if n_barcodes >= n_pairs:  # one cloud per pair (hybrid mode)
        counts = np.ones(n_barcodes, dtype=np.int64)
    else:
        counts = rng.poisson(n_pairs / n_barcodes, size=n_barcodes).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(counts)])
    start = np.minimum(start, n_pairs)
    start[-1] = n_pairs
    abundance = np.random.default_rng(seed).lognormal(0.0, 1.0, size=n_genomes)
    genome = rng.choice(n_genomes, size=n_barcodes, p=abundance / abundance.sum()).astype(np.int32)
    frag_len = int(min(frag_len, genome_len))
    insert = int(min(max(2 * read_len, 350), frag_len))
    n_reads = 2 * n_pairs
    n_bytes = n_reads * (read_len + 1)
    d_start = torch.from_numpy(start).to(dev)
    d_genome = torch.from_numpy(genome).to(dev)
    seq = torch.empty(n_bytes + 64, dtype=torch.uint8, device=dev)
    off = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
    flag = torch.empty(max(n_reads, 1), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx._ck(_lib.lib().pg_synth_generate2(ctx.h, n_pairs, read_len, n_barcodes, d_start.data_ptr(), d_genome.data_ptr(), genome_len,
                                          frag_len, insert, sub_rate, n_rate, seed, bc_base, pair_base, seq.data_ptr(), off.data_ptr(),
                                          flag.data_ptr()))
    nonempty = int((np.diff(start) > 0).sum())
    reads = _lib.make_reads(seq, off, flag, n_reads=n_reads, n_bytes=n_bytes)
    return {"seq": seq[:n_bytes], "_seq_full": seq, "off": off, "flag": flag[:n_reads], "reads": reads, "n_groups": nonempty + 1,
            "n_pairs": n_pairs, "read_len": read_len, "bc_start": start, "n_bytes": n_bytes, "n_reads": n_reads, "bc_base": bc_base,
            "pair_base": pair_base, "n_barcodes": n_barcodes}
