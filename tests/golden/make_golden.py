"""Regenerates tests/golden/* by running the UNMODIFIED reference tools
(oracle/_ref/count_kmer, oracle/_ref/count_tnf, built by oracle/Makefile from
/root/reference/src/cpptools) on small, fully specified inputs.

    python tests/golden/make_golden.py

Every case directory holds the input FASTQ(s), the k-mer dump handed to
`count_kmer -g` (written by the oracle's jellyfish stand-in, or hand-made), the
parameters, and the two tools' decompressed CSV outputs.  Nothing here is read from
/root/reference at test time.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pangaea_b200 import synth  # noqa: E402

GOLD = os.path.dirname(os.path.abspath(__file__))
# `--jellyfish`: take the k-mer dumps from a real jellyfish (absent from this image and from the reference tree, so the
# committed dumps come from the oracle's stand-in); the abundance goldens are then regenerated from those dumps
USE_JELLYFISH = "--jellyfish" in sys.argv and shutil.which("jellyfish") is not None
if "--jellyfish" in sys.argv and not USE_JELLYFISH:
    print("jellyfish is not installed here: dumps come from the oracle's stand-in (parity of step 1a stays unpinned)")


def rand_seq(rng, n):
    return bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))


def run_case(name, files, params, dump_lines=None):
    """files: {'i': path} or {'1': path, '2': path} already written inside the case dir."""
    d = os.path.join(GOLD, name)
    k, tk, mlen, vs, ws = params["k"], params["tnf_k"], params["min_length"], params["vector_size"], params["window_size"]
    dump = os.path.join(d, "kmers.dump")
    if dump_lines is None and USE_JELLYFISH:
        # the real thing (src/feature.py:76-94,103): wherever a jellyfish binary exists this pins step 1a, which is otherwise
        # anchored only on the oracle's restatement of jellyfish's documented behaviour (DESIGN.md §2 "parity unpinned")
        jf = os.path.join(d, "kmers.jf")
        mq = ["--min-qual-char=" + chr(params["min_qual"])] if params.get("min_qual") else []
        subprocess.run(["jellyfish", "count", "-t", "2", "-C", "-m", str(k), "-s", "100M", *mq, "-o", jf, *[files[x] for x in sorted(files)]], check=True)
        with open(dump, "w") as f:
            subprocess.run(["jellyfish", "dump", "-c", "-t", jf], check=True, stdout=f)
        os.remove(jf)
        lines = sorted(open(dump).read().splitlines())
        open(dump, "w").write("\n".join(lines) + ("\n" if lines else ""))
    elif dump_lines is None:
        t = O.count_fastq([files[x] for x in sorted(files)], k, params.get("min_qual", 0))
        t.write_dump(dump, k)
        # keep the fixture deterministic: sort the dump
        lines = sorted(open(dump).read().splitlines())
        open(dump, "w").write("\n".join(lines) + ("\n" if lines else ""))
    else:
        open(dump, "w").write("".join(dump_lines))
    io = ["-i", files["i"]] if "i" in files else ["-1", files["1"], "-2", files["2"]]
    for tool, out, extra in (
        (O.REF_COUNT_KMER, "abundance.csv", ["-g", dump, "-k", str(k), "-w", str(ws), "-v", str(vs)]),
        (O.REF_COUNT_TNF, "tnf.csv", ["-k", str(tk)]),
    ):
        gz = os.path.join(d, out + ".gz")
        subprocess.run([tool, *io, "-l", str(mlen), "-t", "3", "-o", gz, *extra], check=True, stdout=subprocess.DEVNULL)
        open(os.path.join(d, out), "wb").write(gzip.open(gz, "rb").read())
        os.remove(gz)
    params = dict(params, files={x: os.path.basename(p) for x, p in files.items()})
    json.dump(params, open(os.path.join(d, "params.json"), "w"), indent=1, sort_keys=True)
    print(name, "abundance rows:", sum(1 for _ in open(os.path.join(d, "abundance.csv"))),
          "tnf rows:", sum(1 for _ in open(os.path.join(d, "tnf.csv"))))


def fresh(name):
    d = os.path.join(GOLD, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


P = dict(k=15, tnf_k=4, min_length=0, vector_size=400, window_size=10)


def main():
    assert O.have_ref(), "run `make -C oracle ref` first (needs /root/reference)"
    rng = np.random.default_rng(20261018)

    # KAT-1: interleaved, 10x headers, 30 bp, AAAA x3, CCCC x2, GGGG x3
    d = fresh("kat1_interleaved_10x")
    with open(os.path.join(d, "reads.fq"), "wb") as f:
        i = 0
        for bc, n in ((b"AAAA", 3), (b"CCCC", 2), (b"GGGG", 3)):
            for _ in range(n):
                for mate in (1, 2):
                    f.write(b"@r%d BX:Z:%s-1\n%s\n+\n%s\n" % (i, bc, rand_seq(rng, 30), b"I" * 30))
                i += 1
    run_case("kat1_interleaved_10x", {"i": os.path.join(d, "reads.fq")}, P)

    # KAT-2: paired, stLFR headers, 40 bp, 0_0_0 = no barcode, one mismatched pair
    d = fresh("kat2_paired_stlfr")
    with open(os.path.join(d, "r1.fq"), "wb") as f1, open(os.path.join(d, "r2.fq"), "wb") as f2:
        i = 0
        for bc, n in ((b"1_1_1", 3), (b"2_2_2", 2), (b"0_0_0", 2), (b"3_3_3", 3)):
            for j in range(n):
                name2 = b"x%d" % i if (bc == b"3_3_3" and j == 1) else b"r%d" % i
                f1.write(b"@r%d#%s/1\n%s\n+\n%s\n" % (i, bc, rand_seq(rng, 40), b"I" * 40))
                f2.write(b"@%s#%s/2\n%s\n+\n%s\n" % (name2, bc, rand_seq(rng, 40), b"I" * 40))
                i += 1
    run_case("kat2_paired_stlfr", {"1": os.path.join(d, "r1.fq"), "2": os.path.join(d, "r2.fq")}, P)

    # synthetic community, production parameters (-l 2000), N / lower case / unbarcoded tail
    d = fresh("synth_10x_l2000")
    data = synth.generate(n_barcodes=24, mean_pairs=14, read_len=100, n_genomes=3, genome_len=30_000,
                          frag_len=8_000, seed=11, unbarcoded_pairs=5, lower_rate=0.002, n_rate=0.002)
    synth.write_interleaved(os.path.join(d, "reads.fq"), data)
    run_case("synth_10x_l2000", {"i": os.path.join(d, "reads.fq")}, dict(P, min_length=2000))

    # same model, paired files with 10x headers and jellyfish's --min-qual-char=? branch (feature.py:83)
    d = fresh("synth_paired_minqual")
    data = synth.generate(n_barcodes=10, mean_pairs=12, read_len=75, n_genomes=2, genome_len=20_000,
                          frag_len=5_000, seed=12)
    L = data["read_len"]
    with open(os.path.join(d, "r1.fq"), "wb") as f1, open(os.path.join(d, "r2.fq"), "wb") as f2:
        for i in range(data["n_pairs"]):
            q1 = bytes(rng.choice(np.frombuffer(b"I5?>", dtype=np.uint8), size=L, p=[0.9, 0.04, 0.03, 0.03]))
            q2 = bytes(rng.choice(np.frombuffer(b"I5?>", dtype=np.uint8), size=L, p=[0.9, 0.04, 0.03, 0.03]))
            h = b"@r%d\tBX:Z:%s-1\n" % (i, data["barcode"][i])
            f1.write(h + data["seq1"][i].tobytes() + b"\n+\n" + q1 + b"\n")
            f2.write(h + data["seq2"][i].tobytes() + b"\n+\n" + q2 + b"\n")
    run_case("synth_paired_minqual", {"1": os.path.join(d, "r1.fq"), "2": os.path.join(d, "r2.fq")},
             dict(P, min_length=1000, min_qual=ord("?")))

    # ragged / hostile text: reads shorter than k, empty sequence lines, CRLF, no '-' after the
    # barcode, reads without BX tag in the middle, truncated last record, non-ACGT bytes
    d = fresh("edge_ragged")
    recs = []
    lens = [0, 3, 4, 14, 15, 16, 31, 32, 33, 63, 64, 65, 100, 151, 7, 250]
    bcs = [b"AAC", b"AAC", b"AAC", b"ACG", b"ACG", b"", b"", b"CCA", b"CCA", b"CCA", b"CCA", b"GT", b"GT", b"TTT", b"TTT", b"TTT"]
    for i, (ln, bc) in enumerate(zip(lens, bcs)):
        for mate in (1, 2):
            s = bytearray(rand_seq(rng, ln + mate))
            if ln > 20 and i % 3 == 0:
                s[ln // 2] = ord("N")
            if ln > 40 and i % 4 == 1:
                s[5] = ord("a"); s[ln - 3] = ord("."); s[ln - 20] = ord("R")
            tag = (b"\tBX:Z:" + bc + (b"-1" if i % 5 else b"")) if bc else b""
            eol = b"\r\n" if i == 9 else b"\n"
            recs.append(b"@q%d" % i + tag + eol + bytes(s) + eol + b"+" + eol + b"I" * len(s) + eol)
    recs.append(b"@q99\tBX:Z:TTT-1\nACGTACGTACGTACGTACGTAC\n")  # truncated: R1 sequence only
    open(os.path.join(d, "reads.fq"), "wb").write(b"".join(recs))
    run_case("edge_ragged", {"i": os.path.join(d, "reads.fq")}, dict(P, k=5, tnf_k=3, min_length=10, vector_size=7, window_size=2))

    # KAT-4: hand-made dump - counts 0, 9, 10, 3999, 4000, absent keys, a duplicate key (last wins),
    # a reverse-complement spelling (re-canonicalised by the loader)
    d = fresh("kat4_bins")
    seqs = [rand_seq(rng, 60) for _ in range(8)]
    with open(os.path.join(d, "reads.fq"), "wb") as f:
        for i in range(4):
            for bc in (b"AAAA",) if i < 2 else (b"CCCC",):
                f.write(b"@p%d BX:Z:%s-1\n%s\n+\n%s\n@p%d BX:Z:%s-1\n%s\n+\n%s\n"
                        % (i, bc, seqs[2 * i], b"I" * 60, i, bc, seqs[2 * i + 1], b"I" * 60))
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    dump, vals = [], [0, 9, 10, 3999, 4000, 123456789, 39, 40]
    for i, s in enumerate(seqs):
        for j in range(0, 60 - 15 + 1, 3):  # every third 15-mer present
            km = s[j:j + 15]
            if (i + j) % 2:
                km = km.translate(comp)[::-1]
            dump.append(b"%s\t%d\n" % (km, vals[(i + j) % len(vals)]))
    dump.append(dump[0].split(b"\t")[0] + b"\t77\n")
    run_case("kat4_bins", {"i": os.path.join(d, "reads.fq")}, P, dump_lines=[x.decode() for x in dump])


if __name__ == "__main__":
    main()
