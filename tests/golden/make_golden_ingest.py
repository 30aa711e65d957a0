"""Golden vectors for the SURVEY §8f text tools, produced by the UNMODIFIED reference binaries (oracle/_ref, built by
oracle/Makefile from /root/reference/src/cpptools): preprocess_stlfr -n [-l], preprocess_tellseq, extract_reads -i.
    python tests/golden/make_golden_ingest.py
Every case directory holds the inputs and the files the tool wrote; nothing is read from /root/reference at test time."""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLD = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "oracle", "_ref")


def seq(rng, n):
    return bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n, p=[0.24, 0.24, 0.24, 0.24, 0.04]))


def fresh(name):
    d = os.path.join(GOLD, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    return d


def stlfr():
    rng = np.random.default_rng(1)
    d = fresh("ingest_stlfr")
    bcs = ["12_345_6", "0_0_0", "7_0_9", "0_5_5", "1_2_0", "1000_2000_1536", "3_3_3"]
    with open(os.path.join(d, "r1.fq"), "wb") as f1, open(os.path.join(d, "r2.fq"), "wb") as f2:
        for i in range(60):
            bc = bcs[i % len(bcs)].encode()
            l1, l2 = int(rng.integers(20, 60)), int(rng.integers(20, 60))
            extra = b" extra/field" if i % 11 == 0 else b""
            f1.write(b"@V300%d#%s/1%s\n%s\n+\n%s\n" % (i, bc, extra, seq(rng, l1), b"F" * l1))
            f2.write(b"@V300%d#%s/2\n%s\n+r2\n%s\n" % (i, bc, seq(rng, l2), b"G" * l2))
        f2.write(b"@extra#1_1_1/2\nACGT\n+\nIIII\n")  # file 2 longer than file 1: ignored
    for tag, flags in (("n", ["-n"]), ("nl", ["-n", "-l"])):
        subprocess.run([os.path.join(REF, "preprocess_stlfr"), "-1", os.path.join(d, "r1.fq"), "-2", os.path.join(d, "r2.fq"), "-o",
                        os.path.join(d, "out_" + tag), *flags], check=True, stdout=subprocess.DEVNULL)
    # file 1 longer than file 2 (missing lines of file 2 are empty)
    lines = open(os.path.join(d, "r2.fq"), "rb").read().split(b"\n")
    open(os.path.join(d, "r2_short.fq"), "wb").write(b"\n".join(lines[:50]) + b"\n")
    subprocess.run([os.path.join(REF, "preprocess_stlfr"), "-1", os.path.join(d, "r1.fq"), "-2", os.path.join(d, "r2_short.fq"), "-o",
                    os.path.join(d, "out_short"), "-n", "-l"], check=True, stdout=subprocess.DEVNULL)
    print("stlfr", sorted(os.listdir(d)))


def tellseq():
    rng = np.random.default_rng(2)
    d = fresh("ingest_tellseq")
    with open(os.path.join(d, "r1.fq"), "wb") as f1, open(os.path.join(d, "r2.fq"), "wb") as f2, open(os.path.join(d, "i1.fq"), "wb") as fi:
        for i in range(50):
            l1, l2 = int(rng.integers(20, 70)), int(rng.integers(20, 70))
            bl = 18 if i % 7 else int(rng.integers(10, 24))  # some index reads have the wrong length: record dropped
            name = b"@A00:%d:X 1:N:0:ACGT" % i if i % 5 else b"@A00:%d:X" % i
            f1.write(b"%s\n%s\n+comment\n%s\n" % (name, seq(rng, l1), b"F" * l1))
            f2.write(b"%s\n%s\n+\n%s\n" % (name.replace(b" 1:", b" 2:"), seq(rng, l2), b"G" * l2))
            fi.write(b"%s\n%s\n+\n%s\n" % (name, seq(rng, bl).replace(b"N", b"A"), b"I" * bl))
    subprocess.run([os.path.join(REF, "preprocess_tellseq"), "-1", os.path.join(d, "r1.fq"), "-2", os.path.join(d, "r2.fq"), "-l",
                    os.path.join(d, "i1.fq"), "-o", os.path.join(d, "out")], check=True, stdout=subprocess.DEVNULL)
    print("tellseq", sorted(os.listdir(d)))


def extract():
    from pangaea_b200 import synth

    d = fresh("ingest_extract")
    data = synth.generate(n_barcodes=30, mean_pairs=4, read_len=40, n_genomes=2, genome_len=4000, frag_len=800, seed=6, unbarcoded_pairs=6, barcode_len=8)
    synth.write_interleaved(os.path.join(d, "tenx.fq"), data)
    synth.write_interleaved(os.path.join(d, "stlfr.fq"), data, style="stlfr")
    bcs = sorted({b for b in data["barcode"] if b})
    for style in ("tenx", "stlfr"):
        names = bcs if style == "tenx" else bcs
        with open(os.path.join(d, f"clusters_{style}.tsv"), "wb") as f:
            f.write(b"0\t" + b",".join(names[0:7]) + b"\n")
            f.write(b"-1\t" + b",".join(names[7:12]) + b"\n")          # noise cluster: skipped
            f.write(b"7\t" + b",".join(names[12:20] + [names[3]]) + b"\n")  # names[3] moves to this cluster (last line wins)
            f.write(b"2\t" + names[20] + b"\n")
            f.write(b"9\tNOTTHERE\n")                                   # an empty cluster still gets its files
        od = os.path.join(d, "out_" + style)
        os.makedirs(od)
        subprocess.run([os.path.join(REF, "extract_reads"), "-i", os.path.join(d, style + ".fq"), "-c", os.path.join(d, f"clusters_{style}.tsv"),
                        "-o", os.path.join(od, "x")], check=True, stdout=subprocess.DEVNULL)
        print("extract", style, sorted(os.listdir(od)))


if __name__ == "__main__":
    stlfr()
    tellseq()
    extract()
