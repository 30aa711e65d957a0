"""Shared test helpers (CPU side).  Anything here that computes features does so with the
ORACLE's primitives and exists only to check the product; it is never shipped."""
import numpy as np

from oracle import oracle as O


def load_dump_arrays(path, k):
    """jellyfish-style dump text -> (keys uint64 forward-encoded as written, counts uint32), file order."""
    keys, counts = [], []
    for line in open(path):
        km, c = line.rstrip("\n").split("\t")
        if len(km) != k or set(km) - set("ACGT"):
            continue
        keys.append(O.encode(km))
        counts.append(int(c))
    return np.array(keys, dtype=np.uint64), np.array(counts, dtype=np.uint32)


def contract_features(seq, off, flag, keep, labels, table, k=15, tnf_k=4, mlen=2000, vs=400, ws=10):
    """What include/pangaea_b200.h says a batch means, evaluated with oracle primitives:
    cloud(r) = number of CHANGE flags before r; NOFEAT reads belong to no cloud; a cloud is
    emitted iff keep[g] and sum(len+1) > mlen; per cloud the windows of each read separately
    (the separator byte breaks windows exactly like the reference's 'N')."""
    seq = bytes(seq)
    lut, td = O.tnf_lut(tnf_k)
    n_groups = 1 + int((np.asarray(flag) & 1).sum())
    assert n_groups == len(keep) == len(labels)
    reads_of = [[] for _ in range(n_groups)]
    g = 0
    for r in range(len(flag)):
        if not flag[r] & 2:
            reads_of[g].append(seq[off[r]:off[r + 1]])
        if flag[r] & 1:
            g += 1
    names, abd, tnf = [], [], []
    for g in range(n_groups):
        total = sum(len(x) for x in reads_of[g])
        if not keep[g] or mlen < 0 or total <= mlen:
            continue
        a = np.zeros(vs, dtype=np.int64)
        t = np.zeros(td, dtype=np.int64)
        for rd in reads_of[g]:
            for kk, is_abd in ((k, True), (tnf_k, False)):
                val, run, mask = 0, 0, (1 << (2 * kk)) - 1
                for ch in rd:
                    if ch not in b"ACGT":
                        val, run = 0, 0
                        continue
                    val = ((val << 2) & mask) | ((ch >> 1) & 3)
                    run += 1
                    if run >= kk:
                        key = O.canonical(val, kk)
                        if is_abd:
                            c = table.get(key)
                            if c is not None and c // ws < vs:
                                a[c // ws] += 1
                        else:
                            t[lut[key]] += 1
        names.append(labels[g])
        abd.append(a)
        tnf.append(t)
    return names, np.array(abd).reshape(-1, vs), np.array(tnf).reshape(-1, td)
