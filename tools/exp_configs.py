"""Throughput of the other BASELINE.json shapes on one GPU (device-resident input, same step as bench.py):
  C3 TELL-Seq 2x150 bp, 100 pairs per barcode;  C4 hybrid: 2x150 bp, one (virtual) barcode per pair, min_length 0.
    python tools/exp_configs.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pangaea_b200 import _lib
from bench import make_synthetic_batch

def run(label, n_pairs, read_len, n_barcodes, min_length, steps=3):
    ctx = _lib.Context(device=0, min_length=min_length)
    s = make_synthetic_batch(ctx, n_pairs, read_len, n_barcodes=n_barcodes, seed=3)
    keep = np.ones(s["n_groups"], np.uint8); keep[0] = 0
    for it in range(2 + steps):
        if it == 2:
            ctx.synchronize(); ctx.timing_reset(); t0 = time.perf_counter()
        ctx.table_clear()
        b = ctx.adopt(s["reads"])
        ctx.count(b)
        f = ctx.featurize(b, keep)
        f.normalize()
        rows = f.rows
        f.free(); b.free()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / steps
    st = {n: round(ctx.timing(w)[0] / steps, 2) for n, w in (("pack", 0), ("count_scatter", 6), ("count_split", 9), ("count_apply", 1), ("group", 2), ("tnf", 8),
                                                              ("feat_scatter", 7), ("feat_apply", 3), ("normalize", 4))}
    print(f"{label}: {2 * n_pairs / dt / 1e6:.1f} M reads/s, {dt * 1e3:.1f} ms/step, rows {rows}, stages {st}", flush=True)
    ctx.close()

run("C3 TELL-Seq 2x150, 30M pairs, 300k barcodes", 30_000_000, 150, 300_000, 2000)
run("C4 hybrid 2x150, 4M pairs, one barcode per pair, -l 0", 4_000_000, 150, 4_000_000, 0)
