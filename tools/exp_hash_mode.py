"""Throughput of the open-addressing hash-table path (k > 16) on one GPU: 10 M pairs 2x100 bp, k = 21."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pangaea_b200 import _lib
from bench import make_synthetic_batch
for k, cap in ((21, 1 << 31), (31, 1 << 31)):
    ctx = _lib.Context(device=0, k=k, table_capacity=cap)
    s = make_synthetic_batch(ctx, 10_000_000, read_len=100, seed=2)
    keep = np.ones(s["n_groups"], np.uint8); keep[0] = 0
    for it in range(3):
        if it == 1:
            ctx.synchronize(); ctx.timing_reset(); t0 = time.perf_counter()
        ctx.table_clear()
        b = ctx.adopt(s["reads"])
        ctx.count(b)
        f = ctx.featurize(b, keep)
        f.normalize()
        f.free(); b.free()
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / 2
    st = {n: round(ctx.timing(w)[0] / 2, 2) for n, w in (("pack", 0), ("count", 1), ("group", 2), ("featurize", 3), ("normalize", 4))}
    print(f"hash mode k={k}: {2 * 10_000_000 / dt / 1e6:.1f} M reads/s, {dt * 1e3:.1f} ms/step, distinct {ctx.table_size()}, stages {st}", flush=True)
    ctx.close()
