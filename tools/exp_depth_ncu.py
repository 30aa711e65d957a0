"""one featurize at table depth x1, then one at depth x8 (for ncu: apply_feat launches 0-4 and 5-9)"""
import sys, numpy as np
sys.path.insert(0, ".")
from pangaea_b200 import _lib
from bench import make_synthetic_batch
ctx = _lib.Context(device=0)
s = make_synthetic_batch(ctx, 50_000_000, 100, seed=2)
keep = np.ones(s["n_groups"], np.uint8); keep[0] = 0
for reps in (1, 8):
    ctx.table_clear()
    b = ctx.adopt(s["reads"])
    for _ in range(reps):
        ctx.count(b)
    f = ctx.featurize(b, keep)
    ctx.synchronize()
    f.free(); b.free()
print("done")
