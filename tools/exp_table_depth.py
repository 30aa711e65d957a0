"""Look-up sweep (bucket_apply_feat_kernel) against table depth: the batch is counted 1 / 2 / 4 / 8 times before it is featurized
(deeper counts spread over more bins: fewer tallies of a warp fold into one RED).  PG_LIB_PATH selects the library (A/B)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pangaea_b200 import _lib
from bench import make_synthetic_batch
ctx = _lib.Context(device=0)
s = make_synthetic_batch(ctx, 50_000_000, read_len=100, seed=2)
keep = np.ones(s["n_groups"], np.uint8); keep[0] = 0
for reps in (1, 2, 4, 8):
    for it in range(2):
        ctx.table_clear()
        b = ctx.adopt(s["reads"])
        for _ in range(reps):
            ctx.count(b)
        ctx.timing_reset()
        f = ctx.featurize(b, keep)
        ctx.synchronize()
        ms, tnf = ctx.timing(_lib.T_FEAT)[0], ctx.timing(_lib.T_TNF)[0]
        chk = int(f.torch(_lib.ABD_RAW).to(torch.int64).sum())
        f.free(); b.free()
    print("table depth x%d: feat_apply %.2f ms (TNF kernel beside it: %.2f ms), tallies %d" % (reps, ms, tnf, chk), flush=True)
