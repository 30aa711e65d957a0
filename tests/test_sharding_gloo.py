"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path - shard plan at
cloud boundaries, one all-reduce of the count tables, rows gathered in rank order - with the
ORACLE doing the arithmetic the GPUs do in production.  Checked against the unsharded oracle
run over the same file (which itself is pinned to the reference binaries)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import contract_features
from pangaea_b200 import _lib, synth
from pangaea_b200.shard import plan_shards, slice_shard

K, TK, MLEN, VS, WS = 9, 4, 1500, 40, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dense_counts(oracle, seq, off, k):
    """per-shard dense count vector indexed by canonical key (what pg_count leaves in HBM)"""
    t = oracle.Table()
    for r in range(len(off) - 1):
        t.count_read(bytes(seq[off[r]:off[r + 1]]), k)
    dense = torch.zeros(4 ** k, dtype=torch.int64)
    keys, vals = t.items()
    dense[torch.from_numpy(keys.astype(np.int64))] = torch.from_numpy(vals.astype(np.int64))
    return dense


class _DenseTable:
    def __init__(self, dense):
        self.dense = dense

    def get(self, key):
        c = int(self.dense[key])
        return c if c else None


def _worker(rank, world, port, path, out_dir):
    from oracle import oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fq = _lib.Fastq(path)
    seq, off, flag, keep = fq.arrays()
    labels = [fq.label(g) for g in range(fq.n_groups)]
    shard = plan_shards(off, flag, world)[rank]
    s, soff, sflag, skeep, _ = slice_shard(shard, seq, off, flag, keep)
    table = _dense_counts(O, s, soff, K)
    dist.all_reduce(table)                                   # the one exchange step
    n_local = 1 + int((sflag & 1).sum())
    local_labels = (labels[shard.group_lo:shard.group_hi] + [""] * n_local)[:n_local]
    names, abd, tnf = contract_features(s, soff, sflag, skeep, local_labels, _DenseTable(table), K, TK, MLEN, VS, WS)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((names, abd, tnf), gathered, dst=0)
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), names=np.array(sum((g[0] for g in gathered), []), dtype=object),
                 abd=np.concatenate([g[1] for g in gathered]), tnf=np.concatenate([g[2] for g in gathered]), allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(tmp_path, oracle, world):
    data = synth.generate(n_barcodes=40, mean_pairs=12, read_len=90, n_genomes=2, genome_len=30_000, frag_len=5_000, seed=31,
                          unbarcoded_pairs=6, n_rate=0.003)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    want_names, want_abd, want_tnf = oracle.featurize(path, None, k=K, tnf_k=TK, mlen=MLEN, vs=VS, ws=WS)
    mp.spawn(_worker, args=(world, _free_port(), path, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "merged.npz", allow_pickle=True)
    assert list(got["names"]) == list(want_names)
    assert np.array_equal(got["abd"], want_abd) and np.array_equal(got["tnf"], want_tnf)


def test_plan_cuts_only_at_cloud_flushes():
    rng = np.random.default_rng(0)
    for n_ranks in (1, 2, 4, 8, 13):
        lens = rng.integers(1, 200, size=500)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        flag = np.zeros(500, dtype=np.uint8)
        flag[1::2] = rng.random(250) < 0.2          # change flags sit on R2 reads only
        shards = plan_shards(off, flag, n_ranks)
        assert len(shards) == n_ranks
        assert shards[0].read_lo == 0 and shards[-1].read_hi == 500
        assert shards[0].group_lo == 0 and shards[-1].group_hi == 1 + int(flag.sum())
        for a, b in zip(shards, shards[1:]):
            assert a.read_hi == b.read_lo and a.group_hi == b.group_lo
            if a.read_hi not in (0, 500) and a.n_reads:
                assert flag[a.read_hi - 1] & 1, "a shard must end right after a flush"
            assert a.group_hi - a.group_lo == int(flag[a.read_lo:a.read_hi].sum())
    # degenerate: no flags at all -> everything on the last rank
    shards = plan_shards(np.array([0, 5, 9], np.int64), np.zeros(2, np.uint8), 4)
    assert [s.n_reads for s in shards] == [0, 0, 0, 2] and shards[-1].n_groups == 1
    assert [s.n_reads for s in plan_shards(np.zeros(1, np.int64), np.zeros(0, np.uint8), 2)] == [0, 0]


@pytest.mark.parametrize("flags", [[0, 1, 0, 0, 0, 1, 0, 1], [0, 1, 0, 1], [0, 0, 0, 1], [0, 1, 0, 0], [0, 0]])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_slice_keep_length_matches_local_clouds(flags, world):
    """pg_featurize insists on n_groups == 1 + local change flags; the stream may END on a flagged read (the file's last
    barcode has exactly one pair), in which case the trailing cloud is empty and already part of group_keep."""
    flag = np.array(flags, dtype=np.uint8)
    n = len(flag)
    off = np.arange(n + 1, dtype=np.int64) * 11
    seq = np.zeros(int(off[-1]), dtype=np.uint8)
    n_groups = 1 + int(flag.sum())
    keep = (np.arange(n_groups) % 2).astype(np.uint8)       # recognisable pattern
    seen = []
    for shard in plan_shards(off, flag, world):
        s, soff, sflag, skeep, _ = slice_shard(shard, seq, off, flag, keep)
        assert len(skeep) == 1 + int((sflag & 1).sum()), (shard, skeep)
        assert len(soff) == len(sflag) + 1 and (len(s) == soff[-1] if len(sflag) else len(s) == 0)
        own = shard.group_hi - shard.group_lo
        assert list(skeep[:own]) == list(keep[shard.group_lo:shard.group_hi][:len(skeep)])
        assert not skeep[own:].any(), "the cloud a flush opens locally is empty and dropped"
        seen.append(own)
    assert sum(seen) == n_groups
