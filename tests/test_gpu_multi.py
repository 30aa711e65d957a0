"""Multi-rank parity: pangaea_b200.distributed on two or three ranks == the oracle over the whole file.
Two flavours: NCCL on two GPUs (skipped on a one-GPU box) and - so that every box proves the multi-rank data path -
the same code with the ranks sharing cuda:0 and gloo carrying the table all-reduce."""
import os
import socket

import numpy as np
import pytest

from pangaea_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, path, out_dir, backend, mode, k, batch_bytes):
    import torch
    import torch.distributed as dist

    from pangaea_b200.distributed import (extract_features_from_file, extract_features_owner_partitioned, extract_features_sharded, gather_rows,
                                          gather_rows_named)
    from pangaea_b200.shard import plan_shards, slice_shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{dev}"))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _lib.Context(device=dev, k=k, table_capacity=1 << 22 if k > 16 else 0)
    if mode == "owner":  # the owner-partitioned table: keys to their owners by all-to-all, counts back the same way
        fq = _lib.Fastq(path)
        seq, off, flag, keep = fq.arrays()
        shard = plan_shards(off, flag, world)[rank]
        s, soff, sflag, skeep, _ = slice_shard(shard, seq, off, flag, keep)
        reads = _lib.make_reads(np.ascontiguousarray(s), soff, np.ascontiguousarray(sflag))
        feats = extract_features_owner_partitioned(ctx, reads, skeep, seg_words=batch_bytes or (1 << 21))
        assert ctx.table_size() > 0
        merged = gather_rows(feats, shard, fq.label)
    elif mode == "arrays":  # every rank holds the parsed stream and slices its shard
        fq = _lib.Fastq(path)
        seq, off, flag, keep = fq.arrays()
        feats, shard = extract_features_sharded(ctx, seq, off, flag, keep)
        merged = gather_rows(feats, shard, fq.label)
    else:                 # every rank parses only its byte range of the file
        names, feats = extract_features_from_file(ctx, path, batch_seq_bytes=batch_bytes)
        merged = gather_rows_named(names, feats)
    if rank == 0:
        names, abd, tnf = merged
        np.savez(os.path.join(out_dir, "merged.npz"), names=names, abd=abd, tnf=tnf, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, oracle, world, backend, mode, k=15, batch_bytes=None, seed=17):
    import torch.multiprocessing as mp

    data = synth.generate(n_barcodes=300, mean_pairs=15, read_len=100, n_genomes=3, genome_len=80_000, frag_len=10_000, seed=seed,
                          unbarcoded_pairs=25, n_rate=0.002)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    want_names, want_abd, want_tnf = oracle.featurize(path, None, k=k)
    mp.spawn(_worker, args=(world, _free_port(), path, str(tmp_path), backend, mode, k, batch_bytes), nprocs=world, join=True)
    got = np.load(tmp_path / "merged.npz", allow_pickle=True)
    assert list(got["names"]) == list(want_names)
    assert np.array_equal(got["abd"], want_abd) and np.array_equal(got["tnf"], want_tnf)


@pytest.mark.parametrize("mode,k", [("arrays", 15), ("file", 15), ("owner", 21), ("owner", 15)])
def test_two_ranks_equal_one(tmp_path, oracle, mode, k):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(tmp_path, oracle, 2, "nccl", mode, k=k, batch_bytes={"file": 200_000, "owner": 4096}.get(mode))


@pytest.mark.parametrize("world,k", [(2, 21), (3, 31), (2, 13)])
def test_owner_partitioned_table_on_one_gpu(tmp_path, oracle, world, k):
    """The all-to-all form of the table (hash mode k = 21 / 31, dense k = 13) with the ranks sharing cuda:0: every rank's table
    holds only the k-mers it owns, the abundance queries travel to the owners and back, rows == the oracle."""
    _run(tmp_path, oracle, world, "gloo", "owner", k=k, batch_bytes=4096, seed=29)


@pytest.mark.parametrize("world,mode,batch_bytes", [(2, "file", None), (3, "file", 150_000), (2, "arrays", None)])
def test_ranks_sharing_one_gpu_equal_one(tmp_path, oracle, world, mode, batch_bytes):
    """The multi-rank flow on ONE GPU: per-rank byte ranges (or shards), per-rank count tables, clamp, all-reduce (gloo),
    per-rank featurize, rows gathered in rank order.  k = 13 keeps the all-reduced table at 128 MB."""
    _run(tmp_path, oracle, world, "gloo", mode, k=13, batch_bytes=batch_bytes, seed=23)
