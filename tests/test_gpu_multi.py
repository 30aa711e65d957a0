"""2-GPU parity (NCCL): pangaea_b200.distributed.extract_features_sharded on two ranks ==
the single-GPU result == the oracle.  Skipped on boxes with fewer than two GPUs."""
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


def _worker(rank, world, port, path, out_dir):
    import torch
    import torch.distributed as dist

    from pangaea_b200.distributed import extract_features_sharded, gather_rows

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    fq = _lib.Fastq(path)
    seq, off, flag, keep = fq.arrays()
    ctx = _lib.Context(device=rank)
    feats, shard = extract_features_sharded(ctx, seq, off, flag, keep)
    merged = gather_rows(feats, shard, fq.label)
    if rank == 0:
        names, abd, tnf = merged
        np.savez(os.path.join(out_dir, "merged.npz"), names=names, abd=abd, tnf=tnf, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    data = synth.generate(n_barcodes=300, mean_pairs=15, read_len=100, n_genomes=3, genome_len=80_000, frag_len=10_000, seed=17,
                          unbarcoded_pairs=25, n_rate=0.002)
    path = synth.write_interleaved(str(tmp_path / "reads.fq"), data)
    want_names, want_abd, want_tnf = oracle.featurize(path, None)
    mp.spawn(_worker, args=(2, _free_port(), path, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "merged.npz", allow_pickle=True)
    assert list(got["names"]) == list(want_names)
    assert np.array_equal(got["abd"], want_abd) and np.array_equal(got["tnf"], want_tnf)
