"""SURVEY.md §8f.3 - the weighted sampler and the batch gather.  CPU: the oracle's restatement of numpy.random.choice against
golden vectors produced by the reference's own CustomWeightedRandomSampler.  GPU: the device sampler against the same vectors,
and DeviceBatches against plain indexing of Data."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "sampler", "*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_sampler_matches_reference_golden(path):
    from oracle import ingest_oracle as I

    g = np.load(path)
    np.random.seed(int(g["seed"]))
    for want in (g["epoch1"], g["epoch2"]):
        got = I.weighted_choice(g["weights"], int(g["num_samples"]), bool(g["replacement"]))
        assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_device_sampler_matches_reference_golden(path):
    from pangaea_b200.sampler import CustomWeightedRandomSampler

    g = np.load(path)
    s = CustomWeightedRandomSampler(g["weights"], num_samples=int(g["num_samples"]), replacement=bool(g["replacement"]))
    np.random.seed(int(g["seed"]))
    assert np.array_equal(np.array(list(iter(s))), g["epoch1"])
    assert np.array_equal(s.indices_cuda().cpu().numpy(), g["epoch2"])   # the next epoch continues numpy's generator
    assert len(s) == int(g["num_samples"])


@pytest.mark.gpu
def test_device_sampler_large_and_argument_checks():
    from oracle import ingest_oracle as I
    from pangaea_b200.sampler import CustomWeightedRandomSampler

    rng = np.random.default_rng(5)
    w = (rng.random(400_000) * 0.9 + 0.05) ** 2
    for rep, m in ((True, 400_000), (False, 280_000)):
        np.random.seed(7)
        want = I.weighted_choice(w, m, rep)
        np.random.seed(7)
        got = CustomWeightedRandomSampler(w, m, replacement=rep).indices_cuda().cpu().numpy()
        assert np.array_equal(got, want)
        if not rep:
            assert len(np.unique(got)) == m
    with pytest.raises(ValueError):
        CustomWeightedRandomSampler(w, 0)
    with pytest.raises(ValueError):
        CustomWeightedRandomSampler(np.array([1.0, 0.0, 0.0]), 2, replacement=False).indices_cuda()


@pytest.mark.gpu
def test_device_batches_replace_the_dataloader():
    """DeviceBatches(Data, batch_size, sampler) yields what DataLoader(Data, batch_size, sampler=...) yields - the rows the
    sampler drew, in its order - as CUDA tensors gathered on the device."""
    from pangaea_b200 import Data
    from pangaea_b200.sampler import CustomWeightedRandomSampler, DeviceBatches

    rng = np.random.default_rng(2)
    abd = rng.integers(0, 500, size=(3000, 400)) * (rng.random((3000, 400)) < 0.05)
    tnf = rng.integers(0, 900, size=(3000, 136))
    names = np.array([f"bc{i}" for i in range(3000)], dtype=object)
    ds = Data(names, abd, tnf)
    np.random.seed(3)
    sampler = CustomWeightedRandomSampler(ds.weights, num_samples=len(ds))
    batches = list(DeviceBatches(ds, 256, sampler=sampler))
    np.random.seed(3)
    idx = np.array(list(iter(CustomWeightedRandomSampler(ds.weights, num_samples=len(ds)))))
    assert len(batches) == (3000 + 255) // 256 and sum(len(b["bc"]) for b in batches) == 3000
    at = 0
    for b in batches:
        m = len(b["bc"])
        want = [ds[int(i)] for i in idx[at:at + m]]
        assert b["abd"].is_cuda and np.array_equal(b["abd"].cpu().numpy(), np.stack([w["abd"] for w in want]))
        assert np.array_equal(b["tnf"].cpu().numpy(), np.stack([w["tnf"] for w in want])) and b["bc"] == [w["bc"] for w in want]
        at += m
    plain = list(DeviceBatches(ds, 1000))
    assert np.array_equal(np.concatenate([b["abd"].cpu().numpy() for b in plain]), ds.abd)
