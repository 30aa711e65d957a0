import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(d for d in os.listdir(GOLDEN) if os.path.isfile(os.path.join(GOLDEN, d, "params.json")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _csv(path):
    """labels, int64 matrix - parsed the way feature.py:115-121 does (pandas, header=None)."""
    import pandas as pd

    if os.path.getsize(path) == 0:
        return np.array([], dtype=object), None
    df = pd.read_csv(path, header=None, dtype={0: str})
    return df[0].to_numpy(), df.drop(columns=0).to_numpy()


class GoldenCase:
    def __init__(self, name):
        self.name = name
        self.dir = os.path.join(GOLDEN, name)
        self.params = json.load(open(os.path.join(self.dir, "params.json")))
        f = self.params["files"]
        self.interleaved = os.path.join(self.dir, f["i"]) if "i" in f else None
        self.reads1 = os.path.join(self.dir, f["1"]) if "1" in f else None
        self.reads2 = os.path.join(self.dir, f["2"]) if "2" in f else None
        self.dump = os.path.join(self.dir, "kmers.dump")
        self.abd_labels, self.abd = _csv(os.path.join(self.dir, "abundance.csv"))
        self.tnf_labels, self.tnf = _csv(os.path.join(self.dir, "tnf.csv"))

    @property
    def path1(self):
        return self.interleaved or self.reads1


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return GoldenCase(request.param)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build(ref=os.path.isdir("/root/reference"))
    return O
