import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def load_pavillon_weights():
    z = np.load(os.path.join(GOLDEN, "weights_pavillon.npz"), allow_pickle=False)
    sd0 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}
    sd1 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd1/")}
    return sd0, sd1


def case_weights(case):
    """Weights used to generate a golden stage case (see oracle/gen_golden.py)."""
    from oracle import adanerf_oracle as orc
    if case.startswith("pav"):
        return load_pavillon_weights()
    if case.startswith("shaped"):
        return orc.make_weights("shaped", seed=0)
    if case.startswith("ndc"):
        return orc.make_weights("ndc", seed=0)
    return orc.make_weights("rand", seed=0)


@pytest.fixture(scope="session")
def pavillon_weights():
    return load_pavillon_weights()
