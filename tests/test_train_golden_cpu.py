"""CPU: the oracle's training-side restatements (oracle.conv_bwd, l1loss_fwd/bwd, downsample_fwd, flow_augmentation,
generate_augmentation_parameters) against the outputs of the REFERENCE's own layer classes (oracle/_ref, recorded on the B200 by
tests/golden/make_train_golden.py into tests/golden/train_golden.npz).  This is what pins those oracle functions."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import train_cases as TC  # noqa: E402
from tests.refcheck import TRAIN_TOL, train_oracle_eval  # noqa: E402
from tests.util import maxabs  # noqa: E402

GOLD = os.path.join(HERE, "golden", "train_golden.npz")


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.fail("tests/golden/train_golden.npz is missing: run tests/golden/make_train_golden.py on the GPU box")
    return np.load(GOLD)


@pytest.mark.parametrize("name", list(TC.TRAIN_CASES))
def test_oracle_matches_reference(gold, name):
    out = train_oracle_eval(name)
    keys = [k[len("T/%s/" % name):] for k in gold.files if k.startswith("T/%s/" % name)]
    assert keys and set(out) == set(keys), (sorted(out), sorted(keys))
    tol = TRAIN_TOL[TC.TRAIN_CASES[name]["kind"]]
    for k in keys:
        want = gold["T/%s/%s" % (name, k)]
        got = np.asarray(out[k]).reshape(want.shape)
        scale = max(float(np.nanmax(np.abs(want))) if want.size else 0.0, 1e-6)
        assert maxabs(got, want) <= tol * scale, (name, k, maxabs(got, want), scale)
