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


def test_oracle_net_backward_matches_reference_net(gold):
    """OracleNet.forward / backward over the FlowNet2-C graph with five EPE losses against the reference's Net semantics (its layers,
    Split layers inserted like Net::Init, loss weights; oracle.ref.RefNet on the B200): losses and every parameter gradient."""
    import flownet2_b200 as F
    from oracle.net import OracleNet, synth_weights
    if "N/lossnet/loss2" not in gold.files:
        pytest.fail("train_golden.npz has no whole-net gradients: re-run tests/golden/make_train_golden.py")
    proto, ins = TC.loss_net_proto(), TC.loss_net_inputs()
    small = F.fill_template(F.model_template("FlowNet2-C"), 64, 64)
    _, cm = synth_weights(small, TC.LOSS_NET["seed"], F.fill_template(F.model_template("FlowNet2-C"), TC.LOSS_NET["w"], TC.LOSS_NET["h"]))
    net = OracleNet(proto, cm, batch=TC.LOSS_NET["batch"], f64acc=True)
    B = net.forward(**ins)
    for lvl in TC.LOSS_NET["weights"]:
        want = float(gold["N/lossnet/loss%d" % lvl][0])
        assert abs(float(B["flow_loss%d" % lvl][0]) - want) <= 2e-5 * abs(want), lvl
    _, P = net.backward()
    keys = [k for k in gold.files if k.startswith("N/lossnet/grad/")]
    assert len(keys) >= 40
    for k in keys:
        _, _, _, name, i = k.split("/")
        want = gold[k]
        got = TC.grad_signature(name, int(i), P[name][int(i)])
        scale = float(np.abs(want).max())
        assert np.abs(got - want).max() <= 2e-5 * scale, (k, float(np.abs(got - want).max()), scale)      # measured: 1.0e-6
