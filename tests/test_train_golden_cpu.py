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


# ---- size-independent properties of the training-side oracle functions (no fixture needed) ------------------------------------
def test_l1loss_gradient_is_the_derivative_of_the_loss():
    from oracle import oracle as O
    r = np.random.default_rng(5)
    a = r.standard_normal((2, 2, 5, 6)).astype(np.float32)
    b = r.standard_normal((2, 2, 5, 6)).astype(np.float32)
    b[0, :, 1, 2] = np.nan
    for kw in (dict(), dict(l2_per_location=True), dict(l2_per_location=True, normalize_by_num_entries=True, epsilon=1e-3),
               dict(plateau=0.05, normalize_by_num_entries=True)):
        g0, g1 = O.l1loss_bwd(a, b, 1.5, **kw)
        assert np.array_equal(g1, -g0) and not g0[0, :, 1, 2].any()                 # masked where the ground truth is NaN
        eps = 2e-3
        for idx in [(0, 0, 0, 0), (1, 1, 3, 4), (0, 1, 2, 5)]:
            ap, am = a.copy(), a.copy()
            ap[idx] += eps
            am[idx] -= eps
            fd = 1.5 * (float(O.l1loss_fwd(ap, b, **kw)[0]) - float(O.l1loss_fwd(am, b, **kw)[0])) / (2 * eps)
            assert abs(fd - float(g0[idx])) <= 2e-2 * max(1e-2, abs(float(g0[idx]))) + 2e-3, (kw, idx, fd, float(g0[idx]))


def test_downsample_properties():
    from oracle import oracle as O
    x = np.full((1, 2, 32, 48), 3.25, np.float32)
    assert np.abs(O.downsample_fwd(x, 8, 12) - 3.25).max() <= 2e-6                                    # weights are normalised (float32 sums)
    assert np.array_equal(O.downsample_fwd(x, 32, 48), x)                                             # same size: a copy
    x[0, :, :20, :] = np.nan                                                                          # mostly invalid region -> NaN
    y = O.downsample_fwd(x, 8, 12)
    assert np.isnan(y[0, :, :3]).all() and np.abs(y[0, :, 7] - 3.25).max() <= 2e-6


def test_flow_augmentation_identity_and_composition():
    from oracle import oracle as O
    r = TC.rng_for("fa_property")
    p = TC._mild_coeffs(r, 2, 1.0)
    zero = np.zeros((2, 2, 40, 56), np.float32)
    # both images under the SAME transform and no motion: the augmented flow is zero (M2^-1 M1 = identity)
    assert np.abs(O.flow_augmentation(zero, p, p, 32, 24)).max() <= 2e-4
    # identity transforms: the flow field is only cropped (centre crop of the nearest samples)
    ident = TC._mild_coeffs(r, 2, 0.0)
    f = (3 * r.standard_normal((2, 2, 24, 32))).astype(np.float32)
    out = O.flow_augmentation(f, ident, ident, 24, 16)
    assert np.abs(out - f[:, :, 4:20, 4:28]).max() <= 1e-5


def test_generate_augmentation_parameters_modes():
    from oracle import oracle as O
    from tests.refcheck import _gens_of
    gens = _gens_of(TC.TRAIN_CASES["ga_add"]["text"])
    r = TC.rng_for("ga_property")
    inp = TC.coeff_blob(r, 3, effects=False).reshape(3, 42)
    reg = O.generate_augmentation_parameters(None, "regenerate", gens, 48, 32, 64, 48, num=3)
    assert np.array_equal(reg[0], reg[1]) and np.array_equal(reg[0], reg[2])          # deterministic generators, no input
    add = O.generate_augmentation_parameters(inp, "add", gens, 48, 32, 64, 48)
    rep = O.generate_augmentation_parameters(inp, "replace", gens, 48, 32, 64, 48)
    # "replace" discards the incoming spatial coefficients, "add" keeps them (array form: sums)
    assert np.allclose(rep[:, 1:3], reg[:, 1:3], atol=1e-6)
    assert np.allclose(add[:, 6] - inp[:, 6], reg[:, 6], atol=1e-6)                    # gamma: log-domain sum
