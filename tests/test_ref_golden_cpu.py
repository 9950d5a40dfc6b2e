"""Pins the CPU oracle to the REFERENCE: oracle outputs vs tests/golden/ref_golden.npz, vectors written by the reference's
own layer code (oracle/_ref: correlation_layer.cu, resample_layer.cu, data_augmentation_layer.cu, channel_norm_layer.cu,
flow_warp_layer.cu, conv/deconv through im2col + cuBLAS) running on a B200 (tests/golden/make_ref_golden.py).  No GPU and no
/root/reference needed here: inputs are regenerated from seeds (tests/golden/ref_cases.py), only outputs are stored."""
import json
import os

import numpy as np
import pytest

from tests import refcheck as RCK
from tests.refcheck import RC

GOLD_JSON = os.path.join(os.path.dirname(RCK.GOLD), "ref_golden.json")


@pytest.fixture(scope="module")
def gold():
    return RCK.golden()


@pytest.mark.parametrize("name", sorted(RC.LAYER_CASES))
def test_oracle_matches_reference_layer(gold, name):
    out = RCK.oracle_eval(name, gold)
    keys = [k for k in gold.files if k.startswith("L/%s/" % name)]
    assert keys, "no reference vectors for " + name
    for full in keys:
        key = full.split("/")[-1]
        if key == "space":
            if gold[full].any():              # all zero: no chromatic-eigen coefficient was active, the statistics never ran
                RCK.check_eigenspace(gold[full], out["space"])
            continue
        assert key in out, (name, key)
        err = RCK.rel_err(out[key], gold[full])
        assert err <= RCK.tol_for(name, key), (name, key, err)


def test_reference_correlation_was_deterministic():
    """CorrelateData sums its 32 lane partials from shared memory without a __syncwarp (correlation_layer.cu:99-105); on sm_100
    three runs of every case gave identical bits (recorded by make_ref_golden.py)."""
    meta = json.load(open(GOLD_JSON))
    corr = {k: v for k, v in meta["layers"].items() if "deterministic_over_3_runs" in v}
    assert corr and all(v["deterministic_over_3_runs"] for v in corr.values())


def test_integer_index_paths_bit_exact(gold):
    """Index arithmetic (displacement -> channel mapping, NEAREST rounding, identity resample) carries no rounding: bit-exact."""
    for name in ("corr_d9", "corr_s2_1", "corr_sub", "rs_nearest_down", "rs_nearest_up2", "rs_same"):
        assert np.array_equal(RCK.oracle_eval(name)["top0"], gold["L/%s/top0" % name]), name


@pytest.mark.parametrize("cname", sorted(RC.NET_CASES))
def test_oracle_net_matches_reference_net(gold, cname):
    """Whole deploy nets: the float64-accumulating oracle vs the reference's layers chained on the GPU, same synthetic weights
    (oracle.net.synth_weights), 1e-4 max-abs on predict_flow_final (north_star tolerance)."""
    import flownet2_b200 as F
    from oracle.net import OracleNet, synth_weights
    from tests.util import maxabs, rng, smooth_images
    model, w, h, batch = RC.NET_CASES[cname]
    proto = F.fill_template(F.model_template(model), w, h)
    _, blob = synth_weights(F.fill_template(F.model_template(model), 64, 64), 1701, proto)
    img0, img1 = smooth_images(rng(11), batch, h, w)
    flow = OracleNet(proto, blob, batch=batch, f64acc=True).forward(img0=img0, img1=img1)["predict_flow_final"]
    want = gold["N/%s/flow" % cname]
    assert np.abs(want).max() > 1.0, "degenerate: flow ~ 0"
    assert maxabs(flow, want) <= 1e-4, maxabs(flow, want)


def test_insert_splits_matches_the_reference_tests():
    """oracle.ref.insert_splits (used to run the reference's layers as a net that can do Backward) against the reference's own
    InsertSplits expectations (test_split_layer.cpp SplitLayerInsertionTest: TestInsertion, TestInsertionTwoTop, TestWithInPlace;
    extracted by tests/golden/make_split_golden.py).  Pure text processing: needs neither oracle/_ref nor a GPU."""
    import json
    from oracle import ref as R
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "insert_splits_fixtures.json")) as f:
        cases = json.load(f)["cases"]

    def specs_of(text):
        _, blocks = R.split_layers(text)
        return [(R._names(b, "name")[0], R._names(b, "type")[0], b) for b in blocks]

    def sig(specs):
        return [(n, t, R._names(b, "bottom"), R._names(b, "top")) for n, t, b in specs]
    assert len(cases) == 3
    for c in cases:
        assert sig(R.insert_splits(specs_of(c["input"]))) == sig(specs_of(c["expected"])), c["test"]
