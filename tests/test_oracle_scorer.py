"""Pins oracle/mr_eval_oracle.py against fixtures frozen from the reference's own eval/mr_eval.py."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import mr_eval_oracle as mo

CASES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "mr_eval_*.json")) if "docstring" not in p)


def _same(a, b):
    """exact equality with nan == nan"""
    if isinstance(a, dict):
        assert set(a) == set(map(str, b)) or set(a) == set(b)
        return all(_same(a[k], b[k]) for k in a)
    a, b = float(a), float(b)
    return a == b or (np.isnan(a) and np.isnan(b))


def test_docstring_cross_iou():
    # eval/mr_utils.py:49-55
    exp = json.load(open(os.path.join(GOLDEN, "mr_eval_docstring_iou.json")))
    gts = np.array([[0, 0.3], [0.0, 1.0]])
    for i, p in enumerate([[0, 0.2, 0.9], [0.5, 1.0, 0.2]]):
        got = mo.temporal_iou_cross_1xM(p, gts)
        assert np.array_equal(got, np.array(exp["iou"][i]))
    assert np.allclose(exp["iou"], [[2 / 3, 0.2], [0.0, 0.5]])


def test_argsort_tie_order_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(500):
        n = int(rng.integers(1, 4))  # <= 3 elements: numpy's argsort is stable on every platform (see oracle docstring)
        v = rng.integers(0, 4, size=n) / 4.0
        v = v.astype(np.float64)
        if rng.random() < 0.3:
            v[rng.integers(0, n)] = np.nan
        assert mo.stable_desc_order(v) == v.argsort()[::-1].tolist()
    assert mo.stable_desc_order(np.array([.5, .5, .2, .5])) == [3, 1, 0, 2]


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-5] for p in CASES])
def test_oracle_matches_reference_fixture(path):
    fx = json.load(open(path))
    sub, gt = fx["submission"], fx["ground_truth"]
    rec = mo.score_records(sub, gt)
    ref_ap = np.array(fx["per_query_ap"], dtype=np.float64)
    assert rec["ap"].shape == ref_ap.shape
    assert np.array_equal(rec["ap"], ref_ap, equal_nan=True), "per-query AP differs from the reference"
    got = mo.eval_submission(sub, gt)
    exp = fx["eval_submission"]
    for k, v in exp["brief"].items():
        assert _same(got["brief"][k], v), (k, got["brief"][k], v)
    for name in ("short", "middle", "long", "full"):
        for k in ("MR-mAP", "MR-R1"):
            assert _same(got[name][k], exp[name][k]), (name, k)
        assert got[name]["MR-invalid_pred_num"] == exp[name]["MR-invalid_pred_num"]
        assert _same(got[name]["MR-mIoU"], exp[name]["MR-mIoU"])
        assert _same(got[name]["MR-R1-avg"], exp[name]["MR-R1-avg"])


def test_known_answer_vectors():
    # SURVEY.md section 8(a)
    ap = mo.average_precision_one_query([[10, 20], [30, 40]], [[12, 20], [30, 41]])
    assert ap.tolist() == [1, 1, 1, 1, 1, 1, 1, 0.25, 0.25, 0]
    assert np.isnan(mo.temporal_iou_cross_1xM([-1, -1], np.array([[0, 0]]))[0])


def test_match_number_assert():
    sub, gt = mo.synth_submission(4)
    with pytest.raises(AssertionError):
        mo.eval_submission(sub[:3], gt)
    out = mo.eval_submission(sub[:3], gt, match_number=False)
    assert "MR-full-mAP" in out["brief"]
