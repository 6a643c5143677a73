"""Highlight-detection branch of eval_submission (eval/mr_eval.py:219-325, eval/mr_utils.py:174-221)
against fixtures produced by the reference's own functions (tests/golden/make_golden_highlight.py).  Host numpy by
design, so everything except the combined moment-retrieval + highlight case runs without a GPU."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

FX = json.load(open(os.path.join(GOLDEN, "hl_eval.json")))


def test_get_ap_matches_reference_vectors():
    from mraudio_b200 import mr_eval
    n = 0
    for v in FX["get_ap"]:
        y, s = np.array(v["y_true"], dtype=float), np.array(v["y_predict"])
        try:
            got = float(mr_eval.get_ap(y, s, interpolate=v["interpolate"], point_11=v["point_11"]))
        except Exception as e:
            got = type(e).__name__
        assert got == v["ap"], (v, got)
        n += 1
    assert n >= 150


def test_precision_recall_curve_matches_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    from mraudio_b200 import mr_eval
    rng = np.random.default_rng(0)
    for _ in range(200):
        m = int(rng.integers(2, 60))
        y = (rng.random(m) < 0.4).astype(float)
        if y.sum() == 0:
            y[0] = 1.0
        s = np.round(rng.random(m), int(rng.integers(1, 4)))
        p0, r0, t0 = sk.precision_recall_curve(y, s)
        p1, r1, t1 = mr_eval._precision_recall_curve(y, s)
        assert np.array_equal(p0, p1) and np.array_equal(r0, r1) and np.array_equal(t0, t1)


def test_eval_highlight_matches_reference():
    from mraudio_b200 import mr_eval
    for name, case in FX["cases"].items():
        got = mr_eval.eval_highlight(case["submission"], case["ground_truth"], verbose=False)
        assert got == case["eval_highlight"], name
    # a highlight-only submission goes through eval_submission without touching the GPU scorer
    case = FX["cases"]["hl_only_120"]
    res = mr_eval.eval_submission(case["submission"], case["ground_truth"], verbose=False)
    assert json.loads(json.dumps(res)) == case["eval_submission"]
    assert list(res.keys()) == list(case["eval_submission"].keys())
    assert list(res["brief"].keys()) == list(case["eval_submission"]["brief"].keys())


def test_mk_gt_scores_and_hit1_edge_cases():
    from mraudio_b200 import mr_eval
    g = {"qid": 0, "duration": 11, "relevant_clip_ids": [1, 3], "saliency_scores": [[4, 0, 2], [1, 3, 3]]}
    full = mr_eval.mk_gt_scores(g)
    assert full.shape == (5, 3) and full[1].tolist() == [4, 0, 2] and full[3].tolist() == [1, 3, 3] and full[0].sum() == 0
    # the arg-max clip lies beyond the video (more scores than clips): no hit, like the reference's bounds check
    preds = {0: {"pred_saliency_scores": [0, 0, 0, 0, 0, 0, 9.0]}}
    assert mr_eval.compute_hl_hit1(preds, {0: (full >= 2).astype(float)}) == 0.0
    preds = {0: {"pred_saliency_scores": [0, 5.0, 0]}}
    assert mr_eval.compute_hl_hit1(preds, {0: (full >= 2).astype(float)}) == 100.0
    # shorter / longer prediction vectors are zero-padded / truncated to the number of clips
    y = np.array([0, 1, 0, 1, 0], dtype=float)
    assert mr_eval.compute_ap_from_tuple((0, 0, y, np.array([0.1, 0.9])))[2] == mr_eval.get_ap(y, np.array([0.1, 0.9, 0, 0, 0]))
    assert mr_eval.compute_ap_from_tuple((0, 0, y, np.arange(8.0)))[2] == mr_eval.get_ap(y, np.arange(5.0))


@pytest.mark.gpu
def test_eval_submission_with_both_branches_matches_reference():
    from mraudio_b200 import mr_eval
    case = FX["cases"]["hl_and_mr_80"]
    res = mr_eval.eval_submission(case["submission"], case["ground_truth"], verbose=False)
    assert json.loads(json.dumps(res)) == case["eval_submission"]
    assert list(res["brief"].keys()) == list(case["eval_submission"]["brief"].keys())
