"""CPU oracle for the moment-retrieval scorer.  TEST INFRASTRUCTURE ONLY (see oracle/qformer_oracle.py header).

A plain-numpy restatement of the reference's ``eval/mr_eval.py`` + ``eval/mr_utils.py`` moment-retrieval metrics,
written per query so that it can be compared record by record with the CUDA scorer kernel.  Every function cites the
reference lines it follows (paths relative to /root/reference).

Pinning: ``tests/golden/make_golden.py`` imports the reference's own ``eval.mr_eval`` in the build container
(``PYTHONPATH=/root/reference``) and freezes its outputs on seeded random submissions into
``tests/golden/mr_eval_*.json``; ``tests/test_oracle_scorer.py`` checks this restatement against those fixtures, the
docstring example at ``eval/mr_utils.py:49-55`` and the known-answer vectors of SURVEY.md section 8(a).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np

# eval/mr_eval.py:30,99: thresholds are ``float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)``
IOU_THDS = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]


def temporal_iou_cross_1xM(pred: Sequence[float], gts: np.ndarray) -> np.ndarray:
    """eval/mr_utils.py:40-67 with N == 1: true-union IoU; 0/0 -> nan (numpy semantics, RuntimeWarning suppressed)."""
    p0, p1 = float(pred[0]), float(pred[1])
    gts = np.asarray(gts, dtype=np.float64).reshape(-1, 2)
    area1 = p1 - p0
    area2 = gts[:, 1] - gts[:, 0]
    left = np.maximum(p0, gts[:, 0])
    right = np.minimum(p1, gts[:, 1])
    inter = np.clip(right - left, 0, None)
    union = area1 + area2 - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / union


def temporal_iou_paired(pred: Sequence[float], gt: Sequence[float]) -> float:
    """eval/mr_utils.py:16-37 for one pair: hull 'union'; 0 where union == 0."""
    inter = max(0.0, min(float(pred[1]), float(gt[1])) - max(float(pred[0]), float(gt[0])))
    union = max(float(pred[1]), float(gt[1])) - min(float(pred[0]), float(gt[0]))
    return inter / union if union != 0 else 0.0


def interpolated_precision_recall(precision: np.ndarray, recall: np.ndarray) -> float:
    """eval/mr_utils.py:70-86 (VOC-2011 envelope)."""
    mprecision = np.hstack([[0], precision, [0]])
    mrecall = np.hstack([[0], recall, [1]])
    for i in range(len(mprecision) - 1)[::-1]:
        mprecision[i] = max(mprecision[i], mprecision[i + 1])
    idx = np.where(mrecall[1::] != mrecall[0:-1])[0] + 1
    return float(np.sum((mrecall[idx] - mrecall[idx - 1]) * mprecision[idx]))


def stable_desc_order(iou: np.ndarray) -> List[int]:
    """``tiou_arr.argsort()[::-1]`` (eval/mr_utils.py:147).  numpy's default argsort on <= 16 elements is an insertion
    sort (stable ascending; nan last), so reversing puts nan first, then descending IoU with ties in DESCENDING index
    order ([.5,.5,.2,.5] -> [3,1,0,2]).  CAVEAT: on CPUs where numpy dispatches to its AVX-512/AVX2
    x86-simd-sort argsort (numpy >= 2.0) arrays of >= 4 elements are sorted by a bitonic network whose tie order is
    neither stable nor documented -- the reference's result is then platform-dependent whenever one prediction has
    EXACTLY equal IoU with two different ground-truth windows out of >= 4.  This restatement (and the CUDA kernel)
    defines the tie order as the stable one, which is what the reference does for <= 3 ground truths everywhere."""
    n = len(iou)
    keys = [(1 if np.isnan(v) else 0, v if not np.isnan(v) else 0.0, i) for i, v in enumerate(iou)]
    asc = sorted(range(n), key=lambda i: (keys[i][0], keys[i][1], i))
    return asc[::-1]


def average_precision_one_query(preds: Sequence[Sequence[float]], gts: Sequence[Sequence[float]],
                                thds: Sequence[float] = IOU_THDS) -> np.ndarray:
    """eval/mr_utils.py:89-171 for one qid (all GT share the pred's video-id).  Predictions are visited in list order
    (no score sort); each prediction greedily takes the best still-unlocked GT with IoU >= thd."""
    nt, ng, npred = len(thds), len(gts), len(preds)
    ap = np.zeros(nt)
    if npred == 0:
        return ap
    tp = np.zeros((nt, npred))
    fp = np.zeros((nt, npred))
    if ng == 0:
        fp[:] = 1  # eval/mr_utils.py:131-133 (video-id not in ground truth)
    else:
        lock = -np.ones((nt, ng))
        g = np.asarray(gts, dtype=np.float64)[:, :2]
        for idx, p in enumerate(preds):
            iou = temporal_iou_cross_1xM(p, g)
            order = stable_desc_order(iou)
            for t, thd in enumerate(thds):
                for j in order:
                    if iou[j] < thd:           # nan < thd is False: a nan IoU counts as a match (reference quirk)
                        fp[t, idx] = 1
                        break
                    if lock[t, j] >= 0:
                        continue
                    tp[t, idx] = 1
                    lock[t, j] = idx
                    break
                if fp[t, idx] == 0 and tp[t, idx] == 0:
                    fp[t, idx] = 1
    tp_c = np.cumsum(tp, axis=1).astype(float)
    fp_c = np.cumsum(fp, axis=1).astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = tp_c / float(ng)
        precision = tp_c / (tp_c + fp_c)
    for t in range(nt):
        ap[t] = interpolated_precision_recall(precision[t], recall[t])
    return ap


def r1_one_query(preds: Sequence[Sequence[float]], gts: Sequence[Sequence[float]]) -> Tuple[float, bool]:
    """eval/mr_eval.py:97-131 for one qid: top-1 prediction ``[:2]``, the GT with the highest cross IoU (``np.argmax``:
    first maximum, a nan wins), then the paired (hull) IoU.  Returns (iou, invalid) where invalid = ``-1 in pred``."""
    top = [float(preds[0][0]), float(preds[0][1])]
    k = 0
    if len(gts) > 0:
        k = int(np.argmax(temporal_iou_cross_1xM(top, np.asarray(gts, dtype=np.float64)[:, :2])))
    iou = temporal_iou_paired(top, gts[k][:2])
    return iou, (-1 in top)


def score_records(submission: List[dict], ground_truth: List[dict]) -> Dict[str, np.ndarray]:
    """Per-query records in submission order: ap [Q,10], iou [Q], invalid [Q]."""
    gt_by_qid = {d["qid"]: d["relevant_windows"] for d in ground_truth}
    ap = np.zeros((len(submission), len(IOU_THDS)))
    iou = np.zeros(len(submission))
    inv = np.zeros(len(submission), dtype=bool)
    has_pred = np.zeros(len(submission), dtype=bool)
    for i, d in enumerate(submission):
        gts = gt_by_qid[d["qid"]]
        preds = d["pred_relevant_windows"]
        has_pred[i] = len(preds) > 0
        ap[i] = average_precision_one_query(preds, gts)
        iou[i], inv[i] = r1_one_query(preds, gts)
    return {"ap": ap, "iou": iou, "invalid": inv, "has_pred": has_pred}


def reduce_records(rec: Dict[str, np.ndarray]) -> dict:
    """The reductions of eval/mr_eval.py:87-94 (mAP) and :120-136 (R1, mIoU) over per-query records."""
    ap_array = rec["ap"][rec["has_pred"]]  # compute_mr_ap only visits qids that have >= 1 predicted window (:66-68)
    ap_thds = ap_array.mean(0)
    m_ap = dict(zip([str(e) for e in IOU_THDS], ap_thds))
    m_ap["average"] = np.mean(ap_thds)
    m_ap = {k: float(f"{100 * v:.2f}") for k, v in m_ap.items()}
    r1 = {str(t): float(f"{np.mean(rec['iou'] >= t) * 100:.2f}") for t in IOU_THDS}
    return {
        "MR-mAP": m_ap,
        "MR-R1": r1,
        "MR-R1-avg": np.mean(list(r1.values())),
        "MR-mIoU": np.mean(rec["iou"]),
        "MR-invalid_pred_num": int(rec["invalid"].sum()),
    }


def eval_submission(submission: List[dict], ground_truth: List[dict], verbose: bool = False,
                    match_number: bool = True) -> OrderedDict:
    """eval/mr_eval.py:328-414, moment-retrieval part (the four identical short/middle/long/full passes of
    eval_moment_retrieval :179-216 are computed once and replicated)."""
    pred_qids = set(e["qid"] for e in submission)
    gt_qids = set(e["qid"] for e in ground_truth)
    if match_number:
        assert pred_qids == gt_qids, "qids in ground_truth and submission must match. " \
                                     "use `match_number=False` if you wish to disable this check"
    else:
        shared = pred_qids & gt_qids
        submission = [e for e in submission if e["qid"] in shared]
        ground_truth = [e for e in ground_truth if e["qid"] in shared]
    # duplicate qids: the reference's dicts keep the LAST entry for R1 (dict comprehension, :101-103) but
    # concatenate windows for mAP (:33-62); the oracle only supports unique qids and says so.
    assert len(set(e["qid"] for e in submission)) == len(submission) and \
        len(set(e["qid"] for e in ground_truth)) == len(ground_truth), "duplicate qids unsupported"
    metrics = reduce_records(score_records(submission, ground_truth))
    out = {name: metrics for name in ("short", "middle", "long", "full")}
    brief = {
        "MR-full-mAP": metrics["MR-mAP"]["average"],
        "MR-full-mAP@0.5": metrics["MR-mAP"]["0.5"],
        "MR-full-mAP@0.75": metrics["MR-mAP"]["0.75"],
        "MR-short-mAP": metrics["MR-mAP"]["average"],
        "MR-middle-mAP": metrics["MR-mAP"]["average"],
        "MR-long-mAP": metrics["MR-mAP"]["average"],
        "MR-full-R1@0.5": metrics["MR-R1"]["0.5"],
        "MR-full-R1@0.7": metrics["MR-R1"]["0.7"],
        "MR-full-R1-avg": metrics["MR-R1-avg"],
        "MR-full-mIoU": metrics["MR-mIoU"],
        "MR-full-invalid_pred_num": metrics["MR-invalid_pred_num"],
    }
    final = OrderedDict()
    final["brief"] = OrderedDict(sorted(brief.items(), key=lambda x: x[0]))
    final.update(sorted(out.items(), key=lambda x: x[0]))
    return final


def synth_submission(num_queries: int, seed: int = 7, duration: int = 150, clip_len: int = 2, max_pred: int = 5,
                     max_gt: int = 3, invalid_frac: float = 0.02, float_windows: bool = False):
    """SURVEY.md 8(d) cfg5 scorer inputs: integer windows on the 2-s grid in [0, duration], 1..max_pred predictions,
    1..max_gt ground truths, ``invalid_frac`` of the queries predicting [[-1, -1]]."""
    rng = np.random.default_rng(seed)
    sub, gt = [], []
    nclip = duration // clip_len

    def window():
        a, b = sorted(rng.integers(0, nclip + 1, size=2).tolist())
        if a == b:
            b = min(nclip, a + 1)
            a = b - 1
        if float_windows:
            return [round(a * clip_len + float(rng.random()), 3), round(b * clip_len + 1 + float(rng.random()), 3)]
        return [int(a * clip_len), int(b * clip_len)]

    for q in range(num_queries):
        g = [window() for _ in range(int(rng.integers(1, max_gt + 1)))]
        if rng.random() < invalid_frac:
            p = [[-1, -1]]
        else:
            p = [window() for _ in range(int(rng.integers(1, max_pred + 1)))]
            # make exact hits and ties common so the greedy matcher and the tie order are exercised
            if rng.random() < 0.5:
                p[int(rng.integers(0, len(p)))] = list(g[int(rng.integers(0, len(g)))])
            if rng.random() < 0.2 and len(g) > 1:
                g[-1] = list(g[0])
        sub.append({"qid": q, "pred_relevant_windows": p})
        gt.append({"qid": q, "relevant_windows": g})
    return sub, gt
