"""Drop-in for the moment-retrieval part of the reference's ``eval/mr_eval.py`` with the per-query work on the GPU.

Same names, arguments, return structure and error behaviour as the reference:

* ``eval_submission(submission, ground_truth, verbose=True, match_number=True)``  (eval/mr_eval.py:328-414)
* ``eval_moment_retrieval(submission, ground_truth, verbose=True)``               (eval/mr_eval.py:179-216)
* ``compute_mr_ap`` / ``compute_mr_r1``                                            (eval/mr_eval.py:21-138)
* ``eval_main`` CLI: ``--submission_path --gt_path --save_path --not_verbose``    (eval/mr_eval.py:417-439)

Per query the CUDA kernel (``mra_mr_score``) returns the 10 average precisions, the top-1 IoU and the invalid flag;
the final reductions are the reference's own numpy expressions (``ap_array.mean(0)``, ``np.mean(iou >= thd)``,
2-decimal formatting) evaluated on those records in submission order, so rounded and unrounded metrics match.
The four identical short/middle/long/full passes of the reference (:184-192) are computed once.
Multi-GPU: ``score_records_distributed`` scores a rank's shard of queries and gathers fixed-width records to rank 0.
Highlight-detection metrics (``pred_saliency_scores``; eval/mr_eval.py:219-325, eval/mr_utils.py:174-221) are never produced
by mrAudio itself (evaluate.py:50-56) but ``eval_submission`` accepts them: ``eval_highlight`` / ``compute_hl_hit1`` /
``compute_hl_ap`` / ``mk_gt_scores`` / ``get_ap`` are restated here as host numpy (a few thousand scalar operations per
query; SURVEY.md 8a row 21 keeps this branch on the host), pinned by fixtures generated with the reference's own functions.
"""
from __future__ import annotations

import json
import time
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from ._lib import MraError

IOU_THDS = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]  # eval/mr_eval.py:30,99
MAX_PRED, MAX_GT = 256, 64


def load_jsonl(filename):
    with open(filename, "r") as f:
        return [json.loads(l.strip("\n")) for l in f.readlines()]


def pack_windows(submission: List[dict], ground_truth: List[dict]):
    """Host packing: ragged window lists -> padded fp64 arrays in submission order.

    Mirrors the reference's indexing errors: an empty ``pred_relevant_windows`` / ``relevant_windows`` raises
    IndexError like ``d["pred_relevant_windows"][0]`` (eval/mr_eval.py:102) / ``cur_gt_windows[0]`` (:115) would.
    """
    gt_by_qid = {d["qid"]: d["relevant_windows"] for d in ground_truth}
    Q = len(submission)
    npred = np.zeros(Q, dtype=np.int32)
    ngt = np.zeros(Q, dtype=np.int32)
    for i, d in enumerate(submission):
        npred[i] = len(d["pred_relevant_windows"])
        ngt[i] = len(gt_by_qid[d["qid"]])
    if Q and (npred.min() == 0 or ngt.min() == 0):
        raise IndexError("list index out of range")  # reference behaviour for a query without windows
    Pmax, Gmax = int(npred.max()) if Q else 1, int(ngt.max()) if Q else 1
    if Pmax > MAX_PRED or Gmax > MAX_GT:
        raise MraError(f"mr scorer limits exceeded: {Pmax} predictions (max {MAX_PRED}) / {Gmax} GT windows (max {MAX_GT})")
    pred = np.zeros((Q, Pmax, 2), dtype=np.float64)
    gt = np.zeros((Q, Gmax, 2), dtype=np.float64)
    for i, d in enumerate(submission):
        for j, w in enumerate(d["pred_relevant_windows"]):
            pred[i, j, 0], pred[i, j, 1] = w[0], w[1]  # ``[:2]``: a third column (score) is ignored (:102)
        for j, w in enumerate(gt_by_qid[d["qid"]]):
            gt[i, j, 0], gt[i, j, 1] = w[0], w[1]
    return pred, npred, gt, ngt


def score_records(submission: List[dict], ground_truth: List[dict], device=None) -> Dict[str, np.ndarray]:
    """Per-query records in submission order, computed on the GPU: ``ap`` [Q,10], ``iou`` [Q], ``invalid`` [Q], and
    ``tie_ambiguous`` [Q] (queries whose reference result depends on numpy's platform-dependent argsort tie order)."""
    if not torch.cuda.is_available():
        raise MraError("mraudio_b200.mr_eval needs a CUDA device (no CPU fallback)")
    device = torch.device(device if device is not None else "cuda")
    pred, npred, gt, ngt = pack_windows(submission, ground_truth)
    if len(submission) == 0:
        return {"ap": np.zeros((0, 10)), "iou": np.zeros(0), "invalid": np.zeros(0, dtype=bool)}
    to = lambda a: torch.from_numpy(a).pin_memory().to(device, non_blocking=True)
    ap, iou, inv = ops.mr_score(to(pred), to(npred), to(gt), to(ngt), torch.tensor(IOU_THDS, dtype=torch.float64, device=device))
    flags = inv.cpu().numpy()
    # bit 1: the query's result depends on numpy's argsort tie order (see include/mraudio_b200.h, mra_mr_score)
    return {"ap": ap.cpu().numpy(), "iou": iou.cpu().numpy(), "invalid": (flags & 1).astype(bool),
            "tie_ambiguous": (flags & 2).astype(bool)}


def gather_records(rec: Dict[str, np.ndarray], order: np.ndarray, group=None, device=None):
    """The one exchange step of the scoring path: every rank contributes fixed-width records
    ``[ap(10), iou, invalid, order]`` (13 x f64 per query) of ITS shard; rank 0 receives them (``dist.gather``, NCCL
    over NVLink on the GPU box, gloo in the CPU tests), restores the global submission order given by ``order`` and
    returns the merged records; other ranks return None.  Reducing on rank 0 in submission order keeps every metric
    bit-identical to the single-process result."""
    import torch.distributed as dist
    local = np.concatenate([rec["ap"].reshape(-1, 10), rec["iou"].reshape(-1, 1),
                            rec["invalid"].reshape(-1, 1).astype(np.float64), np.asarray(order, dtype=np.float64).reshape(-1, 1)], 1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        allrec = local
    else:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        backend = dist.get_backend(group)
        dev = torch.device(device if device is not None else "cuda") if backend == "nccl" else torch.device("cpu")
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n, group=group)
        nmax = int(max(c.item() for c in counts))
        buf = torch.zeros(nmax, 13, dtype=torch.float64, device=dev)
        buf[: local.shape[0]] = torch.from_numpy(local).to(dev)
        gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0, group=group)
        if rank != 0:
            return None
        allrec = np.concatenate([g[: int(c.item())].cpu().numpy() for g, c in zip(gathered, counts)], 0)
    allrec = allrec[np.argsort(allrec[:, 12], kind="stable")]
    return {"ap": allrec[:, :10], "iou": allrec[:, 10], "invalid": allrec[:, 11].astype(bool)}


def score_records_distributed(submission: List[dict], ground_truth: List[dict], group=None, device=None):
    """Each rank scores ITS shard of queries on its GPU, then ``gather_records``.  ``submission`` entries may carry
    ``"_order"`` (global position) to restore the global submission order on rank 0; rank != 0 returns None."""
    rec = score_records(submission, ground_truth, device=device) if len(submission) else \
        {"ap": np.zeros((0, 10)), "iou": np.zeros(0), "invalid": np.zeros(0, dtype=bool)}
    order = np.array([d.get("_order", i) for i, d in enumerate(submission)], dtype=np.float64)
    return gather_records(rec, order, group=group, device=device)


def _mr_ap_from_records(rec) -> dict:
    # eval/mr_eval.py:87-94
    ap_thds = rec["ap"].mean(0)
    iou_thd2ap = dict(zip([str(e) for e in IOU_THDS], ap_thds))
    iou_thd2ap["average"] = np.mean(ap_thds)
    return {k: float(f"{100 * v:.2f}") for k, v in iou_thd2ap.items()}


def _mr_r1_from_records(rec):
    # eval/mr_eval.py:120-136
    pred_gt_iou = rec["iou"]
    iou_thd2recall_at_one = {str(thd): float(f"{np.mean(pred_gt_iou >= thd) * 100:.2f}") for thd in IOU_THDS}
    invalid_pred_num = int(rec["invalid"].sum())
    r1_avg = np.mean(list(iou_thd2recall_at_one.values()))
    mIoU = np.mean(pred_gt_iou)
    return iou_thd2recall_at_one, r1_avg, mIoU, invalid_pred_num


def compute_mr_ap(submission, ground_truth, iou_thds=None, max_gt_windows=None, max_pred_windows=None, num_workers=8,
                  chunksize=50, _records=None):
    """eval/mr_eval.py:21-94.  ``num_workers`` / ``chunksize`` are accepted for signature compatibility (the GPU kernel
    replaces the multiprocessing pool)."""
    assert iou_thds is None or [float(f"{e:.2f}") for e in iou_thds] == IOU_THDS, "only the standard 10 thresholds"
    if max_gt_windows is not None or max_pred_windows is not None:
        submission = [dict(d, pred_relevant_windows=d["pred_relevant_windows"][:max_pred_windows]) for d in submission]
        ground_truth = [dict(d, relevant_windows=d["relevant_windows"][:max_gt_windows]) for d in ground_truth]
        _records = None
    rec = _records if _records is not None else score_records(submission, ground_truth)
    return _mr_ap_from_records(rec)


def compute_mr_r1(submission, ground_truth, iou_thds=None, _records=None):
    """eval/mr_eval.py:97-138."""
    assert iou_thds is None or [float(f"{e:.2f}") for e in iou_thds] == IOU_THDS, "only the standard 10 thresholds"
    rec = _records if _records is not None else score_records(submission, ground_truth)
    return _mr_r1_from_records(rec)


def _check_unique(submission, ground_truth):
    if len(set(e["qid"] for e in submission)) != len(submission) or \
            len(set(e["qid"] for e in ground_truth)) != len(ground_truth):
        raise MraError("duplicate qids are not supported by the GPU scorer")


def eval_moment_retrieval(submission, ground_truth, verbose=True, _records=None):
    """eval/mr_eval.py:179-216: the reference runs the same computation under four names; it is done once here."""
    _check_unique(submission, ground_truth)
    start_time = time.time()
    rec = _records if _records is not None else score_records(submission, ground_truth)
    ap = _mr_ap_from_records(rec)
    r1, r1_avg, mIoU, invalid = _mr_r1_from_records(rec)
    ret_metrics = {}
    for name in ["short", "middle", "long", "full"]:
        print(f"{name}: {len(ground_truth)}/{len(ground_truth)}={100 * len(ground_truth) / len(ground_truth):.2f} examples.")
        ret_metrics[name] = {"MR-mAP": dict(ap), "MR-R1": dict(r1), "MR-R1-avg": r1_avg, "MR-mIoU": mIoU,
                             "MR-invalid_pred_num": invalid}
        if verbose:
            print(f"[eval_moment_retrieval] [{name}] {time.time() - start_time:.2f} seconds")
    return ret_metrics


# ---------------------------------------------------------------------------------------------- highlight detection
def _precision_recall_curve(y_true: np.ndarray, y_score: np.ndarray):
    """``sklearn.metrics.precision_recall_curve`` for binary labels (the call of eval/mr_utils.py:208): one point per
    distinct score in decreasing-threshold order, returned reversed with the final (precision 1, recall 0) point."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_score = np.asarray(y_score, dtype=np.float64)
    desc = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[desc], y_true[desc]
    threshold_idxs = np.r_[np.where(np.diff(y_score))[0], y_true.size - 1]
    tps = np.cumsum(y_true)[threshold_idxs]
    fps = 1 + threshold_idxs - tps
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=(ps != 0))
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    sl = slice(None, None, -1)
    return np.hstack((precision[sl], 1)), np.hstack((recall[sl], 0)), y_score[threshold_idxs][sl]


def get_ap(y_true, y_predict, interpolate=True, point_11=False):
    """eval/mr_utils.py:174-221: (interpolated) average precision of a binary ranking."""
    assert len(y_true) == len(y_predict), "Prediction and ground truth need to be of the same length"
    if len(set(y_true)) == 1:
        if y_true[0] == 0:
            return 0  # True labels are all zeros
        else:
            return 1
    else:
        assert sorted(set(y_true)) == [0, 1], "Ground truth can only contain elements {0,1}"
    precision, recall, _ = _precision_recall_curve(y_true, y_predict)
    recall = recall.astype(np.float32)
    if interpolate:  # Compute the interpolated precision
        for i in range(1, len(precision)):
            precision[i] = max(precision[i - 1], precision[i])
    if point_11:  # Compute the 11-point approximated AP
        precision_11 = [precision[np.where(recall >= t)[0][-1]] for t in np.arange(0, 1.01, 0.1)]
        return np.mean(precision_11)
    indices = np.where(np.diff(recall))
    return np.mean(precision[indices])


def compute_ap_from_tuple(input_tuple):
    """eval/mr_eval.py:268-281: pad / truncate the predicted scores to the number of clips of the video."""
    idx, w_idx, y_true, y_predict = input_tuple
    if len(y_true) < len(y_predict):
        y_predict = y_predict[: len(y_true)]
    elif len(y_true) > len(y_predict):
        _y_predict = np.zeros(len(y_true))
        _y_predict[: len(y_predict)] = y_predict
        y_predict = _y_predict
    score = get_ap(y_true, y_predict)
    return idx, w_idx, score


def compute_hl_hit1(qid2preds, qid2gt_scores_binary):
    """eval/mr_eval.py:219-234."""
    qid2max_scored_clip_idx = {k: np.argmax(v["pred_saliency_scores"]) for k, v in qid2preds.items()}
    hit_scores = np.zeros((len(qid2preds), 3))
    for idx, qid in enumerate(list(qid2preds.keys())):
        pred_clip_idx = qid2max_scored_clip_idx[qid]
        gt_scores_binary = qid2gt_scores_binary[qid]  # (#clips, 3)
        if pred_clip_idx < len(gt_scores_binary):
            hit_scores[idx] = gt_scores_binary[pred_clip_idx]
    # max over the 3 annotators, mean over the queries
    return float(f"{100 * np.mean(np.max(hit_scores, 1)):.2f}")


def compute_hl_ap(qid2preds, qid2gt_scores_binary, num_workers=8, chunksize=50):
    """eval/mr_eval.py:237-265.  ``num_workers`` / ``chunksize`` are accepted for signature compatibility: the work (3 short
    rankings per query) runs in-process."""
    qid2pred_scores = {k: v["pred_saliency_scores"] for k, v in qid2preds.items()}
    ap_scores = np.zeros((len(qid2preds), 3))  # (#preds, 3)
    for idx, qid in enumerate(list(qid2preds.keys())):
        for w_idx in range(3):  # annotation score idx
            y_true = qid2gt_scores_binary[qid][:, w_idx]
            y_predict = np.array(qid2pred_scores[qid])
            _, _, score = compute_ap_from_tuple((idx, w_idx, y_true, y_predict))
            ap_scores[idx, w_idx] = score
    return float(f"{100 * np.mean(ap_scores):.2f}")


def mk_gt_scores(gt_data, clip_length=2):
    """eval/mr_eval.py:284-294: saliency scores of the relevant clips scattered over the whole video, (#clips, 3)."""
    num_clips = int(gt_data["duration"] / clip_length)
    saliency_scores_full_video = np.zeros((num_clips, 3))
    relevant_clip_ids = np.array(gt_data["relevant_clip_ids"])
    saliency_scores_relevant_clips = np.array(gt_data["saliency_scores"])
    saliency_scores_full_video[relevant_clip_ids] = saliency_scores_relevant_clips
    return saliency_scores_full_video


def eval_highlight(submission, ground_truth, verbose=True):
    """eval/mr_eval.py:297-325: HL-mAP / HL-Hit1 with positives = clips scored >= Fair (2) / Good (3) / VeryGood (4)."""
    qid2preds = {d["qid"]: d for d in submission}
    qid2gt_scores_full_range = {d["qid"]: mk_gt_scores(d) for d in ground_truth}  # scores in range [0, 4]
    highlight_det_metrics = {}
    for gt_saliency_score_min, score_name in zip([2, 3, 4], ["Fair", "Good", "VeryGood"]):
        start_time = time.time()
        qid2gt_scores_binary = {k: (v >= gt_saliency_score_min).astype(float) for k, v in qid2gt_scores_full_range.items()}
        hit_at_one = compute_hl_hit1(qid2preds, qid2gt_scores_binary)
        mean_ap = compute_hl_ap(qid2preds, qid2gt_scores_binary)
        highlight_det_metrics[f"HL-min-{score_name}"] = {"HL-mAP": mean_ap, "HL-Hit1": hit_at_one}
        if verbose:
            print(f"Calculating highlight scores with min score {gt_saliency_score_min} ({score_name})")
            print(f"Time cost {time.time() - start_time:.2f} seconds")
    return highlight_det_metrics


def eval_submission(submission, ground_truth, verbose=True, match_number=True, _records=None):
    """eval/mr_eval.py:328-414: moment retrieval (GPU scorer) and, when the submission carries ``pred_saliency_scores``,
    highlight detection (host)."""
    pred_qids = set([e["qid"] for e in submission])
    gt_qids = set([e["qid"] for e in ground_truth])
    if match_number:
        assert pred_qids == gt_qids, (
            f"qids in ground_truth and submission must match. "
            f"use `match_number=False` if you wish to disable this check"
        )
    else:  # only leave the items that exists in both submission and ground_truth
        shared_qids = pred_qids.intersection(gt_qids)
        submission = [e for e in submission if e["qid"] in shared_qids]
        ground_truth = [e for e in ground_truth if e["qid"] in shared_qids]
        _records = None

    eval_metrics = {}
    eval_metrics_brief = OrderedDict()
    if "pred_relevant_windows" in submission[0]:
        moment_ret_scores = eval_moment_retrieval(submission, ground_truth, verbose=verbose, _records=_records)
        eval_metrics.update(moment_ret_scores)
        moment_ret_scores_brief = {
            "MR-full-mAP": moment_ret_scores["full"]["MR-mAP"]["average"],
            "MR-full-mAP@0.5": moment_ret_scores["full"]["MR-mAP"]["0.5"],
            "MR-full-mAP@0.75": moment_ret_scores["full"]["MR-mAP"]["0.75"],
            "MR-short-mAP": moment_ret_scores["short"]["MR-mAP"]["average"],
            "MR-middle-mAP": moment_ret_scores["middle"]["MR-mAP"]["average"],
            "MR-long-mAP": moment_ret_scores["long"]["MR-mAP"]["average"],
            "MR-full-R1@0.5": moment_ret_scores["full"]["MR-R1"]["0.5"],
            "MR-full-R1@0.7": moment_ret_scores["full"]["MR-R1"]["0.7"],
            "MR-full-R1-avg": moment_ret_scores["full"]["MR-R1-avg"],
            "MR-full-mIoU": moment_ret_scores["full"]["MR-mIoU"],
            "MR-full-invalid_pred_num": moment_ret_scores["full"]["MR-invalid_pred_num"],
        }
        eval_metrics_brief.update(sorted([(k, v) for k, v in moment_ret_scores_brief.items()], key=lambda x: x[0]))

    if "pred_saliency_scores" in submission[0]:
        highlight_det_scores = eval_highlight(submission, ground_truth, verbose=verbose)
        eval_metrics.update(highlight_det_scores)
        highlight_det_scores_brief = dict([(f"{k}-{sub_k.split('-')[1]}", v[sub_k]) for k, v in highlight_det_scores.items()
                                           for sub_k in v])
        eval_metrics_brief.update(highlight_det_scores_brief)

    final_eval_metrics = OrderedDict()
    final_eval_metrics["brief"] = eval_metrics_brief
    final_eval_metrics.update(sorted([(k, v) for k, v in eval_metrics.items()], key=lambda x: x[0]))
    return final_eval_metrics


def eval_main():
    import argparse

    parser = argparse.ArgumentParser(description="Moments and Highlights Evaluation Script")
    parser.add_argument("--submission_path", type=str, help="path to generated prediction file")
    parser.add_argument("--gt_path", type=str, help="path to GT file")
    parser.add_argument("--save_path", type=str, help="path to save the results")
    parser.add_argument("--not_verbose", action="store_true")
    args = parser.parse_args()

    verbose = not args.not_verbose
    submission = load_jsonl(args.submission_path)
    gt = load_jsonl(args.gt_path)
    results = eval_submission(submission, gt, verbose=verbose)
    if verbose:
        print(json.dumps(results, indent=4))

    with open(args.save_path, "w") as f:
        f.write(json.dumps(results, indent=4))


if __name__ == "__main__":
    eval_main()
