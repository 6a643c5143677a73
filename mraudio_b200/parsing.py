"""LLM text -> moment windows: the producer of the scorer's inputs.

Reference surface kept (same names, arguments, return values and exceptions):

* ``post_process(pred: str) -> str``          utils/utils.py:66-132   (repairs "[[0 1],, [7, 4]]"-style generations)
* ``moment_str_to_list(m: str) -> list``      utils/utils.py:364-415  (string -> nested int list, ``[[-1, -1]]`` on failure)
* call sites: ``evaluate.py:48`` / ``utils/trainer.py:168-169`` compose them per generated string and write one jsonl
  record per query (``evaluate.py:50-58``).

This is host-side string work (regular expressions + ``ast.literal_eval``): it stays on the CPU by design, but is batched
here so that a whole evaluation sweep goes from raw strings to the padded arrays ``mra_mr_score`` consumes without a
Python-level record list in between (``windows_to_arrays`` / ``parse_batch``).  Behaviour is pinned by
``tests/golden/parser_cases.json`` (3 080 strings run through the reference's own two functions, including the seven
inputs on which the reference raises TypeError / KeyError).
"""
from __future__ import annotations

import ast
import json
import re
from typing import Iterable, List, Sequence, Tuple

import numpy as np

INVALID = [[-1, -1]]
_NESTED = re.compile(r"\[\[.*\]\]")            # "looks like a nested list" test, anchored at the start only
_WINDOW_GAP = re.compile(r"\s+(?=\[)")         # whitespace that precedes an opening bracket
_TRAILING_COMMAS = re.compile(r",+$")
_DIGIT_SPACE_DIGIT = re.compile(r"(\d) (\d)")
_COMMA_RUN = re.compile(r",+")
_UINT = re.compile(r"\d+")
# canonical output of post_process for a well-formed generation: "[[s, e], [s, e], ...]" with plain non-negative decimal
# integers (no leading zeros: ast.literal_eval rejects "007").  For these the literal_eval round trip below is just
# "take the integers in pairs"; everything else goes through ast.literal_eval as in the reference.
_INT = r"(?:0|[1-9]\d*)"
_CANONICAL = re.compile(r"\[\[" + _INT + ", " + _INT + r"\](?:, \[" + _INT + ", " + _INT + r"\])*\]")


def _repair_window(w: str) -> str:
    """One window of the split: drop trailing commas, "3 7" -> "3, 7", collapse comma runs, order start <= end."""
    w = _TRAILING_COMMAS.sub("", w)
    w = _DIGIT_SPACE_DIGIT.sub(r"\1, \2", w)
    w = _COMMA_RUN.sub(",", w)
    nums = _UINT.findall(w)
    if len(nums) == 2 and int(nums[0]) > int(nums[1]):
        w = f"[{nums[1]}, {nums[0]}]"
    return w


def post_process(pred: str) -> str:
    """utils/utils.py:66-132.  Cut at the first ``</s>``, drop line breaks; anything that does not start like a nested
    list becomes ``"[[-1, -1]]"``; otherwise strip the outer brackets, split into windows at whitespace followed by
    ``[``, repair each window and re-join with ``", "``."""
    text = pred.split("</s>")[0].replace("\n", "").replace("\r", "")
    if _NESTED.match(text) is None:
        return "[[-1, -1]]"
    inner = text[1:-1]
    return "[" + ", ".join(_repair_window(w) for w in _WINDOW_GAP.split(inner)) + "]"


def _sanitise(entry):
    """utils/utils.py:405-413 for one element of the parsed list (duck-typed exactly like the reference, so the same
    inputs raise the same TypeError / KeyError): an int becomes [-1, -1]; a wrong length becomes the 1-element marker
    [-len]; non-int members become -1 (in place)."""
    if isinstance(entry, int):
        entry = [-1, -1]
    n = len(entry)
    if n != 2:
        entry = [-n]
    for j in range(len(entry)):
        if not isinstance(entry[j], int):
            entry[j] = -1
    return entry


def moment_str_to_list(m: str) -> list:
    """utils/utils.py:364-415."""
    if m == "[[-1, -1]]" or _NESTED.match(m) is None:
        return [[-1, -1]]
    if _CANONICAL.fullmatch(m) is not None:      # fast path (about 5x faster than literal_eval), same result
        nums = [int(x) for x in _UINT.findall(m)]
        return [[nums[k], nums[k + 1]] for k in range(0, len(nums), 2)]
    try:
        parsed = ast.literal_eval(m)
    except Exception:   # the reference uses a bare except; KeyboardInterrupt / SystemExit cannot come out of literal_eval
        return [[-1, -1]]
    if not isinstance(parsed, list):
        return [[-1, -1]]
    for i in range(len(parsed)):
        parsed[i] = _sanitise(parsed[i])
    return parsed


def parse_output(raw_out: str) -> list:
    """``moment_str_to_list(post_process(raw_out))`` -- evaluate.py:48, utils/trainer.py:168-169."""
    return moment_str_to_list(post_process(raw_out))


def prediction_record(qid, query, vid, raw_out: str) -> dict:
    """One line of the submission jsonl, keys and order as evaluate.py:50-56."""
    return {"qid": qid, "query": query, "vid": vid, "pred_relevant_windows": parse_output(raw_out), "raw_out": raw_out}


def write_jsonl(records: Iterable[dict], path: str) -> None:
    """evaluate.py:37-60: one ``json.dumps`` per line."""
    with open(path, "w") as f:
        for r in records:
            f.write(json.dumps(r) + "\n")


def windows_to_arrays(window_lists: Sequence[list]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Ragged per-query window lists -> ``(windows f64 [Q, Pmax, 2], counts int32 [Q], malformed bool [Q])``, the padded
    layout ``mra_mr_score`` reads.  A window that is not a pair (the reference's 1-element ``[-len]`` marker,
    utils/utils.py:409-410, which makes ``np.array`` ragged in eval/mr_eval.py:117) is stored as ``[-1, -1]`` and the
    query is flagged in ``malformed``."""
    Q = len(window_lists)
    counts = np.fromiter((len(w) for w in window_lists), dtype=np.int32, count=Q)
    pmax = int(counts.max()) if Q else 1
    out = np.zeros((Q, max(pmax, 1), 2), dtype=np.float64)
    malformed = np.zeros(Q, dtype=bool)
    for i, wins in enumerate(window_lists):
        for j, w in enumerate(wins):
            if len(w) >= 2:
                out[i, j, 0], out[i, j, 1] = w[0], w[1]
            else:
                out[i, j] = -1.0
                malformed[i] = True
    return out, counts, malformed


def parse_batch(raw_outs: Sequence[str]) -> Tuple[List[list], np.ndarray, np.ndarray, np.ndarray]:
    """A whole sweep of generations -> (window lists, padded windows, counts, malformed flags)."""
    lists = [parse_output(s) for s in raw_outs]
    return (lists,) + windows_to_arrays(lists)
