"""Generate the golden fixtures under tests/golden/ from the real reference code.  Run in the BUILD container only:

    python tests/golden/make_golden.py

* ``mr_eval_*.json``  <- the reference's own scorer, imported from /root/reference (``eval.mr_eval.eval_submission``,
  ``eval.mr_utils.compute_average_precision_detection``), on seeded synthetic submissions from
  ``oracle.mr_eval_oracle.synth_submission`` plus hand-written edge cases.
* ``qformer_*.npz``   <- the HuggingFace port of the LAVIS Q-Former arithmetic (the only importable copy of the
  reference's third-party dependency in this image: transformers InstructBlipQFormerModel / Blip2QFormerModel) with
  weights from ``oracle.qformer_oracle.init_qformer_weights`` (seeded), frozen outputs on seeded inputs.

/root/reference does not exist on the GPU box; tests only read the committed fixtures.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle.mr_eval_oracle import synth_submission  # noqa: E402
from oracle import qformer_oracle as qo  # noqa: E402


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, np.ndarray):
        return _jsonable(o.tolist())
    return o


def make_scorer_golden():
    from eval.mr_eval import eval_submission  # the reference itself
    from eval.mr_utils import compute_average_precision_detection, compute_temporal_iou_batch_cross

    cases = {}
    cases["int_300"] = synth_submission(300, seed=7)
    cases["int_ties_200"] = synth_submission(200, seed=11, duration=20, clip_len=2, max_pred=8, max_gt=6)
    cases["float_150"] = synth_submission(150, seed=3, float_windows=True)
    cases["many_invalid_64"] = synth_submission(64, seed=5, invalid_frac=0.5)
    # 17..64 ground-truth windows per query (beyond numpy's insertion-sort range of 16 elements): float windows, so equal
    # IoUs come only from duplicated windows; and the integer grid, where DIFFERENT windows can tie exactly -- there the
    # reference's result depends on the tie order of this numpy build's (AVX-512) argsort, see oracle/mr_eval_oracle.py
    cases["many_gt_float_120"] = synth_submission(120, seed=13, max_pred=10, max_gt=64, float_windows=True)
    cases["many_gt_int_120"] = synth_submission(120, seed=17, duration=150, max_pred=10, max_gt=40)
    # SURVEY.md 8(a) known-answer vectors
    cases["kat_two_queries"] = (
        [{"qid": "a", "pred_relevant_windows": [[10, 20], [30, 40]]}, {"qid": "b", "pred_relevant_windows": [[-1, -1]]}],
        [{"qid": "a", "relevant_windows": [[12, 20], [30, 41]]}, {"qid": "b", "relevant_windows": [[10, 20]]}],
    )
    # nan IoU (pred [-1,-1] vs zero-length GT at the same point; 0/0), scores in a third column, duplicate preds
    cases["edge"] = (
        [{"qid": 0, "pred_relevant_windows": [[-1, -1]]},
         {"qid": 1, "pred_relevant_windows": [[0, 0], [0, 10]]},
         {"qid": 2, "pred_relevant_windows": [[4, 8, 0.9], [4, 8, 0.8], [4, 8, 0.1]]},
         {"qid": 3, "pred_relevant_windows": [[0, 150]]},
         {"qid": 4, "pred_relevant_windows": [[10, 12], [10, 12], [12, 14], [0, 2], [2, 4], [4, 6], [6, 8], [8, 10]]}],
        [{"qid": 0, "relevant_windows": [[-1, -1], [3, 9]]},
         {"qid": 1, "relevant_windows": [[0, 0], [0, 10]]},
         {"qid": 2, "relevant_windows": [[4, 8], [4, 8]]},
         {"qid": 3, "relevant_windows": [[0, 150], [10, 20], [30, 60]]},
         {"qid": 4, "relevant_windows": [[10, 12], [12, 14], [0, 8]]}],
    )
    for name, (sub, gt) in cases.items():
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            res = eval_submission(sub, gt, verbose=False)
            per_query_ap = []
            for s, g in zip(sub, gt):
                gts = [{"video-id": g["qid"], "t-start": w[0], "t-end": w[1]} for w in g["relevant_windows"]]
                prs = [{"video-id": s["qid"], "t-start": w[0], "t-end": w[1]} for w in s["pred_relevant_windows"]]
                thds = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]
                per_query_ap.append(compute_average_precision_detection(gts, prs, tiou_thresholds=thds))
        out = {"submission": sub, "ground_truth": gt, "eval_submission": _jsonable(res),
               "per_query_ap": _jsonable(per_query_ap)}
        with open(os.path.join(HERE, f"mr_eval_{name}.json"), "w") as f:
            json.dump(out, f)
        print("wrote", name, res["brief"])
    # docstring example eval/mr_utils.py:49-55
    iou, union = compute_temporal_iou_batch_cross(np.array([[0, 0.2, 0.9], [0.5, 1.0, 0.2]]),
                                                  np.array([[0, 0.3], [0.0, 1.0]]))
    with open(os.path.join(HERE, "mr_eval_docstring_iou.json"), "w") as f:
        json.dump({"iou": iou.tolist(), "union": union.tolist()}, f)


def _hf_state_dict(w):
    sd = {}
    for k, v in w.items():
        if not k.startswith("bert."):
            continue
        k2 = (k[5:].replace("crossattention.self.", "crossattention.attention.")
              .replace("attention.self.", "attention.attention.")
              .replace("embeddings.LayerNorm", "embeddings.layernorm"))
        sd[k2] = v
    return sd


def make_qformer_golden():
    from transformers.models.instructblip.modeling_instructblip import (InstructBlipQFormerConfig,
                                                                        InstructBlipQFormerModel)
    from transformers.models.blip_2.modeling_blip_2 import Blip2QFormerConfig, Blip2QFormerModel

    # (1) text+query Q-Former, X-InstructBLIP video shape, full depth, 2 rows, one padded prompt
    for name, W, Nk, layers in (("video", 1408, 257, 12), ("audio", 768, 256, 12)):
        cfg = qo.QFormerOracleConfig(encoder_width=W, num_hidden_layers=layers)
        w = qo.init_qformer_weights(cfg, seed=1234, randomize_ln_and_bias=True)
        hcfg = InstructBlipQFormerConfig(vocab_size=cfg.vocab_size, encoder_hidden_size=W,
                                         cross_attention_frequency=2, num_hidden_layers=layers)
        m = InstructBlipQFormerModel(hcfg).eval()
        missing, unexpected = m.load_state_dict(_hf_state_dict(w), strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        g = torch.Generator().manual_seed(99)
        rows, T = 2, 32
        ids = torch.randint(1000, 30000, (rows, T), generator=g)
        tmask = torch.ones(rows, T, dtype=torch.long)
        tmask[1, 20:] = 0
        enc = torch.randn(rows, Nk, W, generator=g)
        atts = torch.cat([torch.ones(rows, 32, dtype=torch.long), tmask], 1)
        qe = w["query_tokens"].expand(rows, -1, -1)
        with torch.no_grad():
            hid = m(ids, attention_mask=atts, query_embeds=qe, encoder_hidden_states=enc,
                    encoder_attention_mask=torch.ones(rows, Nk, dtype=torch.long), return_dict=True).last_hidden_state
            proj = torch.nn.functional.linear(hid[:, :32], w["llm_proj.weight"], w["llm_proj.bias"])
        np.savez_compressed(os.path.join(HERE, f"qformer_{name}.npz"), weight_seed=1234, input_seed=99,
                            input_ids=ids.numpy(), text_mask=tmask.numpy(),
                            last_hidden_state=hid.numpy().astype(np.float32),
                            llm_proj_sample=proj[:, :, ::64].numpy().astype(np.float32))
        print("wrote qformer", name, tuple(hid.shape))

    # (2) query-only 2-layer Q-Former (Video-LLaMA-v1 style), cross-attention every layer, 1024 keys of width 768
    cfg = qo.QFormerOracleConfig(encoder_width=768, num_hidden_layers=2, cross_attention_freq=1, has_text=False)
    w = qo.init_qformer_weights(cfg, seed=4321, randomize_ln_and_bias=True)
    hcfg = Blip2QFormerConfig(num_hidden_layers=2, cross_attention_frequency=1, encoder_hidden_size=768,
                              vocab_size=cfg.vocab_size)
    m = Blip2QFormerModel(hcfg).eval()
    sd = _hf_state_dict(w)
    # Blip2QFormerModel applies its embedding LayerNorm as ``layernorm`` at model level and has no word embeddings in use
    own = m.state_dict()
    remap = {}
    for k, v in sd.items():
        if k in own:
            remap[k] = v
        elif k == "embeddings.layernorm.weight" and "layernorm.weight" in own:
            remap["layernorm.weight"] = v
        elif k == "embeddings.layernorm.bias" and "layernorm.bias" in own:
            remap["layernorm.bias"] = v
    missing, unexpected = m.load_state_dict(remap, strict=False)
    missing = [k for k in missing if "embeddings.word" not in k and "embeddings.position" not in k
               and ".intermediate.dense" not in k and ".output.dense" not in k and ".output.LayerNorm" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    g = torch.Generator().manual_seed(98)
    enc = torch.randn(2, 1024, 768, generator=g)
    qe = w["query_tokens"].expand(2, -1, -1)
    with torch.no_grad():
        hid = m(query_embeds=qe, encoder_hidden_states=enc, return_dict=True).last_hidden_state
    np.savez_compressed(os.path.join(HERE, "qformer_queryonly.npz"), weight_seed=4321, input_seed=98,
                        last_hidden_state=hid.numpy().astype(np.float32))
    print("wrote qformer queryonly", tuple(hid.shape))


if __name__ == "__main__":
    make_scorer_golden()
    make_qformer_golden()
