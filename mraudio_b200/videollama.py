"""VideoLLaMA call surface + the Video-LLaMA-v1-style Q-Former branches named by BASELINE.json (config 3).

* ``VideoLLaMA`` keeps the reference wrapper's surface (``models/videollama.py:3-25``): ``__init__(path)`` with
  ``.model / .processor / .tokenizer`` from ``videollama2.model_init`` and ``generate(samples)`` calling
  ``mm_infer(samples["video"][0], samples["text_input"][0], modal='video', do_sample=False)`` with every exception
  turned into the string ``"error"``.  ``videollama2`` (LLM + STC connector) is out of scope and not installed here;
  the class imports it lazily so that the wrapper itself is importable.
* ``VideoLLaMAQFormers`` is the hot path BASELINE.json describes: a video Q-Former over frame-position-embedded frame
  tokens (keys ``[B, F*32, 768]``, 2 layers, cross-attention in every layer, 32 queries) and an ImageBind-audio Q-Former
  (keys ``[B, 8, 1024]`` + position embedding, 8 queries), each followed by a ``Linear(768, 4096)``.  None of this
  arithmetic exists in the reference or its pinned dependencies: parity is pinned only to the HF
  ``Blip2QFormerModel`` port through the oracle ("unpinned" with respect to the reference).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib, ops
from .qformer import BertConfig, BertLMHeadModel, LLMProjB200


class VideoLLaMA:
    def __init__(self, path):
        from videollama2 import model_init  # noqa: F401  (absent offline: raises ModuleNotFoundError like the reference)
        self.model, self.processor, self.tokenizer = model_init(path)

    def generate(self, samples):
        try:
            from videollama2 import mm_infer
            output = mm_infer(samples["video"][0], samples["text_input"][0], model=self.model, tokenizer=self.tokenizer,
                              modal='video', do_sample=False)
        except Exception:   # models/videollama.py:21-23 swallows everything
            output = "error"
        return output


def _query_only_qformer(num_query_token, width, num_hidden_layers=2):
    cfg = BertConfig.from_pretrained("bert-base-uncased")
    cfg.encoder_width = width
    cfg.add_cross_attention = True
    cfg.cross_attention_freq = 1
    cfg.query_length = num_query_token
    cfg.num_hidden_layers = num_hidden_layers
    q = BertLMHeadModel(cfg)
    q.cls = None
    query_tokens = nn.Parameter(torch.zeros(1, num_query_token, cfg.hidden_size))
    query_tokens.data.normal_(mean=0.0, std=cfg.initializer_range)
    return q, query_tokens


class VideoLLaMAQFormers(nn.Module):
    def __init__(self, max_frame_pos=32, num_video_query_token=32, frame_width=768, audio_width=1024, num_audio_query_token=8,
                 max_audio_pos=8, llm_hidden_size=4096, num_hidden_layers=2):
        super().__init__()
        self.video_frame_position_embedding = nn.Embedding(max_frame_pos, frame_width)
        self.video_Qformer, self.video_query_tokens = _query_only_qformer(num_video_query_token, frame_width, num_hidden_layers)
        self.llama_proj = LLMProjB200(self.video_Qformer.config.hidden_size, llm_hidden_size)
        self.audio_position_embedding = nn.Embedding(max_audio_pos, audio_width)
        self.audio_Qformer, self.audio_query_tokens = _query_only_qformer(num_audio_query_token, audio_width, num_hidden_layers)
        self.audio_llama_proj = LLMProjB200(self.audio_Qformer.config.hidden_size, llm_hidden_size)
        self.last_launches = 0

    def _encode(self, qformer, query_tokens, proj, pos_emb, tokens):
        if not tokens.is_cuda:
            raise _lib.MraError("VideoLLaMAQFormers needs CUDA tensors (no CPU fallback)")
        keys = ops.add_frame_position(tokens.contiguous(), pos_emb.weight.detach().float().contiguous())
        out = qformer.bert(query_embeds=query_tokens, encoder_hidden_states=keys, return_dict=True, llm_proj=proj,
                           need_last_hidden=False)
        self.last_launches = qformer.bert.last_launches + 1
        return out.llm_inputs

    def encode_videoQformer(self, frame_tokens: torch.Tensor) -> torch.Tensor:
        """frame_tokens [B, F, 32, 768] (per-frame Q-Former outputs of the frozen image branch) -> [B, 32, 4096]."""
        return self._encode(self.video_Qformer, self.video_query_tokens, self.llama_proj,
                            self.video_frame_position_embedding, frame_tokens)

    def encode_audioQformer(self, audio_embeds: torch.Tensor) -> torch.Tensor:
        """audio_embeds [B, 8, 1024] (ImageBind clip embeddings) -> [B, 8, 4096]."""
        return self._encode(self.audio_Qformer, self.audio_query_tokens, self.audio_llama_proj,
                            self.audio_position_embedding, audio_embeds.unsqueeze(2))
