"""BASELINE.json config 3 (Video-LLaMA-v1-style): frame-position-embedded video Q-Former (1024 keys) and ImageBind-audio
Q-Former (8 keys of width 1024, 8 queries) against the oracle.  Parity is unpinned w.r.t. the reference (its
models/videollama.py wraps VideoLLaMA2 and contains none of this arithmetic); the oracle is pinned to HF Blip2QFormerModel."""
import pytest
import torch

from oracle import qformer_oracle as qo

pytestmark = pytest.mark.gpu


def _load(qformer, qtok, proj, w):
    # query-only Q-Former: the (unused) word / position embedding tables keep the module's own shapes
    msg = qformer.load_state_dict({k: v for k, v in w.items() if k.startswith("bert.") and "embeddings.word" not in k
                                   and "embeddings.position" not in k}, strict=False)
    assert all("embeddings." in k for k in msg.missing_keys), msg.missing_keys   # only the unused embedding tables
    assert not msg.unexpected_keys, msg.unexpected_keys
    proj.load_state_dict({"weight": w["llm_proj.weight"], "bias": w["llm_proj.bias"]})
    qtok.data.copy_(w["query_tokens"])


@pytest.mark.parametrize("B,F", [(3, 32), (2, 5), (64, 32)])   # (64, 32) = BASELINE.json config 3 at its full size
def test_video_and_audio_qformers(B, F):
    from mraudio_b200.videollama import VideoLLaMAQFormers
    m = VideoLLaMAQFormers(llm_hidden_size=512)
    vcfg = qo.QFormerOracleConfig(encoder_width=768, num_hidden_layers=2, cross_attention_freq=1, has_text=False)
    acfg = qo.QFormerOracleConfig(encoder_width=1024, num_hidden_layers=2, cross_attention_freq=1, has_text=False, query_length=8)
    vw = qo.init_qformer_weights(vcfg, seed=11, llm_dim=512, randomize_ln_and_bias=True)
    aw = qo.init_qformer_weights(acfg, seed=12, llm_dim=512, randomize_ln_and_bias=True)
    _load(m.video_Qformer, m.video_query_tokens, m.llama_proj, vw)
    _load(m.audio_Qformer, m.audio_query_tokens, m.audio_llama_proj, aw)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(5)
    frames = torch.randn(B, F, 32, 768, generator=g)
    audio = torch.randn(B, 8, 1024, generator=g)
    with torch.no_grad():
        yv = m.encode_videoQformer(frames.cuda())
        ya = m.encode_audioQformer(audio.cuda())
        rv = qo.videollama_v1_encode(vw, vcfg, m.video_frame_position_embedding.weight.detach().cpu(), frames)
        ra = qo.videollama_v1_encode(aw, acfg, m.audio_position_embedding.weight.detach().cpu(), audio.unsqueeze(2))
    assert yv.shape == (B, 32, 512) and ya.shape == (B, 8, 512)
    assert ((yv.float().cpu() - rv).abs().max() / rv.abs().max()).item() < 2e-2
    assert ((ya.float().cpu() - ra).abs().max() / ra.abs().max()).item() < 2e-2


def test_wrapper_surface_matches_reference():
    from mraudio_b200.videollama import VideoLLaMA
    with pytest.raises(ModuleNotFoundError):
        VideoLLaMA("some/path")            # videollama2 is not installed offline, exactly like importing the reference
    obj = VideoLLaMA.__new__(VideoLLaMA)
    obj.model = obj.tokenizer = obj.processor = None
    assert obj.generate({"video": [None], "text_input": ["q"]}) == "error"   # models/videollama.py:21-23
