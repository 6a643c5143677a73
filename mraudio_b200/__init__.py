"""mraudio_b200 -- B200-native (sm_100a) Q-Former / llm_proj / moment-scoring hot path of globc/mrAudio.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the arithmetic runs in hand-written CUDA
kernels behind the C-ABI of ``include/mraudio_b200.h`` (``libmraudio_b200.so``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401  (the library is loaded on first use and fails loudly if it is not built)
from ._lib import MraError  # noqa: F401
from .qformer import BertConfig, BertLMHeadModel, BertModelB200, LLMProjB200, QFormerOutput  # noqa: F401

__all__ = ["MraError", "BertConfig", "BertLMHeadModel", "BertModelB200", "LLMProjB200", "QFormerOutput"]
