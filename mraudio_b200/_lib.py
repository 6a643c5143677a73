"""ctypes binding of libmraudio_b200.so (C-ABI declared in include/mraudio_b200.h).

There is no CPU fallback: the first call into the library without the built ``.so`` raises ImportError, and every compute
entry point returns an error on a machine without an sm_100 GPU (``MraError``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRA_LIB=<name> loads libmraudio_b200_<name>.so: the instrumented build (make -C mraudio_b200/csrc INSTRUMENT=1 -> "instr")
# or an A/B variant (make VARIANT=<name> EXTRA=-D...).  Tuning tools only; the product loads libmraudio_b200.so.
_VARIANT = os.environ.get("MRA_LIB", "")
LIB_PATH = os.path.join(_HERE, f"libmraudio_b200_{_VARIANT}.so" if _VARIANT else "libmraudio_b200.so")

MRA_MAX_LAYERS = 16
MRA_NUM_IOU_THDS = 10
FWD_SKIP_DEAD_TEXT_FFN = 1
FWD_SAVE_FOR_BACKWARD = 2
GEMM_IMPL_TCGEN05 = 0
GEMM_IMPL_SIMT_DEBUG = 1
PROFILE_OFF, PROFILE_DOMINANT, PROFILE_ALL = 0, 1, 2
PROFILE_CATS = ("gemm_cross_kv", "gemm", "attention", "layernorm", "other")


class MraError(RuntimeError):
    pass


class QFormerConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden", "layers", "heads", "inter", "enc_width", "cross_freq", "num_query",
                                         "llm_dim", "vocab", "max_pos")] + [("ln_eps", C.c_float)]


_LAYER_FIELDS = ("w_qkv", "b_qkv", "w_ao", "b_ao", "ln_a_g", "ln_a_b", "w_cq", "b_cq", "w_co", "b_co", "ln_c_g", "ln_c_b",
                 "w_fq1", "b_fq1", "w_fq2", "b_fq2", "ln_fq_g", "ln_fq_b", "w_ft1", "b_ft1", "w_ft2", "b_ft2", "ln_ft_g",
                 "ln_ft_b")


class QFormerLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class QFormerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("word_emb", "pos_emb", "ln_e_g", "ln_e_b", "w_ckv", "b_ckv", "w_proj", "b_proj")] + \
               [("layer", QFormerLayerWeights * MRA_MAX_LAYERS)]


class QFormerLayerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class QFormerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("word_emb", "pos_emb", "ln_e_g", "ln_e_b", "w_ckv", "b_ckv", "w_proj", "b_proj",
                                          "query_tokens")] + [("layer", QFormerLayerGrads * MRA_MAX_LAYERS)]


class QFormerIO(C.Structure):
    _fields_ = [("enc", C.c_void_p), ("input_ids", C.c_void_p), ("attn_mask", C.c_void_p), ("enc_mask", C.c_void_p),
                ("query_embeds", C.c_void_p), ("q_rows", C.c_int32), ("rows", C.c_int32), ("T", C.c_int32),
                ("Nk", C.c_int32), ("flags", C.c_uint32), ("last_hidden", C.c_void_p), ("llm_out", C.c_void_p),
                ("llm_frames", C.c_int32), ("reserved0", C.c_int32), ("llm_ld", C.c_int64), ("llm_frame_stride", C.c_int64),
                ("llm_video_stride", C.c_int64), ("dropout_p", C.c_float), ("reserved1", C.c_uint32), ("dropout_seed", C.c_uint64),
                ("enc_ready", C.c_void_p)]


MAX_PROMPT_SEGMENTS = 16


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`make -C mraudio_b200/csrc`.  mraudio_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.mra_last_error.restype = C.c_char_p
    lib.mra_last_error.argtypes = []
    lib.mra_version.restype = C.c_int
    lib.mra_device_check.restype = C.c_int
    lib.mra_qformer_create.argtypes = [C.POINTER(QFormerConfig), C.POINTER(vp)]
    lib.mra_qformer_set_weights.argtypes = [vp, C.POINTER(QFormerWeights)]
    lib.mra_qformer_destroy.argtypes = [vp]
    lib.mra_qformer_destroy.restype = None
    lib.mra_qformer_workspace_bytes.argtypes = [vp, i32, i32, i32, C.c_uint32]
    lib.mra_qformer_workspace_bytes.restype = C.c_size_t
    lib.mra_qformer_forward.argtypes = [vp, C.POINTER(QFormerIO), vp, C.c_size_t, vp]
    lib.mra_qformer_forward_multi.argtypes = [i32, C.POINTER(vp), C.POINTER(C.POINTER(QFormerIO)), C.POINTER(vp),
                                              C.POINTER(C.c_size_t), vp]
    lib.mra_qformer_last_launch_count.argtypes = [vp]
    lib.mra_qformer_backward_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.mra_qformer_backward_workspace_bytes.restype = C.c_size_t
    lib.mra_qformer_backward.argtypes = [vp, C.POINTER(QFormerIO), vp, vp, C.POINTER(QFormerGrads), vp,
                                         C.c_size_t, vp, C.c_size_t, vp]
    lib.mra_qformer_backward_layer_events.argtypes = [vp, C.POINTER(vp), i32]
    lib.mra_adam_step.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp]
    lib.mra_adam_step_fused.argtypes = [vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, i32, vp]
    lib.mra_adam_step_fused_dyn.argtypes = [vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp, vp]
    lib.mra_adam_hyper.argtypes = [f32, f32, f32, i32, f32, C.POINTER(C.c_float)]
    lib.mra_adam_hyper.restype = None
    lib.mra_cast_bf16.argtypes = [vp, vp, i64, vp]
    lib.mra_qformer_profile_mode.argtypes = [vp, i32]
    lib.mra_qformer_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.mra_gemm_bf16.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, vp]
    lib.mra_wgrad_bf16.argtypes = [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp]
    lib.mra_dgrad_bf16.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, vp]
    lib.mra_gemm_ln_bf16.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, vp, vp, i64, vp, i64, i32, i32, i32, f32, vp]
    lib.mra_gemm_ln_split_bf16.argtypes = [vp, i64, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, i64, i32, i32, i32, f32, vp]
    lib.mra_gemm_tile_override.argtypes = [i32]
    lib.mra_gemm_cluster_override.argtypes = [i32]
    lib.mra_attention.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.mra_attention_strided.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.mra_gemm_head_major_bf16.argtypes = [vp, i64, vp, i64, vp, vp, i32, i32, i32, vp]
    lib.mra_qkv_attention_bf16.argtypes = [vp, i64, vp, i64, vp, vp, vp, i64, i32, i32, i32, vp]
    lib.mra_attention_impl_override.argtypes = [i32]
    lib.mra_layernorm.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, vp]
    lib.mra_modality_layernorm.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]
    lib.mra_add_frame_position.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, vp]
    lib.mra_prompt_assemble.argtypes = [vp, i32, i32, i32, vp, i32, vp]
    lib.mra_mr_score.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    return lib


class _LazyLib:
    """Loads libmraudio_b200.so on first use, so that the host-only modules (parsing, prompt layout, record packing) import
    on a checkout where the library has not been built yet; any compute call without it raises the ImportError of
    ``_load`` -- there is still no CPU fallback."""
    _real = None

    def _get(self):
        if _LazyLib._real is None:
            _LazyLib._real = _load()
        return _LazyLib._real

    def __getattr__(self, name):
        return getattr(self._get(), name)


lib = _LazyLib()

# every symbol include/mraudio_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = (
    "mra_last_error", "mra_version", "mra_device_check", "mra_qformer_create", "mra_qformer_set_weights",
    "mra_qformer_destroy", "mra_qformer_workspace_bytes", "mra_qformer_forward", "mra_qformer_forward_multi", "mra_qformer_last_launch_count", "mra_qformer_backward_workspace_bytes", "mra_qformer_backward", "mra_qformer_backward_layer_events", "mra_adam_step", "mra_adam_step_fused", "mra_adam_step_fused_dyn", "mra_adam_hyper",
    "mra_cast_bf16",
    "mra_qformer_profile_mode", "mra_qformer_profile_read",
    "mra_gemm_bf16", "mra_wgrad_bf16", "mra_dgrad_bf16", "mra_gemm_ln_bf16", "mra_gemm_ln_split_bf16", "mra_gemm_tile_override", "mra_gemm_cluster_override", "mra_attention", "mra_attention_strided", "mra_gemm_head_major_bf16", "mra_qkv_attention_bf16", "mra_attention_impl_override", "mra_layernorm", "mra_modality_layernorm", "mra_add_frame_position", "mra_prompt_assemble", "mra_mr_score",
)


def check(rc: int) -> None:
    if rc != 0:
        raise MraError(lib.mra_last_error().decode("utf-8", "replace"))


def ptr(t) -> int:
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def bind_to_gpu_cpus(device_index: int):
    """Bind the calling process to the CPUs NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe root), so
    that pinned host staging buffers allocated afterwards are first-touched on that node.  Host-side plumbing for the
    end-to-end path (``HostPipeline``) on multi-socket boxes with one process per GPU; returns a short description or
    None when NVML / affinity control is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device_index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        try:
            node = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            node = None
        return {"gpu": idx, "cpus": len(cpus), "numa_node": node}
    except Exception:
        return None
