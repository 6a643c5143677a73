"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mraudio_b200.h
declares, and the compute entry points fail loudly (no CPU fallback) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mraudio_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mra_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mraudio_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/mraudio_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(names)
    assert _lib.lib.mra_version() >= 100


def test_struct_layouts_match_header():
    from mraudio_b200 import _lib
    assert ctypes.sizeof(_lib.QFormerConfig) == 44
    assert ctypes.sizeof(_lib.QFormerLayerWeights) == 24 * 8
    assert ctypes.sizeof(_lib.QFormerWeights) == 8 * 8 + _lib.MRA_MAX_LAYERS * 24 * 8


def test_ctypes_mirror_matches_the_header_compiled_by_gcc(tmp_path):
    """The binding a reference maintainer would write is a ctypes mirror of include/mraudio_b200.h: compile the header as
    plain C (it must stay C: no torch / CUDA types in the boundary) and compare struct sizes and the offsets of the io
    struct's fields with mraudio_b200/_lib.py."""
    import shutil
    import subprocess
    from mraudio_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    io_fields = [n for n, _ in _lib.QFormerIO._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {',
            '  printf("cfg %zu\\n", sizeof(mra_qformer_config));', '  printf("layer %zu\\n", sizeof(mra_qformer_layer_weights));',
            '  printf("weights %zu\\n", sizeof(mra_qformer_weights));', '  printf("grads %zu\\n", sizeof(mra_qformer_grads));',
            '  printf("io %zu\\n", sizeof(mra_qformer_io));']
    prog += [f'  printf("io.{n} %zu\\n", offsetof(mra_qformer_io, {n}));' for n in io_fields]
    prog += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out["cfg"]) == ctypes.sizeof(_lib.QFormerConfig)
    assert int(out["layer"]) == ctypes.sizeof(_lib.QFormerLayerWeights)
    assert int(out["weights"]) == ctypes.sizeof(_lib.QFormerWeights)
    assert int(out["grads"]) == ctypes.sizeof(_lib.QFormerGrads)
    assert int(out["io"]) == ctypes.sizeof(_lib.QFormerIO)
    for n in io_fields:
        assert int(out[f"io.{n}"]) == getattr(_lib.QFormerIO, n).offset, n


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback():
    from mraudio_b200 import _lib, ops, mr_eval
    assert _lib.lib.mra_device_check() != 0
    assert b"no CPU fallback" in _lib.lib.mra_last_error() or b"CUDA" in _lib.lib.mra_last_error()
    with pytest.raises(_lib.MraError):
        ops.linear(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(_lib.MraError):
        mr_eval.score_records([{"qid": 0, "pred_relevant_windows": [[0, 2]]}], [{"qid": 0, "relevant_windows": [[0, 2]]}])


def test_handle_argument_validation():
    from mraudio_b200 import _lib
    cfg = _lib.QFormerConfig(hidden=768, layers=99, heads=12, inter=3072, enc_width=1408, cross_freq=2, num_query=32,
                             llm_dim=4096, vocab=30523, max_pos=512, ln_eps=1e-12)
    h = ctypes.c_void_p()
    assert _lib.lib.mra_qformer_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"layers" in _lib.lib.mra_last_error()
    cfg.layers = 12
    assert _lib.lib.mra_qformer_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    assert _lib.lib.mra_qformer_workspace_bytes(h, 8, 32, 257, 0) > 0
    _lib.lib.mra_qformer_destroy(h)
