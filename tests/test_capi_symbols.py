"""CPU gate: the C-ABI library builds, loads without a GPU, exports every symbol include/b2l.h
declares, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gabby_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_cuda()
    return _capi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2l.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2l_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 18
    assert sorted(_capi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_struct_layouts_match_header():
    # b2l_params: 8 int32 + float + 8 int32; b2l_info ends with char[64]
    assert C.sizeof(_capi.B2lParams) == 17 * 4
    # 4 int32 + 5 int64 + 3 int32 (decode_mode, batched_tensor_core, tp_transport) + char[64], padded to 8
    assert C.sizeof(_capi.B2lInfo) == (4 * 4 + 5 * 8 + 3 * 4 + 64 + 7) // 8 * 8


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gabby_b200 import synth
    arch = synth.preset("tiny")
    with pytest.raises(_capi.B2lError, match="no CUDA device"):
        _capi.Engine(arch, np.zeros((64, arch.head_dim // 2, 2), np.float32), max_positions=64)
    with pytest.raises(_capi.B2lError, match="no CPU fallback"):
        _capi.op_gemv(np.zeros((16, 256), np.uint16), np.zeros((1, 256), np.float32))


def test_product_package_never_imports_oracle():
    """The product path may not route through the oracle (grep the package sources)."""
    pkg = os.path.join(ROOT, "gabby_b200")
    bad = []
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|llama_oracle\.h|liboracle|orc_[a-z_]+\(", txt):
                    bad.append(fn)
    assert bad == []
