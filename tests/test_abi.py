"""The C-ABI library: loads on a CPU-only box, exports every symbol include/easyhec_b200.h declares,
and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "easyhec_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"EHB_API\s+[\w\s\*]+?\b(ehb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from easyhec_b200 import _lib
    if not os.path.exists(_lib.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    names = declared_symbols()
    assert len(names) >= 20 and "ehb_render_views_fused" in names and "ehb_render_mask_fwd" in names
    dll = ctypes.CDLL(_lib.library_path())
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing
    # and the Python binding knows the prototype of every one of them
    assert sorted(_lib._PROTOS) == names


def test_version_and_error_string():
    from easyhec_b200 import _lib
    l = _lib.lib()
    assert l.ehb_version() >= 100
    assert l.ehb_ctx_destroy(None) == 0
    assert l.ehb_launch_count(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from easyhec_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.lib().ehb_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in _lib.lib().ehb_last_error()
    with pytest.raises(_lib.EhbError):
        _lib.Context()
    from easyhec_b200.renderer import B200Renderer
    with pytest.raises(_lib.EhbError):
        B200Renderer([48, 64])


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under easyhec_b200/ may import or load it."""
    pkg = os.path.join(ROOT, "easyhec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "libehb_oracle" not in src and "ehb_oracle.c" not in src, f
