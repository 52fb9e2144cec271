"""The C-ABI library builds, loads on a CPU-only box and exports exactly what include/sfb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from self_forcing_b200 import build
    return build.build()


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert "sfb_gemm_bf16" in names and "sfb_attention_fwd" in names and len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sfb200.h but not exported"


def test_binding_covers_header(lib_path):
    from self_forcing_b200 import _lib
    declared = set(_declared()) - {"sfb_last_error", "sfb_abi_version", "sfb_attention_workspace_bytes", "sfb_gemm_workspace_bytes",
                                   "sfb_causal_conv3d_workspace_bytes"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load(lib_path)
    assert lib.sfb_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define SFB_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "sfb200.h")).read()).group(1))
    assert isinstance(lib.sfb_last_error(), bytes)


def test_graft_entry_build_runs(lib_path):
    """The driver's build check: __graft_entry__.build() compiles (or finds) the library and verifies the ABI version."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    entry = importlib.import_module("__graft_entry__")
    entry.build()


def test_argument_errors_are_reported_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is observable on a CPU box."""
    from self_forcing_b200 import _lib
    lib = _lib.load(lib_path)
    rc = lib.sfb_attention_fwd(None, 0, 0, None, None, 0, 0, None, 0, 0, 1, 16, 16, 1, 64, 1.0, None, 0, None)
    assert rc != 0 and b"head_dim" in lib.sfb_last_error()
    rc = lib.sfb_gemm_bf16(None, 8, None, 8, None, 4, 10, 8, 0, None, 8, None, 0, None, 0, 0, None, 0, None, 0, 1, 0, 0, None, 0, None)
    assert rc != 0 and b"multiples of 8" in lib.sfb_last_error()
    with pytest.raises(_lib.SfbError):
        _lib.check(rc, "sfb_gemm_bf16")
    rc = lib.sfb_causal_conv3d_cl(1, 1, 4, 4, 16, 0, 1, 1, None, 8, 1, 3, None, 0, 1, None, 8, 0, None, 0, None)
    assert rc != 0 and b"without a workspace" in lib.sfb_last_error()       # implicit path: no upsampling
    rc = lib.sfb_cfg_unipc_step(1, None, 1, None, None, None, 1, 1, 1, 8, (ctypes.c_float * 12)(), 2, 2, None)
    assert rc != 0 and b"history tensors missing" in lib.sfb_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from self_forcing_b200 import _lib
    from self_forcing_b200.ops import CudaOps
    with pytest.raises(_lib.SfbError):
        CudaOps()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "self_forcing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "_torch_ops" not in src, f
