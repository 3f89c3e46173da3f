"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "peclr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(peclr_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from peclr_b200.build import build

    return build()


def test_header_declares_functions():
    names = header_functions()
    assert len(names) >= 25 and "peclr_ntxent_fused" in names and "peclr_conv2d_fprop" in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.peclr_abi_version() == 3


def test_python_binding_matches_header(lib_path):
    from peclr_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_functions()
    src = open(os.path.join(ROOT, "include", "peclr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(argtypes), (name, n, len(argtypes))
    _lib.load()


def test_workspace_queries_run_without_a_gpu(lib_path):
    """The *_workspace_bytes entry points are host arithmetic (no launch): callable on the CPU-only build box."""
    from peclr_b200 import _lib

    # 56x56 64->64 3x3 at 2B = 256: few output tiles -> many pixel splits -> a workspace; 1 pixel chunk -> none
    big = _lib.call("peclr_conv2d_wgrad_workspace_bytes", 256, 56, 56, 64, 64, 3, 1)
    assert big > 0 and big % (64 * 9 * 64 * 4) == 0
    assert _lib.call("peclr_conv2d_wgrad_workspace_bytes", 1, 8, 8, 64, 64, 1, 1) == 0
    assert _lib.call("peclr_conv2d_wgrad_workspace_bytes", 4, 7, 7, 64, 64, 3, 2) < 0  # odd size at stride 2
    assert _lib.call("peclr_stem_wgrad_workspace_bytes", 256, 224, 224) > 0
    assert _lib.call("peclr_sgemm_workspace_bytes", 256, 512, 2048) > 256 * 512 * 4
    assert _lib.call("peclr_sgemm_workspace_bytes", 4096, 4096, 64) == 0  # enough tiles: K is not split
    small, large = (_lib.call("peclr_ntxent_workspace_bytes", b, w) for b, w in ((8, 1), (128, 8)))
    assert 0 < small < large


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from peclr_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PeclrKernelError):
        _lib.load()


def test_cpu_tensors_are_rejected(lib_path):
    import torch

    from peclr_b200 import _lib, ops

    with pytest.raises(_lib.PeclrKernelError):
        ops.conv2d_fprop(torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16), torch.zeros(64, 1, 64, dtype=torch.bfloat16), 1, 1)
