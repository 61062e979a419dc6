"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/nbm_b200.h declares; host-only entry points work without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from birdsoundclassif_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_header_symbols_are_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "nbm_b200.h")).read()
    declared = set(re.findall(r"\b(nbm_[a-z_0-9]+)\s*\(", header))
    declared -= {"nbm_status"}
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nbm_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_version_and_error_string(lib):
    assert lib.nbm_version() == 100
    assert isinstance(lib.nbm_last_error(), bytes)


def test_make_anchors_host_only(lib):
    from oracle import postproc_oracle as po
    ratios = (C.c_double * 3)(0.5, 1.0, 2.0)
    scales = (C.c_int64 * 5)(1, 2, 4, 8, 16)
    out = np.empty((24 * 64 * 15, 4), dtype=np.float32)
    assert lib.nbm_make_anchors(16, ratios, 3, scales, 5, 64, 24, 16, out.ctypes.data) == 0
    np.testing.assert_array_equal(out, po.make_anchors())
    # 5-ratio variant (setattr_others, nets_utils.py:408-409)
    r5 = (C.c_double * 5)(0.2, 0.5, 1, 2, 5)
    out5 = np.empty((23 * 64 * 20, 4), dtype=np.float32)
    s4 = (C.c_int64 * 4)(1, 2, 4, 8)
    assert lib.nbm_make_anchors(16, r5, 5, s4, 4, 64, 23, 16, out5.ctypes.data) == 0
    np.testing.assert_array_equal(out5, po.make_anchors(16, (0.2, 0.5, 1, 2, 5), (1, 2, 4, 8), 64, 23, 16))


def test_bad_arguments_return_codes(lib):
    assert lib.nbm_make_anchors(16, None, 3, None, 5, 64, 24, 16, None) == -1
    assert b"bad argument" in lib.nbm_last_error()
    assert lib.nbm_nms_workspace_bytes(4, 500) == 4 * 500 * 8 * 8


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "birdsoundclassif_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_ctypes_argument_counts_match_the_header():
    """Every declaration of include/nbm_b200.h against the ctypes table: same number of parameters, pointer parameters
    bound as pointers (a drifted binding would pass garbage without any error at the call)."""
    header = open(os.path.join(ROOT, "include", "nbm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    header = re.sub(r"//[^\n]*", "", header)
    decls = re.findall(r"\b(nbm_[a-z_0-9]+)\s*\(([^;{}]*?)\)\s*;", header, flags=re.S)
    seen = 0
    for name, args in decls:
        if name not in _lib.SIGNATURES:
            continue
        params = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
        _, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, ctypes table {len(argtypes)}"
        for p, t in zip(params, argtypes):
            is_ptr = "*" in p or "cudaStream_t" in p
            bound_ptr = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or getattr(t, "_type_", None) == "P"
            if is_ptr:
                assert bound_ptr or issubclass(t, C._Pointer) or t is C.c_void_p, f"{name}: `{p}` bound as {t}"
        seen += 1
    assert seen == len(_lib.SIGNATURES)
