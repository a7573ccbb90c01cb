"""CPU checks of the drop-in boundary: libpmaf.so builds for sm_100a, loads, exports every symbol
include/pmaf.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

from pmaf_b200 import planner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    planner.build()
    return planner.load_library()


def test_header_and_symbol_list_agree():
    hdr = open(os.path.join(ROOT, "include", "pmaf.h")).read()
    declared = set(re.findall(r"\b(pmaf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(planner.API_SYMBOLS), declared ^ set(planner.API_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in planner.API_SYMBOLS:
        assert hasattr(lib, name), f"libpmaf.so does not export {name}"
    assert b"sm_100a" in lib.pmaf_version()


def test_null_handle_is_an_argument_error(lib):
    assert lib.pmaf_start_prediction(None) == -1
    assert b"null planner" in lib.pmaf_last_error()


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    h = C.c_void_p()
    rc = lib.pmaf_create(C.byref(h), 0)
    assert rc == -3 and not h.value  # PMAF_ERR_CUDA
    assert lib.pmaf_last_error()
    with pytest.raises(planner.PmafError):
        planner.CfManager(0)


def test_only_sm100a_code_is_embedded(lib):
    out = os.popen(f"cuobjdump -lelf {planner.LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
