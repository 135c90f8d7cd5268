"""The C-ABI shared object: loads, exports every symbol include/*.h declares, mirrors the
reference's health check, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        names += re.findall(r"MRB_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_reference_symbols():
    # the five symbols cpp/ls_lib/ls_linux_dll.cpp:8-103 exports
    need = {"set_thread_count", "get_thread_count", "cg_least_squares_from_python",
            "cg_least_squares2_from_python", "als_from_python"}
    assert need <= set(declared_symbols())


def test_library_exports_every_declared_symbol():
    from movie_recommender_b200 import _lib
    for name in declared_symbols():
        assert hasattr(_lib.dll, name), "cpp_ls_lib.so does not export %s" % name


def test_thread_count_round_trip(cpp_ls):
    # python/full_data/cpp_ls.py:23-36
    assert cpp_ls.has_dll_loaded()
    old = cpp_ls.get_thread_count()
    cpp_ls.set_thread_count(17)
    assert cpp_ls.get_thread_count() == 17
    cpp_ls.set_thread_count(old)


def test_python_boundary_signatures(cpp_ls):
    import inspect
    sig = inspect.signature(cpp_ls.als)
    assert list(sig.parameters)[:9] == ["user_ids", "item_ids", "ratings", "num_item_factors",
                                        "num_users", "num_items", "min_r_decrease",
                                        "max_iterations", "algorithm"]
    assert sig.parameters["min_r_decrease"].default == 0.01
    assert sig.parameters["max_iterations"].default == 200
    sig = inspect.signature(cpp_ls.cg_least_squares)
    assert list(sig.parameters)[:8] == ["A_row_indices", "A_col_indices", "A_values",
                                        "A_num_columns", "b", "min_r_decrease", "max_iterations",
                                        "algorithm"]


def test_no_cpu_fallback(cpp_ls):
    """Without a GPU a compute call must raise, never silently compute on the host."""
    from movie_recommender_b200 import _lib
    if _lib.dll.mrb_device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    with pytest.raises(cpp_ls.CppLsError) as e:
        cpp_ls.als(np.zeros(4, np.int32), np.zeros(4, np.int32), np.zeros(4), 2, 1, 1,
                   max_iterations=1)
    assert e.value.code == _lib.ERR_CUDA


def test_product_never_imports_the_oracle():
    """No product file imports, includes, links or loads anything under oracle/ (comments and
    docstrings may cite the oracle's function names)."""
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|"
                     r"(liboracle)|(oracle/_ref)|(oracle[/\\]\w+\.(so|py|c)\b)", re.M)
    for path in glob.glob(os.path.join(ROOT, "movie_recommender_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
            text = open(path, errors="ignore").read()
            # a citation like "oracle/ls_oracle.c:oracle_cosine_topk" in a comment is allowed
            text = re.sub(r"oracle/ls_oracle\.c", "", text)
            assert not bad.search(text), path
