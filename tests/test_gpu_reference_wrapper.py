"""The drop-in claim of INTEGRATION.md section 1, pinned: the reference's OWN, unmodified ctypes
wrapper and test script (cpp/python/cpp_ls.py + cpp_ls_test.py, shipped here only as bytecode
compiled by oracle/Makefile into oracle/_ref/) run against THIS repo's cpp_ls_lib.so placed in the
working directory, exactly how the reference loads its library (cpp_ls.py:5-14,
``os.getcwd() + "/cpp_ls_lib.so"``).  The script prints one line per check: the load / thread
count round trip, its planted least-squares problem, its planted ALS problem."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _stage(tmp_path, library):
    for mod in ("cpp_ls", "cpp_ls_test"):
        src = os.path.join(REF_DIR, mod + ".pyc.bin")
        if not os.path.exists(src):
            pytest.skip("oracle/_ref/%s.pyc.bin was not built (reference sources absent at build time)" % mod)
        shutil.copy(src, tmp_path / (mod + ".pyc"))
    shutil.copy(library, tmp_path / "cpp_ls_lib.so")


def _run(tmp_path):
    # the reference's thresholds are statistical on unseeded data ("an average error of under
    # 0.10 is ideal, but does not always happen", cpp_ls_test.py:139-141): allow a second draw
    out = None
    for _ in range(3):
        out = subprocess.run([sys.executable, "cpp_ls_test.pyc"], cwd=tmp_path, capture_output=True,
                             text=True, timeout=600)
        if out.returncode == 0 and out.stdout.count("pass - ") == 3 and "FAIL" not in out.stdout:
            return out.stdout
    raise AssertionError((out.stdout[-3000:], out.stderr[-3000:]))


@pytest.mark.gpu
def test_reference_wrapper_and_its_test_script_on_our_library(require_gpu, tmp_path):
    _stage(tmp_path, os.path.join(ROOT, "movie_recommender_b200", "cpp_ls_lib.so"))
    stdout = _run(tmp_path)
    assert "pass - has_dll_loaded() is True" in stdout
    assert "pass - cpp_ls.cg_least_squares() passed testing" in stdout
    assert "pass - cpp_ls.als() passed testing" in stdout


def test_staging_recipe_against_the_reference_library(tmp_path):
    """CPU: the same staging with the reference's own library -- proves the recipe (bytecode of
    the unmodified scripts + a cpp_ls_lib.so in the working directory) is what the reference runs."""
    ref_lib = os.path.join(REF_DIR, "cpp_ls_lib.so")
    if not os.path.exists(ref_lib):
        pytest.skip("oracle/_ref/cpp_ls_lib.so not built")
    _stage(tmp_path, ref_lib)
    assert _run(tmp_path).count("pass - ") == 3
