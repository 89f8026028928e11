"""Drop-in proof: the reference's OWN Python package (rustfst-python/rustfst, unmodified) is pointed at
librustfst_b200.so — it loads the first *.so next to its ffi_utils.py (rustfst-python/rustfst/ffi_utils.py:16-22) —
and the reference's own test files are run against it.

Nothing is copied into the repository: package and tests are copied from /root/reference into a temporary directory
at test time, so these tests only run where the reference checkout exists (the development container); on the GPU
box they skip.  Symbol tables are outside the scope of this library (presentation metadata), so
the four test_fst.py cases that build a SymbolTable are expected to fail with a missing `symt_*` symbol and nothing
else is.
"""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/rustfst-python"
LIB = os.path.join(ROOT, "rustfst_b200", "librustfst_b200.so")

needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rustfst")),
                                     reason="the reference checkout is not present on this machine")

# test_fst.py cases that construct rustfst.SymbolTable (symt_new / symt_add_symbol / ... are not exported)
EXPECTED_SYMT_FAILURES = {
    "tests/test_fst.py::test_fst_read_write_with_symt",
    "tests/test_fst.py::test_fst_symt",
    "tests/test_fst.py::test_fst_with_symt_mut_fail",
    "tests/test_fst.py::test_fst_relabel_tables",
}


def _run_reference_tests(tmp_path, selection):
    work = tmp_path / "dropin"
    shutil.copytree(os.path.join(REF, "rustfst"), work / "rustfst")
    shutil.copytree(os.path.join(REF, "tests"), work / "tests")
    for stale in (work / "rustfst").glob("*.so"):
        stale.unlink()
    shutil.copy(LIB, work / "rustfst" / "librustfst_b200.so")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)  # the copy in the temporary directory must be the `rustfst` that gets imported
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-rf", *selection],
                       cwd=work, env=env, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    failed = set(re.findall(r"^FAILED (\S+)", out, flags=re.M))
    m = re.search(r"(\d+) passed", out)
    return failed, int(m.group(1)) if m else 0, out


@needs_reference
def test_reference_python_package_container_and_host_algorithms(tmp_path):
    """VectorFst / Tr / TrsVec / iterators and the host-side algorithms, through the unmodified reference client."""
    failed, passed, out = _run_reference_tests(tmp_path, [
        "tests/test_fst.py", "tests/test_tr.py", "tests/test_trs.py", "tests/test_iterator.py",
        "tests/algorithms/test_tr_sort.py", "tests/algorithms/test_isomorphic.py", "tests/algorithms/test_top_sort.py"])
    assert failed == EXPECTED_SYMT_FAILURES, out[-4000:]
    assert passed == 22, out[-4000:]


@needs_reference
@pytest.mark.gpu
def test_reference_python_package_device_algorithms(tmp_path):
    """The reference's own known-answer tests of the hot path (compose incl. configs and sigma matchers, shortest
    path, connect, reverse) through the unmodified reference client; these entry points run on the GPU."""
    failed, passed, out = _run_reference_tests(tmp_path, [
        "tests/algorithms/test_compose.py", "tests/algorithms/test_shortest_path.py",
        "tests/algorithms/test_connect.py", "tests/algorithms/test_reverse.py"])
    assert not failed, out[-4000:]
    assert passed >= 6, out[-4000:]
