"""fst_reverse computes its property word from one OR-reduction over the reversed arcs' "events" instead of replaying
add_tr_properties arc by arc (mutate_properties.rs:43-100).  tests/cpp/props_events_check.cpp fuzzes the claim that
the two are equal for any arc sequence and any starting word; it is compiled here with the host compiler against the
product header."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arc_event_reduction_equals_sequential_property_replay(tmp_path):
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no host C++ compiler")
    exe = str(tmp_path / "props_check")
    subprocess.check_call([cxx, "-std=c++17", "-O2", os.path.join(ROOT, "tests", "cpp", "props_events_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
