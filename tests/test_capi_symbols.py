"""The C-ABI library loads on a CPU-only box and exports every symbol include/rustfst_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rustfst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(?:RUSTFST_FFI_RESULT|const char\*)\s+([a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "rustfst_b200", "librustfst_b200.so"))
    names = declared_symbols()
    assert len(names) > 90
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_hot_path_symbols_are_the_reference_names():
    names = set(declared_symbols())
    for n in ("fst_compose", "fst_compose_with_config", "fst_compose_config_new", "fst_compose_config_destroy",
              "fst_matcher_config_new", "fst_matcher_config_destroy", "fst_shortest_path",
              "fst_shortest_path_with_config", "fst_shortest_path_config_new", "fst_connect", "fst_tr_sort",
              "vec_fst_from_bytes", "vec_fst_to_bytes", "rustfst_ffi_get_last_error", "rustfst_destroy_string"):
        assert n in names


def test_product_does_not_depend_on_the_oracle():
    """The shipped sources must never include or link anything under oracle/."""
    csrc = os.path.join(ROOT, "rustfst_b200")
    for dirpath, _, files in os.walk(csrc):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle/ ", ""), f"{f} mentions the oracle"


def test_stats_structs_have_the_layout_of_the_public_header(tmp_path):
    """The ctypes mirrors of B200ComposeStats / B200SsspStats (rustfst_b200/ffi.py) against what a C compiler sees in
    include/rustfst_b200.h: size and the offset of every field (tests/cpp/abi_layout.c prints them)."""
    import subprocess
    from rustfst_b200 import ffi
    exe = tmp_path / "abi_layout"
    subprocess.run(["gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "abi_layout.c"),
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        name, value = line.split()
        struct, field = name.split(".")
        mirror = {"B200ComposeStats": ffi.ComposeStats, "B200SsspStats": ffi.SsspStats}[struct]
        if field == "sizeof":
            assert ctypes.sizeof(mirror) == int(value), (name, ctypes.sizeof(mirror), value)
        else:
            assert getattr(mirror, field).offset == int(value), (name, getattr(mirror, field).offset, value)
        seen += 1
    assert seen >= 24
