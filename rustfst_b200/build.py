"""Builds librustfst_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

    python rustfst_b200/build.py            # incremental (run as a script: importing the package dlopens the library)
    python rustfst_b200/build.py --force

The library links the CUDA runtime statically and has no torch dependency; it loads on a CPU-only box (the
container, for the symbol/ABI tests) and reports "no CUDA device" from the compute entry points there.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# B200_LIB_SUFFIX=x builds an experimental variant librustfst_b200_x.so (objects in build_x/) next to the default
# library; rustfst_b200/ffi.py loads it when B200_LIB points at it.
_SUFFIX = os.environ.get("B200_LIB_SUFFIX", "")
OBJ = os.path.join(HERE, "build" + ("_" + _SUFFIX if _SUFFIX else ""))
LIB = os.path.join(HERE, "librustfst_b200" + ("_" + _SUFFIX if _SUFFIX else "") + ".so")
SOURCES = ["capi.cu", "compose.cu", "compose_coop.cu", "compose_ws.cu", "dag_order.cu", "batch.cu", "iso.cu", "connect.cu", "sssp.cu", "nshortest.cu", "device_common.cu", "queue_plan.cpp"]
HEADERS = ["algos.h", "compose_common.cuh", "compose_match.cuh", "bulk_async.cuh", "coop_utils.cuh", "device_common.cuh", "fst_types.h", "host_fst.h", os.path.join("..", "..", "include", "rustfst_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--cudart", "static",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("B200_EXTRA_NVCC_FLAGS", "").split()  # e.g. -DB200_COOP_PROFILE for the in-kernel timeline


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_paths = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs = []
    jobs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(op)
        if force or _newer(sp, op) or any(_newer(h, op) for h in hdr_paths):
            cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", sp, "-o", op]
            jobs.append((src, op, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    rebuilt = bool(jobs)
    failed = []
    for src, op, proc in jobs:
        out, err = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out + err)
        if proc.returncode != 0:
            failed.append(src)
            continue
        with open(op + ".ptxas.log", "w") as f:
            f.write(err)
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
