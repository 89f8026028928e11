#!/bin/bash
# Regenerates tests/golden/fst_NNN_{raw,compose}.fst from the reference's fixture builders.
# Needs /root/reference (this container only); the generated files are committed.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$REPO/tests/golden"
mkdir -p "$OUT" /tmp/oracle_fixtures
g++ -std=c++17 -O1 -w -I"$HERE" -I"$REF/rustfst-tests-data" "$HERE/gen_fixtures.cpp" -o /tmp/oracle_fixtures/gen_fixtures
( cd "$REF/rustfst-tests-data" && /tmp/oracle_fixtures/gen_fixtures "$OUT" )
# An OpenFst-written "const" file (16-byte aligned layout, version 1) as a reader fixture for const_fst_from_path:
# the HCL machine of fst_012 (215 states / 942 arcs), copied verbatim; its vector twin is fst_012_raw.fst above.
cp "$REF/rustfst-tests-data/fst_012/hcl.fst.in" "$OUT/fst_012_hcl_const.fst.in"
