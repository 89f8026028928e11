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
