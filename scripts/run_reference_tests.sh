#!/bin/bash
# Run (a subset of) the upstream unit tests against deepspeed_b200 through an import alias (`deepspeed` -> `deepspeed_b200`).
# usage: scripts/run_reference_tests.sh <path to the upstream checkout> [pytest args / test paths relative to its tests/ dir]
# e.g.   scripts/run_reference_tests.sh /root/reference unit/elasticity unit/runtime/test_ds_initialize.py -n 8
set -e
REF=$1; shift
HERE=$(cd "$(dirname "$0")/.." && pwd)
WORK=${REFTEST_WORK:-/tmp/reftests}
rm -rf "$WORK" && mkdir -p "$WORK" && cp -r "$REF/tests" "$WORK/tests"   # the upstream tree may be read-only
cd "$WORK/tests"
PATH="$HERE/bin:$PATH" PYTHONPATH="$HERE/scripts/reftest_shim:$WORK/tests" \
  python -m pytest -q -p no:cacheprovider --timeout "${REFTEST_TIMEOUT:-120}" "${@:-unit/elasticity}"
