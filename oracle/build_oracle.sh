#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the CPU restatement oracle/cge_oracle.cpp -> oracle/liboracle.so.
# -O2, no -march, -ffp-contract=off: the reference is built without FMA and so must the oracle be.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
g++ -std=c++17 -O2 -fopenmp -fPIC -ffp-contract=off -Wall -Wno-unused-function -I"$HERE/../include" \
    -shared -o "$HERE/liboracle.so" "$HERE/cge_oracle.cpp"
echo "built $HERE/liboracle.so"
