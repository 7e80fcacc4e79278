#!/usr/bin/env bash
# builds the headless C++ entry that uses the mirrored reference interface (host/cge_engine.hpp) on top of libcge.so
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
g++ -std=c++17 -O2 -Wall -I"$REPO/include" -I"$REPO/computer-graphics-engine_b200/host" "$HERE/drop_in_main.cpp" \
    -L"$REPO/computer-graphics-engine_b200" -lcge -Wl,-rpath,"$REPO/computer-graphics-engine_b200" -o "$HERE/drop_in_main"
g++ -std=c++17 -O2 -Wall -pthread -I"$REPO/include" -I"$REPO/computer-graphics-engine_b200/host" \
    "$REPO/computer-graphics-engine_b200/host/cge_cli.cpp" -lz \
    -L"$REPO/computer-graphics-engine_b200" -lcge -Wl,-rpath,"$REPO/computer-graphics-engine_b200" -o "$HERE/cge_cli"
# the mirrored scene loaders (host/cge_scene_io.hpp: OBJ / MTL / PNG -> flat scene file); -ffp-contract=off: the reference's loader
# arithmetic (centring, normals) has no fused multiply-add
g++ -std=c++17 -O2 -Wall -ffp-contract=off -I"$REPO/include" -I"$REPO/computer-graphics-engine_b200/host" "$HERE/scene_export_main.cpp" \
    -L"$REPO/computer-graphics-engine_b200" -lcge -lz -Wl,-rpath,"$REPO/computer-graphics-engine_b200" -o "$HERE/scene_export"
