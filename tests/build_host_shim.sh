#!/usr/bin/env bash
# TEST INFRASTRUCTURE: compiles csrc/tclip_math.cuh as plain C++ (host_math_shim.cpp) so the CPU test-suite can check
# the M-step series arithmetic against SciPy without a GPU.  Output: tests/_build/libtclip_host_math.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/../transductive-clip_b200/csrc"
mkdir -p "$HERE/_build"
g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -I"$SRC" "$SRC/host_math_shim.cpp" -o "$HERE/_build/libtclip_host_math.so"
echo "built $HERE/_build/libtclip_host_math.so"
