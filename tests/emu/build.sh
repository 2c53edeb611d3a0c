#!/bin/bash
# Build the CPU emulation of the fused kernels from the product's own kernel sources (TEST INFRASTRUCTURE ONLY).
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
mkdir -p "$HERE/gen"
# the only source rewrite: "extern __shared__ <type> name[]" (dynamic shared memory) becomes a plain extern of the emulator's array
sed -e 's/extern __shared__/extern/' "$ROOT/pyro_b200/csrc/pyrodp_device.cuh" > "$HERE/gen/pyrodp_device.cuh"
sed -e 's/extern __shared__/extern/' "$ROOT/pyro_b200/csrc/sweep_fused.cuh" > "$HERE/gen/sweep_fused.cuh"
sed -e 's/extern __shared__/extern/' "$ROOT/pyro_b200/csrc/table_kernels.cuh" > "$HERE/gen/table_kernels.cuh"
sed -e 's/extern __shared__/extern/' "$ROOT/pyro_b200/csrc/sweep_mech2.cuh" > "$HERE/gen/sweep_mech2.cuh"
sed -e 's/extern __shared__/extern/' "$ROOT/pyro_b200/csrc/rollout.cuh" > "$HERE/gen/rollout.cuh"
sed -e 's/extern __shared__/extern/' "$ROOT/pyro_b200/csrc/spline.cuh" > "$HERE/gen/spline.cuh"
cp "$ROOT/pyro_b200/csrc/mech2_plan.h" "$HERE/gen/mech2_plan.h"
# -ffp-contract=off: like nvcc -fmad=false, a*b+c is never fused unless the source says fma()
g++ -std=c++17 -O1 -g -ffp-contract=off -fno-fast-math -fPIC -shared -Wno-unused-variable \
    -o "$HERE/libpyrodp_emu.so" "$HERE/emu_main.cpp"
