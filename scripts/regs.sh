#!/bin/bash
# register / spill report and loop instruction mix of the sweep kernels (no GPU needed)
# usage: bash scripts/regs.sh [name-fragment] [extra nvcc flags...]
FRAG=${1:-mech2}; shift
cd "$(dirname "$0")/../pyro_b200/csrc"
OUT=../../gpurun_out/scratch; mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xptxas -v -cubin "$@" -o $OUT/pyrodp.cubin pyrodp.cu > $OUT/ptxas.log 2>&1 || { cat $OUT/ptxas.log; exit 1; }
python - "$FRAG" <<'PY'
import re, sys
frag = sys.argv[1]
name = None
for line in open('../../gpurun_out/scratch/ptxas.log'):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m: name = m.group(1)
    if name and frag in name:
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m: spill = m.groups()
        m = re.search(r"Used (\d+) registers", line)
        if m: print(f"{name[:64]:64s} regs {m.group(1):>3s}  stack/spill {spill}")
PY
