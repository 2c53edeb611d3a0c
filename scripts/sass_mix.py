#!/usr/bin/env python
"""Static instruction mix of a kernel's loops from the SASS of libpyrodp.so (no GPU needed).

    python scripts/sass_mix.py sweep_pendulum_kernelILi1ELb1ELb1ELb1      # mangled-name fragment
    python scripts/sass_mix.py sweep_mech2_kernelILi3ELi1ELb1E

For every backward branch (= loop) it prints the body length, the FP64 / LDG / LDS / select counts and the issue-cycle
cost 2*FP64 + other of one pass through the whole body (an FP64 warp instruction holds the issue
port for two cycles; for the pendulum kernel this count, weighted by ncu's execution frequencies, equals the measured
sweep time to within 1 %).  The body includes rarely taken side paths, so this is an upper bound of the common path;
use it to compare two builds of the same loop before spending GPU time."""
import re
import subprocess
import sys


def main():
    frag = sys.argv[1] if len(sys.argv) > 1 else "sweep_pendulum_kernelILi1ELb1ELb1ELb1"
    lib = sys.argv[2] if len(sys.argv) > 2 else "pyro_b200/libpyrodp.so"
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    ops, on = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            if on:
                break
            on = frag in line
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if on and m:
            ops.append((int(m.group(1), 16), m.group(2).strip()))
    if not ops:
        sys.exit(f"no function matching {frag!r} in {lib}")
    fp64 = re.compile(r"(@!?U?P\d\s+)?D(ADD|MUL|FMA|SETP)")
    print(f"{frag}: {len(ops)} instructions, {sum(bool(fp64.match(i)) for _, i in ops)} FP64")
    for addr, ins in ops:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [i for a, i in ops if tgt <= a <= addr]
            f = sum(bool(fp64.match(i)) for i in body)
            if len(body) < 40:
                continue
            count = lambda pat: sum(1 for i in body if re.search(pat, i))
            print(f"  loop {tgt:#06x}-{addr:#06x}: {len(body):4d} instr, FP64 {f:3d}, LDG {count(r'LDG'):2d}, LDS {count(r'LDS'):2d}, "
                  f"SEL/FSEL {count(r'SEL'):2d}, MOV {count(r'MOV'):3d}, BRA {count(r'BRA'):2d} -> 2F+O = {len(body) + f} issue cycles per pass")


if __name__ == "__main__":
    main()
