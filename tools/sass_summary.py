#!/usr/bin/env python
"""Opcode histogram of every kernel in libkmc_b200.so from `cuobjdump -sass` (no GPU needed): the evidence that the library
holds sm_100a code only, that the hot kernels move data with 256-bit LDG/STG (LDG.E.ENL2.256 / STG.E.ENL2.256), that the staged
variants use the bulk-copy engine (UBLKCP + SYNCS mbarrier ops), and that nothing spills (no LDL/STL).
  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kitti_motion_compensation_b200", "lib", "libkmc_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    print("embedded cubins:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))), f"({len(elf.splitlines())} ELF sections listed)")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    current = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            current = m.group(1)
            kernels[current] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and current:
            kernels[current][m.group(1)] += 1
    names = demangle(list(kernels))
    interesting = re.compile(r"^(LDG|STG|LDL|STL|LDS|STS|UBLKCP|SYNCS|MUFU|SHFL|ATOM|RED|BAR|LDC|ULDC|DFMA|DADD|DMUL|FFMA|FMUL|FADD|F2F|HMMA|UTCMMA)")
    print(f"{len(kernels)} kernels\n")
    for k, hist in kernels.items():
        total = sum(hist.values())
        name = names.get(k, k)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        print(f"== {name[:150]}")
        print(f"   {total} SASS instructions; memory / special ops:")
        keep = {op: c for op, c in hist.items() if interesting.match(op)}
        groups = collections.defaultdict(list)
        for op, c in sorted(keep.items()):
            groups[re.match(r"[A-Z0-9]+", op).group(0)].append(f"{op} x{c}")
        for g in ("LDG", "STG", "LDL", "STL", "LDS", "STS", "UBLKCP", "SYNCS", "LDC", "ULDC", "SHFL", "ATOM", "RED", "BAR", "MUFU", "FFMA", "FMUL", "FADD", "DFMA", "DADD",
                  "DMUL", "F2F", "HMMA", "UTCMMA"):
            if g in groups:
                print(f"     {', '.join(groups[g])}")
        spills = sum(c for op, c in hist.items() if op.startswith(("LDL", "STL")))
        print(f"   local-memory (spill) instructions: {spills}\n")


if __name__ == "__main__":
    sys.exit(main())
